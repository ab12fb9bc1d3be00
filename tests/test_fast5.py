"""CPU: the fast5 reader (warpstr_b200/fast5.py) -- VBZ decoding against an encoder written here
from the published format, and the HDF5 walk against the reference's bundled multi-read file
(build container only; its raw reads are also committed, delta-coded, in golden/c1_bundled.npz)."""
import csv
import ctypes
import os
import struct

import numpy as np
import pytest

from warpstr_b200 import fast5

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
REF = '/root/reference'
BUNDLED = os.path.join(REF, 'test/test_input/test_run1/fast5s/batch_0.fast5')


def _zstd_compress(data: bytes) -> bytes:
    lib = fast5._libzstd()
    lib.ZSTD_compressBound.restype = ctypes.c_size_t
    lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
    lib.ZSTD_compress.restype = ctypes.c_size_t
    lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    cap = lib.ZSTD_compressBound(len(data))
    out = ctypes.create_string_buffer(cap)
    n = lib.ZSTD_compress(out, cap, data, len(data), 1)
    assert not lib.ZSTD_isError(n)
    return out.raw[:n]


def _vbz_encode(x: np.ndarray, version: int, zstd: bool = True) -> bytes:
    """int16 samples -> one VBZ chunk: zig-zag of the deltas, StreamVByte (32-bit codes in
    version 0, the 16-bit variant in version 1), optional zstd, uint32 byte-size header."""
    x = x.astype(np.int64)
    d = np.diff(x, prepend=0)
    if version == 0:
        zz = ((d << 1) ^ (d >> 31)) & 0xFFFFFFFF
        ctrl = bytearray((len(x) + 3) // 4)
        data = bytearray()
        for i, v in enumerate(zz.tolist()):
            nb = 1 if v < 1 << 8 else 2 if v < 1 << 16 else 3 if v < 1 << 24 else 4
            ctrl[i >> 2] |= (nb - 1) << (2 * (i & 3))
            data += int(v).to_bytes(nb, 'little')
    else:
        d = ((d + 32768) % 65536) - 32768      # 16-bit arithmetic wraps
        zz = ((d << 1) ^ (d >> 15)) & 0xFFFF
        ctrl = bytearray((len(x) + 7) // 8)
        data = bytearray()
        for i, v in enumerate(zz.tolist()):
            nb = 1 if v < 1 << 8 else 2
            ctrl[i >> 3] |= (nb - 1) << (i & 7)
            data += int(v).to_bytes(nb, 'little')
    body = bytes(ctrl) + bytes(data)
    if zstd:
        body = _zstd_compress(body)
    return struct.pack('<I', 2 * len(x)) + body


@pytest.mark.parametrize('version', [0, 1])
@pytest.mark.parametrize('n', [0, 1, 3, 4, 5, 8, 9, 1000, 4097])
def test_vbz_round_trip(version, n):
    rng = np.random.default_rng(100 * version + n)
    x = (450 + np.cumsum(rng.integers(-40, 41, size=n))).astype(np.int16)
    if n > 10:
        x[7] = 32767          # extreme jumps need the wide codes
        x[8] = -32768
    for zstd in (True, False):
        chunk = _vbz_encode(x, version, zstd)
        raw = fast5.vbz_decompress(chunk, (version, 2, 1, 1 if zstd else 0))
        assert np.array_equal(np.frombuffer(raw, dtype='<i2'), x)


def test_vbz_rejects_truncated_chunk():
    x = np.arange(100, dtype=np.int16)
    chunk = _vbz_encode(x, 0, zstd=False)
    with pytest.raises(fast5.Fast5FormatError):
        fast5.vbz_decompress(chunk[:-5], (0, 2, 1, 0))
    with pytest.raises(fast5.Fast5FormatError):
        fast5.vbz_decompress(b'\x01', (0, 2, 1, 0))


def test_not_hdf5(tmp_path):
    p = tmp_path / 'x.fast5'
    p.write_bytes(b'not an hdf5 file at all')
    with pytest.raises(fast5.Fast5FormatError):
        fast5.H5File(str(p))
    (tmp_path / 'empty.fast5').write_bytes(b'')
    with pytest.raises(fast5.Fast5FormatError):
        fast5.H5File(str(tmp_path / 'empty.fast5'))


@pytest.mark.reference
@pytest.mark.skipif(not os.path.exists(BUNDLED), reason='reference tree not mounted')
def test_bundled_multi_read_file():
    z = np.load(os.path.join(GOLD, 'c1_bundled.npz'))
    names = fast5.read_names(BUNDLED)
    assert sorted(names) == sorted(str(n) for n in z['names'])
    with fast5.H5File(BUNDLED) as h5:
        assert h5.keys(f'read_{names[0]}') == ['Analyses', 'Raw', 'channel_id', 'context_tags', 'tracking_id']
        assert 'read_nope' not in h5
        with pytest.raises(KeyError):
            h5.dataset('read_nope/Raw/Signal')
    rows = list(csv.DictReader(open(os.path.join(REF, 'test/test_caller_only/example.csv'))))
    for i, r in enumerate(rows):
        raw = fast5.raw_signal(BUNDLED, r['read_name'])
        assert raw.dtype == np.int16
        assert np.array_equal(raw, np.cumsum(z[f'raw_delta{i}'].astype(np.int64)).astype(np.int16))
        assert int(r['r_end_raw']) < len(raw) and 200 < raw.min() and raw.max() < 1000
    with pytest.raises(KeyError):
        fast5.raw_signal(BUNDLED, 'no-such-read')


@pytest.mark.reference
@pytest.mark.skipif(not os.path.exists(BUNDLED), reason='reference tree not mounted')
def test_caller_only_prepare(tmp_path):
    """prepare_caller_only.py's contract: overview.csv per locus with run_id and saved columns."""
    import pandas as pd
    from warpstr_b200 import caller_only
    src = tmp_path / 'reads.csv'
    with open(os.path.join(REF, 'test/test_caller_only/example.csv')) as fh:
        src.write_text(fh.read().replace('test/test_input', os.path.join(REF, 'test/test_input')))
    written = caller_only.prepare(str(tmp_path / 'out'), str(src))
    assert list(written) == ['Human_STR_1108232']
    df = pd.read_csv(written['Human_STR_1108232'])
    assert list(df.columns) == caller_only.REQUIRED + ['run_id', 'saved']
    assert len(df) == 10 and (df.saved == 1).all() and (df.run_id == 'run_0').all()
    bad = tmp_path / 'bad.csv'
    bad.write_text('fast5_path,locus,read_name\n')
    with pytest.raises(ValueError):
        caller_only.prepare(str(tmp_path / 'out2'), str(bad))
    missing = tmp_path / 'missing.csv'
    missing.write_text(src.read_text().replace('0592ed32', 'ffffffff'))
    with pytest.raises(ValueError):
        caller_only.prepare(str(tmp_path / 'out3'), str(missing))
