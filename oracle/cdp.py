"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/dp_oracle.c."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libdp_oracle.so')
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, 'dp_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', _HERE, 'libdp_oracle.so'])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.wso_warp_range.restype = ctypes.c_int
        _lib.wso_warp.restype = ctypes.c_int
        _lib.wso_backtrack.restype = ctypes.c_int
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty)) if a is not None else None


def _tb_arrays(tb):
    return (np.ascontiguousarray(tb.values, dtype=np.float64),
            np.ascontiguousarray(tb.seq_idx, dtype=np.int32),
            np.ascontiguousarray(tb.in_ptr, dtype=np.int32),
            np.ascontiguousarray(tb.in_idx, dtype=np.int32))


def fill(x, tb, mask, mv, flank_length):
    x = np.ascontiguousarray(x, dtype=np.float64)
    v, sq, ip, ii = _tb_arrays(tb)
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    D = np.empty((len(x), len(v)), dtype=np.float64)
    lib().wso_fill(_p(x, ctypes.c_double), len(x), _p(v, ctypes.c_double), _p(sq, ctypes.c_int32),
                   _p(ip, ctypes.c_int32), _p(ii, ctypes.c_int32), len(v), _p(m, ctypes.c_uint8),
                   int(mv), int(flank_length), _p(D, ctypes.c_double))
    return D


def warp(x, tb, mask, mv, flank_length):
    x = np.ascontiguousarray(x, dtype=np.float64)
    v, sq, ip, ii = _tb_arrays(tb)
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    trace = np.empty(len(x), dtype=np.int32)
    rc = lib().wso_warp(_p(x, ctypes.c_double), len(x), _p(v, ctypes.c_double), _p(sq, ctypes.c_int32),
                        _p(ip, ctypes.c_int32), _p(ii, ctypes.c_int32), len(v), int(tb.endstate),
                        _p(m, ctypes.c_uint8), int(mv), int(flank_length), _p(trace, ctypes.c_int32), None)
    if rc == 1:
        raise RuntimeError('Unexpected error during backtracking')
    if rc:
        raise RuntimeError(f'dp_oracle failed rc={rc}')
    return trace.astype(int)


def warp_batch(signals, tb, masks, mv, flank_length, threads=1):
    """signals: list of f64 arrays; masks: list of bool arrays or None.  Reads are split
    over ``threads`` Python threads (the C call releases the GIL).  Returns (traces, threads)."""
    from concurrent.futures import ThreadPoolExecutor
    n = len(signals)
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in signals])
    x = np.ascontiguousarray(np.concatenate(signals), dtype=np.float64)
    m = None if masks is None else np.ascontiguousarray(np.concatenate(masks), dtype=np.uint8)
    v, sq, ip, ii = _tb_arrays(tb)
    trace = np.empty(len(x), dtype=np.int32)
    status = np.zeros(n, dtype=np.int32)
    fn = lib().wso_warp_range

    def work(r):
        fn(_p(x, ctypes.c_double), _p(off, ctypes.c_int64), r, r + 1,
           _p(v, ctypes.c_double), _p(sq, ctypes.c_int32), _p(ip, ctypes.c_int32),
           _p(ii, ctypes.c_int32), len(v), int(tb.endstate), _p(m, ctypes.c_uint8),
           int(mv), int(flank_length), _p(trace, ctypes.c_int32), _p(status, ctypes.c_int32))

    threads = max(1, int(threads))
    if threads == 1:
        for r in range(n):
            work(r)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, range(n)))
    if status.any():
        raise RuntimeError(f'dp_oracle batch failed: {status[status != 0][:5]}')
    return [trace[off[r]:off[r + 1]] for r in range(n)], threads
