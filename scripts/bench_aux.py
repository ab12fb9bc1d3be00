#!/usr/bin/env python
"""Roofline check of the two HBM-bound side kernels: per-read normalisation (kernel 2) and the
pore-model lookup (kernel 1).  Prints one JSON line per kernel (achieved GB/s of algorithmic
bytes against MEASURED_PEAKS.json: hbm_gbs)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from warpstr_b200 import _lib  # noqa: E402
from warpstr_b200.pore_model import get_pore_model  # noqa: E402

peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- normalisation: n reads of N int16 samples, window of T samples each ---------------------------
rng = np.random.default_rng(7)
SHAPES = ((4000, 150000, 3300),) if os.environ.get('AUX_ONLY_NORM') else ((4000, 150000, 3300), (20000, 60000, 3300))
for n, N, T in SHAPES:
    base = (450 + 40 * rng.standard_normal(N)).astype(np.int16)
    raw = np.tile(base, n)
    raw[rng.integers(0, raw.size, raw.size // 10000)] = 1500       # spikes
    raw_off = np.arange(n + 1, dtype=np.int64) * N
    lo = rng.integers(1000, N - T - 1000, n).astype(np.int32)
    hi = (lo + T - 1).astype(np.int32)
    out_off = np.arange(n, dtype=np.int64) * T
    d_raw = torch.from_numpy(raw).cuda()
    d_out = torch.empty(n * T, dtype=torch.float64, device='cuda')
    d_ss = torch.empty(2 * n, dtype=torch.float64, device='cuda')
    ws = torch.empty(_lib.normalize_workspace_bytes(n), dtype=torch.uint8, device='cuda')
    ms = timed(lambda: _lib.normalize_batch(d_raw, raw_off, lo, hi, 1, d_out, out_off, d_ss, ws))
    if os.environ.get('AUX_NORM_PHASES'):        # a -DWSTR_NORM_TIMING build: clock64 ticks per phase, summed over CTAs
        import ctypes
        lib = ctypes.CDLL(_lib.LIB_PATH)
        buf = (ctypes.c_ulonglong * 8)()
        lib.wstr_debug_norm_phases(buf, 1)
        _lib.normalize_batch(d_raw, raw_off, lo, hi, 1, d_out, out_off, d_ss, ws)
        torch.cuda.synchronize()
        lib.wstr_debug_norm_phases(buf, 0)
        tot = float(sum(buf)) or 1.0
        names = ['-', 'meta+estimate', 'scan', 'patch', 'reduce', 'stats', 'convert', 'queue']
        print(json.dumps({'norm_phase_share': {k: round(v / tot, 3) for k, v in zip(names, buf) if k != '-'},
                          'ticks_per_read': {k: int(v / n) for k, v in zip(names, buf) if k != '-'}}))
    alg = raw.nbytes + n * T * 8                                    # 2 B/raw sample read once + 8 B/window sample written
    print(json.dumps({'kernel': 'normalize_kernel', 'reads': n, 'samples_per_read': N, 'window': T, 'ms': ms,
                      'algorithmic_bytes': alg, 'achieved_gbs': alg / ms / 1e6, 'peak_gbs': peak,
                      'frac': alg / ms / 1e6 / peak, 'reads_per_s': n / ms * 1e3}))
    del d_raw, d_out

if os.environ.get('AUX_ONLY_NORM'):
    sys.exit(0)
# ---- pore lookup: one long sequence ---------------------------------------------------------------------
pm = get_pore_model()
L = 1 << 26
seq = rng.integers(0, 4, L).astype(np.uint8)
seq = np.frombuffer(b'ACGT', dtype=np.uint8)[seq]
d_seq = torch.from_numpy(seq).cuda()
d_tab = torch.from_numpy(np.ascontiguousarray(pm.table)).cuda()
d_o = torch.empty(L - 5, dtype=torch.float64, device='cuda')
d_bad = torch.zeros(1, dtype=torch.int32, device='cuda')
ms = timed(lambda: _lib.pore_lookup(d_seq, d_tab, 6, d_o, d_bad))
alg = L + (L - 5) * 8
print(json.dumps({'kernel': 'pore_lookup_kernel', 'bases': L, 'ms': ms, 'algorithmic_bytes': alg,
                  'achieved_gbs': alg / ms / 1e6, 'peak_gbs': peak, 'frac': alg / ms / 1e6 / peak}))
