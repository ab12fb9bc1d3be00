"""6-mer pore-model table and expected-signal lookup.

Mirrors the reference's ``PoreModel`` (squiggler/pore_model.py:13-47) and
``Squiggler._generate_signal`` (squiggler/Squiggler.py:20-28).  The reference finds
a k-mer by a boolean scan over all 4096 table rows per lookup; the table is in
lexicographic ACGT order, so a k-mer's row is its base-4 value and the lookup is a
gather.  Sequence-level lookups (:meth:`PoreModel.generate_signal`) run on the GPU
(kernel ``pore_lookup_kernel`` behind ``wstr_pore_lookup``); the 4096-row MAD
normalisation of the table itself is a one-off host step (float input, see
SURVEY Appendix C last paragraph).
"""
import os
import re
from typing import Dict, Iterable, Optional, Tuple

import numpy as np

from .config import DEFAULT_PORE_MODEL
from .templates import DNA_DICT

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _b in enumerate('ACGT'):
    _CODE[ord(_b)] = _i


def mad_normalize_float(data: np.ndarray) -> np.ndarray:
    """Median/MAD normalisation of a float vector with the reference's arithmetic
    (schemas/fast5.py:104-114): shift = mean of the 46.5 and 53.5 percentiles,
    scale = median(|x - shift|)."""
    data = np.asarray(data, dtype=np.float64)
    shift = np.mean(np.percentile(data, (46.5, 53.5)))
    scale = np.median(np.abs(data - shift))
    return np.asarray((data - shift) / scale)


class PoreModel:
    def __init__(self, pore_model_path: str = DEFAULT_PORE_MODEL) -> None:
        if not os.path.exists(pore_model_path):
            raise FileNotFoundError('Not found pore model table at path', pore_model_path)
        with open(pore_model_path, 'r') as fh:
            header = fh.readline().rstrip('\n').split('\t')
            if 'kmer' not in header or 'level_mean' not in header:
                raise ValueError('Pore model table do not contains "kmer" and "level_mean" columns')
            ck, cv = header.index('kmer'), header.index('level_mean')
            kmers, means = [], []
            for line in fh:
                parts = line.rstrip('\n').split('\t')
                if len(parts) <= max(ck, cv):
                    continue
                kmers.append(parts[ck])
                means.append(float(parts[cv]))
        self.kmersize = len(kmers[0])
        self.kmers = kmers
        self.level_mean = np.array(means, dtype=np.float64)
        self.level_norm = mad_normalize_float(self.level_mean)
        # row of every k-mer by its base-4 value; falls back to a dict when the file is
        # not in lexicographic order
        n = 4 ** self.kmersize
        self._lex = len(kmers) == n and all(self._index(kmers[i]) == i for i in (0, 1, n // 3, n - 1)) \
            and all(self._index(km) == i for i, km in enumerate(kmers))
        self._row = None if self._lex else {km: i for i, km in enumerate(kmers)}
        if self._lex:
            self.table = self.level_norm
        else:
            self.table = np.full(n, np.nan)
            for i, km in enumerate(kmers):
                self.table[self._index(km)] = self.level_norm[i]
        self._dev = {}

    @staticmethod
    def _index(kmer: str) -> int:
        idx = 0
        for ch in kmer:
            c = _CODE[ord(ch)]
            if c == 255:
                raise IndexError(f'k-mer {kmer!r} is not over ACGT')
            idx = idx * 4 + int(c)
        return idx

    def get_value(self, kmer: str) -> float:
        """Normalised level of one k-mer (reference: pore_model.py:45-47; an unknown
        k-mer raises IndexError there as well)."""
        if len(kmer) != self.kmersize:
            raise IndexError(f'k-mer {kmer!r} not in pore model table')
        val = self.table[self._index(kmer)]
        if np.isnan(val):
            raise IndexError(f'k-mer {kmer!r} not in pore model table')
        return val

    def get_values(self, kmers: Iterable[str]) -> np.ndarray:
        return np.array([self.get_value(km) for km in kmers], dtype=np.float64)

    # -- GPU expected-signal generation ---------------------------------------------
    def device_table(self, device):
        """The 4**k level table resident on ``device`` (a torch tensor, f64)."""
        import torch
        key = str(device)
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(np.ascontiguousarray(self.table)).to(device)
        return self._dev[key]

    def generate_signal(self, sequence: str, device='cuda') -> np.ndarray:
        """Expected normalised signal of a plain ACGT sequence: one level per sliding
        k-mer (reference: Squiggler.py:20-28).  Runs ``wstr_pore_lookup`` on the GPU."""
        from . import _lib
        import torch
        seq = np.frombuffer(sequence.encode('ascii'), dtype=np.uint8)
        n_out = len(seq) - self.kmersize + 1
        if n_out <= 0:
            return np.zeros(0, dtype=np.float64)
        dev = torch.device(device)
        d_seq = torch.from_numpy(seq.copy()).to(dev)
        d_out = torch.empty(n_out, dtype=torch.float64, device=dev)
        d_bad = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.pore_lookup(d_seq, self.device_table(dev), self.kmersize, d_out, d_bad)
        if int(d_bad.item()) != 0:
            raise IndexError('sequence contains a k-mer that is not in the pore model table')
        return d_out.cpu().numpy()

    # -- state-similarity report (reference: pore_model.py:35-71) --------------------
    def _get_consecutive_diff(self, pattern: str) -> Tuple[float, float]:
        rep = pattern * self.kmersize
        levels = [self.get_value(rep[i:i + self.kmersize]) for i in range(len(pattern) + 1)]
        steps = np.abs(np.diff(levels))
        return np.mean(steps), np.median(steps)

    def get_diffs_for_all(self, sequence: str) -> Dict[str, Tuple[float, float]]:
        diffs: Dict[str, Tuple[float, float]] = {}
        for group in re.findall(r'[\(\{].*?[\)\}]', sequence):
            letters = [c for c in group if c not in '(){}']
            expanded = ['']
            for ch in letters:
                if ch in DNA_DICT:
                    expanded = [p + alt for alt in DNA_DICT[ch] for p in expanded]
                else:
                    expanded = [p + ch for p in expanded]
            for p in expanded:
                diffs[p] = self._get_consecutive_diff(p)
        return diffs


_default: Optional[PoreModel] = None


def get_pore_model(path: Optional[str] = None) -> PoreModel:
    """Process-wide table (the reference keeps a module-level singleton,
    pore_model.py:74)."""
    global _default
    if path is not None:
        return PoreModel(path)
    if _default is None:
        _default = PoreModel(DEFAULT_PORE_MODEL)
    return _default
