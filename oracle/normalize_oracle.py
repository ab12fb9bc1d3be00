"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's per-read normalisation
(/root/reference/src/schemas/fast5.py:45-57, 68-77, 90-114).  numpy's percentile / median and
scipy's medfilt are the reference's own third-party calls (numpy pinned 1.20, scipy 1.6.3 there;
2.3.5 / 1.18.1 in this container, which is what the goldens in tests/golden/normalize.npz were
made with).  Pinned by tests/test_oracle_golden.py against those goldens and, in the build
container, against the unmodified functions."""
import numpy as np
from scipy.signal import medfilt

SPIKE_MODES = {'None': 0, 'Brute': 1, 'median3': 3, 'median5': 5}


def brute_remove(data: np.ndarray) -> np.ndarray:
    # fast5.py:90-101: sequential, in place on the copy, int16 truncation on store
    out = data.copy()
    for i in np.where((data > 1000) | (data < 250))[0]:
        if i > 2:
            out[i] = np.median(out[i - 2:i + 3])
    return out


def remove_spikes(data: np.ndarray, mode: str) -> np.ndarray:
    # fast5.py:68-77
    if mode == 'median3':
        return medfilt(data, 3)
    if mode == 'median5':
        return medfilt(data, 5)
    if mode == 'Brute':
        return brute_remove(data)
    return data


def normalize_signal_mad(data: np.ndarray) -> np.ndarray:
    # fast5.py:104-114
    shift = np.mean(np.percentile(data, (46.5, 53.5)))
    scale = np.median(np.abs(data - shift))
    return np.asarray((data - shift) / scale)


def get_data_processed(raw: np.ndarray, position, mode: str = 'Brute') -> np.ndarray:
    # fast5.py:45-57
    norm = normalize_signal_mad(remove_spikes(raw, mode))
    return norm[position[0]:position[1] + 1]
