"""CPU: the C-ABI shared library loads and exports every symbol include/warpstr_b200.h
declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'warpstr_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(wstr_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_path():
    names = _declared()
    for must in ('wstr_pore_lookup', 'wstr_normalize_batch', 'wstr_automaton_create', 'wstr_warp_batch',
                 'wstr_call_batch'):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in _declared():
        assert hasattr(lib, name), f'{name} is declared in the header but not exported'


def test_python_binding_covers_the_header(built_lib):
    from warpstr_b200 import _lib
    assert sorted(_lib.exported_symbols()) == _declared()
    assert _lib.lib().wstr_version() >= 100
    assert _lib.lib().wstr_error_string(-3).decode().startswith('automaton too large')


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from warpstr_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_lib.WarpstrError):
        _lib.lib()


def test_engine_refuses_to_run_without_cuda(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from warpstr_b200 import _lib
    from warpstr_b200.caller import CallerEngine
    with pytest.raises(_lib.WarpstrError):
        CallerEngine()


def test_host_only_entry_points_validate_arguments(built_lib):
    from warpstr_b200 import _lib
    import numpy as np
    with pytest.raises(_lib.WarpstrError):
        _lib.automaton_plan(np.array([0, 0, 1], dtype=np.int32), np.array([0], dtype=np.int32), 2, 1)   # config.py:115
    # nothing the reference accepts is refused: an unusual dwell goes to the catch-all kernel (all-zero plan)
    info, sop = _lib.automaton_plan(np.array([0, 0, 1], dtype=np.int32), np.array([0], dtype=np.int32), 2, 99)
    assert info['chain_slots'] == 0 and info['generic_slots'] == 0 and len(sop) == 0
