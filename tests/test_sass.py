"""CPU: what the compiled fill kernel must look like (cuobjdump on the sm_100a object, no GPU needed).

Two properties of the hot loop depend on choices ptxas makes and were lost once without any test noticing
(DESIGN 4.1): with one-warp CTAs and a counted read loop the 3-row cycle of the (7,1) first-pass kernel is
~263 SASS instructions (297 otherwise) and holds no local-memory access.  Also checked: the DP is FP64
add/compare work (no FMA contraction, no tensor-core instructions) and the tiles move through the TMA engine."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL = 'dtw_fill_kernelILi7ELi1ELi2ELi4ELb0E'     # (7,1) layout, in-degree 2, mv 4, first pass


def _sass(obj, pattern):
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True, check=True).stdout
    for blk in out.split('Function : ')[1:]:
        if pattern in blk.split('\n', 1)[0]:
            ins = []
            for line in blk.split('\n'):
                m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
                if m:
                    ins.append((int(m.group(1), 16), m.group(2).strip()))
            return ins
    return None


@pytest.fixture(scope='module')
def fill_sass(built_lib):
    if shutil.which('cuobjdump') is None:
        pytest.skip('no cuobjdump')
    obj = os.path.join(ROOT, 'build', 'dtw.cu.p0.o')
    if not os.path.exists(obj):
        import __graft_entry__ as ge
        ge.build()
    ins = _sass(obj, KERNEL)
    assert ins, 'the (7,1,2,4) first-pass fill kernel is not in build/dtw.cu.p0.o'
    return ins


def _loops(ins):
    """(length, first index, last index) of every backward-branch loop body."""
    at = {a: i for i, (a, _) in enumerate(ins)}
    out = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r'\bBRA(?:\.U)?(?:\.ANY)?\s+(?:[!U]*P\d,\s*)?0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in at:
            out.append((i - at[int(m.group(1), 16)] + 1, at[int(m.group(1), 16)], i))
    return out


def _op(text):
    parts = text.split()
    return (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]


def test_row_loop_is_the_short_form_without_local_memory(fill_sass):
    # the unbanded, unmasked 3-row cycle: the loop with 123 DADD (41 per row) and nothing but the DP in it
    cands = []
    for n, a, b in _loops(fill_sass):
        ops = collections.Counter(_op(t) for _, t in fill_sass[a:b + 1])
        if ops['DADD'] == 123 and ops['DSETP'] == 27 and n < 320:
            cands.append((n, ops))
    assert cands, 'no 3-row cycle with 123 DADD + 27 DSETP found'
    n, ops = min(cands, key=lambda c: c[0])
    assert n <= 270, f'the row loop grew to {n} instructions (262-264 expected): {dict(ops)}'
    assert ops['LDL'] == 0 and ops['STL'] == 0, 'the first-pass row loop touches local memory'
    assert ops['FSEL'] == 54 and ops['LOP3'] <= 30 and ops['LDS'] == 12 and ops['STS'] == 6


def test_fill_kernel_instruction_classes(fill_sass):
    ops = collections.Counter(_op(t) for _, t in fill_sass)
    assert ops['DFMA'] == 0 and ops['DMUL'] == 0, 'the DP must be add/compare only (-fmad=false, no contraction)'
    assert not any(o.startswith(('HMMA', 'UTCMMA', 'UTCHMMA', 'IMMA', 'DMMA')) for o in ops), 'tensor-core code in the DP'
    assert ops['UBLKCP'] >= 2, 'signal tiles and traceback windows are bulk async copies (TMA engine)'
    assert ops['SYNCS'] >= 4, 'mbarrier waits/arrivals expected'


def test_normalize_kernel_stages_its_samples_with_bulk_copies(built_lib):
    """The normalisation kernel's scan: samples arrive by bulk asynchronous copy completing on
    mbarriers, are counted with shared-memory reductions, and the common step makes no call -- the only calls
    in the kernel are the out-of-line general step's (window and value-indexed form)."""
    if shutil.which('cuobjdump') is None:
        pytest.skip('no cuobjdump')
    obj = os.path.join(ROOT, 'build', 'aux.cu.o')
    if not os.path.exists(obj):
        import __graft_entry__ as ge
        ge.build()
    ins = _sass(obj, 'normalize_kernel')
    assert ins, 'normalize_kernel is not in build/aux.cu.o'
    ops = collections.Counter(_op(t) for _, t in ins)
    assert ops['UBLKCP'] >= 2                       # prologue + refill
    assert any('SYNCS.PHASECHK' in t for _, t in ins) and any('SYNCS.ARRIVE' in t for _, t in ins)
    assert ops['ATOMS'] + ops['REDS'] + sum('RED' in _op(t) for _, t in ins) >= 16
    # the loop the stages are consumed in: from the first wait on a stage to the refill, no call in between
    waits = [i for i, (_, t) in enumerate(ins) if 'SYNCS.PHASECHK' in t]
    copies = [i for i, (_, t) in enumerate(ins) if _op(t) == 'UBLKCP']
    first_wait = min(w for w in waits if w > copies[0])
    refill = min(c for c in copies if c > first_wait)
    assert not any(_op(t) == 'CALL' for _, t in ins[first_wait:refill])
