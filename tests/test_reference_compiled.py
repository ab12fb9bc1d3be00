"""CPU: the byte-compiled unmodified reference under oracle/_ref (built by oracle/build_ref.py, the
timed arm of `bench.py --impl reference`) loads without the source tree and reproduces a committed
golden read -- the goldens were made by the same functions imported from /root/reference."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
from oracle import refshim
ref = refshim.load()
assert ref.root.endswith('oracle/_ref'), ref.root
import json
g = np.load(%(gold)r, allow_pickle=True)
case = json.loads(str(g['cases']))[0]
rev = bool(case['reverse'])
sta = ref.StateAutomata(case['reverse_regex'] if rev else case['template_regex'])
rs = ref.wrapper.ReadSignal(name=case['key'], reverse=rev, signal=g[case['key'] + '_signal'])
res = ref.wrapper.warpstr_call_sequential(rs, int(case['flank']), sta, None)
assert res.seq == case['seq'] and res.resc_seq == case['resc_seq']
assert res.cost == case['cost'] and res.resc_cost == case['resc_cost']
print('ok')
'''


def test_compiled_reference_reproduces_a_golden_read():
    from oracle import build_ref, refshim
    build_ref.build()
    if not refshim.compiled_available():
        pytest.skip('oracle/_ref not built (no reference tree in this environment)')
    env = dict(os.environ, WSTR_REF_COMPILED='1')
    script = _SCRIPT % {'root': ROOT, 'gold': os.path.join(ROOT, 'tests', 'golden', 'caller.npz')}
    out = subprocess.run([sys.executable, '-c', script], env=env, capture_output=True, text=True, cwd='/tmp')
    assert out.returncode == 0 and 'ok' in out.stdout, out.stderr[-2000:]
