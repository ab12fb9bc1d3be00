"""TEST INFRASTRUCTURE ONLY.  Import the *unmodified* reference (fmfi-compbio/warpstr)
so that its own functions can be used to pin the oracle, to generate golden vectors
(tests/golden/, see oracle/make_golden.py) and as the timed CPU arm of bench.py.
In this container the modules come from /root/reference; on the GPU box, where that
tree does not exist, from oracle/_ref/ -- the same modules byte-compiled by
oracle/build_ref.py (source-less bytecode, git-ignored, shipped like a built .so).

The reference parses argv and its YAML config at import time
(src/config.py:174-210), needs h5py / pysam / Biopython / matplotlib / seaborn
(absent here) and uses ``np.bool8`` (removed in numpy 2; caller.py:60,392,402).
All of that is shimmed here without touching the reference tree.
"""
import os
import sys
import tempfile
import types

import numpy as np

REF_ROOT = '/root/reference'
_HERE = os.path.dirname(os.path.abspath(__file__))
COMPILED_ROOT = os.path.join(_HERE, '_ref')

_CFG = """\
reference_path: /nonexistent/GRCh38.fa
output: {out}
pore_model_path: {pore}
single_read_extraction: False
guppy_annotation: False
exp_signal_generation: False
tr_region_extraction: False
tr_region_calling: True
genotyping: False
tr_calling_config:
  visualize_alignment: False
  visualize_phase: False
  visualize_strand: False
  visualize_cost: False
loci:
  - name: ORACLE
    coord: chr1:1-10
    sequence: (AAAT)
"""

_loaded = None


def available() -> bool:
    return source_available() or compiled_available()


def source_available() -> bool:
    # WSTR_REF_COMPILED=1: use oracle/_ref even where the source tree exists (to test that path here)
    return os.path.isdir(os.path.join(REF_ROOT, 'src', 'caller')) and not os.environ.get('WSTR_REF_COMPILED')


def compiled_available() -> bool:
    return os.path.exists(os.path.join(COMPILED_ROOT, 'src', 'caller', 'wrapper.refbc'))


def _compiled_hook(path):
    """sys.path hook: directories under oracle/_ref hold modules as source-less bytecode, suffix .refbc."""
    from importlib.machinery import FileFinder, SourcelessFileLoader
    if not os.path.abspath(path).startswith(COMPILED_ROOT):
        raise ImportError
    return FileFinder(path, (SourcelessFileLoader, ['.refbc']))


def _stub(name, **kw):
    mod = types.ModuleType(name)
    mod.__dict__.update(kw)
    sys.modules[name] = mod
    return mod


class _Seq:
    """Stand-in for Bio.Seq.Seq: only reverse_complement()/str() are used
    (src/squiggler/dna_sequence.py:26-28)."""
    _tbl = str.maketrans('ACGTN', 'TGCAN')

    def __init__(self, s):
        self.s = s

    def reverse_complement(self):
        return _Seq(self.s.translate(self._tbl)[::-1])

    def __str__(self):
        return self.s


def load():
    """Returns a namespace with the reference's own hot-path callables."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f'reference neither at {REF_ROOT} nor compiled under {COMPILED_ROOT}')
    tmp = tempfile.mkdtemp(prefix='warpstr_oracle_')
    cfg = os.path.join(tmp, 'oracle_cfg.yaml')
    old_cwd, old_argv = os.getcwd(), sys.argv
    if source_available():
        root, cwd = REF_ROOT, REF_ROOT
        pore = os.path.join(REF_ROOT, 'example', 'deps', 'template_median68pA.model')
    else:
        # the compiled modules; src/config.py:13-26 reads <cwd>/src/default.yaml at import: write this
        # repo's table of the same defaults there, and use the bundled copy of the k-mer table
        import yaml
        from warpstr_b200 import config as our_config
        root, cwd = COMPILED_ROOT, tmp
        if _compiled_hook not in sys.path_hooks:
            sys.path_hooks.insert(0, _compiled_hook)
            sys.path_importer_cache.clear()
        os.makedirs(os.path.join(tmp, 'src'), exist_ok=True)
        with open(os.path.join(tmp, 'src', 'default.yaml'), 'w') as fh:
            yaml.safe_dump(our_config.DEFAULTS, fh)
        pore = our_config.DEFAULT_PORE_MODEL
    with open(cfg, 'w') as fh:
        fh.write(_CFG.format(out=tmp, pore=pore))
    os.chdir(cwd)
    sys.path.insert(0, root)
    sys.argv = ['WarpSTR.py', cfg]
    try:
        if not hasattr(np, 'bool8'):
            np.bool8 = np.bool_
        for name in ('h5py', 'pysam', 'seaborn'):
            _stub(name)
        mpl = _stub('matplotlib')
        mpl.pyplot = _stub('matplotlib.pyplot')
        bio = _stub('Bio')
        bio.Seq = _stub('Bio.Seq', Seq=_Seq)
        bio.pairwise2 = _stub('Bio.pairwise2')
        bio.SeqIO = _stub('Bio.SeqIO')
        _stub('Bio.Align')
        _stub('Bio.Align.Applications', MuscleCommandline=None)
        import src  # noqa: F401  (the reference's real package)
        ext = _stub('src.extractor')
        ext.__path__ = []
        # tr_extractor.py:51-52 has mutable dataclass defaults (py>=3.11 rejects them)
        ext.tr_extractor = _stub('src.extractor.tr_extractor', Flanks=object, load_flanks=None)
        from src import config as ref_config
        from src.caller import caller as ref_caller
        from src.caller import wrapper as ref_wrapper
        from src.caller.automata import StateAutomata
        from src.schemas.fast5 import Fast5, normalize_signal_mad
        from src.squiggler.pore_model import pore_model
        from src.squiggler.Squiggler import Squiggler
    finally:
        os.chdir(old_cwd)
        sys.argv = old_argv
    ns = types.SimpleNamespace(
        root=root, config=ref_config, caller=ref_caller, wrapper=ref_wrapper,
        StateAutomata=StateAutomata, WarpSTR=ref_caller.WarpSTR,
        Fast5=Fast5, normalize_signal_mad=normalize_signal_mad,
        pore_model=pore_model, Squiggler=Squiggler)
    _loaded = ns
    return ns
