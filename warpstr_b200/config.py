"""Configuration for the caller hot path: the same YAML keys WarpSTR reads.

The reference parses ``sys.argv`` and its YAML at import time and exposes module
globals (src/config.py:174-210).  Here the same keys and defaults
(src/default.yaml:1-38) are loaded explicitly with :func:`load_config`, and the
hot-path knobs travel as plain dataclasses carrying the reference's names:
``CallerConfig`` (config.py:104-119) and ``RescalerConfig`` (config.py:91-101).
Unknown keys of the other pipeline steps are kept untouched in ``Config.raw`` so a
WarpSTR YAML loads unchanged.
"""
import copy
import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_PORE_MODEL = os.path.join(_HERE, 'data', 'r9.4_450bps_6mer_template_median68pA.tsv')

# defaults of src/default.yaml
DEFAULTS: Dict[str, Any] = {
    'verbose': 0,
    'threads': 2,
    'force_overwrite': False,
    'flank_length': 110,
    'pore_model_path': 'example/deps/template_median68pA.model',
    'tr_calling_config': {
        'spike_removal': 'Brute',
        'min_values_per_state': 4,
        'states_in_segment': 6,
        'min_state_similarity': 0.75,
        'visualize_alignment': True,
        'visualize_phase': True,
        'visualize_strand': True,
        'visualize_cost': True,
    },
    'rescaling': {
        'reps_as_one': False,
        'threshold': 0.5,
        'max_std': 0.5,
        'method': 'mean',
    },
    'genotyping_config': {
        'min_weight': 0.2,
        'std_filter': 2,
        'visualize': True,
        'msa': False,
    },
    'alignment': {
        'accuracy_factor': 1.15,
        'identity_factor': 0.85,
        'match_score': 2,
        'mismatch_score': -3,
        'gap_open_score': -3,
        'gap_extend_score': -3,
    },
}


@dataclass
class RescalerConfig:
    reps_as_one: bool = False
    threshold: float = 0.5
    max_std: float = 0.5
    method: str = 'mean'

    def __post_init__(self):
        # same validity rules as the reference (config.py:98-101)
        if not self.threshold > 0:
            raise AssertionError('rescaling.threshold must be > 0')
        if not self.max_std > 0:
            raise AssertionError('rescaling.max_std must be > 0')
        if self.method not in ('mean', 'median'):
            raise AssertionError('rescaling.method must be mean or median')


@dataclass
class CallerConfig:
    spike_removal: str = 'Brute'
    min_values_per_state: int = 4
    states_in_segment: int = 6
    min_state_similarity: float = 0.75
    visualize_alignment: bool = True
    visualize_phase: bool = True
    visualize_strand: bool = True
    visualize_cost: bool = True

    def __post_init__(self):
        # same validity rules as the reference (config.py:115-119)
        if not self.min_values_per_state > 1:
            raise AssertionError('min_values_per_state must be > 1')
        if not self.states_in_segment > 1:
            raise AssertionError('states_in_segment must be > 1')
        if not self.min_state_similarity > 0:
            raise AssertionError('min_state_similarity must be > 0')
        if self.spike_removal not in ('None', 'median3', 'median5', 'Brute'):
            raise AssertionError('spike_removal must be None, median3, median5 or Brute')


@dataclass
class LocusConfig:
    """One entry of ``loci:`` (reference: schemas/locus.py:8-46).  ``sequence`` is
    the automaton regex; ``motif`` loci need the reference genome and are resolved
    by the caller of this package (out of the hot path)."""
    name: str
    coord: str = ''
    sequence: Optional[str] = None
    motif: Optional[str] = None
    noting: Optional[str] = None
    flank_length: Optional[int] = None

    def __post_init__(self):
        if self.sequence:
            self.sequence = self.sequence.upper()


@dataclass
class Config:
    output: str = ''
    reference_path: str = ''
    pore_model_path: str = DEFAULT_PORE_MODEL
    flank_length: int = 110
    threads: int = 2
    verbose: int = 0
    force_overwrite: bool = False
    caller_config: CallerConfig = field(default_factory=CallerConfig)
    rescaler_config: RescalerConfig = field(default_factory=RescalerConfig)
    loci: List[LocusConfig] = field(default_factory=list)
    raw: Dict[str, Any] = field(default_factory=dict)

    def locus_flank_length(self, locus: LocusConfig) -> int:
        """Per-locus flank length with the global default (schemas/locus.py:33-37)."""
        return locus.flank_length if locus.flank_length else self.flank_length


def add_defaults(cfg: Dict[str, Any], default: Dict[str, Any]) -> None:
    """Recursive defaults merge with the reference's rules (config.py:31-59):
    nested dicts are merged key by key, present keys win."""
    for key, val in default.items():
        if isinstance(val, dict):
            sub = cfg.setdefault(key, {})
            if sub is None:
                sub = cfg[key] = {}
            add_defaults(sub, val)
        elif key not in cfg:
            cfg[key] = copy.deepcopy(val)


def config_from_dict(raw: Dict[str, Any], base_dir: Optional[str] = None) -> Config:
    raw = copy.deepcopy(raw) if raw else {}
    add_defaults(raw, DEFAULTS)
    pore = raw.get('pore_model_path')
    if pore and not os.path.isabs(pore) and base_dir:
        cand = os.path.join(base_dir, pore)
        pore = cand if os.path.exists(cand) else pore
    if not pore or (not os.path.exists(pore) and
                    os.path.normpath(str(raw.get('pore_model_path'))) == os.path.normpath(DEFAULTS['pore_model_path'])):
        # the reference's default (example/deps/template_median68pA.model, relative to its checkout):
        # the same table is bundled with this package
        pore = DEFAULT_PORE_MODEL
    elif not os.path.exists(pore):
        # a user's own model that is not there: fail like the reference (pore_model.py:22-24), do not
        # silently call with the levels of another chemistry
        raise FileNotFoundError(f'Not found pore model table at path {pore}')
    loci = [LocusConfig(**{k: v for k, v in item.items()
                           if k in LocusConfig.__dataclass_fields__})
            for item in (raw.get('loci') or [])]
    return Config(
        output=raw.get('output', '') or '',
        reference_path=raw.get('reference_path', '') or '',
        pore_model_path=pore,
        flank_length=int(raw['flank_length']),
        threads=int(raw['threads']),
        verbose=int(raw['verbose']),
        force_overwrite=bool(raw['force_overwrite']),
        caller_config=CallerConfig(**raw['tr_calling_config']),
        rescaler_config=RescalerConfig(**raw['rescaling']),
        loci=loci,
        raw=raw,
    )


def load_config(path: str) -> Config:
    """Load a WarpSTR YAML (reference: config.py:10-28,69-85)."""
    if not os.path.exists(path):
        raise FileNotFoundError(f'The config file {path} does not exist!')
    with open(path, 'r') as fh:
        try:
            raw = yaml.safe_load(fh)
        except yaml.YAMLError as exc:
            raise ValueError(f'Incorrect YAML format in config path {path},err={exc}')
    if raw is None:
        raise ValueError(f'Error when loading config file from {path}')
    if 'loci' not in raw:
        raise KeyError('No loci defined in the config')
    return config_from_dict(raw, base_dir=os.path.dirname(os.path.abspath(path)))
