"""Constant tables shared by the automaton builder and the caller seam.

Mirrors the parts of the reference's ``src/templates.py`` that the hot path
consumes: the IUPAC expansion used by the automaton builder
(templates.py:32-44, consumed at caller/automata.py:104-131) and the
per-character reverse-strand map of the locus regex (templates.py:45-65,
consumed at caller/wrapper.py:78-84).  Output sub-directory / file names are
the ones the reference's caller reads and writes (templates.py:8-22).
"""

# output layout names (templates.py:8-22)
FAST5_SUBDIR = 'fast5'
ANNOT_SUBDIR = 'annot'
OVERVIEW_NAME = 'overview.csv'
PREDICTIONS_SUBDIR = 'predictions'
LOCUS_INFO_SUBDIR = 'expected_signals'
SUMMARY_SUBDIR = 'summaries'
COMPLEX_SUBDIR = 'complexSTR_analysis'
LOCUS_FLANKS = 'sequences.csv'
LOCUS_NAMES = ['left_flank_template', 'right_flank_template', 'left_flank_reverse',
               'right_flank_reverse', 'temp_ref_pattern', 'rev_ref_pattern']

BASES = 'ACGT'

# IUPAC ambiguity code -> alternatives, in the order the reference expands them
# (the order fixes the state numbering of the automaton).
_IUPAC = 'R:AG Y:CT S:GC W:AT K:GT M:AC B:CGT D:AGT H:ACT V:ACG N:ACGT'
DNA_DICT = {item[0]: list(item[2:]) for item in _IUPAC.split()}

# complement of every symbol the locus regex may contain; brackets flip because the
# regex is also reversed (wrapper.py:78-84).
_PAIRS = '() {} AT GC MK RY BV DH'
ENCODING_DICT = {}
for _p in _PAIRS.split():
    ENCODING_DICT[_p[0]] = _p[1]
    ENCODING_DICT[_p[1]] = _p[0]
for _s in 'NWS':
    ENCODING_DICT[_s] = _s
del _p, _s

_COMP = str.maketrans('ACGTN', 'TGCAN')


def reverse_complement(seq: str) -> str:
    """Plain-base reverse complement (reference: squiggler/dna_sequence.py:26-28)."""
    return seq.translate(_COMP)[::-1]


def reverse_uniq_sequence(sequence: str) -> str:
    """Locus regex on the reverse strand (reference: caller/wrapper.py:78-84)."""
    return ''.join(ENCODING_DICT[c] for c in reversed(sequence))
