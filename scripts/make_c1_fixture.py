#!/usr/bin/env python
"""Fixture of BASELINE config 1 (the reference's bundled caller-only test), made in the build
container where /root/reference is mounted; the GPU box only sees the committed .npz.

Inputs (reference tree, read-only): test/test_caller_only/example.csv (10 reads: strand and
raw window), test/test_input/test_run1/fast5s/batch_0.fast5 (raw signals, VBZ) and
mapping/mapping.bam.  GRCh38 is not available, so the 110-base flanks of
chr4:183178378-183178421 are the pile-up consensus of the 10 aligned reads
(scripts/bam_consensus.py).  Output: tests/golden/c1_bundled.npz with the raw reads
(delta-coded int16), windows, flanks and what the oracle (numpy normalisation + C restatement
of the caller) returns for them.  The reference's README.md:55 states the genotype (44, 40)."""
import csv
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
import bam_consensus as bc  # noqa: E402
from oracle import caller_oracle as co, normalize_oracle as no  # noqa: E402
from warpstr_b200 import fast5, templates as tmpl  # noqa: E402
from warpstr_b200.automata import StateAutomata  # noqa: E402

REF = '/root/reference'
F = 110
LOCUS = ('chr4', 183178378, 183178421, '(AAAT)')

chrom, a, b, seq = LOCUS
cons, depth = bc.consensus(os.path.join(REF, 'test/test_input/test_run1/mapping/mapping.bam'), chrom, a - 1 - F, b + F)
left, right = cons[:F], cons[-F:]
tseq = left + seq + right
rseq = tmpl.reverse_complement(right) + tmpl.reverse_uniq_sequence(seq) + tmpl.reverse_complement(left)
tbs = [co.tables_from(StateAutomata(tseq)), co.tables_from(StateAutomata(rseq))]

rows = list(csv.DictReader(open(os.path.join(REF, 'test/test_caller_only/example.csv'))))
out = dict(left=left, right=right, sequence=seq, flank_length=F, consensus_depth=np.array(depth),
           names=np.array([r['read_name'] for r in rows]), reverse=np.array([r['reverse'] == 'TRUE' for r in rows]),
           l_start_raw=np.array([int(r['l_start_raw']) for r in rows]), r_end_raw=np.array([int(r['r_end_raw']) for r in rows]))
len1, len2, c1, c2, seqs, rseqs, ss = [], [], [], [], [], [], []
for i, r in enumerate(rows):
    raw = fast5.raw_signal(os.path.join(REF, r['fast5_path']), r['read_name'])
    out[f'raw_delta{i}'] = np.diff(raw, prepend=np.int16(0)).astype(np.int16)
    fixed = no.remove_spikes(raw, 'Brute')
    shift = np.mean(np.percentile(fixed, (46.5, 53.5)))
    scale = np.median(np.abs(fixed - shift))
    x = np.ascontiguousarray(no.get_data_processed(raw, (int(r['l_start_raw']), int(r['r_end_raw']))))
    res = co.run_read(x, tbs[int(r['reverse'] == 'TRUE')], F, r['reverse'] == 'TRUE', impl='c')
    len1.append(len(res.seq)); len2.append(len(res.resc_seq)); c1.append(res.cost); c2.append(res.resc_cost)
    seqs.append(res.seq); rseqs.append(res.resc_seq); ss.append((shift, scale))
out.update(len1=np.array(len1), len2=np.array(len2), cost1=np.array(c1), cost2=np.array(c2), seq=np.array(seqs),
           resc_seq=np.array(rseqs), shift_scale=np.array(ss))
path = os.path.join(ROOT, 'tests', 'golden', 'c1_bundled.npz')
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), 'bytes; lengths', len2, 'costs', np.round(c2, 3))
