#!/usr/bin/env python
"""Pile-up consensus of a BAM over a reference interval (fixture generation only).

The bundled WarpSTR test needs the GRCh38 flanks of chr4:183178378-183178421, and GRCh38 is
not available here; the bundled mapping.bam holds the basecalled reads aligned to it (no MD
tags), so the best reconstruction is a per-column majority vote.  Pure Python: BGZF blocks
through zlib, BAM records by hand.  Usage: bam_consensus.py BAM chr:start-end"""
import struct
import sys
import zlib
from collections import Counter


def bgzf_bytes(path):
    raw = open(path, 'rb').read()
    out = []
    pos = 0
    while pos < len(raw):
        d = zlib.decompressobj(31)
        out.append(d.decompress(raw[pos:]))
        pos = len(raw) - len(d.unused_data)
    return b''.join(out)


def records(path):
    b = bgzf_bytes(path)
    assert b[:4] == b'BAM\x01'
    (l_text,) = struct.unpack_from('<i', b, 4)
    p = 8 + l_text
    (n_ref,) = struct.unpack_from('<i', b, p)
    p += 4
    refs = []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from('<i', b, p)
        name = b[p + 4:p + 4 + l_name - 1].decode()
        (l_ref,) = struct.unpack_from('<i', b, p + 4 + l_name)
        refs.append((name, l_ref))
        p += 8 + l_name
    while p < len(b):
        (block,) = struct.unpack_from('<i', b, p)
        q = p + 4
        ref_id, pos, l_rn, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from('<iiBBHHHi', b, q)
        q += 32
        name = b[q:q + l_rn - 1].decode()
        q += l_rn
        cigar = [(v & 15, v >> 4) for v in struct.unpack_from('<%dI' % n_cig, b, q)]
        q += 4 * n_cig
        packed = b[q:q + (l_seq + 1) // 2]
        seq = ''.join('=ACMGRSVTWYHKDBN'[(packed[i >> 1] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
        yield refs[ref_id][0] if ref_id >= 0 else '*', pos, flag, mapq, cigar, seq, name
        p += 4 + block


def consensus(path, chrom, start, end):
    """0-based half-open [start, end) -> (sequence, per-column depth)"""
    cols = [Counter() for _ in range(end - start)]
    for ref, pos, flag, mapq, cigar, seq, _ in records(path):
        if ref != chrom or flag & 0x904:
            continue
        r, s = pos, 0
        for op, ln in cigar:
            if op in (0, 7, 8):
                for k in range(ln):
                    if start <= r + k < end:
                        cols[r + k - start][seq[s + k]] += 1
                r += ln
                s += ln
            elif op in (2, 3):
                for k in range(ln):
                    if start <= r + k < end:
                        cols[r + k - start]['-'] += 1
                r += ln
            elif op in (1, 4):
                s += ln
    out = []
    for c in cols:
        bases = [(n, b) for b, n in c.items() if b in 'ACGT']
        out.append(max(bases)[1] if bases else 'N')
    return ''.join(out), [sum(c.values()) for c in cols]


if __name__ == '__main__':
    chrom, rng = sys.argv[2].split(':')
    a, b = (int(x) for x in rng.split('-'))
    seq, depth = consensus(sys.argv[1], chrom, a - 1, b)
    print(seq)
    print('depth min/max', min(depth), max(depth), file=sys.stderr)
