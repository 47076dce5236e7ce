"""Shared helpers of the test-suite."""
import os

import numpy as np

from hylight_b200 import synth
from oracle import orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def read_fasta(path):
    seqs, cur = [], None
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            cur = []
            seqs.append(cur)
        elif cur is not None:
            cur.append(line)
    return ["".join(s) for s in seqs]


def golden_args(tag):
    min_len, min_id, o, r, n = open(os.path.join(GOLDEN, tag + ".args")).read().split()
    return dict(min_ovlp_len=int(min_len), min_identity=float(min_id), max_overhang=int(o), overhang_ratio=float(r))


def overlap_tuples(path_or_lines):
    """canonical tuples (ID1, ID2, POS1, ORI1, ORI2, LEN1) of a 13-column file."""
    lines = open(path_or_lines) if isinstance(path_or_lines, (str, os.PathLike)) else path_or_lines
    out = set()
    for line in lines:
        a = line.rstrip("\n").split("\t")
        assert len(a) == 13, a
        out.add((int(a[0]), int(a[1]), int(a[2]), a[5], a[6], int(a[9])))
    return out


def jaccard(a, b):
    if not a and not b:
        return 1.0
    return len(a & b) / float(len(a | b))


def to_orc_batch(b, ids=None):
    return orc.Batch.from_arrays(b.bases, b.base_off, b.cluster_off, ids=ids)


def random_reads(rng, n, lo, hi, alphabet=b"ACGT", n_rate=0.0, lower_rate=0.0):
    seqs = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        s = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), size=L)].copy()
        if n_rate:
            s[rng.random(L) < n_rate] = ord("N")
        if lower_rate:
            m = rng.random(L) < lower_rate
            s[m] = s[m] | 0x20
        seqs.append(s.tobytes())
    return seqs


def small_cluster(seed, **kw):
    args = dict(n_fragments=200, read_len=150, n_strains=3, strain_cov=8.0, snp_rate=0.01, err_rate=0.005)
    args.update(kw)
    return synth.cluster_batch(1, seed0=seed, **args)


def sort_recs(rec, *extra):
    order = np.lexsort((rec["K"], rec["OHB"], rec["OHA"], rec["ori"], rec["idB"], rec["idA"]))
    return (rec[order],) + tuple(e[order] for e in extra)
