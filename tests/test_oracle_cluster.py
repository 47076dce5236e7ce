"""Oracle (CPU restatement) as a whole chain vs the reference converters' output on the same reads
(tests/golden/cluster*.overlaps.txt, produced by the unmodified reference scripts, make_golden_cluster.py)."""
import os

import numpy as np
import pytest

from hylight_b200 import synth
from oracle import orc

from util import GOLDEN, golden_args, jaccard, overlap_tuples, read_fasta


@pytest.mark.parametrize("tag", ["cluster1", "cluster2", "cluster3"])
def test_oracle_chain_vs_reference_scripts(tag, tmp_path):
    seqs = read_fasta(os.path.join(GOLDEN, tag + ".fasta"))
    b = orc.Batch([seqs])
    p = orc.default_params(**golden_args(tag))
    rec, blk, coff, st = orc.overlap_batch(b, p)
    out = tmp_path / "o.txt"
    orc.write_overlaps(rec, str(out))
    got, want = overlap_tuples(str(out)), overlap_tuples(os.path.join(GOLDEN, tag + ".overlaps.txt"))
    # the only licensed difference is the per-pair dedup scope (global here, 1000-line chunks in
    # filter_ovlp_inline.py:97): the canonical rule never adds a pair, it only drops second hits
    assert got <= want
    assert jaccard(got, want) >= 0.999
    assert {(t[0], t[1]) if t[0] < t[1] else (t[1], t[0]) for t in got} == \
           {(t[0], t[1]) if t[0] < t[1] else (t[1], t[0]) for t in want}


def test_oracle_sketch_known_answers():
    # hash64 is invertible on 2k bits and maps 0 -> a fixed value; pin a few values computed from the published
    # formula (minimap2 sketch.c hash64) with python big ints
    def h(key, mask):
        key = (~key + (key << 21)) & mask
        key = key ^ key >> 24
        key = ((key + (key << 3)) + (key << 8)) & mask
        key = key ^ key >> 14
        key = ((key + (key << 2)) + (key << 4)) & mask
        key = key ^ key >> 28
        key = (key + (key << 31)) & mask
        return key
    mask = (1 << 42) - 1
    for key in [0, 1, 12345678901, mask, mask - 1, 0x2AAAAAAAAAA]:
        assert orc.lib().orc_hash64(key, mask) == h(key, mask)
    # a read shorter than k has no minimizer; a read of exactly w+k-1 bases has at least one
    assert len(orc.sketch_one("ACGTACGTAC")) == 0
    rng = np.random.default_rng(1)
    s = "".join("ACGT"[i] for i in rng.integers(0, 4, size=31))
    mz = orc.sketch_one(s)
    assert len(mz) >= 1 and all((m["x"] & 0xff) == 21 for m in mz)
    # positions are the last base of the k-mer and strictly increasing... not necessarily; but within range
    assert all(20 <= ((int(m["y"]) & 0xffffffff) >> 1) < 31 for m in mz)


def test_oracle_sketch_window_guarantee():
    """every window of w consecutive k-mers (no N, no palindromes) holds at least one sampled minimizer"""
    rng = np.random.default_rng(2)
    for _ in range(20):
        L = int(rng.integers(60, 300))
        s = "".join("ACGT"[i] for i in rng.integers(0, 4, size=L))
        pos = sorted(((int(m["y"]) & 0xffffffff) >> 1) for m in orc.sketch_one(s))
        assert pos, s
        ends = list(range(20, L))
        for i in range(0, len(ends) - 11 + 1):
            win = ends[i:i + 11]
            assert any(win[0] <= p <= win[-1] for p in pos)
