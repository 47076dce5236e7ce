"""Pins of the CPU oracle that need no GPU.

1. oracle/hlo_oracle.c:orc_overlap_score against the reference's OWN code: oracle/_ref/libecref.so holds
   EdgeCalculator::score / phred_to_prob / overlap_score compiled from /root/reference/tools/HaploConduct/src/
   EdgeCalculator.cpp:25-140 (oracle/Makefile, oracle/ec_ref_harness.cpp).  Bit-equal doubles are required: both run
   the same IEEE operations in the same order on the same libm.
2. the data-derived occurrence cut-off (minimap2 mm_idx_cal_max_occ) against a numpy restatement built on the
   oracle's own sketches.
"""
import numpy as np
import pytest

from hylight_b200 import synth
from oracle import orc

from util import small_cluster, to_orc_batch, random_reads


def need_ecref():
    if orc.ecref() is None:
        pytest.skip("oracle/_ref/libecref.so not built (needs /root/reference at build time)")


def test_phred_to_prob_is_the_reference_expression():
    need_ecref()
    L = orc.ecref()
    import ctypes, math
    libm = ctypes.CDLL("libm.so.6")
    libm.pow.restype = ctypes.c_double
    libm.pow.argtypes = [ctypes.c_double, ctypes.c_double]
    for q in range(0, 100):  # (the reference asserts 0 <= P <= 1: a phred byte below '!' aborts ViralQuasispecies)
        assert L.ecref_phred_to_prob(q) == libm.pow(10.0, -q / 10.0)


@pytest.mark.parametrize("seed,n_rate,mismatch_prob,min_read_len", [(1, 0.0, 0.0, 0), (2, 0.01, 0.0, 0), (3, 0.0, 1e-3, 0),
                                                                   (4, 0.005, 1e-4, 120), (5, 0.0, 0.0, 200)])
def test_overlap_score_port_equals_the_reference_code(seed, n_rate, mismatch_prob, min_read_len):
    need_ecref()
    rng = np.random.default_rng(seed)
    b = synth.cluster_batch(2, seed0=40 + seed, n_fragments=150, read_len=150, n_strains=3, strain_cov=8.0, n_rate=n_rate)
    ob = to_orc_batch(b)
    rec, blk, coff, st = orc.overlap_batch(ob)
    assert len(rec) > 500
    quals = (rng.integers(2, 42, size=len(ob.bases)) + 33).astype(np.uint8)
    clu = np.repeat(np.arange(ob.n_clusters, dtype=np.uint32), np.diff(coff).astype(np.int64))
    got = orc.score_records(ob, quals, rec, clu, mismatch_prob, min_read_len)
    pick = rng.choice(len(rec), size=400, replace=False)
    zeros = 0
    for i in pick:
        sc, mr = orc.ecref_score_record(ob, quals, rec[i], int(ob.cluster_off[clu[i]]), mismatch_prob, min_read_len)
        assert sc == got[i, 0] and mr == got[i, 1], (i, sc, got[i, 0], mr, got[i, 1])
        zeros += sc == 0.0
    if mismatch_prob > 0 or min_read_len > 150:
        assert zeros > 0  # the threshold / length rules were exercised


def test_overlap_score_ragged_lengths_and_containments():
    need_ecref()
    rng = np.random.default_rng(9)
    base = random_reads(rng, 1, 400, 400)[0]
    seqs = [base[i:i + int(rng.integers(60, 200))] for i in rng.integers(0, 200, size=120)]
    ob = orc.Batch([seqs])
    rec, blk, coff, st = orc.overlap_batch(ob, orc.default_params(min_ovlp_len=30))
    assert len(rec) > 200
    quals = (rng.integers(0, 60, size=len(ob.bases)) + 33).astype(np.uint8)
    got = orc.score_records(ob, quals, rec)
    for i in range(0, len(rec), 7):
        sc, mr = orc.ecref_score_record(ob, quals, rec[i])
        assert sc == got[i, 0] and mr == got[i, 1]


def numpy_mid_occ(ob, p):
    mz, roff = orc.sketch_all(ob, p)
    out = []
    for c in range(ob.n_clusters):
        m0, m1 = int(roff[ob.cluster_off[c]]), int(roff[ob.cluster_off[c + 1]])
        h = mz["x"][m0:m1] >> np.uint64(8)
        if len(h) == 0:
            out.append(10)
            continue
        _, cnt = np.unique(h, return_counts=True)
        cnt = np.sort(cnt)
        k = int(np.uint32((1.0 - float(np.float32(2e-4))) * len(cnt)))
        out.append(int(min(max(int(cnt[k]) + 1, 10), 1000000)))
    return np.array(out, dtype=np.uint32)


def test_mid_occ_is_minimap2s_quantile_rule():
    p = orc.default_params()
    assert p.mid_occ == 0
    # shallow clusters: fewer than 5000 distinct minimizers -> k = n - 1 -> the largest bucket + 1: nothing is filtered
    b = synth.cluster_batch(3, seed0=70, n_fragments=300, read_len=150, n_strains=3, strain_cov=8.0)
    ob = to_orc_batch(b)
    assert np.array_equal(orc.mid_occ(ob, p), numpy_mid_occ(ob, p))
    # a deep, repetitive cluster: many distinct minimizers plus a few very frequent ones -> the cut-off bites
    rng = np.random.default_rng(3)
    genome = random_reads(rng, 1, 400000, 400000)[0]  # ~65 000 distinct minimizers: the top 0.02 % are the repeat's
    rep = random_reads(rng, 1, 60, 60)[0]
    reads = []
    for i in range(8000):
        s = int(rng.integers(0, len(genome) - 150))
        r = bytearray(genome[s:s + 150])
        if i % 3 == 0:
            r[40:100] = rep  # the same 60-mer in a third of the reads: its minimizers occur 2000 times
        reads.append(bytes(r))
    ob2 = orc.Batch([reads])
    mo = orc.mid_occ(ob2, p)
    assert np.array_equal(mo, numpy_mid_occ(ob2, p))
    assert 10 <= mo[0] < 2000
    c_auto, hits_auto = orc.candidates(ob2, p)
    c_all, hits_all = orc.candidates(ob2, orc.default_params(mid_occ=100000))
    assert hits_auto < hits_all  # the repeat's buckets seed nothing under the data-derived cut-off
    # an explicit value overrides
    assert orc.mid_occ(ob2, orc.default_params(mid_occ=77))[0] == 77
