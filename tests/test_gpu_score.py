"""the consumer's phred-weighted overlap score (EdgeCalculator.cpp:26-139) on the device vs its
CPU restatement (oracle/hlo_oracle.c:orc_overlap_score).  FP64: mismatch_rate must be bit-equal (a ratio of small
integers), score within 1e-12 relative (the only difference is the last ulp of log()/exp())."""
import numpy as np
import pytest

from hylight_b200 import api, synth
from oracle import orc

from util import small_cluster, to_orc_batch

pytestmark = pytest.mark.gpu
REL_TOL = 1e-12


def random_quals(rng, n, lo=2, hi=41):
    return (rng.integers(lo, hi + 1, size=n) + 33).astype(np.uint8)


def check(b, mismatch_prob=0.0, min_read_len=0, seed=0):
    rng = np.random.default_rng(seed)
    ob = to_orc_batch(b)
    quals = random_quals(rng, len(ob.bases))
    with api.Context(0) as c:
        R = c.pack(ob)
        rec, aux, coff = c.overlap_batch(R)
        got = c.score_records(R, quals, rec, aux["cluster"], mismatch_prob, min_read_len)
        R.free()
    want = orc.score_records(ob, quals, rec, aux["cluster"], mismatch_prob, min_read_len)
    assert len(got) == len(rec) > 100
    assert np.array_equal(got["mismatch_rate"], want[:, 1])
    assert np.allclose(got["score"], want[:, 0], rtol=REL_TOL, atol=0.0)
    assert np.array_equal(got["score"] == 0.0, want[:, 0] == 0.0)
    return got


def test_scores_small_cluster():
    got = check(small_cluster(61), seed=1)
    assert (got["score"] > 0).all() and (got["score"] <= 1).all()
    assert (got["mismatch_rate"] <= 0.05).all()  # <= 2 % in the block plus at most one clipped terminal mismatch


def test_scores_multi_cluster_with_ambiguous_bases_and_ragged_lengths():
    rng = np.random.default_rng(2)
    parts = []
    for s in (62, 63):
        b = small_cluster(s, n_rate=0.004, n_fragments=150)
        parts.append(synth.ReadBatch.from_seqs([[r[:int(rng.integers(90, 151))] for r in b.cluster_reads(0)]]))
    check(synth.ReadBatch.concat(parts), seed=3)


def test_scores_unusual_phred_bytes():
    """reads with phred bytes outside 33..126 leave the shared-memory table (per-read flag) and take the general path"""
    b = small_cluster(67, n_fragments=120)
    ob = to_orc_batch(b)
    rng = np.random.default_rng(6)
    quals = random_quals(rng, len(ob.bases))
    off = np.asarray(b.base_off)
    for rd in rng.choice(b.n_reads, size=40, replace=False):
        lo, hi = int(off[rd]), int(off[rd + 1])
        quals[rng.integers(lo, hi, size=3)] = rng.integers(0, 128, size=3).astype(np.uint8)
    with api.Context(0) as c:
        R = c.pack(ob)
        rec, aux, coff = c.overlap_batch(R)
        got = c.score_records(R, quals, rec, aux["cluster"])
        R.free()
    want = orc.score_records(ob, quals, rec, aux["cluster"], 0.0, 0)
    assert np.array_equal(got["mismatch_rate"], want[:, 1])
    fin = np.isfinite(want[:, 0])
    assert np.array_equal(np.isfinite(got["score"]), fin)
    assert np.allclose(got["score"][fin], want[fin, 0], rtol=REL_TOL, atol=0.0)


def test_scores_mismatch_threshold_and_min_read_len():
    b = small_cluster(64, err_rate=0.01)
    got = check(b, mismatch_prob=1e-3, seed=4)       # a low-quality mismatch has p < 1e-3 -> score 0
    assert (got["score"] == 0).any() and (got["score"] > 0).any()
    got = check(b, min_read_len=151, seed=5)         # every read is too short
    assert (got["score"] == 0).all() and (got["mismatch_rate"] == 1).all()


def test_scores_need_plain_names():
    b = small_cluster(65, n_fragments=30)
    ids = (np.arange(b.n_reads, dtype=np.uint32) * 2 + 1)
    with api.Context(0) as c:
        R = c.pack(to_orc_batch(b, ids))
        rec, aux, coff = c.overlap_batch(R)
        with pytest.raises(api.HloError) as e:
            c.score_records(R, np.full(b.n_bases, 70, dtype=np.uint8), rec, aux["cluster"])
        assert e.value.status == -5
        R.free()


def test_host_binary_scored_file(tmp_path):
    """hylight-ovlp --scores: FASTQ in, <out>.scored = the 13 columns + score and mismatch rate (%.17g)"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "hylight_b200", "hylight-ovlp")
    rng = np.random.default_rng(9)
    b = small_cluster(66, n_fragments=80)
    quals = random_quals(rng, b.n_bases)
    fq = tmp_path / "s_p1_p2.fastq"
    with open(fq, "wb") as f:
        for i in range(b.n_reads):
            lo, hi = int(b.base_off[i]), int(b.base_off[i + 1])
            f.write(b"@%d\n" % i + b.bases[lo:hi].tobytes() + b"\n+\n" + quals[lo:hi].tobytes() + b"\n")
    out = tmp_path / "o.txt"
    r = subprocess.run([exe, "--in", str(fq), "--out", str(out), "--scores"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ob = to_orc_batch(b)
    rec, blk, coff, st = orc.overlap_batch(ob)
    want = orc.score_records(ob, quals, rec)
    plain = open(out).read().splitlines()
    scored = open(str(out) + ".scored").read().splitlines()
    assert len(plain) == len(scored) == len(rec)
    for k, (p, s) in enumerate(zip(plain, scored)):
        f = s.split("\t")
        assert len(f) == 15 and "\t".join(f[:13]) == p
        assert float(f[14]) == want[k, 1]
        assert abs(float(f[13]) - want[k, 0]) <= REL_TOL * want[k, 0]
    # FASTA input has no qualities
    fa = tmp_path / "r.fa"
    synth.write_fasta(b, 0, str(fa))
    r = subprocess.run([exe, "--in", str(fa), "--out", str(tmp_path / "x.txt"), "--scores"], capture_output=True, text=True)
    assert r.returncode != 0 and "FASTQ" in r.stderr
