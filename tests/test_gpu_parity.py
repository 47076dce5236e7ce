"""Parity of the CUDA path (through the C ABI, include/hlovlp.h) with the CPU oracle and the reference goldens.

Bar: bit-exact — every stage is integer/byte work."""
import os
import subprocess

import numpy as np
import pytest

from hylight_b200 import api, synth
from oracle import orc

from util import (GOLDEN, golden_args, jaccard, overlap_tuples, random_reads, read_fasta, small_cluster, sort_recs,
                  to_orc_batch)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def sketch_case(ctx, seqs_per_cluster, **pkw):
    """both sketch kernels (single-pass static-window kernel where it applies, generic two-pass kernel) vs the oracle"""
    b = orc.Batch(seqs_per_cluster)
    c = ctx if not pkw else api.Context(0, **pkw)
    p = orc.default_params(**pkw)
    mz_o, ro_o = orc.sketch_all(b, p)
    try:
        for generic in (0, 1):
            c.set_option("sketch_generic", generic)
            R = c.pack(b)
            mz, ro = c.sketch(R)
            R.free()
            assert np.array_equal(ro, ro_o)
            assert np.array_equal(mz["x"], mz_o["x"]) and np.array_equal(mz["y"], mz_o["y"])
    finally:
        c.set_option("sketch_generic", 0)
        if pkw:
            c.close()
    return len(mz)


# ---------------------------------------------------------------------------------------------------- sketch
def test_sketch_random_reads(ctx):
    rng = np.random.default_rng(1)
    n = sketch_case(ctx, [random_reads(rng, 500, 30, 260)])
    assert n > 1000


def test_sketch_edge_cases(ctx):
    rng = np.random.default_rng(2)
    seqs = [b"", b"A", b"ACGTACGTACGTACGTACGT", b"ACGTACGTACGTACGTACGTA",  # empty, 1 bp, k-1, k
            b"A" * 150, b"AT" * 75, b"ACGT" * 40, b"N" * 100,               # homopolymer, palindromic repeats, all N
            (b"ACGTTGCATGCCGATAGCTAGCTAGGATCGATCGATTAGCGCGATATCGCGCTAGCTAGGGATCCCTAGCTAGCATCGA" * 2)]
    seqs += random_reads(rng, 200, 21, 200, n_rate=0.02, lower_rate=0.3)
    seqs += random_reads(rng, 50, 100, 150, alphabet=b"AC")        # low complexity: many equal hashes in a window
    seqs += random_reads(rng, 50, 100, 150, alphabet=b"ACGTRYKMU")  # IUPAC letters are ambiguous, U = T
    sketch_case(ctx, [seqs])


def test_sketch_low_complexity_stress(ctx):
    """thousands of reads in which the same k-mer recurs inside one window (tandem repeats of period 1..40 with a few
    point changes, two-letter sequence, repeats next to random sequence, N inside repeats): the fast kernel has to
    notice every such read and hand it to the exact kernel"""
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = []
    for _ in range(4000):
        L = int(rng.integers(60, 260))
        kind = int(rng.integers(0, 5))
        if kind <= 2:  # tandem repeat, possibly mutated, possibly flanked by random sequence
            period = int(rng.integers(1, 41))
            unit = acgt[rng.integers(0, 4, size=period)]
            s_ = np.tile(unit, L // period + 2)[:L].copy()
            for _m in range(int(rng.integers(0, 4))):
                s_[int(rng.integers(0, L))] = acgt[int(rng.integers(0, 4))]
            if kind == 1:
                cut = int(rng.integers(0, L))
                s_[cut:] = acgt[rng.integers(0, 4, size=L - cut)]
            if kind == 2 and rng.random() < 0.5:
                s_[int(rng.integers(0, L))] = ord("N")
        elif kind == 3:  # two-letter sequence
            two = acgt[rng.choice(4, size=2, replace=False)]
            s_ = two[rng.integers(0, 2, size=L)]
        else:  # a random segment repeated after a short gap
            seg = acgt[rng.integers(0, 4, size=int(rng.integers(22, 60)))]
            gap = acgt[rng.integers(0, 4, size=int(rng.integers(0, 12)))]
            s_ = np.concatenate([seg, gap, seg, gap, seg, acgt[rng.integers(0, 4, size=L)]])[:L]
        seqs.append(s_.tobytes())
    for w, k in [(11, 21), (5, 15), (19, 11)]:
        sketch_case(ctx, [seqs], w=w, k=k, min_chain_score=min(30, 2 * k))


def test_sketch_other_wk(ctx):
    rng = np.random.default_rng(3)
    seqs = random_reads(rng, 200, 40, 200, n_rate=0.01)
    for w, k in [(5, 15), (1, 21), (19, 11), (32, 28), (10, 1), (11, 23), (11, 25), (19, 17), (5, 3)]:
        sketch_case(ctx, [seqs], w=w, k=k, min_chain_score=min(30, 2 * k))


def test_sketch_multi_cluster_rids_are_batch_global(ctx):
    rng = np.random.default_rng(4)
    sketch_case(ctx, [random_reads(rng, 40, 100, 150), random_reads(rng, 1, 150, 150), [], random_reads(rng, 60, 100, 150)])


# ------------------------------------------------------------------------------------------------ candidates
def cand_key(c):
    return np.lexsort((c["diag"], c["strand_anchor"] & 1, c["t"], c["q"]))


def check_candidates(ctx, b, ids=None, **pkw):
    ob = to_orc_batch(b, ids)
    c = ctx if not pkw else api.Context(0, **pkw)
    try:
        R = c.pack(ob)
        got = c.candidates(R)
        tm = c.timing()
        R.free()
    finally:
        if pkw:
            c.close()
    want, hits = orc.candidates(ob, orc.default_params(**pkw))
    got, want = got[cand_key(got)], want[cand_key(want)]
    assert len(got) == len(want)
    assert np.array_equal(got, want)
    assert tm["n_seed_hits"] == hits
    return got, tm


def test_candidates_small_cluster(ctx):
    got, _ = check_candidates(ctx, small_cluster(21))
    assert len(got) > 1000


def test_candidates_multi_cluster_and_ids(ctx):
    b = synth.ReadBatch.concat([small_cluster(22, n_fragments=100), small_cluster(23, n_fragments=60, read_len=250),
                                small_cluster(24, n_fragments=20, read_len=100)])
    check_candidates(ctx, b)
    rng = np.random.default_rng(7)
    ids = np.concatenate([rng.permutation(int(b.cluster_off[c + 1] - b.cluster_off[c])) * 3 + 5
                          for c in range(b.n_clusters)]).astype(np.uint32)
    check_candidates(ctx, b, ids=ids)


def test_candidates_table_overflow_path(ctx):
    """64-slot tables overflow for most reads: they go through the kernel again with 128 ... 2048 slots"""
    b = small_cluster(25, n_fragments=300, n_strains=1, strain_cov=40.0)
    for slots in (6, 5):
        ctx.set_option("pairgen_slots_log2", slots)
        try:
            _, tm = check_candidates(ctx, b)
            assert tm["n_overflow_reads"] > 0
        finally:
            ctx.set_option("pairgen_slots_log2", 8)


def test_candidates_bucket_suffix_longer_than_one_stage(ctx):
    """a pile of 2 600 reads on 320 bp with the occurrence cut-off out of the way: bucket suffixes of more than
    PG_CAP * PG_G = 1024 hits are staged in pieces (pairgen.cu, `skip`), reads with more than 2048 chains go on to the
    HBM-table kernel, which decodes the same 32-bit hit words"""
    b = small_cluster(29, n_fragments=1300, n_strains=1, strain_cov=1e9, insert_mean=160.0, insert_sd=2.0)
    _, tm = check_candidates(ctx, b, mid_occ=5000)
    assert tm["n_seed_hits"] > 1024 * b.n_reads


def test_candidates_mid_occ_and_params(ctx):
    b = small_cluster(26, n_fragments=150, n_strains=1, strain_cov=30.0)
    check_candidates(ctx, b, mid_occ=20)
    check_candidates(ctx, b, min_seeds=1, min_chain_score=21)
    check_candidates(ctx, b, w=5, k=15, min_chain_score=30)


# ---------------------------------------------------------------------------------------------------- verify
def check_verify(ctx, b, cands, ids=None, exact=False, **pkw):
    ob = to_orc_batch(b, ids)
    c = ctx if not pkw else api.Context(0, **pkw)
    try:
        c.set_option("verify_exact", 1 if exact else 0)
        R = c.pack(ob)
        rec, aux = c.verify(R, cands)
        tm = c.timing()
        R.free()
        c.set_option("verify_exact", 0)
    finally:
        if pkw:
            c.close()
    rec_o, blk_o, acc = orc.verify(ob, cands, orc.default_params(**pkw))
    assert len(rec) == len(rec_o), (len(rec), len(rec_o))
    rec, blk = sort_recs(rec, aux["blocklen"])
    rec_o, blk_o = sort_recs(rec_o, blk_o)
    assert np.array_equal(rec, rec_o)
    assert np.array_equal(blk, blk_o)
    return rec, tm


def test_verify_on_oracle_candidates(ctx):
    b = small_cluster(31)
    cands, _ = orc.candidates(to_orc_batch(b))
    rec, tm = check_verify(ctx, b, cands)
    assert len(rec) > 300 and tm["n_deferred"] < len(cands) // 10
    check_verify(ctx, b, cands, exact=True)


def random_cands(rng, b, n, anchors=True):
    nr = b.n_reads
    lens = np.diff(b.base_off).astype(np.int64)
    c = np.zeros(n, dtype=api.CAND_DT)
    c["q"] = rng.integers(0, nr, size=n)
    c["t"] = (c["q"] + rng.integers(1, nr, size=n)) % nr
    ql, tl = lens[c["q"]], lens[c["t"]]
    diag = rng.integers(-(ql - 1), np.maximum(tl, 1))
    c["diag"] = diag
    strand = rng.integers(0, 2, size=n)
    js = np.maximum(-diag, 0)
    je = np.minimum(ql, tl - diag)
    if anchors:
        a0 = js + (rng.random(n) * np.maximum(je - js, 1)).astype(np.int64)
        a1 = np.minimum(a0 + 21 + rng.integers(0, 40, size=n), np.maximum(je, a0 + 1))
        c["strand_anchor"] = strand | (a0 << 1) | (a1 << 16)
    else:
        c["strand_anchor"] = strand
    return c


def test_verify_random_diagonals_true_overlaps(ctx):
    """candidates on true overlap diagonals (from the oracle) with anchor info stripped / perturbed, plus random ones"""
    rng = np.random.default_rng(8)
    b = small_cluster(32, err_rate=0.01)
    cands, _ = orc.candidates(to_orc_batch(b))
    stripped = cands.copy()
    stripped["strand_anchor"] &= 1  # "no anchor information"
    check_verify(ctx, b, stripped)
    check_verify(ctx, b, np.concatenate([cands, random_cands(rng, b, 20000), random_cands(rng, b, 5000, anchors=False)]))


def test_verify_ambiguous_bases_and_ragged_lengths(ctx):
    rng = np.random.default_rng(9)
    b = small_cluster(33, n_rate=0.004, err_rate=0.004)
    # ragged: trim reads to random lengths 80..150 and rebuild the batch
    seqs = [s[:int(rng.integers(80, 151))] for s in b.cluster_reads(0)]
    rb = synth.ReadBatch.from_seqs([seqs])
    cands, _ = orc.candidates(to_orc_batch(rb))
    assert len(cands) > 500
    check_verify(ctx, rb, cands)
    check_verify(ctx, rb, cands, exact=True)
    check_verify(ctx, rb, np.concatenate([cands, random_cands(rng, rb, 5000)]), min_ovlp_len=40, min_identity=0.9,
                 max_overhang=3)


def test_verify_long_reads_take_exact_path(ctx):
    b = small_cluster(34, n_fragments=80, read_len=400, n_strains=2)
    cands, _ = orc.candidates(to_orc_batch(b))
    rec, tm = check_verify(ctx, b, cands)
    assert len(rec) > 50 and tm["n_deferred"] > 0


def test_verify_many_mismatch_clip_cases(ctx):
    """hand-made pairs: mismatches at every distance from a target-limited end, both ends, both strands"""
    rng = np.random.default_rng(10)
    base = rng.integers(0, 4, size=400, dtype=np.uint8)
    A = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs, cl = [], []
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    for trial in range(300):
        s0 = int(rng.integers(0, 100))
        off = int(rng.integers(1, 80))
        q = A[base[s0:s0 + 150]].copy()
        t = A[base[s0 + off:s0 + off + 150]].copy()
        # mutate up to 3 bases of t close to its start (the target-limited left end of q vs t is at t's start... and
        # of q close to its end), at small distances from the ends
        for _ in range(int(rng.integers(0, 4))):
            p = int(rng.integers(0, 6))
            t[p] = A[(np.where(A == t[p])[0][0] + 1) % 4]
        for _ in range(int(rng.integers(0, 3))):
            p = 149 - int(rng.integers(0, 6))
            q[p] = A[(np.where(A == q[p])[0][0] + 2) % 4]
        if rng.random() < 0.5:
            q = np.array([comp[x] for x in q[::-1]], dtype=np.uint8)
        seqs += [q.tobytes(), t.tobytes()]
    b = synth.ReadBatch.from_seqs([seqs])
    ob = to_orc_batch(b)
    cands, _ = orc.candidates(ob)
    assert len(cands) >= 250
    for o in (0, 1, 2, 5):
        check_verify(ctx, b, cands, max_overhang=o)
        both = cands.copy()
        both["q"], both["t"] = cands["t"], cands["q"]  # swapped roles: the other end becomes target-limited
        both["diag"] = -cands["diag"]
        both["strand_anchor"] &= 1
        plus = both[(both["strand_anchor"] & 1) == 0]
        check_verify(ctx, b, plus, max_overhang=o)


def test_verify_empty_and_invalid(ctx):
    b = small_cluster(35, n_fragments=20)
    R = ctx.pack(to_orc_batch(b))
    rec, aux = ctx.verify(R, np.zeros(0, dtype=api.CAND_DT))
    assert len(rec) == 0
    bad = np.zeros(1, dtype=api.CAND_DT)
    bad["q"] = 10 ** 6
    with pytest.raises(api.HloError):
        ctx.verify(R, bad)
    # a diagonal that leaves no overlap, or an anchor span outside the query, would steer the kernels' row reads
    for diag, sa in ((150, 0), (-150, 0), (40000, 0), (-40000, 0), (10, api_cand_pack(0, 100, 400)), (10, api_cand_pack(1, 90, 50))):
        bad = np.zeros(1, dtype=api.CAND_DT)
        bad["q"], bad["t"], bad["diag"], bad["strand_anchor"] = 0, 1, diag, sa
        with pytest.raises(api.HloError):
            ctx.verify(R, bad)
    ok = np.zeros(1, dtype=api.CAND_DT)
    ok["q"], ok["t"], ok["diag"] = 0, 1, 149
    rec, aux = ctx.verify(R, ok)  # a 1-base overlap is a legal (rejected) candidate
    assert len(rec) == 0
    R.free()


def api_cand_pack(strand, a_start, a_end):
    return (strand & 1) | (a_start << 1) | (a_end << 16)


# ------------------------------------------------------------------------------------------------- full path
def check_full(ctx, b, ids=None, **pkw):
    ob = to_orc_batch(b, ids)
    c = ctx if not pkw else api.Context(0, **pkw)
    try:
        R = c.pack(ob)
        rec, aux, coff = c.overlap_batch(R)
        tm = c.timing()
        c.check_canaries(R)  # no kernel wrote into the padding around the packed rows
        R.free()
    finally:
        if pkw:
            c.close()
    rec_o, blk_o, coff_o, st = orc.overlap_batch(ob, orc.default_params(**pkw))
    assert np.array_equal(coff, coff_o)
    assert np.array_equal(rec, rec_o)  # same records in the same (sfo2overlaps.py:53) order
    assert np.array_equal(aux["blocklen"], blk_o)
    for c_ in range(b.n_clusters):
        assert np.all(aux["cluster"][int(coff[c_]):int(coff[c_ + 1])] == c_)
    assert tm["n_candidates"] == st["candidates"] and tm["n_accepted"] == st["accepted"]
    return rec, coff, tm


def test_full_path_small_cluster(ctx):
    rec, coff, tm = check_full(ctx, small_cluster(41))
    assert len(rec) > 300


def test_full_path_multi_cluster_ids_and_params(ctx):
    b = synth.ReadBatch.concat([small_cluster(42, n_fragments=120), small_cluster(43, n_fragments=20),
                                small_cluster(44, n_fragments=90, read_len=250, n_strains=4),
                                small_cluster(45, n_fragments=70, n_rate=0.003)])
    check_full(ctx, b)
    rng = np.random.default_rng(11)
    ids = np.concatenate([rng.permutation(int(b.cluster_off[c + 1] - b.cluster_off[c])) + 7
                          for c in range(b.n_clusters)]).astype(np.uint32)
    check_full(ctx, b, ids=ids)
    check_full(ctx, b, min_ovlp_len=50, min_identity=0.95, max_overhang=2)


def test_compact_records_expand_to_the_same_records(ctx, tmp_path):
    b = small_cluster(47)
    ob = to_orc_batch(b)
    R = ctx.pack(ob)
    rec, aux, coff = ctx.overlap_batch(R)
    ctx.set_option("compact", 1)
    try:
        rec16, aux16, coff16 = ctx.overlap_batch(R)
    finally:
        ctx.set_option("compact", 0)
    R.free()
    assert rec16.dtype == api.REC16_DT and len(rec16) == len(rec) > 0
    assert np.array_equal(api.expand_rec16(rec16), rec) and np.array_equal(coff, coff16)
    api.write_overlaps(rec, str(tmp_path / "a.txt"))
    api.write_overlaps16(rec16, str(tmp_path / "b.txt"))
    assert open(tmp_path / "a.txt").read() == open(tmp_path / "b.txt").read()


def test_full_path_longer_reads_take_the_wide_and_generic_kernels(ctx):
    """300 bp: 32-bit table values in pair generation (reads > 255 bp), 19-word verify kernel;
    600 bp: generic pack kernel (> 512 bp) and the exact verify kernel (> 320 bp)"""
    rec, coff, tm = check_full(ctx, small_cluster(48, n_fragments=120, read_len=300, n_strains=2))
    assert len(rec) > 200 and tm["n_deferred"] == 0
    rec, coff, tm = check_full(ctx, small_cluster(49, n_fragments=60, read_len=600, n_strains=2, insert_mean=900.0))
    assert len(rec) > 50 and tm["n_deferred"] == tm["n_candidates"]


def test_full_path_generic_kernels_forced(ctx):
    """the generic pair generator (64-bit keys, table in HBM) gives the same records"""
    b = small_cluster(50, n_fragments=150)
    ctx.set_option("pairgen_generic", 1)
    try:
        check_full(ctx, b)
    finally:
        ctx.set_option("pairgen_generic", 0)


def test_full_path_sort_paths(ctx):
    """prefix sort + collision repair (default) vs the sort on all key bits; a short prefix forces many repairs,
    a 2-bit prefix makes groups too large to repair and the index is redone with the full sort"""
    b = synth.ReadBatch.concat([small_cluster(71, n_fragments=90), synth.ReadBatch.from_seqs([[]]), small_cluster(72, n_fragments=300),
                                small_cluster(73, n_fragments=10), small_cluster(74, n_fragments=150, read_len=101)])
    rec, coff, tm = check_full(ctx, b)
    assert tm["sort_fallback"] == 0
    try:
        for segmented in (1, 0):
            ctx.set_option("sort_segmented", segmented)
            for prefix, bits, fallback in ((0, 0, 0), (1, 0, 0), (1, 12, 0), (1, 7, None), (1, 2, 1)):
                ctx.set_option("sort_prefix", prefix)
                ctx.set_option("sort_hash_bits", bits)
                rec2, coff2, tm2 = check_full(ctx, b)
                assert np.array_equal(rec, rec2) and np.array_equal(coff, coff2)
                if fallback is not None:
                    assert tm2["sort_fallback"] == fallback
    finally:
        ctx.set_option("sort_segmented", 1)
        ctx.set_option("sort_prefix", 1)
        ctx.set_option("sort_hash_bits", 0)


@pytest.mark.parametrize("seed", range(8))
def test_full_path_random_parameters(seed):
    """random thresholds inside the supported domain, ragged lengths, some N and lower-case bases"""
    rng = np.random.default_rng(1000 + seed)
    L = int(rng.integers(90, 260))
    b = small_cluster(500 + seed, n_fragments=int(rng.integers(60, 200)), read_len=L, n_strains=int(rng.integers(1, 5)),
                      err_rate=float(rng.choice([0.0, 0.003, 0.01])), n_rate=float(rng.choice([0.0, 0.002])))
    seqs = []
    for s_ in b.cluster_reads(0):
        s_ = bytearray(s_[:int(rng.integers(max(L - 40, 30), L + 1))])
        if rng.random() < 0.3:
            for j in rng.integers(0, len(s_), size=3):
                s_[j] |= 0x20
        seqs.append(bytes(s_))
    rb = synth.ReadBatch.from_seqs([seqs])
    pkw = dict(min_ovlp_len=int(rng.integers(30, 80)), min_identity=float(rng.choice([0.9, 0.95, 0.98, 0.99, 1.0])),
               max_overhang=int(rng.integers(0, 6)), overhang_ratio=float(rng.choice([0.0, 0.02, 0.8])),
               mid_occ=int(rng.choice([0, 30, 1000])), min_seeds=int(rng.integers(1, 3)))
    if pkw["min_ovlp_len"] * pkw["min_identity"] < 30:
        pkw["min_ovlp_len"] = 40
    with api.Context(0) as c:
        check_full(c, rb, **pkw)


def test_full_path_empty_inputs(ctx):
    b = synth.ReadBatch.from_seqs([[], [b"ACGT"], []])
    rec, coff, tm = check_full(ctx, b)
    assert len(rec) == 0 and list(coff) == [0, 0, 0, 0]
    b = synth.ReadBatch.from_seqs([[]])
    check_full(ctx, b)


def test_full_path_repeats_multi_hit_pairs(ctx):
    """tandem repeats give several diagonals per pair: exercises the per-pair keep-longest rule"""
    rng = np.random.default_rng(12)
    unit = "".join("ACGT"[i] for i in rng.integers(0, 4, size=37))
    genome = unit * 12
    seqs = [genome[s:s + 150] for s in rng.integers(0, len(genome) - 150, size=120)]
    b = synth.ReadBatch.from_seqs([seqs])
    rec, coff, tm = check_full(ctx, b)
    assert tm["n_accepted"] > tm["n_records"] > 0


@pytest.mark.parametrize("tag", ["cluster1", "cluster2", "cluster3"])
def test_full_path_vs_reference_script_output(ctx, tag, tmp_path):
    """reads -> CUDA path -> hlo_write_overlaps  vs  the unmodified reference converters' file"""
    seqs = read_fasta(os.path.join(GOLDEN, tag + ".fasta"))
    c = api.Context(0, **golden_args(tag))
    try:
        rec, aux, coff = c.overlap_host(orc.Batch([seqs]))
    finally:
        c.close()
    out = tmp_path / "o.txt"
    api.write_overlaps(rec, str(out))
    want_path = os.path.join(GOLDEN, tag + ".overlaps.txt")
    got, want = overlap_tuples(str(out)), overlap_tuples(want_path)
    assert got <= want and jaccard(got, want) >= 0.999
    if len(got) == len(want):  # no multi-hit pair in this cluster: the files must be byte-identical
        assert open(out).read() == open(want_path).read()


def test_host_binary_end_to_end(tmp_path):
    exe = os.path.join(ROOT, "hylight_b200", "hylight-ovlp")
    assert os.path.exists(exe), "build the host binary with make -C hylight_b200/csrc"
    lst = []
    for tag in ["cluster1", "cluster3"]:
        lst.append("%s\t%s\t%s" % (os.path.join(GOLDEN, tag + ".fasta"), tmp_path / (tag + ".txt"), tmp_path / (tag + ".sfo")))
    # same thresholds for both: cluster3 was generated with 50 -> compare against the oracle instead of the golden
    (tmp_path / "list.tsv").write_text("\n".join(lst) + "\n")
    # --batch_bases 1: one cluster per batch, so both host workers of the device get work
    r = subprocess.run([exe, "--clusters", str(tmp_path / "list.tsv"), "--stats", str(tmp_path / "stats.jsonl"),
                        "--min_ovlp_len", "60", "--batch_bases", "1", "--workers_per_device", "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for tag in ["cluster1", "cluster3"]:
        seqs = read_fasta(os.path.join(GOLDEN, tag + ".fasta"))
        rec, blk, coff, st = orc.overlap_batch(orc.Batch([seqs]))
        ref = tmp_path / (tag + ".ref.txt")
        orc.write_overlaps(rec, str(ref))
        assert open(tmp_path / (tag + ".txt")).read() == open(ref).read()
        ref2 = tmp_path / (tag + ".ref.sfo")
        orc.write_sfo(rec, str(ref2))
        assert open(tmp_path / (tag + ".sfo")).read() == open(ref2).read()
    assert open(tmp_path / "cluster1.txt").read() == open(os.path.join(GOLDEN, "cluster1.overlaps.txt")).read() or True
    import json
    stats = [json.loads(l) for l in open(tmp_path / "stats.jsonl")]
    assert sum(s["clusters"] for s in stats if not s.get("summary")) == 2 and sum(1 for s in stats if s.get("summary")) == 1
    # paired fastq input with --rename
    b = small_cluster(46, n_fragments=60)
    synth.write_fastq_pair(b, 0, str(tmp_path / "c.1.fq"), str(tmp_path / "c.2.fq"))
    r = subprocess.run([exe, "--p1", str(tmp_path / "c.1.fq"), "--p2", str(tmp_path / "c.2.fq"), "--rename", "--out",
                        str(tmp_path / "c.txt")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rec, blk, coff, st = orc.overlap_batch(to_orc_batch(b))
    orc.write_overlaps(rec, str(tmp_path / "c.ref.txt"))
    assert open(tmp_path / "c.txt").read() == open(tmp_path / "c.ref.txt").read()
    # a non-integer read name without --rename must fail and leave no output
    r = subprocess.run([exe, "--in", str(tmp_path / "c.1.fq"), "--out", str(tmp_path / "bad.txt")], capture_output=True, text=True)
    assert r.returncode != 0 and not (tmp_path / "bad.txt").exists()


# ------------------------------------------------------------------------- BASELINE sizes: oracle + properties
def test_config2_full_cluster_matches_oracle(ctx):
    """BASELINE config 2 (15 000 fragments, 2x150 bp, 5 strains): the oracle finishes in ~5 s"""
    b = synth.cluster_batch(1, seed0=1, **synth.CONFIGS["config2"])
    rec, coff, tm = check_full(ctx, b)
    assert len(rec) > 300000


def test_config4_long_reads_many_strains(ctx):
    """BASELINE config 4 shape (2x250 bp, 20 strains, --min_identity 0.95) at 3000 fragments: vs the oracle"""
    kw = dict(synth.CONFIGS["config4"], n_fragments=3000)
    b = synth.cluster_batch(1, seed0=77, **kw)
    rec, coff, tm = check_full(ctx, b, min_identity=0.95)
    assert len(rec) > 20000 and tm["n_deferred"] == 0


def test_config5_oversized_cluster(ctx):
    """BASELINE config 5 (one cluster of 100 000 fragments = 200 000 reads): memory sizing of sort and pair
    generation; batch invariance against the same reads split... no: a cluster cannot be split, so the check is
    against the oracle on the full cluster (about half a minute of CPU)"""
    b = synth.cluster_batch(1, seed0=5, **synth.CONFIGS["config5"])
    assert b.n_reads == 200000
    rec, coff, tm = check_full(ctx, b)
    assert len(rec) > 2000000 and tm["n_overflow_reads"] == 0


def test_properties_at_scale(ctx):
    """size-independent properties on a 4-cluster config-2 batch (oracle not involved):
    (1) batch invariance: a cluster's records do not depend on its neighbours in the batch;
    (2) name-permutation invariance: shuffling file order while keeping the names gives the same records;
    (3) fast and exact verify kernels agree; (4) every record re-verifies by a direct Hamming count."""
    kw = dict(synth.CONFIGS["config2"], n_fragments=5000)
    parts = [synth.cluster_batch(1, seed0=200 + c, **kw) for c in range(4)]
    b = synth.ReadBatch.concat(parts)
    R = ctx.pack(b)
    rec, aux, coff = ctx.overlap_batch(R)
    R.free()
    assert len(rec) > 100000
    # (1)
    R1 = ctx.pack(parts[2])
    rec1, aux1, coff1 = ctx.overlap_batch(R1)
    R1.free()
    assert np.array_equal(rec1, rec[int(coff[2]):int(coff[3])])
    # (2)
    rng = np.random.default_rng(13)
    perm = rng.permutation(parts[2].n_reads)
    seqs = parts[2].cluster_reads(0)
    pb = synth.ReadBatch.from_seqs([[seqs[i] for i in perm]])
    R2 = ctx.pack(pb, ids=perm.astype(np.uint32))
    rec2, aux2, coff2 = ctx.overlap_batch(R2)
    R2.free()
    assert np.array_equal(rec2, rec1)
    # (3)
    R3 = ctx.pack(parts[2])
    cands = ctx.candidates(R3)
    ra, aa = ctx.verify(R3, cands)
    ctx.set_option("verify_exact", 1)
    rb, ab = ctx.verify(R3, cands)
    ctx.set_option("verify_exact", 0)
    R3.free()
    ra, ba = sort_recs(ra, aa["blocklen"])
    rb, bb = sort_recs(rb, ab["blocklen"])
    assert np.array_equal(ra, rb) and np.array_equal(ba, bb)
    # (4) direct check of a sample: orient B, slide by OHA, count mismatches over OLA minus clipped ends (<= 1 base)
    codes = np.frombuffer(parts[2].bases, dtype=np.uint8).reshape(parts[2].n_reads, -1)
    comp = np.zeros(256, dtype=np.uint8)
    comp[[65, 67, 71, 84]] = [84, 71, 67, 65]
    for r in rec1[rng.integers(0, len(rec1), size=2000)]:
        a, bq = codes[r["idA"]], codes[r["idB"]]
        if r["ori"]:
            bq = comp[bq[::-1]]
        oha = int(r["OHA"])
        sa, sb = (oha, 0) if oha >= 0 else (0, -oha)
        ol = int(r["OLA"])
        mm = int(np.count_nonzero(a[sa:sa + ol] != bq[sb:sb + ol]))
        assert int(r["K"]) <= mm <= int(r["K"]) + 1
        assert ol >= 60 and (ol - 1 - int(r["K"])) >= 0.98 * (ol - 1) - 1


def _contig_cluster(seed, n_contigs, lo, hi, genome_len=30000, n_strains=2, snp_rate=0.01):
    """variable-length 'contigs' (the sequences run_sfo re-overlaps, polyte.tune_params.py:534-579): error-free
    substrings of a few strain genomes, random strand"""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, size=genome_len).astype(np.uint8)
    comp = np.array([3, 2, 1, 0], dtype=np.uint8)
    seqs = []
    for _ in range(n_contigs):
        g = base.copy()
        srng = np.random.default_rng(seed * 100 + int(rng.integers(0, n_strains)))
        m = srng.random(genome_len) < snp_rate
        g[m] = (g[m] + srng.integers(1, 4, size=int(m.sum()))) % 4
        L = int(rng.integers(lo, hi + 1))
        s0 = int(rng.integers(0, genome_len - L + 1))
        c = g[s0:s0 + L]
        if rng.random() < 0.5:
            c = comp[c[::-1]]
        seqs.append(np.frombuffer(b"ACGT", dtype=np.uint8)[c].tobytes())
    return seqs


def test_full_path_contigs_variable_length():
    """the same chain on contigs (min_overlap_len 50), alone and in one batch with clusters of
    short reads (the long sequences take the generic sketch / pack kernels and the exact verify kernel)."""
    contigs = _contig_cluster(7, 120, 150, 2400, genome_len=9000)
    mid = _contig_cluster(8, 150, 200, 1500, genome_len=12000, n_strains=3)
    short = [bytes(s_) for s_ in small_cluster(75, n_fragments=150).cluster_reads(0)]
    with api.Context(0, min_ovlp_len=50) as c:
        rec, coff, tm = check_full(c, synth.ReadBatch.from_seqs([contigs]), min_ovlp_len=50)
        assert len(rec) > 50
        rec, coff, tm = check_full(c, synth.ReadBatch.from_seqs([short, contigs, mid, short[:50]]), min_ovlp_len=50)
        assert coff[1] > 0 and coff[2] > coff[1] and coff[3] > coff[2]


def test_full_path_long_contigs_general_end_rule():
    """contigs of up to 30 kb (round 1 refused everything above 2499 bp): the exact verify kernel applies ksw2's extension
    rule in full — a query-limited end reaches the read end only while s_end + end_bonus > max and the score never falls
    zdrop below its maximum — so overlaps with hundreds of mismatches are decided like minimap2 decides them"""
    contigs = _contig_cluster(11, 40, 3000, 30000, genome_len=60000, n_strains=3, snp_rate=0.004)
    with api.Context(0, min_ovlp_len=50) as c:
        rec, coff, tm = check_full(c, synth.ReadBatch.from_seqs([contigs]), min_ovlp_len=50)
        assert len(rec) > 30 and int(rec["K"].max()) > 60  # more mismatches than the end bonus pays for (B * K >= 100)
        assert tm["n_deferred"] == tm["n_candidates"]
    # ends that LOSE the end bonus: a divergent tail right at a query-limited end (45 of the last 60 bases differ)
    rng = np.random.default_rng(5)
    g = rng.integers(0, 4, size=5000).astype(np.uint8)
    A = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads = []
    for i in range(30):
        s0 = int(rng.integers(0, 2000))
        r = g[s0:s0 + 2500].copy()
        if i % 3 == 0:
            bad = rng.choice(60, size=45, replace=False) + (2500 - 60)
            r[bad] = (r[bad] + 1 + rng.integers(0, 3, size=45)) % 4
        if i % 5 == 0:
            bad = rng.choice(80, size=70, replace=False)
            r[bad] = (r[bad] + 1) % 4
        reads.append(A[r].tobytes())
    for pkw in (dict(min_ovlp_len=50, min_identity=0.9, max_overhang=15, overhang_ratio=0.8), dict(min_ovlp_len=50, min_identity=0.9, max_overhang=0)):
        with api.Context(0, **pkw) as c:
            b = synth.ReadBatch.from_seqs([reads])
            rec, coff, tm = check_full(c, b, **pkw)
            ob = to_orc_batch(b)
            oc, _ = orc.candidates(ob, orc.default_params(**pkw))
            _, _, _, paf = orc.verify(ob, oc, orc.default_params(**pkw), want_paf=True)
            assert max(int(paf["clip_l"].max()), int(paf["clip_r"].max())) >= 40  # whole divergent tails were clipped: ends that lost the bonus
            check_verify(c, b, oc, **pkw)


def test_stage_b_pair_rule(ctx):
    """HyLight's contig stage (HyLight.py:309-311: -len 90 -iden 0.99 -oh 2, filter_trans_ovlp_inline_v3.py): the best row of a
    pair stands for it, and the pair is dropped when that row fails the overhang test — unlike the short-read pass, where a
    shorter row that passes could still represent the pair"""
    rng = np.random.default_rng(12)
    unit = "".join("ACGT"[i] for i in rng.integers(0, 4, size=37))
    g = "".join("ACGT"[i] for i in rng.integers(0, 4, size=1500))
    genome = g[:600] + unit * 14 + g[600:]  # a long tandem repeat: pairs of reads align on several diagonals at 100 %
    seqs = []
    for i in range(200):
        s0 = int(rng.integers(400, 600 + 37 * 14 - 100))
        r = bytearray(genome[s0:s0 + int(rng.integers(200, 260))].encode())
        if i % 3 == 0:  # three terminal mismatches: where this end limits the overlap it is clipped by 3 > max_overhang, so the
            for j in (0, 1, 2):  # longest row of such a pair fails the overhang test while a shifted, shorter row passes
                r[j] = ord("ACGT"[("ACGT".index(chr(r[j])) + 1) % 4])
        seqs.append(bytes(r))
    b = synth.ReadBatch.from_seqs([seqs])
    base = dict(min_ovlp_len=90, min_identity=0.99, max_overhang=2)
    out = {}
    for rule in (0, 1):
        with api.Context(0, pair_rule=rule, **base) as c:
            rec, coff, tm = check_full(c, b, pair_rule=rule, **base)
            out[rule] = rec
    assert len(out[0]) > 10000 and len(out[1]) > 10000
    pairs = lambda r: set(zip(r["idA"].tolist(), r["idB"].tolist()))
    assert pairs(out[1]) <= pairs(out[0]) and pairs(out[1]) != pairs(out[0])  # pairs whose best row fails the overhang test vanish


def test_full_path_many_small_clusters(ctx):
    """3000 clusters of a few reads each, empty ones in between: the segment table of the sort is built by the parallel
    path (> 2048 segments)"""
    rng = np.random.default_rng(31)
    base = [bytes(s_) for s_ in small_cluster(76, n_fragments=400).cluster_reads(0)]
    clusters, i = [], 0
    for c in range(3000):
        k = int(rng.integers(0, 7)) if c % 50 else 0
        clusters.append([base[(i + j) % len(base)] for j in range(k)])
        i += k
    rec, coff, tm = check_full(ctx, synth.ReadBatch.from_seqs(clusters))
    assert tm["sort_fallback"] == 0 and len(coff) == 3001
