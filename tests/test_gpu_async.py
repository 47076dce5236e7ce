"""The asynchronous batch pipeline (hlo_submit / hlo_wait), its rerun path, the compact record forms, the fused score,
the record digest and the data-derived occurrence cut-off — all against the CPU oracle, bit-exact."""
import os
import threading

import numpy as np
import pytest

from hylight_b200 import api, synth
from oracle import orc

from util import random_reads, small_cluster, to_orc_batch

pytestmark = pytest.mark.gpu


def oracle_full(b, **pkw):
    ob = to_orc_batch(b)
    rec, blk, coff, st = orc.overlap_batch(ob, orc.default_params(**pkw))
    return ob, rec, blk, coff, st


def test_submit_wait_equals_the_oracle_and_waits_twice():
    b = synth.ReadBatch.concat([small_cluster(201, n_fragments=150), small_cluster(202, n_fragments=40),
                                small_cluster(203, n_fragments=90, n_rate=0.004)])
    ob, rec_o, blk_o, coff_o, st = oracle_full(b)
    with api.Context(0) as c:
        for it in range(3):
            c.submit(b)
            rec, aux, coff = c.wait()
            tm = c.timing()
            c.check_canaries()
            assert np.array_equal(rec, rec_o) and np.array_equal(coff, coff_o) and np.array_equal(aux["blocklen"], blk_o)
            assert tm["n_candidates"] == st["candidates"] and tm["n_accepted"] == st["accepted"]
            if it > 0:  # buffer estimates are settled: one wait for the counters, one for the download
                assert tm["n_reruns"] == 0 and tm["n_syncs"] == 2, tm
        with pytest.raises(api.HloError):
            c.wait()  # nothing in flight
        c.submit(b)
        with pytest.raises(api.HloError):
            c.submit(b)  # one batch per ctx
        R = None
        with pytest.raises(api.HloError):
            R = c.pack(b)  # the synchronous entry points refuse while a batch is in flight
        rec, aux, coff = c.wait()
        assert np.array_equal(rec, rec_o)


def test_rerun_path_when_the_estimates_are_too_small():
    """first-batch buffer estimates scaled down to 3 %: minimizers, candidates and records all overflow once and the
    batch is enqueued again with what the counters said"""
    b = synth.ReadBatch.concat([small_cluster(211, n_fragments=400, n_strains=1, strain_cov=30.0), small_cluster(212, n_fragments=100)])
    ob, rec_o, blk_o, coff_o, st = oracle_full(b)
    with api.Context(0) as c:
        c.set_option("cap_scale_pct", 3)
        c.set_option("digest", 1)
        c.submit(b)
        rec, aux, coff = c.wait()
        tm = c.timing()
        assert tm["n_reruns"] >= 3, tm  # minimizers, then candidates, then records
        assert np.array_equal(rec, rec_o) and np.array_equal(coff, coff_o)
        assert tm["digest"] == api.records_digest(rec)
        c.submit(b)
        rec, aux, coff = c.wait()
        assert c.timing()["n_reruns"] == 0 and np.array_equal(rec, rec_o)  # learned
        # the parity entry points take the same path
        c.set_option("forget_sizes", 1)
        R = c.pack(ob)
        cands = c.candidates(R)
        oc, _ = orc.candidates(ob)
        key = lambda a: np.lexsort((a["diag"], a["strand_anchor"] & 1, a["t"], a["q"]))
        assert np.array_equal(cands[key(cands)], oc[key(oc)])
        c.set_option("forget_sizes", 1)
        mz, ro = c.sketch(R)
        mz_o, ro_o = orc.sketch_all(ob)
        assert np.array_equal(mz, mz_o) and np.array_equal(ro, ro_o)
        R.free()


def test_compact_record_forms():
    b = synth.ReadBatch.concat([small_cluster(221, n_fragments=120, read_len=150), small_cluster(222, n_fragments=80, read_len=250)])
    ob, rec_o, blk_o, coff_o, st = oracle_full(b)
    lens = np.diff(b.base_off).astype(np.int64)
    with api.Context(0) as c:
        c.set_option("aux", 0)
        for form, dt in ((1, api.REC16_DT), (2, api.REC8_DT)):
            c.set_option("compact", form)
            c.submit(b)
            rec, aux, coff = c.wait()
            assert rec.dtype == dt and aux is None and np.array_equal(coff, coff_o)
            if form == 1:
                full = api.expand_rec16(rec)
            else:
                full = np.concatenate([api.expand_rec8(rec[int(coff[k]):int(coff[k + 1])],
                                                       lens[int(b.cluster_off[k]):int(b.cluster_off[k + 1])])
                                       for k in range(b.n_clusters)])
            assert np.array_equal(full, rec_o)
        # names of 65536 and above do not fit the 8-byte form: the batch comes back in the 16-byte form
        ids = (np.arange(b.n_reads, dtype=np.uint32) - np.repeat(b.cluster_off[:-1], np.diff(b.cluster_off))) + 70000
        c.submit(b, ids=ids)
        rec, aux, coff = c.wait()
        assert rec.dtype == api.REC16_DT
        rec_i, _, coff_i, _ = orc.overlap_batch(to_orc_batch(b, ids))
        assert np.array_equal(api.expand_rec16(rec), rec_i)


def test_fused_score_equals_the_standalone_call_and_the_oracle():
    b = synth.ReadBatch.concat([small_cluster(231, n_fragments=150), small_cluster(232, n_fragments=60, n_rate=0.004)])
    ob, rec_o, blk_o, coff_o, st = oracle_full(b)
    rng = np.random.default_rng(4)
    quals = (rng.integers(2, 42, size=b.n_bases) + 33).astype(np.uint8)
    clu = np.repeat(np.arange(b.n_clusters, dtype=np.uint32), np.diff(coff_o).astype(np.int64))
    for mp, mrl in ((0.0, 0), (1e-3, 0), (0.0, 151)):
        want = orc.score_records(ob, quals, rec_o, clu, mp, mrl)
        with api.Context(0) as c:
            c.set_option("scores", 1)
            c.set_option("compact", 2)
            c.set_score_params(mp, mrl)
            with pytest.raises(api.HloError):
                c.submit(b)  # scores without qualities
            c.submit(b, quals=quals)
            rec, aux, coff = c.wait()
            got = c.last_score
            assert got is not None and len(got) == len(rec_o)
            assert np.array_equal(got["mismatch_rate"], want[:, 1])
            assert np.allclose(got["score"], want[:, 0], rtol=1e-12, atol=0.0)
            assert np.array_equal(got["score"] == 0.0, want[:, 0] == 0.0)
            tm = c.timing()
            assert tm["score_ms"] > 0
            # and the reference's own code, on a sample (oracle/_ref/libecref.so, when it was built)
            if orc.ecref() is not None:
                for i in range(0, len(rec_o), 97):
                    sc, mr = orc.ecref_score_record(ob, quals, rec_o[i], int(ob.cluster_off[clu[i]]), mp, mrl)
                    assert mr == got["mismatch_rate"][i] and abs(sc - got["score"][i]) <= 1e-12 * abs(sc)


def repetitive_cluster(seed, n_reads=8000, genome_len=400000):
    rng = np.random.default_rng(seed)
    genome = random_reads(rng, 1, genome_len, genome_len)[0]
    rep = random_reads(rng, 1, 60, 60)[0]
    reads = []
    for i in range(n_reads):
        s = int(rng.integers(0, len(genome) - 150))
        r = bytearray(genome[s:s + 150])
        if i % 3 == 0:
            r[40:100] = rep
        reads.append(bytes(r))
    return reads


def test_data_derived_mid_occ_matches_the_oracle():
    """minimap2's default occurrence cut-off (quantile of the index) on a cluster where it bites, next to a shallow one"""
    b = synth.ReadBatch.from_seqs([repetitive_cluster(3), small_cluster(241, n_fragments=100).cluster_reads(0)])
    ob = to_orc_batch(b)
    mo = orc.mid_occ(ob)
    assert mo[0] < 2000 and mo[0] != mo[1]
    rec_o, blk_o, coff_o, st = orc.overlap_batch(ob)
    with api.Context(0) as c:
        c.submit(b)
        rec, aux, coff = c.wait()
        tm = c.timing()
        assert (tm["mid_occ_lo"], tm["mid_occ_hi"]) == (int(mo.min()), int(mo.max()))
        assert tm["n_seed_hits"] == st["seed_hits"] and tm["n_candidates"] == st["candidates"]
        assert np.array_equal(rec, rec_o) and np.array_equal(coff, coff_o)
    # an explicit cut-off overrides it, on both sides
    with api.Context(0, mid_occ=500) as c:
        c.submit(b)
        rec, aux, coff = c.wait()
        rec_o2, _, _, st2 = orc.overlap_batch(ob, orc.default_params(mid_occ=500))
        assert np.array_equal(rec, rec_o2) and c.timing()["n_seed_hits"] == st2["seed_hits"] != st["seed_hits"]


def test_three_contexts_pipelined_from_threads():
    batches = [synth.ReadBatch.concat([small_cluster(250 + 3 * i + j, n_fragments=60 + 20 * j) for j in range(3)]) for i in range(6)]
    want = [oracle_full(b)[1] for b in batches]
    got = [None] * len(batches)
    ctxs = [api.Context(0) for _ in range(3)]
    for c in ctxs:
        c.set_option("digest", 1)
    nxt, lock, errs = [0], threading.Lock(), []

    def work(c):
        try:
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= len(batches) * 3:
                    return
                c.submit(batches[i % len(batches)])
                rec, aux, coff = c.wait()
                assert c.timing()["digest"] == api.records_digest(rec)
                got[i % len(batches)] = rec
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(c,)) for c in ctxs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for c in ctxs:
        c.close()
    assert not errs, errs
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_fast_verify_survives_a_long_read_batch_on_the_same_ctx():
    """the threshold tables of a ctx only grow; the fast verify kernel must still be chosen by the BATCH's longest read"""
    long_b = small_cluster(261, n_fragments=40, read_len=1100, n_strains=1, insert_mean=1500.0, err_rate=0.001)
    short_b = small_cluster(262, n_fragments=150)
    ob, rec_o, blk_o, coff_o, st = oracle_full(short_b)
    with api.Context(0) as c:
        c.submit(long_b)
        c.wait()
        assert c.timing()["n_deferred"] == c.timing()["n_candidates"] > 0  # reads > 320 bp: the exact kernel
        c.submit(short_b)
        rec, aux, coff = c.wait()
        tm = c.timing()
        assert np.array_equal(rec, rec_o)
        assert tm["n_deferred"] < tm["n_candidates"] // 10 and tm["verify_fast_ms"] > 0


def test_documented_stub_on_device(tmp_path):
    from test_abi import documented_stub
    ns = documented_stub()
    b = small_cluster(271, n_fragments=80)
    ob, rec_o, blk_o, coff_o, st = oracle_full(b)
    n = ns["gpu_overlaps"](b.cluster_reads(0), str(tmp_path / "o.txt"))
    orc.write_overlaps(rec_o, str(tmp_path / "c.txt"))
    assert n == len(rec_o) and open(tmp_path / "o.txt", "rb").read() == open(tmp_path / "c.txt", "rb").read()


def test_config4_full_size_cluster_matches_the_oracle():
    """BASELINE configs[3] at full size: one cluster of 15 000 fragments, 2 x 250 bp, 20 strains, min_identity 0.95 — the
    regime where the pair generator's table cascade and the large idA groups of the records tail fire"""
    kw = dict(synth.CONFIGS["config4"])
    b = synth.cluster_batch(1, seed0=4004, **kw)
    ob = to_orc_batch(b)
    p = orc.default_params(min_identity=0.95)
    rec_o, blk_o, coff_o, st = orc.overlap_batch(ob, p)
    with api.Context(0, min_identity=0.95) as c:
        c.set_option("digest", 1)
        c.submit(b)
        rec, aux, coff = c.wait()
        tm = c.timing()
    assert tm["n_seed_hits"] == st["seed_hits"] and tm["n_candidates"] == st["candidates"] and tm["n_accepted"] == st["accepted"]
    assert len(rec) == len(rec_o) > 3_000_000
    assert np.array_equal(rec, rec_o) and np.array_equal(aux["blocklen"], blk_o)
    assert tm["digest"] == api.records_digest(rec)


def test_packed_2bit_input_gives_the_same_records():
    """hlo_batch.packed: the reads arrive 2-bit packed (a quarter of the upload), ambiguous bases as a list"""
    rng = np.random.default_rng(8)
    b1 = small_cluster(281, n_fragments=150, n_rate=0.004)
    ragged = []
    for s_ in small_cluster(282, n_fragments=90, read_len=250).cluster_reads(0):
        s_ = bytearray(s_[:int(rng.integers(100, 251))])
        if rng.random() < 0.3:
            s_[int(rng.integers(0, len(s_)))] |= 0x20
        ragged.append(bytes(s_))
    b = synth.ReadBatch.concat([b1, synth.ReadBatch.from_seqs([ragged]), synth.ReadBatch.from_seqs([[]]), small_cluster(283, n_fragments=40)])
    ob, rec_o, blk_o, coff_o, st = oracle_full(b)
    pk = api.pack2b(b)
    assert len(pk[2]) > 0
    with api.Context(0) as c:
        c.set_option("digest", 1)
        for it in range(2):
            c.submit_packed(pk, b.cluster_off)
            rec, aux, coff = c.wait()
            assert np.array_equal(rec, rec_o) and np.array_equal(coff, coff_o) and np.array_equal(aux["blocklen"], blk_o)
        assert c.timing()["digest"] == api.records_digest(rec_o)
        c.submit(b)  # and the ASCII form on the same ctx right after
        rec, aux, coff = c.wait()
        assert np.array_equal(rec, rec_o)


def test_submit_wait_degenerate_batches():
    """no reads, empty clusters, a single read, reads shorter than k: nothing to report, nothing to crash on"""
    with api.Context(0) as c:
        c.set_option("compact", 2)
        for seqs in ([[]], [[], [b"ACGT"], []], [[b"A" * 150]], [[b"ACGTACGTAC", b"ACGTACGTAC"]], [[b"N" * 200, b"N" * 200]]):
            b = synth.ReadBatch.from_seqs(seqs)
            c.submit(b)
            rec, aux, coff = c.wait()
            assert rec is None or len(rec) == 0
            assert coff is None or list(coff) == [0] * (len(seqs) + 1)
            pk = api.pack2b(b)
            c.submit_packed(pk, b.cluster_off)
            rec, aux, coff = c.wait()
            assert rec is None or len(rec) == 0
        # and a real batch right after on the same context
        b = small_cluster(291, n_fragments=60)
        ob, rec_o, blk_o, coff_o, st = oracle_full(b)
        c.set_option("compact", 0)
        c.submit(b)
        rec, aux, coff = c.wait()
        assert np.array_equal(rec, rec_o)
