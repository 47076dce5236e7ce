#!/usr/bin/env python
"""Randomised parity sweep: hlo_submit / hlo_wait on the GPU against the CPU oracle, record for record, on batches the
fixed tests do not enumerate — ragged read lengths (30 .. 600 bp in one batch), several clusters of very different
depth, ambiguous and lower-case bases, low-complexity inserts, random (k, w), thresholds, pair rules, occurrence
cut-offs and pair-generator table sizes.

    python tools/fuzz_parity.py [--seconds 150] [--seed 1] > profiles/rN_fuzz.log

Exit status 0 and a final "N batches, all identical" line, or the first differing batch with its seed.
(oracle/ is the checker here, as in tests/: this is test infrastructure.)
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hylight_b200 import api, synth  # noqa: E402
from oracle import orc  # noqa: E402
from util import to_orc_batch  # noqa: E402

def mutate(rng, b, n_rate, lower_rate, lowc_rate):
    """ambiguous bases, lower case, and low-complexity stretches written over the reads of a batch"""
    bases = np.array(b.bases, dtype=np.uint8, copy=True)
    if n_rate:
        bases[rng.random(len(bases)) < n_rate] = ord("N")
    if lower_rate:
        m = rng.random(len(bases)) < lower_rate
        bases[m] |= 0x20
    if lowc_rate:
        off = np.asarray(b.base_off, dtype=np.int64)
        for r in np.nonzero(rng.random(len(off) - 1) < lowc_rate)[0]:
            L = int(off[r + 1] - off[r])
            if L < 40:
                continue
            unit = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(rng.integers(1, 7)))]
            n = int(rng.integers(20, L))
            s = int(rng.integers(0, L - n + 1))
            bases[off[r] + s: off[r] + s + n] = np.resize(unit, n)
    return synth.ReadBatch(bases, b.base_off, b.cluster_off)


def ragged(rng, b, lo):
    """cut every read at a random length >= lo (a batch then mixes many lengths; offsets are rebuilt)"""
    off = np.asarray(b.base_off, dtype=np.int64)
    keep, new_off = [], [0]
    for r in range(len(off) - 1):
        L = int(off[r + 1] - off[r])
        n = L if rng.random() < 0.5 else int(rng.integers(min(lo, L), L + 1))
        keep.append(np.arange(off[r], off[r] + n))
        new_off.append(new_off[-1] + n)
    return synth.ReadBatch(np.asarray(b.bases, dtype=np.uint8)[np.concatenate(keep)], np.asarray(new_off, dtype=np.uint64), b.cluster_off)


def make_case(seed, gen=1):
    """the batch, the parameters and the pair-generator table size of one seed (tests/test_gpu_fuzz_cases.py replays the seeds
    that once differed).  gen 2 adds what generation 1 did not draw: reads of 1 and 3 kb, read names of the caller's own,
    the 16- and 8-byte record forms — behind a switch, so that the seeds recorded under generation 1 keep their batches."""
    r = np.random.default_rng(seed)
    k = int(r.choice([15, 17, 19, 21, 21, 21, 23]))
    w = int(r.choice([5, 8, 11, 11, 11, 14]))
    if gen >= 4 and seed % 2 == 0:  # gen 4, every other seed: even k (palindromic k-mers are skipped), windows off the specialised set
        k, w = int(np.random.default_rng(seed ^ 0x5bd1e995).choice([14, 16, 18, 20, 22])), int(np.random.default_rng(seed ^ 0x27d4eb2f).choice([3, 7, 19, 25, 32]))
    pkw = dict(k=k, w=w, min_ovlp_len=int(r.choice([30, 50, 60, 60, 90])), min_identity=float(r.choice([0.9, 0.95, 0.98, 0.98, 0.99])),
               max_overhang=int(r.choice([0, 1, 1, 2, 4])), pair_rule=int(r.choice([0, 0, 1])),
               mid_occ=int(r.choice([0, 0, 0, 10, 25, 1000])), min_seeds=int(r.choice([1, 2, 2, 3])),
               min_chain_score=int(r.choice([k, 30, 30, 40])))
    clusters = []
    deep = gen >= 3 and seed % 4 == 0  # gen 3, every fourth seed: one deep pile (hundreds of chains per read: the table-size cascade)
    for c in range(int(r.integers(1, 5))):
        read_len = int(r.choice([60, 100, 150, 150, 250, 400, 600] + ([1000, 3000] if gen >= 2 else [])))
        if deep and c == 0:
            clusters.append(synth.cluster_batch(1, seed0=seed, n_fragments=int(r.integers(300, 1500)), read_len=int(r.choice([100, 150, 250])),
                                                n_strains=int(r.integers(1, 4)), strain_cov=float(r.choice([100.0, 300.0, 1e9])),
                                                snp_rate=0.01, err_rate=0.005, insert_mean=float(r.choice([260.0, 450.0])), insert_sd=5.0))
            continue
        clusters.append(synth.cluster_batch(1, seed0=seed + c, n_fragments=int(r.integers(3, 260 if read_len < 1000 else 50)), read_len=read_len,
                                            n_strains=int(r.integers(1, 5)), strain_cov=float(r.choice([4.0, 8.0, 15.0, 40.0])),
                                            snp_rate=float(r.choice([0.0, 0.005, 0.01, 0.03])), err_rate=float(r.choice([0.0, 0.002, 0.005, 0.02])),
                                            insert_mean=float(max(read_len + 20, r.choice([200.0, 450.0, 900.0])))))
    b = synth.ReadBatch.concat(clusters)
    if r.random() < 0.5:
        b = ragged(r, b, 30)
    b = mutate(r, b, float(r.choice([0.0, 0.0, 0.002, 0.02])), float(r.choice([0.0, 0.0, 0.1])), float(r.choice([0.0, 0.0, 0.05, 0.3])))
    slots = int(r.choice([8, 8, 8, 6, 5]))
    if gen < 2:
        return b, pkw, slots
    ids = None
    if r.random() < 0.3:  # names of the caller's own: any distinct integers below 65 536 per cluster
        ids = np.concatenate([r.permutation(int(b.cluster_off[c + 1] - b.cluster_off[c])) * int(r.integers(1, 4)) + int(r.integers(0, 9))
                              for c in range(b.n_clusters)]).astype(np.uint32)
    form = int(r.choice([0, 1, 2]))
    if gen < 3:
        return b, pkw, slots, ids, form
    quals = (r.integers(2, 42, size=int(b.base_off[-1])) + 33).astype(np.uint8) if r.random() < 0.4 else None  # gen 3: fused scores
    if ids is not None:
        quals = None  # (the consumer indexes its reads by name: scores exist for names 0..N-1 only, as rename_fas.py leaves them)
    return b, pkw, slots, ids, form, quals, (float(r.choice([0.0, 1e-3])), int(r.choice([0, 0, 120])))


def full_records(rec, coff, b, ids):
    """the 32-byte form of whatever form came back (hlo_rec16_expand / hlo_rec8_expand of the header, cluster by cluster:
    the 8-byte form needs the lengths of the cluster's reads by name)"""
    if rec.dtype == api.REC_DT:
        return rec
    if rec.dtype == api.REC16_DT:
        return api.expand_rec16(rec)
    lens = (np.asarray(b.base_off[1:], dtype=np.int64) - np.asarray(b.base_off[:-1], dtype=np.int64))
    out = []
    for c in range(b.n_clusters):
        r0, r1 = int(b.cluster_off[c]), int(b.cluster_off[c + 1])
        names = np.arange(r1 - r0) if ids is None else np.asarray(ids[r0:r1], dtype=np.int64)
        by_name = np.zeros(int(names.max()) + 1 if len(names) else 1, dtype=np.int64)
        by_name[names] = lens[r0:r1]
        out.append(api.expand_rec8(rec[int(coff[c]):int(coff[c + 1])], by_name))
    return np.concatenate(out) if out else np.zeros(0, dtype=api.REC_DT)


def run_case(seed, gen=1, ctx=None, pkw_fixed=None):
    """(same?, details) for one seed; raises api.HloError when the library refuses the parameters.
    ctx + pkw_fixed: run the case's batch on a context that lives across cases (its parameters replace the case's): buffer
    estimates, table-size hints and counters then carry over from whatever batch came before"""
    case = make_case(seed, gen)
    b, pkw, slots = case[:3]
    if pkw_fixed is not None:
        pkw = pkw_fixed
    ids, form = (case[3], case[4]) if gen >= 2 else (None, 0)
    quals, score_prm = (case[5], case[6]) if gen >= 3 else (None, None)
    ob = to_orc_batch(b, ids)
    rec_o, blk_o, coff_o, st = orc.overlap_batch(ob, orc.default_params(**pkw))
    import contextlib
    with (contextlib.nullcontext(ctx) if ctx is not None else api.Context(0, **pkw)) as c:
        c.set_option("pairgen_slots_log2", slots)
        c.set_option("compact", 0)
        c.submit(b, ids=ids)
        rec, aux, coff = c.wait()
        tm = c.timing()
        c.set_option("compact", form)
        if ids is None:
            pk = api.pack2b(b)
            c.submit_packed(pk, b.cluster_off)
        else:
            c.submit(b, ids=ids)
        rec_p, _, coff_p = c.wait()
        rec_p = full_records(rec_p, coff_p, b, ids)
        score_ok = True
        if quals is not None:  # the consumer's overlap score, fused into the batch, against the CPU port of EdgeCalculator::overlap_score
            c.set_option("scores", 1)
            c.set_score_params(*score_prm)
            c.submit(b, quals=quals, ids=ids)
            rec_s, _, coff_s = c.wait()
            got = c.last_score
            c.set_option("scores", 0)
            clu = np.repeat(np.arange(b.n_clusters, dtype=np.uint32), np.diff(coff_o).astype(np.int64))
            want = orc.score_records(ob, quals, rec_o, clu, *score_prm)
            score_ok = (got is not None and len(got) == len(rec_o) and np.array_equal(full_records(rec_s, coff_s, b, ids), rec_o) and
                        np.array_equal(got["mismatch_rate"], want[:, 1]) and np.allclose(got["score"], want[:, 0], rtol=1e-12, atol=0.0) and
                        np.array_equal(got["score"] == 0.0, want[:, 0] == 0.0))
    same = score_ok and (np.array_equal(rec, rec_o) and np.array_equal(coff, coff_o) and np.array_equal(aux["blocklen"], blk_o) and
            np.array_equal(rec_p, rec_o) and np.array_equal(coff_p, coff_o) and
            tm["n_candidates"] == st["candidates"] and tm["n_accepted"] == st["accepted"] and tm["n_seed_hits"] == st["seed_hits"])
    info = dict(b=b, pkw=pkw, slots=slots, rec=rec, rec_o=rec_o, rec_p=rec_p, tm=tm, st=st)
    return same, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=150.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--replay", type=int, default=0, help="run the batch of this seed only and print what differs")
    ap.add_argument("--gen", type=int, default=2, help="case generator: 1 = the first sweep's, 2 = + long reads, own names, compact record forms, 3 = + fused scores, deep piles, 4 = + even k, other windows")
    ap.add_argument("--reuse", type=int, default=0, help="N > 0: one context (the parameters of its first case) serves N consecutive cases")
    args = ap.parse_args()
    shared, shared_pkw, shared_left = None, None, 0
    rng = np.random.default_rng(args.seed)
    t_end = time.time() + args.seconds
    n_batches = n_records = 0
    while time.time() < t_end:
        seed = args.replay or int(rng.integers(1, 1 << 30))
        try:
            if args.reuse and not args.replay:
                if shared_left == 0:
                    if shared is not None:
                        shared.close()
                        shared = None
                    shared_pkw = make_case(seed, args.gen)[1]
                    shared, shared_left = api.Context(0, **shared_pkw), args.reuse
                shared_left -= 1
                same, x = run_case(seed, args.gen, shared, shared_pkw)
            else:
                same, x = run_case(seed, args.gen)
        except api.HloError as e:  # a refusal with a message is a result too
            shared_left = 0
            print("seed %d: refused: %s" % (seed, e))
            if args.replay:
                break
            continue
        b, pkw, slots, rec, rec_o, rec_p, tm, st = (x[k] for k in ("b", "pkw", "slots", "rec", "rec_o", "rec_p", "tm", "st"))
        if not same:
            print("DIFFERENT: gen %d seed %d params %s slots %d: %d vs %d records, candidates %d vs %d, hits %d vs %d, accepted %d vs %d, packed input %d records"
                  % (args.gen, seed, pkw, slots, len(rec), len(rec_o), tm["n_candidates"], st["candidates"], tm["n_seed_hits"], st["seed_hits"],
                     tm["n_accepted"], st["accepted"], len(rec_p)))
            got = {tuple(int(v) for v in r_) for r_ in rec.tolist()}
            want = {tuple(int(v) for v in r_) for r_ in rec_o.tolist()}
            print("fields:", rec_o.dtype.names)
            for r_ in sorted(want - got)[:10]:
                print("  only the oracle:", r_)
            for r_ in sorted(got - want)[:10]:
                print("  only the GPU:   ", r_)
            sys.exit(1)
        if args.replay:
            print("seed %d: identical (%d records)" % (seed, len(rec_o)))
            return
        n_batches += 1
        n_records += len(rec_o)
        if n_batches % 20 == 0:
            print("%d batches, %d records so far (last: %d reads, %d clusters, k=%d w=%d, %d overflow reads, %d deferred)"
                  % (n_batches, n_records, b.n_reads, b.n_clusters, pkw["k"], pkw["w"], tm["n_overflow_reads"], tm["n_deferred"]), flush=True)
    print("%d batches, %d records, all identical to the oracle (ASCII and packed input; generator %d)" % (n_batches, n_records, args.gen))


if __name__ == "__main__":
    main()
