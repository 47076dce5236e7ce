#!/usr/bin/env python
"""SURVEY.md §8f row 1, measurement: HaploConduct's phred-weighted overlap_score (EdgeCalculator.cpp:26-139) on the
device (hlo_score_records) vs its CPU restatement, on the records of the bench workload.

    python tools/score_bench.py [clusters] > profiles/rN_score.json
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from hylight_b200 import api, synth  # noqa: E402
from oracle import orc  # noqa: E402


def main():
    n_clusters = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    b = synth.cluster_batch(n_clusters, seed0=5000, **bench.CLUSTER_KW)
    ob = orc.Batch.from_arrays(b.bases, b.base_off, b.cluster_off)
    rng = np.random.default_rng(1)
    quals = (rng.integers(2, 42, size=b.n_bases) + 33).astype(np.uint8)
    with api.Context(0) as c:
        R = c.pack(ob)
        rec, aux, coff = c.overlap_batch(R)
        times = []
        for _ in range(5):
            t0 = time.perf_counter()
            got = c.score_records(R, quals, rec, aux["cluster"])
            t1 = time.perf_counter()
            tm = c.timing()
            times.append(((t1 - t0) * 1e3, tm["h2d_ms"], tm["verify_ms"], tm["d2h_ms"]))
        R.free()
    # CPU restatement on the first cluster's records (single thread)
    n1 = int(coff[1])
    ob1 = orc.Batch.from_arrays(b.bases[:int(b.base_off[b.cluster_off[1]])], b.base_off[:int(b.cluster_off[1]) + 1], b.cluster_off[:2])
    t0 = time.perf_counter()
    want = orc.score_records(ob1, quals[:len(ob1.bases)], rec[:n1], aux["cluster"][:n1], 0.0, 0)
    cpu_s = time.perf_counter() - t0
    ok_rate = bool(np.array_equal(got["mismatch_rate"][:n1], want[:, 1]))
    ok_score = bool(np.allclose(got["score"][:n1], want[:, 0], rtol=1e-12, atol=0.0))
    best = min(times[1:], key=lambda t: t[0])
    ovl = rec["OLA"].astype(np.int64)
    alg_bytes = int((32 + 16 + 2 * ((ovl + 3) // 4) + 2 * ovl).sum())  # record in, score out, 2-bit bases and phreds of both reads
    print(json.dumps({
        "metric": "scored_overlaps_per_sec", "records": int(len(rec)), "clusters": n_clusters,
        "kernel_ms": best[2], "kernel_records_per_s": len(rec) / best[2] * 1e3,
        "kernel_algorithmic_GBps": alg_bytes / best[2] / 1e6, "algorithmic_bytes": alg_bytes,
        "call_ms_host_buffers": best[0], "h2d_ms": best[1], "d2h_ms": best[3],
        "call_records_per_s": len(rec) / best[0] * 1e3,
        "cpu_port": {"records": n1, "seconds": cpu_s, "records_per_s": n1 / cpu_s, "threads": 1},
        "parity_on_cluster0": {"mismatch_rate_bit_equal": ok_rate, "score_rel_1e-12": ok_score},
        "mean_overlap": float(ovl.mean()),
    }))


if __name__ == "__main__":
    main()
