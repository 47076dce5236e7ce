#!/usr/bin/env python
"""Stage times of the device-resident batch on one GPU: N config-2 clusters per batch, a few distinct batches in turn.

    python tools/stage_times.py [--clusters 32] [--batches 2] [--reps 10] [--config config2]
    HLO_LIB=hylight_b200/libhlovlp_x.so python tools/stage_times.py      # a variant build of the library
    ncu --set full -k regex:pairgen_group -s 4 -c 1 -o gpurun_out/pg python tools/stage_times.py --reps 2

The small driver for A/B runs and for ncu captures (bench.py builds the whole 1000-cluster job first).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hylight_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clusters", type=int, default=32)
ap.add_argument("--batches", type=int, default=2)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--config", default="config2")
args = ap.parse_args()

names = ["pack_ms", "sketch_ms", "sort_ms", "index_ms", "pairgen_ms", "verify_ms", "verify_fast_ms", "dedup_ms", "total_ms"]
with api.Context(0) as ctx:
    comps = [synth.cluster_batch(args.clusters, seed0=1000 + i * args.clusters, **synth.CONFIGS[args.config]) for i in range(args.batches)]
    Rs = [ctx.pack(c) for c in comps]
    ctx.set_option("repack", 1)
    ctx.set_option("d2h", 0)
    ctx.set_option("compact", 2)
    ctx.set_option("digest", 1)
    log = []
    for rep in range(args.warmup + args.reps):
        for R in Rs:
            r = api.Recs()
            ctx._check(api.lib().hlo_overlap_batch(ctx.h, R.handle, r))
            if rep >= args.warmup:
                log.append(ctx.timing())
    out = {k: round(float(np.mean([t[k] for t in log if k in t])), 4) for k in names if k in log[0]}
    out["n_seed_hits"] = log[0]["n_seed_hits"]
    out["n_records"] = log[0]["n_records"]
    out["digests"] = ["%016x" % t["digest"] for t in log[:args.batches]]
    out["lib"] = os.environ.get("HLO_LIB", "default")
    print(json.dumps(out))
    for R in Rs:
        R.free()
