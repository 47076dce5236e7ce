#!/usr/bin/env python
"""BASELINE configs 2, 4 and 5 on one GPU: device time per stage, records, and the device memory the context holds
afterwards (config 5 = memory sizing of sort + pair generation, budget 180 GB).

    python tools/configs_report.py > profiles/rN_configs.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from hylight_b200 import api, synth  # noqa: E402

CASES = [
    ("config2 x32", "config2", 32, {}),
    ("config4 x16 (250 bp, 20 strains, min_identity 0.95)", "config4", 16, dict(min_identity=0.95)),
    ("config5 (one cluster of 200 000 reads)", "config5", 1, {}),
]


def main():
    out = []
    for name, key, n_clusters, pkw in CASES:
        b = synth.cluster_batch(n_clusters, seed0=7000, **synth.CONFIGS[key])
        torch.cuda.synchronize()
        free0, total = torch.cuda.mem_get_info()
        with api.Context(0, **pkw) as c:
            c.set_option("d2h", 0)
            R = c.pack(b)
            best = None
            for _ in range(4):
                c.overlap_batch(R)
                tm = c.timing()
                if best is None or tm["total_ms"] < best["total_ms"]:
                    best = tm
            torch.cuda.synchronize()
            free1, _ = torch.cuda.mem_get_info()
            R.free()
        out.append({"case": name, "reads": int(b.n_reads), "bases": int(b.n_bases), "records": int(best["n_records"]),
                    "candidates": int(best["n_candidates"]), "seed_hits": int(best["n_seed_hits"]),
                    "device_ms": {k: round(v, 3) for k, v in best.items() if k.endswith("_ms") and v},
                    "compute_ms": round(sum(best[k] for k in ("pack_ms", "sketch_ms", "sort_ms", "index_ms", "pairgen_ms", "verify_ms", "dedup_ms")), 3),
                    "overlaps_per_s_device": best["n_records"] / sum(best[k] for k in ("pack_ms", "sketch_ms", "sort_ms", "index_ms", "pairgen_ms", "verify_ms", "dedup_ms")) * 1e3,
                    "device_bytes_held": int(free0 - free1), "device_bytes_total": int(total),
                    "sort_fallback": best["sort_fallback"], "overflow_reads": int(best["n_overflow_reads"])})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
