#!/usr/bin/env python
"""Wall-clock of the host binary on N synthetic config-2 clusters (FASTA in, .overlaps out)."""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hylight_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
extra = sys.argv[2:]
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
lines = []
for c in range(n):
    b = synth.cluster_batch(1, seed0=7000 + c, **synth.CONFIGS["config2"])
    synth.write_fasta(b, 0, os.path.join(d, "c%d.fa" % c))
    lines.append("%s\t%s" % (os.path.join(d, "c%d.fa" % c), os.path.join(d, "c%d.overlaps" % c)))
open(os.path.join(d, "list.tsv"), "w").write("\n".join(lines) + "\n")
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hylight_b200", "hylight-ovlp")
for rep in range(2):
    t0 = time.time()
    r = subprocess.run([exe, "--clusters", os.path.join(d, "list.tsv"), "--stats", os.path.join(d, "stats.jsonl"), "--batch_bases", str(40 << 20)] + extra,
                       capture_output=True, text=True)
    dt = time.time() - t0
    nl = sum(1 for c in range(n) for _ in open(os.path.join(d, "c%d.overlaps" % c)))
    print("rc", r.returncode, "clusters", n, "wall %.2f s" % dt, "lines", nl, "overlaps/s %.3g" % (nl / dt), r.stderr[-200:])
import json
st = [json.loads(l) for l in open(os.path.join(d, "stats.jsonl"))]
print("summary", [x for x in st if x.get("summary")])
st = [x for x in st if not x.get("summary")]
for x in st[:3] + st[-3:]:
    print(x["batch"], x["device"], "wall %.0f ms" % x["ms"]["wall"], {k: round(v, 1) for k, v in x["ms"].items() if k != "wall"})
print("sum over batches (ms): " + ", ".join("%s %.0f" % (k, sum(x["ms"].get(k, 0) for x in st)) for k in ("wall", "host_parse", "host_call", "host_write", "ctx_create", "device_total")))
import shutil
shutil.rmtree(d)
