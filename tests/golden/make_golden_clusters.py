#!/usr/bin/env python
"""Golden vectors for hylight-cluster (host half): the UNMODIFIED reference scripts
bin_pointer_limited_filechunks_shortpath2.py -> getclusters.py -> get_fq_cluster.py (HyLight.py:215-220), run in this
container on a seeded synthetic overlap file, digested into tests/golden/clusters.json.

    python tests/golden/make_golden_clusters.py            # needs /root/reference; rewrites tests/golden/clusters.json

The inputs are not committed (8.6 MB of PAF: the reference cuts its input into 2.6 MB chunks, and the order-dependent
batching only shows with several of them); `inputs()` below regenerates them bit for bit — the test checks the SHA-256.
"""
import hashlib
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/script"
CASES = [dict(size="200", threads=2), dict(size="1000", threads=3), dict(size="inf", threads=5), dict(size="150", threads=1)]


def inputs(d, n_frag=3000, n_lines=160000, seed=5):
    """reads.fq (interleaved /1 /2), readnames.txt (get_readnames.py's rule), shortr2.paf (12 columns, sorted by column 12
    descending like utils.py:58 leaves it): communities of 150 fragments with 3 % links across"""
    rnd = random.Random(seed)
    names = ["frag%d" % i for i in range(n_frag)]
    with open(os.path.join(d, "reads.fq"), "w") as f:
        for nm in names:
            for m in (1, 2):
                s = "".join(rnd.choice("ACGT") for _ in range(30))
                f.write("@%s/%d\n%s\n+\n%s\n" % (nm, m, s, "I" * 30))
    with open(os.path.join(d, "readnames.txt"), "w") as f:
        f.write("".join(nm + "\n" for nm in names))
    lines = []
    for _ in range(n_lines):
        a = rnd.randrange(n_frag)
        b = (a // 150) * 150 + rnd.randrange(150) if rnd.random() < 0.97 else rnd.randrange(n_frag)
        if a == b:
            continue
        ma, mb = rnd.choice("12"), rnd.choice("12")
        sc = rnd.randrange(60, 150)
        lines.append(("%s/%s\t150\t0\t%d\t+\t%s/%s\t150\t%d\t150\t%d\t%d\t%d" % (names[a], ma, sc, names[b], mb, 150 - sc, sc, sc, sc), sc))
    lines.sort(key=lambda x: -x[1])
    with open(os.path.join(d, "shortr2.paf"), "w") as f:
        f.write("".join(l + "\n" for l, _ in lines))
    return hashlib.sha256(open(os.path.join(d, "shortr2.paf"), "rb").read()).hexdigest()


def digest_grouped(grouped):
    return {k: [len(v), hashlib.sha1("\n".join(sorted(v)).encode()).hexdigest()] for k, v in grouped.items()}


def digest_tree(root):
    out = {}
    for dp, _, fs in os.walk(root):
        for fn in fs:
            p = os.path.join(dp, fn)
            out[os.path.relpath(p, root)] = hashlib.sha1(open(p, "rb").read()).hexdigest()
    return out


if __name__ == "__main__":
    d = tempfile.mkdtemp()
    out = {"paf_sha256": inputs(d), "cases": []}
    for i, c in enumerate(CASES):
        rid = "g%d" % i
        subprocess.check_call([sys.executable, REF + "/bin_pointer_limited_filechunks_shortpath2.py", "shortr2.paf", "readnames.txt", c["size"], rid,
                               str(c["threads"])], cwd=d)
        pre = "%s_max%s_final" % (rid, c["size"])
        subprocess.check_call([sys.executable, REF + "/getclusters.py", pre, str(c["threads"])], cwd=d, stdout=subprocess.DEVNULL)
        subprocess.check_call([sys.executable, REF + "/get_fq_cluster.py", pre + "_clusters_grouped.json", "reads.fq", "fq_" + rid], cwd=d,
                              stdout=subprocess.DEVNULL)
        grouped = json.load(open(os.path.join(d, pre + "_clusters_grouped.json")))
        out["cases"].append(dict(c, grouped=digest_grouped(grouped), clustersizes=json.load(open(os.path.join(d, pre + "_clustersizes.json"))),
                                 fastq=digest_tree(os.path.join(d, "fq_" + rid))))
        print(c, len(grouped), "clusters")
    json.dump(out, open(os.path.join(HERE, "clusters.json"), "w"), indent=0, sort_keys=True)
