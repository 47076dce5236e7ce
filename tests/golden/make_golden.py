#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference scripts.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It executes, exactly as polyte.tune_params.py:507-517 does,

    cat in.paf | cut -f 1-12 | python filter_ovlp_inline.py <len> <id> <o> <r> > filtered.paf
    python minimap22sfo.py --in filtered.paf --out sfo.out -m 0 -p 0
    python sfo2overlaps.py --in sfo.out --out overlaps.txt --num_singles N --num_pairs 0

under LC_ALL=C (GNU sort's last-resort comparison is locale dependent, sfo2overlaps.py:53) and commits
input and outputs as small text fixtures.  tests/test_oracle_golden.py pins oracle/hlo_oracle.c to them.
"""
import os
import random
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/script"

APPENDIX_A = """\
0\t150\t50\t150\t+\t1\t150\t0\t100\t99\t100\t60
0\t150\t0\t90\t-\t2\t150\t60\t150\t90\t90\t60
10\t150\t0\t120\t+\t9\t150\t30\t150\t118\t120\t60
3\t150\t20\t150\t-\t4\t150\t20\t150\t130\t130\t60
5\t150\t0\t150\t+\t6\t150\t0\t150\t147\t150\t60
5\t150\t0\t100\t+\t6\t150\t50\t150\t100\t100\t60
7\t150\t40\t150\t+\t8\t150\t0\t109\t107\t111\t60
3\t150\t0\t110\t-\t4\t150\t0\t110\t110\t110\t60
12\t150\t0\t100\t+\t7\t140\t40\t140\t100\t100\t60
20\t100\t0\t100\t+\t21\t150\t25\t125\t99\t100\t60
22\t150\t30\t130\t-\t23\t100\t0\t100\t100\t100\t60
"""


def run_chain(paf_text, outdir, tag, min_len, min_id, o, r, n_singles):
    env = dict(os.environ, LC_ALL="C")
    tmp = tempfile.mkdtemp()
    try:
        with open(os.path.join(tmp, "in.paf"), "w") as f:
            f.write(paf_text)
        cmd = "cat in.paf | cut -f 1-12 | python %s/filter_ovlp_inline.py %s %s %s %s > overlaps.paf" % (REF, min_len, min_id, o, r)
        subprocess.check_call(cmd, shell=True, cwd=tmp, env=env)
        subprocess.check_call("python %s/minimap22sfo.py --in overlaps.paf --out sfoverlaps.out -m 0 -p 0" % REF,
                              shell=True, cwd=tmp, env=env)
        subprocess.check_call("python %s/sfo2overlaps.py --in sfoverlaps.out --out original_overlaps.txt --num_singles %d --num_pairs 0"
                              % (REF, n_singles), shell=True, cwd=tmp, env=env)
        for src, dst in (("in.paf", "in.paf"), ("overlaps.paf", "filtered.paf"), ("sfoverlaps.out", "sfo.out"),
                         ("original_overlaps.txt", "overlaps.txt")):
            shutil.copy(os.path.join(tmp, src), os.path.join(outdir, "%s.%s" % (tag, dst)))
        with open(os.path.join(outdir, "%s.args" % tag), "w") as f:
            f.write("%s %s %s %s %d\n" % (min_len, min_id, o, r, n_singles))
    finally:
        shutil.rmtree(tmp)


def random_paf(n_rows, n_ids, seed):
    """Geometrically consistent ungapped PAF rows covering every strand/order/containment case, clipped
    target ends, identities around the threshold, duplicate pairs and ids whose string order differs
    from their numeric order."""
    rnd = random.Random(seed)
    rows = []
    lens = [100, 120, 136, 150, 150, 150, 250]
    while len(rows) < n_rows:
        a, b = rnd.randrange(n_ids), rnd.randrange(n_ids)
        if a == b and rnd.random() < 0.9:
            continue
        if str(a) > str(b) and rnd.random() < 0.85:  # minimap2 -X: query is the smaller string, mostly
            a, b = b, a
        ql = lens[(a * 7) % len(lens)]
        tl = lens[(b * 7) % len(lens)]
        strand = rnd.choice("+-")
        D = rnd.randrange(-ql + 20, tl - 20)
        js, je = max(0, -D), min(ql, tl - D)
        if je - js < 17:
            continue
        cl = cr = 0
        if D < 0 and rnd.random() < 0.3:
            cl = rnd.choice([1, 1, 1, 2, 3, 30])
        if tl - D < ql and rnd.random() < 0.3:
            cr = rnd.choice([1, 1, 1, 2, 3, 30])
        if rnd.random() < 0.05:  # internal match on a query-limited end as well
            cl += rnd.choice([0, 5, 40])
        bs, be = js + cl, je - cr
        if be - bs < 5:
            continue
        blen = be - bs
        k = rnd.choice([0, 0, 0, 1, 1, 2, 2, 3, 4, 6, 10])
        k = min(k, blen)
        if strand == "+":
            qs, qe = bs, be
        else:
            qs, qe = ql - be, ql - bs
        rows.append((a, ql, qs, qe, strand, b, tl, bs + D, be + D, blen - k, blen, 60))
        if rnd.random() < 0.15:  # a second hit of the same pair nearby (same chunk, usually)
            rows.append(rows[-1][:9] + (max(blen - k - 3, 0), max(blen - 2, 1), 60))
    rnd.shuffle(rows)
    return "".join("\t".join(str(x) for x in r) + "\n" for r in rows[:n_rows])


def main():
    if not os.path.isdir(REF):
        sys.exit("reference scripts not found at %s (run this in the build container)" % REF)
    run_chain(APPENDIX_A, HERE, "appendixA", 60, 0.98, 1, 0.8, 24)
    run_chain(random_paf(2600, 140, seed=7), HERE, "random1", 60, 0.98, 1, 0.8, 140)
    run_chain(random_paf(1200, 1200, seed=11), HERE, "random2", 50, 0.95, 2, 0.8, 1200)
    run_chain(random_paf(600, 40, seed=13), HERE, "random3", 17, 0.9, 3, 0.02, 40)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
