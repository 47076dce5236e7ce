#!/usr/bin/env python
"""Golden fixtures for HyLight's contig stage b (HyLight.py:309-318), from the UNMODIFIED reference scripts:

    cat in.paf | python filter_trans_ovlp_inline_v3.py -len 90 -iden 0.99 -oh 2 -sfo > sfoverlaps.out
    python sfo2overlaps.py --in sfoverlaps.out --out sfoverlap.out.savage --num_singles N --num_pairs 0

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_v3.py
tests/test_oracle_golden.py pins oracle/hlo_oracle.c:orc_filter_chunked_v3 / orc_write_sfo_v3 to the outputs.
"""
import os
import random
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/script"


def run_chain(paf_text, tag, min_len, min_id, o, n_singles):
    env = dict(os.environ, LC_ALL="C")
    tmp = tempfile.mkdtemp()
    try:
        with open(os.path.join(tmp, "in.paf"), "w") as f:
            f.write(paf_text)
        subprocess.check_call("cat in.paf | python %s/filter_trans_ovlp_inline_v3.py -len %s -iden %s -oh %s -sfo > sfoverlaps.out"
                              % (REF, min_len, min_id, o), shell=True, cwd=tmp, env=env)
        subprocess.check_call("python %s/sfo2overlaps.py --in sfoverlaps.out --out sfoverlap.out.savage --num_singles %d --num_pairs 0"
                              % (REF, n_singles), shell=True, cwd=tmp, env=env)
        for src, dst in (("in.paf", "in.paf"), ("sfoverlaps.out", "sfo.out"), ("sfoverlap.out.savage", "overlaps.txt")):
            shutil.copy(os.path.join(tmp, src), os.path.join(HERE, "%s.%s" % (tag, dst)))
        with open(os.path.join(HERE, "%s.args" % tag), "w") as f:
            f.write("%s %s %s 0.8 %d\n" % (min_len, min_id, o, n_singles))
    finally:
        shutil.rmtree(tmp)


def contig_paf(n_rows, n_ids, seed):
    """Ungapped PAF rows between contigs of 200 .. 6000 bp: every strand / order / containment case, clipped target
    ends (overhang 0..4 and internal matches), identities around 0.99, and pairs with a second, shorter or longer, row
    right behind the first (the script keeps the FIRST row of a pair that passes length and identity)."""
    rnd = random.Random(seed)
    lens = {i: rnd.choice([200, 350, 800, 1500, 3000, 6000]) + rnd.randrange(50) for i in range(n_ids)}
    rows = []
    while len(rows) < n_rows:
        a, b = rnd.randrange(n_ids), rnd.randrange(n_ids)
        if a == b and rnd.random() < 0.9:
            continue
        if str(a) > str(b) and rnd.random() < 0.85:
            a, b = b, a
        ql, tl = lens[a], lens[b]
        strand = rnd.choice("+-")
        D = rnd.randrange(-ql + 60, tl - 60)
        js, je = max(0, -D), min(ql, tl - D)
        if je - js < 40:
            continue
        cl = cr = 0
        if rnd.random() < 0.35:
            cl = rnd.choice([1, 1, 2, 2, 3, 4, 50])
        if rnd.random() < 0.35:
            cr = rnd.choice([1, 1, 2, 2, 3, 4, 50])
        bs, be = js + cl, je - cr
        if be - bs < 20:
            continue
        blen = be - bs
        k = min(int(blen * rnd.choice([0, 0, 0.002, 0.005, 0.009, 0.0101, 0.012, 0.03])), blen)
        qs, qe = (bs, be) if strand == "+" else (ql - be, ql - bs)
        rows.append((a, ql, qs, qe, strand, b, tl, bs + D, be + D, blen - k, blen, 60))
        if rnd.random() < 0.2:  # a second row of the same pair: different diagonal / length
            D2 = D + rnd.choice([-7, 5, 12])
            js2, je2 = max(0, -D2), min(ql, tl - D2)
            if je2 - js2 >= 40:
                bl2 = je2 - js2
                q2 = (js2, je2) if strand == "+" else (ql - je2, ql - js2)
                rows.append((a, ql, q2[0], q2[1], strand, b, tl, js2 + D2, je2 + D2, bl2 - rnd.choice([0, 1, 2]), bl2, 60))
    return "".join("\t".join(str(x) for x in r) + "\n" for r in rows[:n_rows])


def main():
    if not os.path.isdir(REF):
        sys.exit("reference scripts not found at %s (run this in the build container)" % REF)
    run_chain(contig_paf(2400, 90, seed=21), "stageb1", 90, 0.99, 2, 90)
    run_chain(contig_paf(700, 700, seed=22), "stageb2", 90, 0.99, 2, 700)
    run_chain(contig_paf(900, 25, seed=23), "stageb3", 50, 0.95, 0, 25)
    print("stage-b golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
