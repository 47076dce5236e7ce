#!/usr/bin/env python
"""One-off stress of the lean sketch automaton's equal-hash detection: many low-complexity reads, several (w, k),
GPU (fast and generic kernels) against the CPU restatement.  python tools/sketch_stress.py [reads] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hylight_b200 import api  # noqa: E402
from oracle import orc  # noqa: E402


def reads(rng, n):
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []
    for _ in range(n):
        L = int(rng.integers(40, 300))
        kind = int(rng.integers(0, 6))
        if kind <= 2:
            period = int(rng.integers(1, 60))
            unit = acgt[rng.integers(0, 4, size=period)]
            s = np.tile(unit, L // period + 2)[:L].copy()
            for _m in range(int(rng.integers(0, 5))):
                s[int(rng.integers(0, L))] = acgt[int(rng.integers(0, 4))]
            if kind == 1:
                cut = int(rng.integers(0, L))
                s[cut:] = acgt[rng.integers(0, 4, size=L - cut)]
            if kind == 2 and rng.random() < 0.5:
                s[int(rng.integers(0, L))] = ord("N")
        elif kind == 3:
            two = acgt[rng.choice(4, size=2, replace=False)]
            s = two[rng.integers(0, 2, size=L)]
        elif kind == 4:
            seg = acgt[rng.integers(0, 4, size=int(rng.integers(15, 70)))]
            gap = acgt[rng.integers(0, 4, size=int(rng.integers(0, 14)))]
            s = np.concatenate([seg, gap, seg, gap, seg, acgt[rng.integers(0, 4, size=L)]])[:L]
        else:  # inverted repeat: a segment followed by its reverse complement
            seg = acgt[rng.integers(0, 4, size=int(rng.integers(15, 70)))]
            rc = np.frombuffer(b"TGCA", dtype=np.uint8)[np.searchsorted(acgt, seg)][::-1]
            s = np.concatenate([seg, acgt[rng.integers(0, 4, size=int(rng.integers(0, 9)))], rc, acgt[rng.integers(0, 4, size=L)]])[:L]
        out.append(s.tobytes())
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    seqs = reads(rng, n)
    b = orc.Batch([seqs])
    for w, k in [(11, 21), (5, 15), (19, 11), (11, 15), (5, 27)]:
        pkw = dict(w=w, k=k, min_chain_score=min(30, 2 * k))
        mz_o, ro_o = orc.sketch_all(b, orc.default_params(**pkw))
        with api.Context(0, **pkw) as c:
            R = c.pack(b)
            mz, ro = c.sketch(R)
            R.free()
        ok = np.array_equal(ro, ro_o) and np.array_equal(mz["x"], mz_o["x"]) and np.array_equal(mz["y"], mz_o["y"])
        print("w=%d k=%d reads=%d minimizers=%d %s" % (w, k, n, len(mz_o), "OK" if ok else "MISMATCH"), flush=True)
        if not ok:
            sys.exit(1)


if __name__ == "__main__":
    main()
