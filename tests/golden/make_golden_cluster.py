#!/usr/bin/env python
"""Cluster-level golden fixtures: reads in, reference-script `.overlaps` out.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_cluster.py

For each small seeded cluster: write the reads (FASTA, names 0..N-1 as rename_fas.py gives them), take the
oracle's restatement of minimap2 (oracle/hlo_oracle.c: candidates -> extension) to produce the PAF that the
absent minimap2 binary would have printed (queries in file order), and push that PAF through the UNMODIFIED
reference scripts exactly as polyte.tune_params.py:507-517 does (via make_golden.run_chain).  The committed
`clusterN.overlaps.txt` is therefore the reference converters' own output; the GPU tests compare the CUDA
path's file against it.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from hylight_b200 import synth  # noqa: E402
from oracle import orc  # noqa: E402
import make_golden  # noqa: E402

CASES = {
    # tag: (generator kwargs, min_len, min_id)
    "cluster1": (dict(n_fragments=150, read_len=150, n_strains=2, strain_cov=8.0, snp_rate=0.01, err_rate=0.005, seed=11), 60, 0.98),
    "cluster2": (dict(n_fragments=120, read_len=250, n_strains=3, strain_cov=6.0, snp_rate=0.01, err_rate=0.01, seed=12), 60, 0.95),
    "cluster3": (dict(n_fragments=100, read_len=100, n_strains=2, strain_cov=10.0, snp_rate=0.02, err_rate=0.003, seed=13,
                      n_rate=0.002), 50, 0.98),
}


def paf_text(paf):
    out = []
    for r in paf:
        if not r["valid"]:
            continue
        out.append("%d\t%d\t%d\t%d\t%s\t%d\t%d\t%d\t%d\t%d\t%d\t60\n" % (
            r["qid"], r["ql"], r["qs"], r["qe"], "-" if r["strand"] else "+", r["tid"], r["tl"], r["ts"], r["te"],
            r["matches"], r["blocklen"]))
    return "".join(out)


def main():
    for tag, (kw, min_len, min_id) in CASES.items():
        b = synth.cluster_batch(1, seed0=kw.pop("seed"), **kw)
        synth.write_fasta(b, 0, os.path.join(HERE, tag + ".fasta"))
        ob = orc.Batch.from_arrays(b.bases, b.base_off, b.cluster_off)
        p = orc.default_params(min_ovlp_len=min_len, min_identity=min_id)
        cands, _ = orc.candidates(ob, p)
        _, _, _, paf = orc.verify(ob, cands, p, want_paf=True)
        # rows with blocklen 0 would make filter_ovlp_inline.py divide by zero; minimap2 never prints them
        paf = paf[(paf["valid"] != 0) & (paf["blocklen"] > 0)]
        make_golden.run_chain(paf_text(paf), HERE, tag, min_len, min_id, 1, 0.8, b.n_reads)
        n = sum(1 for _ in open(os.path.join(HERE, tag + ".overlaps.txt")))
        print(tag, "reads", b.n_reads, "paf rows", len(paf), "overlaps", n)
        # the unfiltered PAF is large and reproducible from the fasta; keep only reads + outputs
        os.remove(os.path.join(HERE, tag + ".in.paf"))


if __name__ == "__main__":
    main()
