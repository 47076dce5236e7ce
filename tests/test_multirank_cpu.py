"""The N>1 host logic on CPU (gloo, world_size 2): the job's batches are pulled dynamically from an atomic counter
on the rendezvous store (bench.py: WorkCounter) — no data-path collective; the only communication is the
max-over-ranks of the time and the sums of counts and digests."""
import os
import stat
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import threading
    import time
    import bench
    store = dist.distributed_c10d._get_default_store()
    counter = bench.WorkCounter(store)
    batches = bench.job_batches(1000, 16)
    mine = []

    def pull():  # two host workers per rank, like the e2e leg
        for p in range(2):
            while True:
                j = counter.next("t_%d" % p)
                if j >= len(batches):
                    break
                mine.append((p, j))
                time.sleep(0.0005 * (rank + 1))  # ranks of different speed: the faster one takes more batches

    th = [threading.Thread(target=pull) for _ in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_pull_the_job_dynamically():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import bench
    n_b = len(bench.job_batches(1000, 16))
    a, b = gathered
    assert len(a) + len(b) == 2 * n_b and set(a).isdisjoint(b)          # every batch of both passes, exactly once
    assert set(a) | set(b) == {(p, j) for p in range(2) for j in range(n_b)}
    assert len(a) > len(b) > 0                                          # the faster rank took more: dynamic, not static


def test_job_batches_cover_the_work_list():
    import bench
    bt = bench.job_batches(1000, 16)
    assert len(bt) == 63 and sum(n for _, n in bt) == 1000 and bt[-1] == (992, 8)
    assert [s for s, _ in bt] == list(range(0, 1000, 16))
    c = bench.WorkCounter(None)
    assert [c.next("k") for _ in range(3)] == [0, 1, 2] and c.next("other") == 0


def test_minimap2_probe_and_chain_with_a_stand_in_binary(tmp_path, monkeypatch):
    """bench.py looks for a real minimap2 (PATH, baseline/_ref/, mappy) and, when it finds one, runs the exact chain of
    polyte.tune_params.py:507-517 on the concordance cluster.  No minimap2 exists in the image, so the plumbing is
    exercised with a stand-in executable that prints a PAF (the oracle's own): the chain must then reproduce the oracle's
    .overlaps tuples through the chunked filter and the converters."""
    import bench
    from hylight_b200 import synth
    from oracle import orc
    b = synth.cluster_batch(1, seed0=5, n_fragments=120, read_len=150, n_strains=2, strain_cov=6.0)
    ob = orc.Batch.from_arrays(b.bases, b.base_off, b.cluster_off)
    cands, _ = orc.candidates(ob)
    rec, blk, acc, paf = orc.verify(ob, cands, want_paf=True)
    orc.write_paf(paf, str(tmp_path / "all.paf"))
    fake = tmp_path / "minimap2"
    fake.write_text("#!/bin/sh\nif [ \"$1\" = \"--version\" ]; then echo 2.99-standin; exit 0; fi\ncat %s\n" % (tmp_path / "all.paf"))
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    pr = bench.probe_minimap2()
    assert pr["found"] == "binary" and pr["version"] == "2.99-standin"
    fa = tmp_path / "s.fasta"
    synth.write_fasta(b, 0, str(fa))
    tup, n_paf = bench.minimap2_chain_tuples(pr["path"], str(fa), orc.default_params(), str(tmp_path))
    assert n_paf == int(paf["valid"].sum())
    # the reference chain's own dedup is chunk-local: compare with the oracle run the same way
    keep = orc.filter_chunked(paf[paf["valid"] == 1], orc.default_params())
    want = orc.paf_rows_to_recs(paf[paf["valid"] == 1], keep)
    orc.write_overlaps(want, str(tmp_path / "w.txt"))
    wt = {(a[0], a[1], a[2], a[5], a[6], a[9]) for a in (l.rstrip("\n").split("\t") for l in open(tmp_path / "w.txt"))}
    assert tup == wt and len(tup) > 500
