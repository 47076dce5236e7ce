"""The N>1 host logic on CPU (gloo, world_size 2): clusters are sharded across ranks with no data-path
collective; the only communication is the max-over-ranks of the time and the sum of the units (bench.py)."""
import os
import sys

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
    import bench
    from hylight_b200 import synth
    seeds = bench.rank_cluster_seeds(rank, 3)
    # every rank generates ITS clusters only; the generator is seeded, so the shards are disjoint and reproducible
    b = synth.cluster_batch(3, seed0=seeds[0], n_fragments=30, read_len=100, n_strains=2, strain_cov=4.0)
    gathered = [None] * world
    dist.all_gather_object(gathered, (seeds, int(b.n_reads), b.bases[:64].tobytes()))
    ms, n = bench.whole_job(dist, torch, "cpu", 10.0 * (rank + 1), 100 * (rank + 1))
    if rank == 0:
        out.put((gathered, ms, n))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_clusters_and_reduce():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, ms, n = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (s0, n0, h0), (s1, n1, h1) = gathered
    assert set(s0).isdisjoint(s1) and len(s0) == len(s1) == 3
    assert n0 == n1 == 180 and h0 != h1      # same shape (weak scaling), different reads
    assert ms == 20.0 and n == 300.0         # max over ranks, sum over ranks


def test_single_rank_needs_no_process_group():
    import bench
    assert bench.whole_job(None, torch, "cpu", 5.0, 7) == (5.0, 7.0)
    assert bench.rank_cluster_seeds(0, 2) == [1000, 1001] and bench.rank_cluster_seeds(3, 2) == [1006, 1007]
