#!/usr/bin/env python
"""bench.py — verified read overlaps / second of the short-read overlap stage (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the CPU chain on the box's host cores

One step = one pass of the whole hot path (2-bit pack -> minimizer sketch -> radix sort -> seed buckets ->
pair/diagonal generation -> verify -> per-pair dedup -> sorted records) over one batch of C synthetic
config-2 clusters (15 000 fragments = 30 000 reads x 150 bp, 5 strains, 0.5 % errors; BASELINE.json
configs[1], batched as configs[2] shards them) per GPU.  Clusters are independent, so N GPUs = N ranks each
working on its own clusters (weak scaling, no data-path collective; torch.distributed is used only for the
barrier and the max-over-ranks of the timing).

`value`  : inputs (ASCII reads) already resident in HBM, records left in HBM; CUDA events on the library's stream.
`e2e`    : hlo_overlap_host() from pinned HOST buffers, H2D of the reads and D2H of the records inside the timing.
`roofline`: the verify kernel (north-star kernel): algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json.
`cpu_baseline`: the CPU restatement of the reference chain (oracle/, kind "port": minimap2 is not shipped with the
reference and is absent from the image) on all host cores, one single-threaded chain per cluster like the
reference's `xargs -P` fan-out (HyLight.py:245).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "verified_read_overlaps_per_sec"
UNIT = "overlaps/s"
CLUSTER_KW = dict(n_fragments=15000, read_len=150, n_strains=5, strain_cov=10.0, snp_rate=0.01, err_rate=0.005)


def rank_cluster_seeds(rank, c_per_gpu):
    """independent clusters per rank (weak scaling): rank r works on seeds 1000 + r*C .. 1000 + (r+1)*C - 1"""
    return [1000 + rank * c_per_gpu + c for c in range(c_per_gpu)]


def whole_job(dist, torch, device, ms_local, n_local):
    """value of the whole job: units of all ranks / max-over-ranks time (no data-path collective: clusters are
    independent; these two reductions are the only communication)"""
    if dist is None:
        return ms_local, float(n_local)
    t = torch.tensor([ms_local], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = torch.tensor([float(n_local)], device=device, dtype=torch.float64)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), float(c.item())


def workload_config(c_per_gpu):
    return {"workload": "BASELINE configs[1] cluster (15000 fragments = 30000 reads x 150 bp, 5 strains, 0.5%% error), "
                        "%d such clusters per step per GPU (a shard of configs[2])" % c_per_gpu,
            "clusters_per_step_per_gpu": c_per_gpu, "reads_per_cluster": 30000, "read_len": 150, "strains": 5,
            "error_rate": 0.005, "min_ovlp_len": 60, "min_identity": 0.98, "k": 21, "w": 11}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU chain
def cpu_chain(clusters, cores):
    """records/s of the oracle's whole chain over `clusters` (list of ReadBatch), `cores` chains in flight."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import orc
    orc.lib()

    def one(b):
        ob = orc.Batch.from_arrays(b.bases, b.base_off, b.cluster_off)
        rec, blk, coff, st = orc.overlap_batch(ob)  # ctypes releases the GIL: threads run in parallel
        return len(rec)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        n = sum(ex.map(one, clusters))
    dt = time.perf_counter() - t0
    return n, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from hylight_b200 import synth
    cores = os.cpu_count() or 1
    n_cl = min(cores, 64)
    clusters = [synth.cluster_batch(1, seed0=5000 + c, **CLUSTER_KW) for c in range(n_cl)]
    for _ in range(args.warmup):
        cpu_chain(clusters[:max(1, min(n_cl, 2))], cores)  # warm-up: page in the library, small
    tot_n, tot_t = 0, 0.0
    for _ in range(args.steps):
        n, dt = cpu_chain(clusters, min(cores, n_cl))
        tot_n += n
        tot_t += dt
    v = tot_n / tot_t
    sample = "%d config-2 clusters per step, one single-threaded chain per cluster, %d in flight" % (n_cl, min(cores, n_cl))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args.clusters),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": min(cores, n_cl), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock + throttle reasons sampled through NVML every few ms DURING the timed region (the same counters
    `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints)."""

    def __init__(self, index, period=0.005):
        self.index, self.period, self.rows, self.t, self.stop_flag = index, period, [], None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index(index))
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, str(e)

    @staticmethod
    def _nvml_index(local):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip() != ""]
            if local < len(ids) and ids[local].strip().isdigit():
                return int(ids[local])
        return local

    def start(self):
        if not self.nv:
            return
        self.stop_flag = False
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, rs))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + self.err]}
        self.stop_flag = True
        self.t.join(timeout=2)
        nv = self.nv
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            mx = None
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": mx, "reasons": ["no samples"]}
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                "hw_power_brake_slowdown": 0x80}
        reasons = sorted(k for k, b in bits.items() if any(r & b for _, r in self.rows))
        return {"sm_mhz": float(np.median([c for c, _ in self.rows])), "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ ours
def verify_algorithmic_bytes(cands, lens, min_len, n_acc):
    """16 B candidate + 2*ceil(OL/4) B of 2-bit bases (both reads, inside the overlap, only when
    the geometric overlap reaches min_ovlp_len) + 16 B compact record per accepted candidate."""
    ql = lens[cands["q"]].astype(np.int64)
    tl = lens[cands["t"]].astype(np.int64)
    d = cands["diag"].astype(np.int64)
    ol = np.minimum(ql, tl - d) - np.maximum(-d, 0)
    ol = np.where(ol >= min_len, ol, 0)
    return int(16 * len(cands) + (2 * ((ol + 3) // 4)).sum() + 16 * n_acc), float(ol[ol > 0].mean()) if (ol > 0).any() else 0.0


def concordance_check(api, synth):
    """Correctness reported with the timing: one config-2 cluster through the CUDA path vs the CPU
    chain (oracle/): (1) records bit-exact; (2) Jaccard of the (ID1,ID2,POS1,ORI1,ORI2,LEN1) tuples of the two
    13-column files, floor 0.999."""
    import tempfile
    from oracle import orc
    b = synth.cluster_batch(1, seed0=4242, **CLUSTER_KW)
    with api.Context(0) as c:
        R = c.pack(b)
        cands = c.candidates(R)
        rec_v, aux_v = c.verify(R, cands)          # hlo_verify on the candidate list
        rec, aux, coff = c.overlap_batch(R)        # full path
        R.free()
    ob = orc.Batch.from_arrays(b.bases, b.base_off, b.cluster_off)
    oc, _ = orc.candidates(ob)
    key = lambda a: np.lexsort((a["diag"], a["strand_anchor"] & 1, a["t"], a["q"]))
    cands_equal = len(oc) == len(cands) and bool(np.array_equal(oc[key(oc)], cands[key(cands)]))
    rv_o, _, _ = orc.verify(ob, oc)
    srt = lambda r: r[np.lexsort((r["K"], r["OHB"], r["OHA"], r["ori"], r["idB"], r["idA"]))]
    verify_equal = len(rv_o) == len(rec_v) and bool(np.array_equal(srt(rv_o), srt(rec_v)))
    rec_o, _, _, _ = orc.overlap_batch(ob)
    with tempfile.TemporaryDirectory() as d:
        api.write_overlaps(rec, os.path.join(d, "g.txt"))
        orc.write_overlaps(rec_o, os.path.join(d, "c.txt"))
        tup = lambda p: {(a[0], a[1], a[2], a[5], a[6], a[9]) for a in (l.rstrip("\n").split("\t") for l in open(p))}
        g, cset = tup(os.path.join(d, "g.txt")), tup(os.path.join(d, "c.txt"))
        byte_identical = open(os.path.join(d, "g.txt"), "rb").read() == open(os.path.join(d, "c.txt"), "rb").read()
    jac = len(g & cset) / float(len(g | cset)) if (g or cset) else 1.0
    return {"cluster": "config-2, seed 4242", "candidates": int(len(cands)), "candidates_bit_exact": cands_equal,
            "verified_records": int(len(rec_v)), "verify_records_bit_exact": verify_equal, "records": int(len(rec)),
            "overlaps_file_byte_identical": byte_identical, "jaccard": jac, "jaccard_floor": 0.999,
            "against": "CPU restatement of the minimap2->filter->sfo->overlaps chain (oracle/; minimap2 binary unavailable)"}


def bind_to_gpu_cpus(local):
    """Run this rank (and the host workers it starts) on the CPUs NVML reports as closest to its GPU, so that the
    pinned buffers are first-touched on that NUMA node and the PCIe traffic does not cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(ClockSampler._nvml_index(local))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1 and 64 * w + b < n_cpu}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:  # noqa: BLE001
        pass
    return 0


def run_ours(args):
    import torch
    from hylight_b200 import api, synth
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or api.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the overlap stage has no CPU fallback")
    torch.cuda.set_device(local)
    bound = 0 if args.no_bind else bind_to_gpu_cpus(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0's stdout carries exactly one JSON line: NCCL prints its banner ("NCCL version ...") on stdout when the
        # communicator comes up, so file descriptor 1 points at stderr until the first collective is through
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    C = args.clusters
    seeds = rank_cluster_seeds(rank, C)
    batch = synth.cluster_batch(C, seed0=seeds[0], **CLUSTER_KW)
    lens = np.diff(batch.base_off).astype(np.int64)
    ctx = api.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, warm, steps, sampler=None):
        for _ in range(warm):
            fn()
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = [fn() for _ in range(steps)]
        e1.record(stream)
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out, clocks

    # ---- device-resident leg: ASCII reads in HBM -> records in HBM (packer included via "repack")
    R = ctx.pack(batch)
    ctx.set_option("repack", 1)
    ctx.set_option("d2h", 0)

    def step_dev():
        r = api.Recs()
        ctx._check(api.lib().hlo_overlap_batch(ctx.h, R.handle, r))
        return int(r.n), ctx.timing()

    ms_dev, outs, clocks = timed(step_dev, args.warmup, args.steps, ClockSampler(local))
    n_rec = sum(o[0] for o in outs)
    tms = [o[1] for o in outs]
    launches = sum(t["total_launches"] for t in tms)
    stage_names = ["pack_ms", "sketch_ms", "sort_ms", "index_ms", "pairgen_ms", "verify_ms", "verify_fast_ms", "dedup_ms"]
    stages = {k: float(np.mean([t[k] for t in tms])) for k in stage_names}
    counts = {k: int(tms[-1][k]) for k in ("n_reads", "n_bases", "n_minimizers", "n_seed_hits", "n_candidates", "n_accepted",
                                           "n_records", "n_deferred", "n_overflow_reads")}
    # ---- roofline of the verify kernel (candidate list fetched outside the timed region)
    ctx.set_option("repack", 0)
    ctx.set_option("d2h", 1)
    cands = ctx.candidates(R)
    vbytes, mean_ol = verify_algorithmic_bytes(cands, lens, ctx.params.min_ovlp_len, counts["n_accepted"])
    del cands
    R.free()
    peak, peak_src = peaks()
    ach = vbytes / (stages["verify_fast_ms"] * 1e-3) / 1e9 if stages["verify_fast_ms"] > 0 else 0.0
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture
    tj = os.path.join(ROOT, "profiles", "verify_traffic.json")
    if os.path.exists(tj):
        t = json.load(open(tj))
        traffic = int((t["dram_bytes_read"] + t["dram_bytes_write"]) * counts["n_candidates"] / float(t["candidates_per_launch"]))
    roofline = {"kernel": "verify_fast_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "algorithmic_bytes_per_launch": vbytes,
                "ms_per_launch": stages["verify_fast_ms"], "mean_overlap_len": mean_ol, "peak_source": peak_src,
                "note": "traffic < algorithmic bytes: the rows of a cluster (1.4 MB per strand) are re-read from L2; the kernel is "
                        "bound by the integer (ALU) pipe, 72% utilised (profiles/r1k_final.md)"}

    # ---- end-to-end leg: pinned host buffers -> hlo_overlap_host -> records on the host
    pb = api.PinnedArray(batch.bases.nbytes)
    po = api.PinnedArray(batch.base_off.nbytes)
    pc = api.PinnedArray(batch.cluster_off.nbytes)
    pb.array[:] = batch.bases
    po.array[:] = batch.base_off.view(np.uint8)
    pc.array[:] = batch.cluster_off.view(np.uint8)
    # A few host workers, each with its own context (stream + workspaces), take the steps alternately: the H2D copy
    # of one step, the kernels of another and the D2H copy of a third overlap, which is how the host binary drives a
    # GPU (main.cpp, --workers_per_device).  Every step still copies its own inputs in and its own records out.
    n_workers = args.e2e_workers
    ctxs = [ctx] + [api.Context(local) for _ in range(n_workers - 1)]
    for c_ in ctxs:
        c_.set_option("aux", 0)      # what hylight-ovlp asks for: records + per-cluster offsets,
        c_.set_option("compact", 1)  # in the 16-byte form (hlo_rec16)
    d2h = [0]

    def run_steps(n_steps):
        nxt = [0]
        tot = [0] * n_workers
        lock = threading.Lock()

        def work(wi):
            c = ctxs[wi]
            while True:
                with lock:
                    if nxt[0] >= n_steps:
                        return
                    nxt[0] += 1
                r = c.overlap_host_raw(pb.ptr, po.ptr, batch.n_reads, pc.ptr, batch.n_clusters)
                tot[wi] += int(r.n)
                d2h[0] = int(r.n) * api.REC16_DT.itemsize + 8 * (batch.n_clusters + 1)

        th = [threading.Thread(target=work, args=(wi,)) for wi in range(n_workers)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        return sum(tot)

    run_steps(max(2 * n_workers, min(args.warmup, 3)))  # warm-up: sizes every worker's buffers
    barrier()
    t0 = time.perf_counter()
    n_rec2 = run_steps(args.steps)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3  # wall clock: the e2e leg includes the host side of every call
    if dist is not None:
        t = torch.tensor([ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    for c in ctxs[1:]:
        c.close()
    h2d = batch.bases.nbytes + batch.base_off.nbytes + batch.cluster_off.nbytes
    _, n_rec_all = whole_job(dist, torch, "cuda", ms_dev, n_rec)
    _, n_rec2_all = whole_job(dist, torch, "cuda", ms_e2e, n_rec2)
    ctx.close()
    for p in (pb, po, pc):
        p.free()

    line = None
    if rank == 0:
        cfg = workload_config(C)
        cfg["l2"] = "inputs larger than L2: %.0f MB ASCII + %.0f MB packed rows per step per GPU (L2 = 126 MB)" % (
            batch.bases.nbytes / 1e6, batch.n_reads * 256 / 1e6)
        cfg["counts_per_step_rank0"] = counts
        line = {"metric": METRIC, "value": n_rec_all / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": {"value": n_rec2_all / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h[0]), "ms_per_step": ms_e2e / args.steps, "host_workers": n_workers, "cpus_bound": bound,
                        "timing": "wall clock between device synchronisations, max over ranks"},
                "gpu_launches": int(launches), "roofline": roofline, "stages_ms": stages}
        if world == 1 and not args.no_cpu:
            line["concordance"] = concordance_check(api, synth)
            cores = os.cpu_count() or 1
            n_cl = max(1, min(cores, 32))
            cl = [synth.cluster_batch(1, seed0=5000 + c, **CLUSTER_KW) for c in range(n_cl)]
            n, dt = cpu_chain(cl, n_cl)
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": n_cl, "kind": "port",
                                    "sample": "%d config-2 clusters, one single-threaded oracle chain per cluster, all in "
                                              "flight at once (%.1f s wall)" % (n_cl, dt)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clusters", type=int, default=32, help="clusters per step per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-bind", action="store_true", help="do not bind the rank to the CPUs closest to its GPU")
    ap.add_argument("--e2e-workers", type=int, default=3, help="pipelined host workers (context + stream each) in the e2e leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
