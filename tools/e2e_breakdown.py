import sys, time, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from hylight_b200 import api, synth
b = synth.cluster_batch(32, seed0=1000, n_fragments=15000, read_len=150, n_strains=5, strain_cov=10.0, snp_rate=0.01, err_rate=0.005)
ctx = api.Context(0)
ctx.set_option("aux", 0)
pb = api.PinnedArray(b.bases.nbytes); po = api.PinnedArray(b.base_off.nbytes); pc = api.PinnedArray(b.cluster_off.nbytes)
pb.array[:] = b.bases; po.array[:] = b.base_off.view(np.uint8); pc.array[:] = b.cluster_off.view(np.uint8)
for i in range(4):
    t0 = time.perf_counter()
    r = ctx.overlap_host_raw(pb.ptr, po.ptr, b.n_reads, pc.ptr, b.n_clusters)
    t1 = time.perf_counter()
    tm = ctx.timing()
    print("call %.2f ms" % ((t1 - t0) * 1e3), {k: round(v, 2) for k, v in tm.items() if k.endswith("_ms")})
# raw pinned copy bandwidth
x = torch.empty(400 * 1024 * 1024, dtype=torch.uint8, pin_memory=True)
d = torch.empty_like(x, device="cuda")
for name, src, dst in (("h2d", x, d), ("d2h", d, x)):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(name, "%.1f GB/s" % (x.numel() / dt / 1e9))
# two pipelined workers
import threading
ctx2 = api.Context(0); ctx2.set_option("aux", 0)
ctxs = [ctx, ctx2]
for c in ctxs:
    c.overlap_host_raw(pb.ptr, po.ptr, b.n_reads, pc.ptr, b.n_clusters)
log = []
def work(wi):
    for i in range(5):
        t0 = time.perf_counter()
        ctxs[wi].overlap_host_raw(pb.ptr, po.ptr, b.n_reads, pc.ptr, b.n_clusters)
        t1 = time.perf_counter()
        tm = ctxs[wi].timing()
        log.append((wi, i, round((t0 - T0) * 1e3, 1), round((t1 - t0) * 1e3, 1), {k[:-3]: round(v, 1) for k, v in tm.items() if k.endswith("_ms")}))
torch.cuda.synchronize(); T0 = time.perf_counter()
th = [threading.Thread(target=work, args=(w,)) for w in range(2)]
[t.start() for t in th]; [t.join() for t in th]
print("2 workers, 10 steps: %.1f ms/step" % ((time.perf_counter() - T0) * 1e3 / 10))
for l in sorted(log, key=lambda x: x[2]): print(l)
