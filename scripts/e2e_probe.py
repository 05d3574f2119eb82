"""Host-tick timing only: pinned host arrays -> lpe_bh_update_host -> pinned host arrays (what bench.py reports as e2e),
next to the bare PCIe legs of the same arrays (torch copies of the same pinned buffers) for comparison."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "little-physics-engine_b200"))
import torch
import lpe_bh, bench
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
n = wl["n"]
x, y, vx, vy, m = lpe_bh.workload(wl["kind"], n, wl["seed"], bench.U)
bh = lpe_bh.BarnesHut(0)
p = lpe_bh.make_params(bench.U, bench.EPS, theta=bench.THETA, dt_kick=bench.DT, dt_drift=bench.DT)
host = [torch.from_numpy(a.copy()).pin_memory() for a in (x, y, vx, vy, m)]
ptrs = [t.data_ptr() for t in host]
dev = [torch.empty_like(t, device="cuda") for t in host]
for leg, pairs in (("h2d 40 B/body", [(d, h) for d, h in zip(dev, host)]), ("d2h 32 B/body", [(h, d) for d, h in zip(dev[:4], host[:4])])):
    ts = []
    for it in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for dst, src in pairs:
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("%s: %.3f ms (%.1f GB/s)" % (leg, min(ts), sum(s.numel() * 8 for _, s in pairs) / min(ts) / 1e6))
for rep in range(3):
    for it in range(3):
        bh.update_host_ptrs(p, n, *ptrs)
    ts = []
    for it in range(20):
        t0 = time.perf_counter()
        bh.update_host_ptrs(p, n, *ptrs)
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    print("e2e ms/tick mean %.3f  min %.3f  median %.3f  max %.3f" % (sum(ts) / len(ts), ts[0], ts[len(ts) // 2], ts[-1]))
bh.set_instrumentation(timing=True)
bh.update_host_ptrs(p, n, *ptrs)
print({k: round(v, 4) for k, v in bh.stats().items() if k.startswith("ms_")})
