"""Host-tick timing only: pinned host arrays -> lpe_bh_update_host -> pinned host arrays (what bench.py reports as e2e)."""
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
