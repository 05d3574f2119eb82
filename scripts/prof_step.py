"""Minimal driver for ncu: C2 (or another workload) resident on the device, a few steps, no torch."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "little-physics-engine_b200"))
import lpe_bh
import bench
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
x, y, vx, vy, m = lpe_bh.workload(wl["kind"], wl["n"], wl["seed"], bench.wl_params(wl)["U"])
bh = lpe_bh.BarnesHut(0)
bh.set_instrumentation(timing=True, warp_only=os.environ.get("LPE_WARP_ONLY") == "1")
bh.upload(x, y, vx, vy, m)
p = bench.make_gpu_params(lpe_bh, wl)
for s in range(steps):
    bh.step(p, 1)
    st = bh.stats()
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()})
