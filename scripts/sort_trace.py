"""Where a sort tile's lifetime goes (variant build -DSORT_TRACE): summed clock64 deltas of thread 0 per phase."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "little-physics-engine_b200"))
import lpe_bh, bench
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
x, y, vx, vy, m = lpe_bh.workload(wl["kind"], wl["n"], wl["seed"], bench.wl_params(wl)["U"])
bh = lpe_bh.BarnesHut(0)
bh.set_instrumentation(timing=True)
bh.upload(x, y, vx, vy, m)
p = bench.make_gpu_params(lpe_bh, wl)
lib = lpe_bh.load_library()
out = (C.c_ulonglong * 8)()
bh.step(p, 3)
lib.lpe_bh_debug_sort_trace(out, 1)
bh.step(p, 5)
lib.lpe_bh_debug_sort_trace(out, 1)
tiles = out[7]
names = ["load issue", "arrive + rank", "publish + scan", "scatter smem", "look-back", "store"]
tot = sum(out[k] for k in range(6))
print("tiles", tiles, "cycles/tile", tot / tiles, "us/tile", tot / tiles / 1965.0)
for k, nm in enumerate(names):
    print(f"  {nm:16s} {out[k] / tiles:9.0f} cycles  {100.0 * out[k] / tot:5.1f} %")
print(bh.stats()["ms_sort"])
