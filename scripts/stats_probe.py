"""Walk statistics of the two-phase traversal (counting run): list entries and node kinds per warp, both walk modes."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "little-physics-engine_b200"))
import lpe_bh, bench
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
x, y, vx, vy, m = lpe_bh.workload(wl["kind"], wl["n"], wl["seed"], bench.wl_params(wl)["U"])
p = bench.make_gpu_params(lpe_bh, wl)
for warp_only in (True, False):
    bh = lpe_bh.BarnesHut(0)
    bh.set_instrumentation(counts=True, warp_only=warp_only)
    bh.upload(x, y, vx, vy, m)
    bh.step(p, 1)
    st = bh.stats()
    warps = (wl["n"] + 31) // 32
    k = st["t2_kinds"]
    print("warp_only" if warp_only else "cta      ", "interactions/body %.1f" % (st["interactions"] / wl["n"]),
          "list entries/warp %.1f" % (st["warp_visits"] / warps),
          "| per warp: A-clean %.1f A-dirty %.1f O-dirty %.1f M->acc %.1f M->open %.1f M->split %.1f rounds %.1f nodes %.1f" %
          tuple(v / warps for v in k))
    bh.close()
