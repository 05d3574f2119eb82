"""Driver for ncu: R ranks of a domain-decomposed run played on ONE device (phases rank after rank), a few steps.
usage: python scripts/prof_dd.py [workload] [ranks] [steps] [bodies]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "little-physics-engine_b200"))
import bench, lpe_bh
key = sys.argv[1] if len(sys.argv) > 1 else "c3"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
wl = dict(bench.WORKLOADS[key])
if len(sys.argv) > 4:
    wl["n"] = int(sys.argv[4])
n = wl["n"]
b = lpe_bh.workload(wl["kind"], n, wl["seed"], bench.U)
p = bench.make_gpu_params(lpe_bh, wl)
g = lpe_bh.DDGroup([0] * R, int(n / R * 1.5) + 65536)
for c in g.ranks:
    c.set_instrumentation(timing=True)
g.upload(p, *b)
g.step(p, steps)
for s in g.stats():
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.items() if k.startswith("ms_") or k in ("n_live", "n_cells", "n_roots", "export_rounds", "exported_blocks")})
g.close()
