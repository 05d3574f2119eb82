"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the path on ~100 k bodies.
Single-GPU resident steps (FAST and STRICT), the host tick, forced frontier overflow, and a 3-rank decomposed run played
on one device. usage: compute-sanitizer --tool memcheck python scripts/sanitize_step.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "little-physics-engine_b200"))
import numpy as np
import lpe_bh
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
U = 2.0 ** 20
x, y, vx, vy, m = lpe_bh.workload("plummer", n, 7, U)
bh = lpe_bh.BarnesHut(0)
for prec in (lpe_bh.PREC_FAST, lpe_bh.PREC_STRICT):
    p = lpe_bh.make_params(U, U / 2 ** 14, precision=prec)
    bh.upload(x, y, vx, vy, m)
    bh.step(p, 2)
    bh.download()
p = lpe_bh.make_params(U, U / 2 ** 14)
bh.set_instrumentation(counts=True, force_overflow=True)
bh.upload(x, y, vx, vy, m); bh.step(p, 1); bh.counts()
bh.set_instrumentation()
hx, hy, hvx, hvy = x.copy(), y.copy(), vx.copy(), vy.copy()
bh.update_host(p, hx, hy, hvx, hvy, m)
p0 = lpe_bh.make_params(U, 0.0)      # eps = 0: depth 30, self-leaf bookkeeping
bh.upload(x[:20000], y[:20000], vx[:20000], vy[:20000], m[:20000]); bh.step(p0, 1); bh.download()
bh.boundary(U)
bh.close()
g = lpe_bh.DDGroup([0, 0, 0], n // 2)
rng = np.random.default_rng(1)
g.upload(lpe_bh.make_params(U, U / 2 ** 14, dt_drift=0.05), x, y, rng.normal(0, 2e3, n), rng.normal(0, 2e3, n), m)
g.step(lpe_bh.make_params(U, U / 2 ** 14, dt_drift=0.05), 3)     # with migration
g.rebalance()
g.step(lpe_bh.make_params(U, U / 2 ** 14, dt_drift=0.05), 1)
out = g.download()
assert np.all(np.isfinite(out["vx"]))
g.close()
print("sanitize_step: done")
