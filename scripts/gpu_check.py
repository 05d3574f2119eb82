"""First-light GPU check (development aid; the real tests live in tests/)."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from parity import *

port = oracle_py.PortLib()
bh = lpe_bh.BarnesHut(0)
print(lpe_bh.load_library().lpe_bh_version().decode())

def case(name, x, y, vx, vy, m, U, eps, theta=0.5, thr=0.0, quirk=True, tree=True):
    n = len(x)
    po = oracle_py.make_params(U, eps, theta=theta, thr=thr, dt_kick=1/120, dt_drift=0.006, quirk=quirk)
    ref = port.run(po, x, y, vx, vy, m, nsteps=1, threads=8, per_body=True)
    out = {}
    for prec in (lpe_bh.PREC_STRICT, lpe_bh.PREC_FAST):
        pg = lpe_bh.make_params(U, eps, theta=theta, thr=thr, dt_kick=1/120, dt_drift=0.006, quirk=quirk, precision=prec)
        bh.set_instrumentation(timing=True, counts=True)
        bh.upload(x, y, vx, vy, m)
        bh.step(pg, 1)
        got = bh.download()
        st = bh.stats()
        acc, vis = bh.counts()
        dv = rel_err((got['vx']-vx, got['vy']-vy), (ref['vx']-vx, ref['vy']-vy))
        dx = float(np.max(np.hypot(got['x']-ref['x'], got['y']-ref['y'])) / U)
        cnt_ok = bool(np.array_equal(acc, ref['accepted']))
        out['strict' if prec else 'fast'] = dict(dv=dv, dx=dx, counts_equal=cnt_ok, n_mismatch=int(np.sum(acc != ref['accepted'])))
        if prec == lpe_bh.PREC_STRICT and tree and n <= 200000:
            dump = bh.dump_tree()
            check_preorder(dump)
            nodes, _ = port.tree(po, x, y, m)
            out['tree'] = compare_tree(dump, nodes, U)
    out['stats'] = {k: st[k] for k in ('n_in_tree','n_terminals','n_nodes','depth','sort_passes','interactions','ms_keygen','ms_sort','ms_build','ms_traverse','ms_total')}
    out['oracle_acc_per_body'] = ref['stats']['accepted']/max(n,1)
    print(name, json.dumps(out, default=float)); sys.stdout.flush()

U = 1024.0
x = np.array([100., 600., 520., 530.]); y = np.array([100., 300., 250., 260.]); m = np.array([1e6, 2e6, 4e6, 8e6]); z = np.zeros(4)
case("four", x, y, z, z, m, U, 1e-3)
case("four_noquirk", x, y, z, z, m, U, 1e-3, quirk=False)
case("one", x[:1], y[:1], z[:1], z[:1], m[:1], U, 1e-3)
case("two", x[:2], y[:2], z[:2], z[:2], m[:2], U, 1e-3)
for n, seed in ((10, 1), (1000, 2), (20000, 3), (200000, 4)):
    x, y, vx, vy, m = gen_uniform(n, U, seed)
    case(f"uni{n}", x, y, vx, vy, m, U, U / 2**14)
    case(f"uni{n}_thr", x, y, vx, vy, m, U, U / 2**14, thr=1.2e6)
x, y, vx, vy, m = gen_uniform(5000, U, 7)
case("eps_big", x, y, vx, vy, m, U, U / 2**6)          # shallow depth bound -> aggregated terminals
case("eps_zero", x, y, vx, vy, m, U, 0.0)              # no bound: depth 30
case("theta03", x, y, vx, vy, m, U, U / 2**14, theta=0.3)
case("theta10", x, y, vx, vy, m, U, U / 2**14, theta=1.0)
# keplerian
U = 6e9
xk = lpe_bh.workload("keplerian", 10000, 5, U)
case("kepler10k", xk[0], xk[1], xk[2], xk[3], xk[4], U, 2e7, thr=1e3)
# out of bounds + component mix
U = 1024.0
x, y, vx, vy, m = gen_uniform(3000, U, 9); x[::7] -= 600.0
case("oob", x, y, vx, vy, m, U, U / 2**14)
# 1M disk timing
U = float(2**20)
xd = lpe_bh.workload("disk", 1_000_000, 42, U)
pg = lpe_bh.make_params(U, 64.0)
bh.set_instrumentation(timing=True, counts=False)
bh.upload(*xd)
for it in range(5):
    bh.step(pg, 1); st = bh.stats()
    print("disk1M", {k: round(st[k], 4) if isinstance(st[k], float) else st[k] for k in st})
t0 = time.time(); bh.step(pg, 10); bh.synchronize(); print("10 steps wall ms/step", (time.time()-t0)*100)
xp = lpe_bh.workload("plummer", 16_000_000, 43, U)
bh.upload(*xp)
for it in range(3):
    bh.step(pg, 1); st = bh.stats()
    print("plummer16M", {k: round(st[k], 4) if isinstance(st[k], float) else st[k] for k in st})
