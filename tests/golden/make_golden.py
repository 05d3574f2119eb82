"""Generates tests/golden/*.npz (and tests/golden/boundary/*.npz) from the REFERENCE ITSELF
(oracle/_ref/libref_bh.so = the reference's barnes_hut.cpp + movement.cpp + boundary.cpp compiled unmodified from
/root/reference, see oracle/Makefile).

Run here (the container with /root/reference):   python tests/golden/make_golden.py
The reference ships no golden vectors of its own, so these are the pinned known answers:
inputs, the reference's outputs after `steps` ticks of {BarnesHutSystem, MovementSystem}, its insertion
order and its full tree dump. They travel to the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "little-physics-engine_b200"))
import oracle_py as O  # noqa: E402


def cases():
    rng = np.random.default_rng(20261018)
    out = {}
    # 1. the 4-body probe of SURVEY.md Q2: masses {1,2,4,8}e6 -> root mass 2.3e7 (true 1.5e7)
    out["four_body"] = dict(
        x=np.array([100.0, 600.0, 520.0, 530.0]), y=np.array([100.0, 300.0, 250.0, 260.0]),
        vx=np.zeros(4), vy=np.zeros(4), m=np.array([1e6, 2e6, 4e6, 8e6]), comp=None,
        U=1024.0, eps=1e-3, theta=0.5, thr=0.0, dt_kick=1 / 120, dt_drift=1 / 120, steps=1)
    # 2. small uniform squares, threshold off / on (unequal masses expose the double count)
    for name, n, thr in (("uniform_64", 64, 0.0), ("uniform_1000", 1000, 0.0), ("uniform_1000_thr", 1000, 1.2e6)):
        U = 1024.0
        out[name] = dict(x=rng.random(n) * U, y=rng.random(n) * U, vx=rng.standard_normal(n),
                         vy=rng.standard_normal(n), m=1e6 * (0.5 + rng.random(n)), comp=None, U=U, eps=U / 2 ** 14,
                         theta=0.5, thr=thr, dt_kick=1 / 120, dt_drift=0.006, steps=2)
    # 3. Keplerian-disk law (reference keplerian_disk.cpp) at the scenario's own config: U=6e9, eps=2e7, thr=1e3
    import lpe_bh
    try:
        x, y, vx, vy, m = lpe_bh.workload("keplerian", 2000, 5, 6e9)
    except Exception as e:  # library not built: skip this case rather than invent numbers
        print("skipping keplerian case:", e)
    else:
        out["keplerian_2000"] = dict(x=x, y=y, vx=vx, vy=vy, m=m, comp=None, U=6e9, eps=2e7, theta=0.5, thr=1e3,
                                     dt_kick=1 / 120, dt_drift=6.756e-3, steps=2)
    # 4. component mix: bodies out of bounds, boundaries, liquids, massless movers, velocity-less sources
    n = 600
    U = 1024.0
    comp = np.full(n, O.HAS_MASS | O.HAS_VELOCITY, np.uint8)
    comp[::11] |= O.BOUNDARY
    comp[5::13] |= O.LIQUID
    comp[3::17] = O.HAS_VELOCITY            # no mass: moves, neither source nor target
    comp[7::19] = O.HAS_MASS                # no velocity: source only
    x = rng.random(n) * U
    y = rng.random(n) * U
    x[2::23] -= 700.0                       # outside [0,U): feels force, exerts none
    y[4::29] += 900.0
    out["component_mix"] = dict(x=x, y=y, vx=rng.standard_normal(n), vy=rng.standard_normal(n),
                                m=1e6 * (0.5 + rng.random(n)), comp=comp, U=U, eps=U / 2 ** 12, theta=0.5, thr=0.0,
                                dt_kick=1 / 120, dt_drift=0.01, steps=2)
    # 5. theta and depth extremes on one input
    n = 400
    base = dict(x=rng.random(n) * U, y=rng.random(n) * U, vx=np.zeros(n), vy=np.zeros(n),
                m=1e6 * (0.5 + rng.random(n)), comp=None, U=U, thr=0.0, dt_kick=1 / 120, dt_drift=1 / 120, steps=1)
    out["theta_0p3"] = dict(base, eps=U / 2 ** 14, theta=0.3)
    out["theta_1p0"] = dict(base, eps=U / 2 ** 14, theta=1.0)
    out["eps_zero"] = dict(base, eps=0.0, theta=0.5)
    out["eps_large"] = dict(base, eps=U / 2 ** 5, theta=0.5)
    # 6. all masses below the threshold: the reference returns before doing anything (barnes_hut.cpp:55-71)
    out["all_small"] = dict(base, eps=U / 2 ** 14, theta=0.5, thr=1e9)
    return out


def boundary_cases():
    """Inputs for BoundarySystem: bodies inside, outside each edge and exactly on the clamp
    edges, fast and slow, with velocity-less and asleep entities mixed in."""
    rng = np.random.default_rng(20261019)
    out = {}
    for name, n, U, margin, damping, vmax, vscale in (("default_cfg", 4000, 1000.0, 15.0, 0.7, 1.0, 3.0),
                                                     ("no_damping_big_cap", 1500, 6e9, 2.5e7, 1.0, 1e4, 5e3),
                                                     ("zero_margin", 1500, 1024.0, 0.0, 0.5, 0.25, 1.0)):
        x = rng.uniform(-0.1 * U, 1.1 * U, n)
        y = rng.uniform(-0.1 * U, 1.1 * U, n)
        x[:8] = [margin, U - margin, np.nextafter(margin, -1), np.nextafter(U - margin, 2 * U), 0.0, U, -U, 2 * U]
        y[4:12] = [margin, U - margin, np.nextafter(margin, -1), np.nextafter(U - margin, 2 * U), 0.0, U, -U, 2 * U]
        vx = rng.standard_normal(n) * vscale
        vy = rng.standard_normal(n) * vscale
        vx[::37] = 0.0
        vy[::41] = -0.0
        comp = np.full(n, O.HAS_MASS | O.HAS_VELOCITY, np.uint8)
        comp[3::7] = O.HAS_MASS                      # no Velocity: not in the view
        comp[5::9] |= O.ASLEEP                       # asleep: skipped
        out[name] = dict(x=x, y=y, vx=vx, vy=vy, comp=comp, U=U, margin=margin, damping=damping, vmax=vmax)
    return out


def main():
    ref = O.RefLib()
    print(ref.describe())
    os.makedirs(os.path.join(HERE, "boundary"), exist_ok=True)
    for name, c in boundary_cases().items():
        r = ref.boundary(c["U"], c["x"], c["y"], c["vx"], c["vy"], comp=c["comp"], margin=c["margin"],
                         damping=c["damping"], max_speed=c["vmax"])
        path = os.path.join(HERE, "boundary", name + ".npz")
        np.savez_compressed(path, x=c["x"], y=c["y"], vx=c["vx"], vy=c["vy"], comp=c["comp"],
                            cfg=np.array([c["U"], c["margin"], c["damping"], c["vmax"]]),
                            out_x=r["x"], out_y=r["y"], out_vx=r["vx"], out_vy=r["vy"])
        moved = int(np.sum((r["x"] != c["x"]) | (r["y"] != c["y"])))
        print(f"boundary/{name}: n={len(c['x'])} clamped={moved} -> {os.path.getsize(path)} B")
    for name, c in cases().items():
        p = O.make_params(c["U"], c["eps"], theta=c["theta"], thr=c["thr"], dt_kick=c["dt_kick"],
                          dt_drift=c["dt_drift"])
        r = ref.run(p, c["x"], c["y"], c["vx"], c["vy"], c["m"], comp=c["comp"], nsteps=c["steps"])
        tree, st = ref.tree(p, c["x"], c["y"], c["m"], comp=c["comp"])
        rank = ref.view_rank(len(c["x"]), c["comp"])
        comp = c["comp"] if c["comp"] is not None else np.full(len(c["x"]), O.HAS_MASS | O.HAS_VELOCITY, np.uint8)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(
            path, x=c["x"], y=c["y"], vx=c["vx"], vy=c["vy"], m=c["m"], comp=comp, rank=rank,
            cfg=np.array([c["U"], c["eps"], c["theta"], c["thr"], c["dt_kick"], c["dt_drift"], c["steps"]]),
            out_x=r["x"], out_y=r["y"], out_vx=r["vx"], out_vy=r["vy"], tree=tree,
            tree_stats=np.array([st["pool_nodes"], st["nonempty_nodes"], st["internal_nodes"], st["max_depth"]]))
        print(f"{name}: n={len(c['x'])} pool={st['pool_nodes']} nonempty={st['nonempty_nodes']} "
              f"depth={st['max_depth']} -> {os.path.getsize(path)} B")


if __name__ == "__main__":
    main()
