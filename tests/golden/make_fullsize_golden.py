"""Golden vectors at BASELINE.json's FULL sizes: the oracle builds the whole tree once and evaluates
every k-th body; the GPU box then needs no minutes-long CPU step.

    python tests/golden/make_fullsize_golden.py        (this container; about ten minutes, ~10 GB of RAM)

For each workload the file tests/golden/full/<name>.npz holds the sampled creation indices, the oracle's accepted-
interaction count and velocity change (one step, kick only matters: v0 = generator's v) of those bodies, and the
parameters. Inputs are regenerated on the GPU box from the same deterministic generators (csrc/workloads.cpp).
The oracle here is the plain-C restatement (oracle/bh_oracle.c, 8 threads), itself pinned bit for bit to the compiled
reference (tests/test_oracle_pinning.py).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "little-physics-engine_b200"))
import oracle_py as O  # noqa: E402
import workloads  # noqa: E402

U = float(2 ** 20)
EPS = U / 2 ** 14
CASES = {
    # name: kind, n, seed, every k-th body is a target
    "c3_plummer_16m": ("plummer", 16_000_000, 43, 256),
    "c4_two_galaxies_4m": ("two_galaxies", 4_000_000, 44, 64),
}


def main():
    out_dir = os.path.join(ROOT, "tests", "golden", "full")
    os.makedirs(out_dir, exist_ok=True)
    port = O.PortLib()
    for name, (kind, n, seed, k) in CASES.items():
        t0 = time.time()
        x, y, vx, vy, m = workloads.workload(kind, n, seed, U)
        comp = np.full(n, O.HAS_MASS, np.uint8)
        idx = np.arange(0, n, k, dtype=np.uint32)
        comp[idx] |= O.HAS_VELOCITY        # targets = bodies with Velocity (barnes_hut.cpp:89); every body is a source
        p = O.make_params(U, EPS, theta=0.5, dt_kick=1 / 120, dt_drift=1 / 120)
        ref = port.run(p, x, y, vx, vy, m, comp=comp, threads=8, per_body=True)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), index=idx, accepted=ref["accepted"][idx],
                            dvx=(ref["vx"] - vx)[idx], dvy=(ref["vy"] - vy)[idx], kind=kind, n=n, seed=seed, every=k,
                            U=U, eps=EPS, theta=0.5, dt=1 / 120)
        print(f"{name}: {len(idx)} targets of {n} bodies, {ref['accepted'][idx].mean():.1f} accepted per target, "
              f"{time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
