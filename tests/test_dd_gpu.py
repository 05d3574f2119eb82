"""GPU: the domain-decomposed multi-GPU step (include/lpe_bh.h, lpe_bh_dd_*) against the single-GPU step and the oracle.

Several ranks are played by several contexts on ONE device (DDGroup runs the three phases rank after rank with a host
synchronisation in between; peer "windows" are then plain device pointers), so the whole machinery — key ranges,
migration into the owner's arrays, published roots, exported child blocks, the top of the tree — is exercised on a
one-GPU box. Gates: every body is owned by exactly one rank; per-body accepted-interaction counts identical to the
oracle's (every theta decision is the reference's, whatever the decomposition); velocities within 1e-4 of the oracle
and within fp32 summation noise of the single-GPU step (the warps group other bodies, so the fp32 partial sums are
added in another order; the decisions are not allowed to change).
"""
import numpy as np
import pytest

import lpe_bh
import oracle_py as O
from parity import gen_uniform, rel_err

pytestmark = pytest.mark.gpu

FAST_TOL = 1e-4


def single_gpu(params, x, y, vx, vy, m, steps, comp=None, rank=None):
    bh = lpe_bh.BarnesHut(0)
    bh.set_instrumentation(counts=True)
    bh.upload(x, y, vx, vy, m, rank=rank, comp=comp)
    bh.step(params, steps)
    out = bh.download()
    out["accepted"], _ = bh.counts()
    bh.close()
    return out


def run_dd(R, params, x, y, vx, vy, m, steps, comp=None, rank=None, capacity=None, splitters=None, import_blocks=0):
    n = len(x)
    g = lpe_bh.DDGroup([0] * R, capacity or (n + 64), import_blocks)
    try:
        for c in g.ranks:
            c.set_instrumentation(counts=True)
        g.upload(params, x, y, vx, vy, m, rank=rank, comp=comp)
        if splitters is not None:
            for c in g.ranks:
                c.dd_set_splitters(splitters(c.dd_get_splitters()))
        g.step(params, steps)
        out = g.download(counts=True)
        out["stats"] = g.stats()
        return out
    finally:
        g.close()


CASES = [
    # name, n, seed, R, eps, theta, thr, quirk
    ("two_ranks", 20000, 5, 2, 0.0625, 0.5, 0.0, True),
    ("three_ranks", 20000, 6, 3, 0.0625, 0.5, 0.0, True),
    ("eight_ranks", 30000, 7, 8, 0.0625, 0.5, 0.0, True),
    ("eight_ranks_few_bodies", 300, 8, 8, 0.0625, 0.5, 0.0, True),
    ("eps_zero_self_leaves", 8000, 9, 4, 0.0, 0.5, 0.0, True),
    ("small_mass_threshold", 8000, 10, 4, 0.0625, 0.5, 1.2e6, True),
    ("textbook_tree", 8000, 11, 4, 0.0625, 0.5, 0.0, False),
    ("theta_03", 8000, 12, 4, 0.0625, 0.3, 0.0, True),
    ("aggregated_terminals", 8000, 13, 4, 16.0, 0.5, 0.0, True),
    ("one_rank", 5000, 14, 1, 0.0625, 0.5, 0.0, True),
]


@pytest.mark.parametrize("name,n,seed,R,eps,theta,thr,quirk", CASES, ids=[c[0] for c in CASES])
def test_one_step_matches_oracle_and_single_gpu(port, name, n, seed, R, eps, theta, thr, quirk):
    U = 1024.0
    x, y, vx, vy, m = gen_uniform(n, U, seed)
    kw = dict(theta=theta, thr=thr, dt_kick=1 / 120, dt_drift=0.006)
    pg = lpe_bh.make_params(U, eps, quirk=quirk, **kw)
    ref = port.run(O.make_params(U, eps, quirk=quirk, **kw), x, y, vx, vy, m, threads=8, per_body=True)
    one = single_gpu(pg, x, y, vx, vy, m, 1)
    got = run_dd(R, pg, x, y, vx, vy, m, 1)
    assert sum(s["n_live"] for s in got["stats"]) == n
    assert np.array_equal(got["accepted"], ref["accepted"]), f"{name}: decisions differ from the oracle"
    assert np.array_equal(got["accepted"], one["accepted"])
    dv = rel_err((got["vx"] - vx, got["vy"] - vy), (ref["vx"] - vx, ref["vy"] - vy))
    assert dv["max"] <= FAST_TOL and dv["norm"] <= FAST_TOL, (name, dv)
    d1 = rel_err((got["vx"] - vx, got["vy"] - vy), (one["vx"] - vx, one["vy"] - vy))
    assert d1["max"] <= 2e-5, (name, d1)
    assert np.max(np.hypot(got["x"] - ref["x"], got["y"] - ref["y"])) / U <= 1e-9


@pytest.mark.parametrize("R", [2, 5])
def test_strict_precision_is_bit_identical_to_one_gpu(port, R):
    """STRICT precision sums every body's interactions in fp64 in tree order, which does not depend on how the bodies
    are grouped into warps or spread over ranks: a decomposed run must reproduce the single-GPU run bit for bit —
    every shared cell's aggregate, every exported record, every decision."""
    U, n = 1024.0, 15000
    x, y, vx, vy, m = gen_uniform(n, U, 15)
    pg = lpe_bh.make_params(U, 0.0625, dt_drift=0.006, precision=lpe_bh.PREC_STRICT)
    one = single_gpu(pg, x, y, vx, vy, m, 2)
    got = run_dd(R, pg, x, y, vx, vy, m, 2)
    assert np.array_equal(got["accepted"], one["accepted"])
    for k in ("x", "y", "vx", "vy"):
        assert np.array_equal(got[k], one[k]), k
    ref = port.run(O.make_params(U, 0.0625, dt_drift=0.006), x, y, vx, vy, m, nsteps=2, threads=8)
    assert rel_err((got["vx"] - vx, got["vy"] - vy), (ref["vx"] - vx, ref["vy"] - vy))["max"] <= 1e-8


def test_bodies_migrate_between_ranks(port):
    """Fast bodies and a long drift: every step hundreds of bodies leave their rank's key range (some leave the
    universe and come back). Ownership stays exact and the trajectory stays on the oracle's."""
    U, n, R, steps = 1024.0, 12000, 4, 6
    x, y, vx, vy, m = gen_uniform(n, U, 21)
    rng = np.random.default_rng(3)
    vx = rng.normal(0, 300.0, n); vy = rng.normal(0, 300.0, n)
    kw = dict(theta=0.5, dt_kick=1 / 120, dt_drift=0.05)
    pg = lpe_bh.make_params(U, 0.25, **kw)
    ref = port.run(O.make_params(U, 0.25, **kw), x, y, vx, vy, m, nsteps=steps, threads=8)
    g = lpe_bh.DDGroup([0] * R, n + 64)
    g.upload(pg, x, y, vx, vy, m)
    owner0 = np.empty(n, np.int32)
    for r, c in enumerate(g.ranks):
        owner0[c.dd_download()["index"]] = r
    g.step(pg, steps)
    owner1 = np.empty(n, np.int32)
    for r, c in enumerate(g.ranks):
        owner1[c.dd_download()["index"]] = r
    got = g.download()
    g.close()
    assert np.count_nonzero(owner0 != owner1) > 100          # the test does exercise migration
    assert np.count_nonzero((ref["x"] < 0) | (ref["x"] >= U) | (ref["y"] < 0) | (ref["y"] >= U)) > 10   # and bodies outside the tree
    dv = rel_err((got["vx"] - vx, got["vy"] - vy), (ref["vx"] - vx, ref["vy"] - vy))
    assert dv["norm"] <= FAST_TOL and dv["max"] <= 10 * FAST_TOL, dv
    assert np.max(np.hypot(got["x"] - ref["x"], got["y"] - ref["y"])) / U <= 1e-9


def test_new_splitters_move_the_bodies(port):
    """Re-balancing = new splitters: the bodies that now belong elsewhere move in the next step, results unchanged."""
    U, n, R = 1024.0, 16000, 4
    x, y, vx, vy, m = gen_uniform(n, U, 22)
    pg = lpe_bh.make_params(U, 0.25, dt_drift=0.004)
    ref = port.run(O.make_params(U, 0.25, dt_drift=0.004), x, y, vx, vy, m, nsteps=2, threads=8)
    g = lpe_bh.DDGroup([0] * R, n + 64)
    g.upload(pg, x, y, vx, vy, m)
    g.step(pg, 1)
    before = [s["n_live"] for s in g.stats()]
    top = 1 << 60
    new = [0, top // 16, top // 2, top // 2 + top // 64, top]     # very uneven on purpose
    for c in g.ranks:
        c.dd_set_splitters(new)
    g.step(pg, 1)
    after = [s["n_live"] for s in g.stats()]
    got = g.download()
    g.close()
    assert sum(before) == n and sum(after) == n and before != after
    dv = rel_err((got["vx"] - vx, got["vy"] - vy), (ref["vx"] - vx, ref["vy"] - vy))
    assert dv["norm"] <= FAST_TOL and dv["max"] <= 10 * FAST_TOL, dv


def test_component_mix_and_bodies_outside_the_universe(port):
    """Sources that are not targets, massless movers, boundary entities and bodies outside [0,U)^2 (targets that feel
    the tree from anywhere: the last rank's domain must cover them)."""
    U, n, R = 1024.0, 9000, 3
    x, y, vx, vy, m = gen_uniform(n, U, 23)
    rng = np.random.default_rng(5)
    comp = np.full(n, O.HAS_MASS | O.HAS_VELOCITY, np.uint8)
    comp[rng.choice(n, 500, replace=False)] = O.HAS_MASS                      # sources only
    comp[rng.choice(n, 300, replace=False)] = O.HAS_VELOCITY                  # massless movers
    comp[rng.choice(n, 200, replace=False)] |= O.BOUNDARY
    out = rng.choice(n, 400, replace=False)
    x[out[:200]] = U + rng.uniform(1, 3000, 200); y[out[200:]] = -rng.uniform(1, 3000, 200)
    kw = dict(theta=0.5, dt_kick=1 / 120, dt_drift=0.006)
    pg = lpe_bh.make_params(U, 0.25, **kw)
    ref = port.run(O.make_params(U, 0.25, **kw), x, y, vx, vy, m, comp=comp, threads=8, per_body=True)
    got = run_dd(R, pg, x, y, vx, vy, m, 1, comp=comp)
    assert np.array_equal(got["accepted"], ref["accepted"])
    dv = rel_err((got["vx"] - vx, got["vy"] - vy), (ref["vx"] - vx, ref["vy"] - vy))
    assert dv["max"] <= FAST_TOL, dv
    assert np.max(np.hypot(got["x"] - ref["x"], got["y"] - ref["y"])) / U <= 1e-9


def test_clustered_input_locally_essential_tree_is_small(port):
    """Two-galaxy input (deep, clustered tree): each rank imports only a small part of the other ranks' cells."""
    U, n, R = float(2 ** 20), 200000, 4
    x, y, vx, vy, m = lpe_bh.workload("two_galaxies", n, 44, U)
    pg = lpe_bh.make_params(U, U / 2 ** 14)
    ref = port.run(O.make_params(U, U / 2 ** 14), x, y, vx, vy, m, threads=8, per_body=True)
    got = run_dd(R, pg, x, y, vx, vy, m, 1)
    assert np.array_equal(got["accepted"], ref["accepted"])
    dv = rel_err((got["vx"] - vx, got["vy"] - vy), (ref["vx"] - vx, ref["vy"] - vy))
    assert dv["max"] <= FAST_TOL, dv
    cells = sum(s["n_cells"] for s in got["stats"])
    exported = sum(sum(s["exported_blocks"]) for s in got["stats"])
    assert exported < 0.5 * cells * (R - 1) / R, (exported, cells)


def test_errors_are_reported():
    U = 1024.0
    x, y, vx, vy, m = gen_uniform(4000, U, 31)
    pg = lpe_bh.make_params(U, 0.25)
    c = lpe_bh.BarnesHut(0)
    with pytest.raises(RuntimeError, match="domain-decomposed"):
        c.dd_step(pg, 1)
    c.dd_init(0, 2, 4096)
    with pytest.raises(RuntimeError, match="window"):
        c.dd_upload(pg, x, y, vx, vy, m); c.dd_step(pg, 1)
    with pytest.raises(RuntimeError, match="domain-decomposed"):
        c.step(pg, 1)
    c.close()
    # capacity too small for the rank's share
    g = lpe_bh.DDGroup([0, 0], 1500)
    with pytest.raises(RuntimeError, match="capacity"):
        g.upload(pg, x, y, vx, vy, m)
    g.close()
    # a plain upload takes the context out of the mode again
    c = lpe_bh.BarnesHut(0)
    c.dd_init(0, 2, 4096)
    c.upload(x, y, vx, vy, m); c.step(pg, 1)
    assert np.all(np.isfinite(c.download()["vx"]))
    c.close()
