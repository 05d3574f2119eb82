"""SURVEY.md §8(f) N2 — BoundarySystem (reference src/systems/boundary.cpp:13-69) as a device pass.

CPU: the C restatement (oracle/bh_oracle.c: orc_boundary) against golden vectors produced by the compiled reference
(tests/golden/boundary, made by tests/golden/make_golden.py) and against the compiled reference itself when present.
GPU: lpe_bh_boundary through the C ABI against the oracle, bit for bit (fp64, elementwise).
"""
import os

import numpy as np
import pytest

from parity import lpe_bh, oracle_py, ROOT

GDIR = os.path.join(ROOT, "tests", "golden", "boundary")
NAMES = sorted(f[:-4] for f in os.listdir(GDIR) if f.endswith(".npz"))
KEYS = ("x", "y", "vx", "vy")


def load(name):
    d = np.load(os.path.join(GDIR, name + ".npz"))
    U, margin, damping, vmax = (float(v) for v in d["cfg"])
    return d, dict(margin=margin, damping=damping, max_speed=vmax), U


def same_bits(a, b):
    return np.array_equal(np.asarray(a).view(np.uint64), np.asarray(b).view(np.uint64))


@pytest.mark.parametrize("name", NAMES)
def test_port_matches_golden_bit_exact(port, name):
    d, cfg, U = load(name)
    r = port.boundary(U, d["x"], d["y"], d["vx"], d["vy"], comp=d["comp"], **cfg)
    for k in KEYS:
        assert same_bits(r[k], d["out_" + k]), k
    # the fixtures do exercise the pass
    assert np.any(d["out_x"] != d["x"]) and np.any(d["out_vy"] != d["vy"])


def test_port_matches_compiled_reference(port, reflib):
    rng = np.random.default_rng(5)
    n, U = 50000, 2048.0
    x, y = rng.uniform(-300, U + 300, n), rng.uniform(-300, U + 300, n)
    vx, vy = rng.standard_normal(n) * 4, rng.standard_normal(n) * 4
    comp = rng.choice(np.array([1, 3, 3, 3, 19, 2, 18], dtype=np.uint8), n)
    a = port.boundary(U, x, y, vx, vy, comp=comp, margin=20.0, damping=0.6, max_speed=2.0)
    b = reflib.boundary(U, x, y, vx, vy, comp=comp, margin=20.0, damping=0.6, max_speed=2.0)
    for k in KEYS:
        assert same_bits(a[k], b[k]), k


def test_semantics_spelled_out(port):
    """One body per rule of boundary.cpp: left/right/top/bottom, corner, speed cap, no-velocity, asleep, inside."""
    U, mg = 100.0, 10.0
    x = np.array([5.0, 95.0, 50.0, 50.0, 1.0, 5.0, 5.0, 5.0, 50.0])
    y = np.array([50.0, 50.0, 5.0, 99.0, 99.0, 50.0, 50.0, 50.0, 50.0])
    vx = np.array([-0.5, 0.5, 0.1, 0.1, -3.0, -8.0, -0.5, -0.5, 7.0])
    vy = np.array([0.2, 0.2, -0.4, 0.4, 4.0, 6.0, 0.2, 0.2, 7.0])
    comp = np.array([3, 3, 3, 3, 3, 3, 1, 19, 3], dtype=np.uint8)
    r = port.boundary(U, x, y, vx, vy, comp=comp, margin=mg, damping=0.5, max_speed=1.0)
    assert r["x"][0] == mg and r["vx"][0] == 0.25 and r["vy"][0] == 0.2            # left: |vx| * damping
    assert r["x"][1] == U - mg and r["vx"][1] == -0.25                             # right: -|vx| * damping
    assert r["y"][2] == mg and r["vy"][2] == 0.2 and r["vx"][2] == 0.1             # top
    assert r["y"][3] == U - mg and r["vy"][3] == -0.2                              # bottom
    assert r["x"][4] == mg and r["y"][4] == U - mg                                 # corner: both, then the cap
    assert np.isclose(np.hypot(r["vx"][4], r["vy"][4]), 1.0, rtol=1e-15)
    assert np.isclose(np.hypot(r["vx"][5], r["vy"][5]), 1.0, rtol=1e-15)           # speed cap after a bounce
    assert r["x"][6] == 5.0 and r["vx"][6] == -0.5                                 # no Velocity: not in the view
    assert r["x"][7] == 5.0 and r["vx"][7] == -0.5                                 # asleep: skipped
    assert r["vx"][8] == 7.0 and r["vy"][8] == 7.0                                 # inside: fast but never capped


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_pass_matches_golden_bit_exact(bh, name):
    d, cfg, U = load(name)
    n = len(d["x"])
    bh.upload(d["x"], d["y"], d["vx"], d["vy"], np.ones(n), comp=d["comp"])
    bh.boundary(U, cfg["margin"], cfg["damping"], cfg["max_speed"])
    got = bh.download()
    for k in KEYS:
        assert same_bits(got[k], d["out_" + k]), k


@pytest.mark.gpu
def test_device_pass_matches_oracle_large_and_ragged(bh, port):
    rng = np.random.default_rng(9)
    for n in (1, 255, 257, 100003):
        U = 2.0 ** 20
        x, y = rng.uniform(-0.2 * U, 1.2 * U, n), rng.uniform(-0.2 * U, 1.2 * U, n)
        vx, vy = rng.standard_normal(n) * 10, rng.standard_normal(n) * 10
        comp = rng.choice(np.array([1, 3, 3, 3, 19, 2, 11], dtype=np.uint8), n)
        want = port.boundary(U, x, y, vx, vy, comp=comp, margin=1000.0, damping=0.7, max_speed=1.0)
        bh.upload(x, y, vx, vy, np.ones(n), comp=comp)
        bh.boundary(U, 1000.0, 0.7, 1.0)
        got = bh.download()
        for k in KEYS:
            assert same_bits(got[k], want[k]), (n, k)


@pytest.mark.gpu
def test_device_pass_on_reordered_state_and_in_the_tick_order(port):
    """Resident tick = Boundary -> BarnesHut -> Movement (the reference's system order), three ticks, against the
    oracle running the same sequence; the state is in key order on the device after the first tick."""
    n, U = 30000, 1024.0
    rng = np.random.default_rng(21)
    x, y = rng.uniform(-40, U + 40, n), rng.uniform(-40, U + 40, n)
    vx, vy = rng.standard_normal(n) * 30, rng.standard_normal(n) * 30
    m = 1e6 * (0.5 + rng.random(n))
    pg = lpe_bh.make_params(U, U / 2 ** 12, dt_drift=0.01, precision=lpe_bh.PREC_STRICT)
    po = oracle_py.make_params(U, U / 2 ** 12, dt_drift=0.01)
    ctx = lpe_bh.BarnesHut(0)
    ctx.upload(x, y, vx, vy, m)
    ox, oy, ovx, ovy = x, y, vx, vy
    for _ in range(3):
        ctx.boundary(U, 15.0, 0.7, 1.0)
        ctx.step(pg, 1)
        b = port.boundary(U, ox, oy, ovx, ovy, margin=15.0, damping=0.7, max_speed=1.0)
        r = port.run(po, b["x"], b["y"], b["vx"], b["vy"], m, nsteps=1)
        ox, oy, ovx, ovy = r["x"], r["y"], r["vx"], r["vy"]
    ctx.boundary(U, 15.0, 0.7, 1.0)   # one more clamp so that the final state has bodies exactly on the edges
    b = port.boundary(U, ox, oy, ovx, ovy, margin=15.0, damping=0.7, max_speed=1.0)
    got = ctx.download()
    ctx.close()
    assert np.abs(got["x"] - b["x"]).max() <= 1e-9 * U and np.abs(got["y"] - b["y"]).max() <= 1e-9 * U
    assert np.abs(got["vx"] - b["vx"]).max() <= 1e-9 * np.abs(b["vx"]).max()
    # the clamp itself is exact: every body the oracle put on an edge is on the same edge
    edge = (b["x"] == 15.0) | (b["x"] == U - 15.0)
    assert edge.any() and np.array_equal(got["x"][edge], b["x"][edge])
