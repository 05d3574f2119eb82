"""CPU: pins the oracle (oracle/bh_oracle.c) to the reference.

(1) bit-for-bit against golden vectors generated from the reference's own compiled sources
    (tests/golden/make_golden.py); (2) bit-for-bit against that compiled reference itself when it is present
    (oracle/_ref, built from /root/reference in this container); (3) the reference quirks of SURVEY.md §8(a).
"""
import numpy as np
import pytest

import oracle_py as O
from conftest import golden_names, load_golden
from parity import gen_uniform


@pytest.mark.parametrize("name", golden_names())
def test_port_matches_golden_bit_exact(port, name):
    d, c = load_golden(name)
    p = O.make_params(c["U"], c["eps"], theta=c["theta"], thr=c["thr"], dt_kick=c["dt_kick"], dt_drift=c["dt_drift"])
    r = port.run(p, d["x"], d["y"], d["vx"], d["vy"], d["m"], comp=d["comp"], rank=d["rank"], nsteps=c["steps"], threads=2)
    for k in ("x", "y", "vx", "vy"):
        assert np.array_equal(r[k], d["out_" + k]), f"{name}: {k} differs from the reference's output"
    tree, st = port.tree(p, d["x"], d["y"], d["m"], comp=d["comp"], rank=d["rank"])
    assert tree.tobytes() == d["tree"].tobytes(), f"{name}: tree dump differs from the reference's"
    assert [st["pool_nodes"], st["nonempty_nodes"], st["internal_nodes"], st["max_depth"]] == list(d["tree_stats"])


@pytest.mark.parametrize("name", golden_names())
def test_default_rank_is_entt_view_order(port, name):
    """rank=None in the port means 'newest entity first' — what the reference's view iteration does."""
    d, c = load_golden(name)
    p = O.make_params(c["U"], c["eps"], theta=c["theta"], thr=c["thr"], dt_kick=c["dt_kick"], dt_drift=c["dt_drift"])
    a, _ = port.tree(p, d["x"], d["y"], d["m"], comp=d["comp"], rank=d["rank"])
    b, _ = port.tree(p, d["x"], d["y"], d["m"], comp=d["comp"], rank=None)
    assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("n,seed,thr", [(1, 1, 0.0), (2, 2, 0.0), (37, 3, 0.0), (5000, 4, 0.0), (5000, 5, 1.1e6)])
def test_port_matches_compiled_reference(port, reflib, n, seed, thr):
    rng = np.random.default_rng(seed)
    U = 4096.0
    x, y = rng.random(n) * U, rng.random(n) * U
    vx, vy = rng.standard_normal(n), rng.standard_normal(n)
    m = 1e6 * (0.5 + rng.random(n))
    p = O.make_params(U, U / 2 ** 13, theta=0.5, thr=thr, dt_kick=1 / 120, dt_drift=0.004)
    a = reflib.run(p, x, y, vx, vy, m, nsteps=3)
    b = port.run(p, x, y, vx, vy, m, nsteps=3, threads=4)
    for k in ("x", "y", "vx", "vy"):
        assert np.array_equal(a[k], b[k])
    ta, _ = reflib.tree(p, x, y, m)
    tb, _ = port.tree(p, x, y, m)
    assert ta.tobytes() == tb.tobytes()


def test_reference_view_order_is_newest_first(reflib):
    r = reflib.view_rank(5)
    assert list(r) == [4, 3, 2, 1, 0]


def test_first_occupant_double_count(port):
    """masses {1,2,4,8}e6 give a root mass of 2.3e7 (true 1.5e7): the first inserted body (the newest
    entity, 8e6) is counted twice in every internal cell that contains it."""
    d, c = load_golden("four_body")
    p = O.make_params(c["U"], c["eps"], thr=0.0)
    tree, _ = port.tree(p, d["x"], d["y"], d["m"])
    assert tree[0]["mass"] == 2.3e7 and tree[0]["is_leaf"] == 0 and tree[0]["single"] == 3
    p.quirk = 0
    tree, _ = port.tree(p, d["x"], d["y"], d["m"])
    assert tree[0]["mass"] == 1.5e7


def test_all_small_early_exit_leaves_velocities_untouched(port):
    d, c = load_golden("all_small")
    assert np.array_equal(d["out_vx"], d["vx"]) and np.array_equal(d["out_vy"], d["vy"])
    assert np.array_equal(d["out_x"], d["x"] + d["vx"] * c["dt_drift"])  # MovementSystem still runs


def test_out_of_bounds_bodies_feel_but_do_not_exert(port):
    """SURVEY.md Q6."""
    U = 1024.0
    x = np.array([200.0, 800.0, -50.0])
    y = np.array([500.0, 500.0, 500.0])
    m = np.array([1e9, 1e9, 1e12])
    z = np.zeros(3)
    p = O.make_params(U, 1.0, run_movement=False)
    r = port.run(p, x, y, z, z, m)
    assert r["vx"][2] > 0.0                       # pulled towards the in-bounds pair
    ax, _ = port.direct(p, x, y, m)
    only_pair = O.G_REAL * 1e9 * 600.0 / (600.0 ** 2 + 1.0) ** 1.5
    assert abs(ax[0] - only_pair) < 1e-12 * only_pair  # the heavy outsider exerts nothing


def test_direct_sum_vs_textbook_tree(port):
    """The quirk-free tree approximates the direct sum at the few-% level; the reference tree does not (K3)."""
    rng = np.random.default_rng(11)
    n, U = 3000, 1024.0
    x, y = rng.random(n) * U, rng.random(n) * U
    m = 1e6 * (0.5 + rng.random(n))
    z = np.zeros(n)
    ax, ay = port.direct(p := O.make_params(U, U / 2 ** 10, run_movement=False, dt_kick=1.0), x, y, m, threads=4)
    errs = {}
    for quirk in (0, 1):
        p.quirk = quirk
        r = port.run(p, x, y, z, z, m, threads=4)
        e = np.hypot(r["vx"] - ax, r["vy"] - ay) / np.hypot(ax, ay)
        errs[quirk] = float(np.median(e))
    assert errs[0] < 2e-2 and errs[1] > 2 * errs[0], errs


def test_debugstats_force_statistics_port_equals_compiled_reference(port, reflib):
    """A10: DebugStats::updateForce (core/debug.hpp:37-41) is fed force = G*M*m/distSq once per accepted node
    (barnes_hut.cpp:278). The port's max / sum / count against the reference's own process globals."""
    x, y, vx, vy, m = gen_uniform(1500, 1024.0, 61)
    p = O.make_params(1024.0, 0.25)
    a = port.run(p, x, y, vx, vy, m, threads=1)["stats"]
    b = reflib.run(p, x, y, vx, vy, m)["stats"]
    assert a["force_count"] == b["force_count"] > 0
    assert a["force_max"] == b["force_max"]
    assert abs(a["force_sum"] - b["force_sum"]) <= 1e-12 * b["force_sum"]
