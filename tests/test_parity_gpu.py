"""GPU: the CUDA path, called through the C ABI, against the oracle.

Gates: Morton keys / sort order / compressed topology bit-exact; node aggregates <= 1e-12 relative;
per-body accepted-interaction counts bit-exact (every theta decision is the reference's); velocity change and
post-step positions within 1e-4 relative in FAST precision (bodies whose |dv| is below 1e-3 of the median are judged
against 1e-3 * median) and within 1e-8 in STRICT precision.
"""
import numpy as np
import pytest

import lpe_bh
import oracle_py as O
from conftest import golden_names, load_golden
from parity import check_preorder, compare_tree, gen_uniform, keys_to_xy, rel_err

pytestmark = pytest.mark.gpu

FAST_TOL = 1e-4     # north_star: "within a stated relative tolerance (e.g. 1e-4 fp32)"
STRICT_TOL = 1e-8


def run_gpu(bh, d, c, precision, quirk=True, counts=True, key_order=lpe_bh.KEYS_AUTO):
    pg = lpe_bh.make_params(c["U"], c["eps"], theta=c["theta"], thr=c["thr"], dt_kick=c["dt_kick"],
                            dt_drift=c["dt_drift"], quirk=quirk, precision=precision, key_order=key_order)
    bh.set_instrumentation(timing=False, counts=counts)
    bh.upload(d["x"], d["y"], d["vx"], d["vy"], d["m"], rank=d.get("rank"), comp=d.get("comp"))
    bh.step(pg, c["steps"])
    return bh.download()


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("precision", [lpe_bh.PREC_STRICT, lpe_bh.PREC_FAST])
def test_golden_vectors(bh, name, precision):
    """Outputs of the reference itself (tests/golden, generated from the compiled reference sources)."""
    g, c = load_golden(name)
    d = {k: g[k] for k in ("x", "y", "vx", "vy", "m", "rank", "comp")}
    got = run_gpu(bh, d, c, precision)
    tol = STRICT_TOL if precision == lpe_bh.PREC_STRICT else FAST_TOL
    dv = rel_err((got["vx"] - g["vx"], got["vy"] - g["vy"]), (g["out_vx"] - g["vx"], g["out_vy"] - g["vy"]))
    assert dv["max"] <= tol and dv["norm"] <= tol, (name, dv)
    dx = np.max(np.hypot(got["x"] - g["out_x"], got["y"] - g["out_y"])) / c["U"]
    assert dx <= 1e-9, (name, dx)


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("key_order", [lpe_bh.KEYS_MORTON, lpe_bh.KEYS_HILBERT], ids=["morton", "hilbert"])
def test_golden_tree_topology_and_aggregates(bh, name, key_order):
    g, c = load_golden(name)
    if c["thr"] > 0 and name == "all_small":
        pass  # the tree is still built on the device; the reference returned before building it: nothing to compare
    d = {k: g[k] for k in ("x", "y", "vx", "vy", "m", "rank", "comp")}
    c1 = dict(c, steps=1)
    run_gpu(bh, d, c1, lpe_bh.PREC_STRICT, key_order=key_order)
    dump = bh.dump_tree()
    assert dump["stats"]["hilbert"] == (1 if key_order == lpe_bh.KEYS_HILBERT else 0)
    check_preorder(dump)
    rep = compare_tree(dump, g["tree"], c["U"])
    assert rep["worst_rel"] <= 1e-12


CASES = [
    # name, n, seed, eps_div (eps = U / 2^k; 0 -> eps = 0), theta, thr
    ("n1", 1, 1, 14, 0.5, 0.0),
    ("n2", 2, 2, 14, 0.5, 0.0),
    ("n33_ragged", 33, 3, 14, 0.5, 0.0),
    ("n2049_tile_edge", 2049, 4, 14, 0.5, 0.0),
    ("n20000", 20000, 5, 14, 0.5, 0.0),
    ("n20000_thr", 20000, 6, 14, 0.5, 1.2e6),
    ("theta03", 5000, 7, 14, 0.3, 0.0),
    ("theta10", 5000, 7, 14, 1.0, 0.0),
    ("eps_zero_depth30", 5000, 8, 0, 0.5, 0.0),
    ("eps_big_aggregated_terminals", 5000, 9, 6, 0.5, 0.0),
    ("theta_zero_every_cell_opened", 2000, 10, 14, 0.0, 0.0),     # direct sum through the leaves; frontier overflow path
    ("theta_huge_root_children_only", 2000, 11, 14, 50.0, 0.0),
]


@pytest.mark.parametrize("name,n,seed,epsdiv,theta,thr", CASES, ids=[c[0] for c in CASES])
def test_seeded_cases_vs_oracle(bh, port, name, n, seed, epsdiv, theta, thr):
    U = 1024.0
    x, y, vx, vy, m = gen_uniform(n, U, seed)
    eps = U / 2 ** epsdiv if epsdiv else 0.0
    c = dict(U=U, eps=eps, theta=theta, thr=thr, dt_kick=1 / 120, dt_drift=0.006, steps=1)
    po = O.make_params(U, eps, theta=theta, thr=thr, dt_kick=c["dt_kick"], dt_drift=c["dt_drift"])
    ref = port.run(po, x, y, vx, vy, m, threads=8, per_body=True)
    d = dict(x=x, y=y, vx=vx, vy=vy, m=m)
    for precision, tol in ((lpe_bh.PREC_STRICT, STRICT_TOL), (lpe_bh.PREC_FAST, FAST_TOL)):
        got = run_gpu(bh, d, c, precision)
        acc, _ = bh.counts()
        assert np.array_equal(acc, ref["accepted"]), f"{name}: per-body accepted counts differ from the oracle"
        dv = rel_err((got["vx"] - vx, got["vy"] - vy), (ref["vx"] - vx, ref["vy"] - vy))
        assert dv["max"] <= tol and dv["norm"] <= tol, (name, precision, dv)
        assert np.max(np.hypot(got["x"] - ref["x"], got["y"] - ref["y"])) / U <= 1e-9
        if precision == lpe_bh.PREC_STRICT:
            dump = bh.dump_tree()
            check_preorder(dump)
            nodes, _ = port.tree(po, x, y, m)
            compare_tree(dump, nodes, U)
            # keys are sorted and the permutation is a permutation
            nin = dump["stats"]["n_in_tree"]
            keys = dump["sorted_keys"]
            assert np.all(keys[:-1] <= keys[1:])
            assert np.array_equal(np.sort(dump["sorted_index"]), np.arange(n))
            # keys are bit-exact: recompute from the fp64 positions with the reference's comparisons
            D = dump["stats"]["depth"]
            h = U / 2 ** D
            ix, iy = keys_to_xy(keys[:nin], D, dump["stats"]["hilbert"])
            b = dump["sorted_index"][:nin]
            assert np.all(ix * h <= x[b]) and np.all(x[b] < (ix + 1) * h)
            assert np.all(iy * h <= y[b]) and np.all(y[b] < (iy + 1) * h)


@pytest.mark.parametrize("key_order", [lpe_bh.KEYS_MORTON, lpe_bh.KEYS_HILBERT], ids=["morton", "hilbert"])
def test_key_order_does_not_change_decisions(bh, port, key_order):
    """Morton or Hilbert sort keys: same cells, same tree, so the same per-body accept/open decisions in both
    precisions (only the summation order differs)."""
    x, y, vx, vy, m = gen_uniform(20000, 1024.0, 51)
    ref = port.run(O.make_params(1024.0, 0.25, dt_drift=0.004), x, y, vx, vy, m, threads=8, per_body=True)
    for precision, tol in ((lpe_bh.PREC_STRICT, STRICT_TOL), (lpe_bh.PREC_FAST, FAST_TOL)):
        bh.set_instrumentation(counts=True)
        bh.upload(x, y, vx, vy, m)
        bh.step(lpe_bh.make_params(1024.0, 0.25, dt_drift=0.004, precision=precision, key_order=key_order), 1)
        got = bh.download()
        acc, _ = bh.counts()
        assert np.array_equal(acc, ref["accepted"])
        assert rel_err((got["vx"] - vx, got["vy"] - vy), (ref["vx"] - vx, ref["vy"] - vy))["max"] <= tol
    bh.set_instrumentation()


def test_empty_and_no_source_inputs(bh):
    pg = lpe_bh.make_params(1024.0, 0.1)
    z = np.zeros(0)
    bh.upload(z, z, z, z, z)
    bh.step(pg, 1)
    assert bh.download()["x"].shape == (0,)
    # bodies but none inside the universe: nothing exerts, everything still drifts
    x = np.array([-5.0, 2000.0, -1.0]); y = np.array([10.0, 10.0, -3.0]); v = np.array([1.0, 2.0, 3.0]); m = np.ones(3)
    bh.upload(x, y, v, v, m)
    bh.step(pg, 1)
    got = bh.download()
    assert np.array_equal(got["vx"], v) and np.allclose(got["x"], x + v / 120, rtol=0, atol=1e-12)
    assert bh.stats()["n_in_tree"] == 0


def test_coincident_bodies_do_not_hang(bh):
    """The reference recurses forever on coincident points (defect D3); the device tree buckets them at depth 30."""
    x = np.array([100.0, 100.0, 100.0, 900.0]); y = np.array([200.0, 200.0, 200.0, 900.0]); z = np.zeros(4)
    bh.upload(x, y, z, z, np.full(4, 1e6))
    bh.step(lpe_bh.make_params(1024.0, 0.5), 1)
    got = bh.download()
    assert np.all(np.isfinite(got["vx"])) and bh.stats()["n_terminals"] == 2


def test_strict_mode_drop_in_kick_only(bh, port):
    """do_drift=0 is BarnesHutSystem::update alone: velocities change, positions do not (barnes_hut.cpp:285-286)."""
    x, y, vx, vy, m = gen_uniform(3000, 1024.0, 21)
    pg = lpe_bh.make_params(1024.0, 0.25, do_drift=False)
    bh.upload(x, y, vx, vy, m)
    bh.step(pg, 1)
    got = bh.download()
    assert np.array_equal(got["x"], x) and np.array_equal(got["y"], y)
    po = O.make_params(1024.0, 0.25, run_movement=False)
    ref = port.run(po, x, y, vx, vy, m, threads=4)
    assert rel_err((got["vx"] - vx, got["vy"] - vy), (ref["vx"] - vx, ref["vy"] - vy))["max"] <= FAST_TOL


def test_update_host_round_trip(bh, port):
    """The single call the ECS drop-in makes (lpe_bh_update_host), host buffers in and out."""
    x, y, vx, vy, m = gen_uniform(4000, 1024.0, 22)
    po = O.make_params(1024.0, 0.25, dt_drift=0.004)
    ref = port.run(po, x, y, vx, vy, m, threads=4)
    pg = lpe_bh.make_params(1024.0, 0.25, dt_drift=0.004)
    gx, gy, gvx, gvy = x.copy(), y.copy(), vx.copy(), vy.copy()
    bh.update_host(pg, gx, gy, gvx, gvy, m)
    assert rel_err((gvx - vx, gvy - vy), (ref["vx"] - vx, ref["vy"] - vy))["max"] <= FAST_TOL
    assert np.max(np.hypot(gx - ref["x"], gy - ref["y"])) / 1024.0 <= 1e-9


def test_multi_step_trajectory(bh, port):
    """20 resident steps stay on the oracle's trajectory (decisions can only diverge through 1e-7-level position noise)."""
    x, y, vx, vy, m = lpe_bh.workload("keplerian", 3000, 9, 6e9)
    kw = dict(theta=0.5, thr=1e3, dt_kick=1 / 120, dt_drift=6.756e-3)
    ref = port.run(O.make_params(6e9, 2e7, **kw), x, y, vx, vy, m, nsteps=20, threads=8)
    bh.upload(x, y, vx, vy, m)
    bh.step(lpe_bh.make_params(6e9, 2e7, **kw), 20)
    got = bh.download()
    dv = rel_err((got["vx"] - vx, got["vy"] - vy), (ref["vx"] - vx, ref["vy"] - vy))
    assert dv["norm"] <= FAST_TOL and dv["max"] <= 10 * FAST_TOL, dv
    assert np.max(np.hypot(got["x"] - ref["x"], got["y"] - ref["y"])) / 6e9 <= 1e-9


@pytest.mark.parametrize("n,eps", [(30000, 0.25), (30000, 0.0), (3000, 0.25)])
def test_traversal_kernels_agree(bh, port, n, eps):
    """FAST precision has two kernels: the two-phase production kernel — with the far field shared by the four warps of a
    CTA (default) or walked per warp only — and the depth-first kernel (also its overflow path). All four routes must take
    identical per-body decisions (= the oracle's) and agree to fp32 rounding."""
    x, y, vx, vy, m = gen_uniform(n, 1024.0, 41)
    pg = lpe_bh.make_params(1024.0, eps, dt_drift=0.004)
    ref = port.run(O.make_params(1024.0, eps, dt_drift=0.004), x, y, vx, vy, m, threads=8, per_body=True)
    outs = {}
    for name, kw in (("two_phase", {}), ("two_phase_warp_only", dict(warp_only=True)), ("depth_first", dict(force_dfs=True)),
                     ("overflow", dict(force_overflow=True))):
        bh.set_instrumentation(counts=True, **kw)
        bh.upload(x, y, vx, vy, m)
        bh.step(pg, 1)
        outs[name] = bh.download()
        acc, _ = bh.counts()
        st = bh.stats()
        assert np.array_equal(acc, ref["accepted"]), name
        if name == "overflow":
            assert st["overflow_chunks"] == ((n + 2047) // 2048) * 64   # every 32-body chunk of every 2048-body block
        if name.startswith("two_phase"):
            assert st["overflow_chunks"] == 0
        dv = rel_err((outs[name]["vx"] - vx, outs[name]["vy"] - vy), (ref["vx"] - vx, ref["vy"] - vy))
        assert dv["max"] <= FAST_TOL, (name, dv)
    bh.set_instrumentation()
    # the overflow route IS the depth-first kernel: bit-identical
    for k in ("x", "y", "vx", "vy"):
        assert np.array_equal(outs["depth_first"][k], outs["overflow"][k])


def test_two_rank_sharded_step_on_one_gpu(port):
    """Two contexts on one device play ranks 0 and 1 (launched one after the other: no kernel waits on another);
    the allgather is done through the host. Result must equal the unsharded GPU step bit for bit."""
    n = 3 * 2048 + 100
    x, y, vx, vy, m = gen_uniform(n, 1024.0, 31)
    pg = lpe_bh.make_params(1024.0, 0.25, dt_drift=0.004)
    one = lpe_bh.BarnesHut(0)
    one.upload(x, y, vx, vy, m); one.step(pg, 2); ref = one.download(); one.close()
    ranks = [lpe_bh.BarnesHut(0) for _ in range(2)]
    for r, c in enumerate(ranks):
        c.set_shard(r, 2)
        c.upload(x, y, vx, vy, m)
    for _ in range(2):
        for c in ranks:
            c.step_begin(pg)
        slices = [c.xchg_read_send() for c in ranks]
        for c in ranks:
            for r, s in enumerate(slices):
                c.xchg_write_recv(r, s)
            c.step_finish()
    for c in ranks:
        got = c.download()
        for k in ("x", "y", "vx", "vy"):
            assert np.array_equal(got[k], ref[k]), k
        c.close()


def test_two_rank_direct_exchange_on_one_gpu(port):
    """Same two-context emulation, but the exchange is the fused one: each context's traversal kernel stores its
    slice straight into BOTH contexts' receive buffers (lpe_bh_xchg_set_peer, the same-process form of the CUDA IPC
    handles); no slice is copied by the host. Three steps so that both generations of the buffer are used."""
    n = 5 * 2048 + 77
    x, y, vx, vy, m = gen_uniform(n, 1024.0, 32)
    pg = lpe_bh.make_params(1024.0, 0.25, dt_drift=0.004)
    one = lpe_bh.BarnesHut(0)
    one.upload(x, y, vx, vy, m); one.step(pg, 3); ref = one.download(); one.close()
    ranks = [lpe_bh.BarnesHut(0) for _ in range(2)]
    for r, c in enumerate(ranks):
        c.set_shard(r, 2)
        c.upload(x, y, vx, vy, m)
    views = [c.device_view() for c in ranks]
    for c in ranks:
        assert not c.xchg_p2p_ready()
        for r, v in enumerate(views):
            c.xchg_set_peer(r, v.xchg_recv)
        assert c.xchg_p2p_ready()
    for _ in range(2):
        for c in ranks:
            c.step_begin(pg)
        for c in ranks:
            c.synchronize()          # the barrier a real run gets from a one-element allreduce
        for c in ranks:
            c.step_finish()
    # back to the collective exchange in mid-run (what a caller does when not every rank could open every handle)
    for c in ranks:
        c.xchg_reset()
        assert not c.xchg_p2p_ready()
    for c in ranks:
        c.step_begin(pg)
    slices = [c.xchg_read_send() for c in ranks]
    for c in ranks:
        for r, s in enumerate(slices):
            c.xchg_write_recv(r, s)
        c.step_finish()
    for c in ranks:
        got = c.download()
        for k in ("x", "y", "vx", "vy"):
            assert np.array_equal(got[k], ref[k]), k
        c.close()


def test_reordered_state_keeps_the_creation_order_interface(port):
    """Resident steps keep the device state in key order; everything the ABI hands out stays in creation order:
    download, per-body counts, the direct sum (kick only, so positions do not move and the sums must agree
    before and after the steps), and partial uploads."""
    n = 20000
    x, y, vx, vy, m = gen_uniform(n, 1024.0, 77)
    pg = lpe_bh.make_params(1024.0, 0.25, do_drift=False)
    bh = lpe_bh.BarnesHut(0)
    bh.set_instrumentation(counts=True)
    bh.upload(x, y, vx, vy, m)
    ax0, ay0 = bh.direct_accel(pg, 100, 5000)
    bh.step(pg, 1)
    acc1, _ = bh.counts()
    bh.step(pg, 1)
    acc2, _ = bh.counts()
    ax1, ay1 = bh.direct_accel(pg, 100, 5000)
    # (the sources are summed in state order, so the two sums differ by fp64 rounding only)
    scale = np.abs(np.concatenate([ax0, ay0])).max()
    assert np.abs(ax0 - ax1).max() <= 1e-12 * scale and np.abs(ay0 - ay1).max() <= 1e-12 * scale
    assert np.array_equal(acc1, acc2)          # same positions -> same decisions, reported per creation index
    got = bh.download()
    assert np.array_equal(got["x"], x) and np.array_equal(got["y"], y)
    # partial upload in creation order lands in the right slots
    bh.upload_velocities(vx, vy)
    back = bh.download()
    assert np.array_equal(back["vx"], vx) and np.array_equal(back["vy"], vy)
    bh.close()


def test_context_reuse_across_sizes_and_modes(port):
    """One context: small upload, resident steps, a larger upload (every device buffer is re-allocated, the key-order
    state and the look-back epochs start over), the host tick, then sharded mode — each result equal to a fresh
    context's."""
    U = 1024.0
    pg = lpe_bh.make_params(U, 0.25, dt_drift=0.004)

    def fresh(n, seed, steps):
        x, y, vx, vy, m = gen_uniform(n, U, seed)
        c = lpe_bh.BarnesHut(0)
        c.upload(x, y, vx, vy, m); c.step(pg, steps); out = c.download(); c.close()
        return (x, y, vx, vy, m), out

    ctx = lpe_bh.BarnesHut(0)
    for n, seed, steps in ((700, 1, 3), (9000, 2, 2), (300, 3, 4), (9000, 2, 2)):
        (x, y, vx, vy, m), want = fresh(n, seed, steps)
        ctx.upload(x, y, vx, vy, m)
        ctx.step(pg, steps)
        got = ctx.download()
        for k in ("x", "y", "vx", "vy"):
            assert np.array_equal(got[k], want[k]), (n, k)
    # host tick on the same context: one step, arrays updated in place
    (x, y, vx, vy, m), want = fresh(5000, 4, 1)
    hx, hy, hvx, hvy = x.copy(), y.copy(), vx.copy(), vy.copy()
    ctx.update_host(pg, hx, hy, hvx, hvy, m)
    for a, k in ((hx, "x"), (hy, "y"), (hvx, "vx"), (hvy, "vy")):
        assert np.array_equal(a, want[k]), k
    ctx.close()


def test_error_paths_report_and_leave_the_context_usable():
    """Bad calls return an error with a message (lpe_bh_last_error) and never touch the state: the context keeps
    working afterwards. (The reference's style: report and skip the update, barnes_hut.cpp:76-79.)"""
    U = 1024.0
    x, y, vx, vy, m = gen_uniform(500, U, 12)
    c = lpe_bh.BarnesHut(0)
    c.upload(x, y, vx, vy, m)
    good = lpe_bh.make_params(U, 0.25)
    for bad, word in ((lpe_bh.make_params(-1.0, 0.25), "universe_size"), (lpe_bh.make_params(U, 0.25, theta=-0.5), "theta"),
                      (lpe_bh.make_params(U, 0.25, precision=7), "precision")):
        with pytest.raises(RuntimeError, match=word):
            c.step(bad, 1)
    with pytest.raises(RuntimeError, match="out of bounds"):
        c.direct_accel(good, 400, 200)
    with pytest.raises(RuntimeError, match="sharded"):
        c.xchg_export()
    with pytest.raises(RuntimeError):
        c.dump_tree()                      # no step has run yet
    c.step(good, 1)
    got = c.download()
    ref = lpe_bh.BarnesHut(0)
    ref.upload(x, y, vx, vy, m); ref.step(good, 1); want = ref.download(); ref.close()
    for k in ("x", "y", "vx", "vy"):
        assert np.array_equal(got[k], want[k]), k
    # a sharded context refuses the unsharded entry points
    c.set_shard(0, 2)
    c.upload(x, y, vx, vy, m)
    with pytest.raises(RuntimeError, match="sharded"):
        c.step(good, 1)
    c.close()



def test_debugstats_force_statistics_and_largest_mass(bh, port):
    """what the reference feeds DebugStats::updateForce with (max / sum / count of G*M*m/distSq over
    the accepted nodes, barnes_hut.cpp:277-278) reduced on the device, and the largest source mass the two per-tick
    Mass scans ask for (barnes_hut.cpp:55-71, gravity.cpp:41-49)."""
    x, y, vx, vy, m = gen_uniform(20000, 1024.0, 71)
    comp = np.full(len(x), O.HAS_MASS | O.HAS_VELOCITY, np.uint8)
    comp[np.argmax(m)] |= O.BOUNDARY                       # the heaviest body is a Boundary: it must not count
    ref = port.run(O.make_params(1024.0, 0.25), x, y, vx, vy, m, comp=comp, threads=8)["stats"]
    for precision, tol in ((lpe_bh.PREC_FAST, 1e-5), (lpe_bh.PREC_STRICT, 1e-10)):
        bh.set_instrumentation(counts=True)
        bh.upload(x, y, vx, vy, m, comp=comp)
        bh.step(lpe_bh.make_params(1024.0, 0.25, precision=precision), 1)
        st = bh.stats()
        assert st["interactions"] == ref["force_count"]
        assert abs(st["force_sum"] - ref["force_sum"]) <= tol * ref["force_sum"], (precision, st["force_sum"], ref["force_sum"])
        assert abs(st["force_max"] - ref["force_max"]) <= tol * ref["force_max"]
        want = np.max(m[(comp & O.BOUNDARY) == 0])
        assert bh.max_source_mass() == want
    bh.set_instrumentation()


def test_scenario_bodies_made_on_the_device(bh, port):
    """the Keplerian-disk scenario's entity law (keplerian_disk.cpp:45-146) evaluated one thread per body
    on the device equals its host restatement (same counter-based streams) to libm rounding, is reproducible, and steps
    like uploaded bodies do."""
    n, Uk = 50000, 6e9
    hx, hy, hvx, hvy, hm = lpe_bh.workload("keplerian_counter", n, 17, Uk)
    bh.generate("keplerian_counter", n, 17, Uk)
    d = bh.download()
    for got, want, scale in ((d["x"], hx, Uk), (d["y"], hy, Uk), (d["vx"], hvx, np.abs(hvx).max()), (d["vy"], hvy, np.abs(hvy).max())):
        assert np.max(np.abs(got - want)) <= 1e-11 * scale
    assert bh.max_source_mass() == 1e36 and hm[0] == 1e36
    assert 0.09 * Uk < hx.min() and hx.max() < 0.91 * Uk           # outer radius = ScreenLength / 2.5 = 240 of 300 px
    kw = dict(theta=0.5, thr=1e3, dt_kick=1 / 120, dt_drift=6.756e-3)
    bh.step(lpe_bh.make_params(Uk, 2e7, **kw), 1)
    got = bh.download()
    ref = port.run(O.make_params(Uk, 2e7, **kw), d["x"], d["y"], d["vx"], d["vy"], hm, threads=8)
    assert rel_err((got["vx"] - d["vx"], got["vy"] - d["vy"]), (ref["vx"] - d["vx"], ref["vy"] - d["vy"]))["max"] <= FAST_TOL
    bh.generate("keplerian_counter", n, 17, Uk)
    again = bh.download()
    assert np.array_equal(again["x"], d["x"]) and np.array_equal(again["vy"], d["vy"])


def test_replayed_cuda_graph_of_the_step_equals_plain_launches():
    """Resident steps are captured as a CUDA graph the second time their key comes up and replayed afterwards: the result
    must be the plain launches' bit for bit, the replay counter must move, and the launch count must be the same."""
    U = 1024.0
    x, y, vx, vy, m = gen_uniform(20000, U, 77)
    pg = lpe_bh.make_params(U, 0.25, dt_drift=0.004)
    out = {}
    for name, plain in (("plain", True), ("graph", False)):
        c = lpe_bh.BarnesHut(0)
        c.set_instrumentation(plain_launches=plain)
        c.upload(x, y, vx, vy, m)
        for _ in range(9):
            c.step(pg, 1)
        out[name] = (c.download(), c.graph_replays(), c.launch_count())
        c.close()
    assert out["plain"][1] == 0
    assert out["graph"][1] >= 4, "steps 6..9 at the latest must be replays"
    assert out["graph"][2] == out["plain"][2]
    for k in ("x", "y", "vx", "vy"):
        assert np.array_equal(out["graph"][0][k], out["plain"][0][k]), k


def test_host_tick_on_page_locked_arrays_replays_one_graph_and_equals_resident_steps():
    """lpe_bh_update_host on page-locked arrays: copies, step, deferred kick and downloads are one CUDA graph from the third
    tick on. Six ticks through the host arrays == six resident steps of a fresh context, bit for bit (FAST precision: the
    kick is deferred to a creation-order pass, same roundings)."""
    U = 1024.0
    n = 12000
    x, y, vx, vy, m = gen_uniform(n, U, 78)
    pg = lpe_bh.make_params(U, 0.25, dt_drift=0.004)
    ref = lpe_bh.BarnesHut(0)
    ref.upload(x, y, vx, vy, m)
    ref.step(pg, 6)
    want = ref.download()
    ref.close()
    host = [lpe_bh.pinned_array(n) for _ in range(5)]
    for h, a in zip(host, (x, y, vx, vy, m)):
        h[:] = a
    c = lpe_bh.BarnesHut(0)
    for _ in range(6):
        c.update_host_ptrs(pg, n, *[h.ctypes.data for h in host])
    assert c.graph_replays() >= 2
    for h, k in zip(host, ("x", "y", "vx", "vy")):
        assert np.array_equal(h, want[k]), k
    # the resident state followed the host arrays
    got = c.download()
    for k in ("x", "y", "vx", "vy"):
        assert np.array_equal(got[k], want[k]), k
    c.close()
