"""GPU: the ECS drop-in. oracle/_ref/dropin_check (built in the container that has /root/reference) drives a real
entt::registry through OUR Systems::BarnesHutSystem + the reference's MovementSystem and compares with the
reference's own BarnesHutSystem. The binary travels to the GPU box; /root/reference does not."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_check")


@pytest.mark.parametrize("n,kind,steps", [(2000, "keplerian", 3), (10000, "keplerian", 2), (5000, "uniform", 3)])
def test_registry_drop_in_matches_reference(n, kind, steps):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/dropin_check not built (needs /root/reference at build time)")
    r = subprocess.run([BIN, str(n), "7", kind, str(steps)], capture_output=True, text=True, timeout=600)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "{}"
    rep = json.loads(line)
    assert r.returncode == 0 and rep.get("ok"), (rep, r.stderr[-500:])
    assert rep["dv_norm_rel"] <= 1e-4 and rep["dv_max_rel"] <= 1e-4


def test_registry_drop_in_restarts_the_tick_when_the_pools_turn_out_misaligned():
    """The page-wise staging checks the pools' entity order while the device already works on the tick. With a Velocity
    pool in another packed order the tick has to be started over entity by entity — and still match the reference."""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/dropin_check not built (needs /root/reference at build time)")
    r = subprocess.run([BIN, "6000", "7", "keplerian", "3", "misalign"], capture_output=True, text=True, timeout=600)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "{}"
    rep = json.loads(line)
    assert r.returncode == 0 and rep.get("ok") and rep["staging_path"] == 0, (rep, r.stderr[-500:])
    assert rep["dv_norm_rel"] <= 1e-4 and rep["dv_max_rel"] <= 1e-4


@pytest.mark.parametrize("n,seed", [(1, 1), (4097, 2), (50000, 3)])
def test_registry_boundary_drop_in_is_bit_exact(n, seed):
    """OUR Systems::BoundarySystem (host/systems/boundary.{hpp,cpp} -> lpe_bh_boundary) on a real registry with
    sleepers and velocity-less entities, against the reference's own BoundarySystem: identical bits."""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/dropin_check not built (needs /root/reference at build time)")
    r = subprocess.run([BIN, str(n), str(seed), "boundary"], capture_output=True, text=True, timeout=600)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "{}"
    rep = json.loads(line)
    if n == 1:      # a single body may or may not be outside the margin: only the comparison matters
        assert rep.get("mismatches") == 0, (rep, r.stderr[-500:])
    else:
        assert r.returncode == 0 and rep.get("ok") and rep["mismatches"] == 0 and rep["clamped_x"] > 0, (rep, r.stderr[-500:])


BENCH = os.path.join(ROOT, "little-physics-engine_b200", "host", "_build", "dropin_bench")


def test_pool_page_staging_equals_entity_by_entity_staging():
    """Systems::BarnesHutSystem stages aligned EnTT pools page by page into page-locked buffers; the result must be
    the one the entity-by-entity walk of the reference's views gives (same bodies, same insertion ranks)."""
    if not os.path.exists(BENCH):
        pytest.skip("host/_build/dropin_bench not built (needs the reference headers at build time)")
    out = {}
    for mode in ("pagewise", "per_entity"):
        r = subprocess.run([BENCH, "150000", "2", mode], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-800:]
        out[mode] = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["pagewise"]["staging_path_taken"] == 1 and out["per_entity"]["staging_path_taken"] == 0
    # (the two paths hand the bodies over in opposite orders; bodies that share a finest cell may then sit in another
    # warp, which changes the order of fp32 partial sums, never a decision)
    a, b = out["pagewise"]["velocity_checksum"], out["per_entity"]["velocity_checksum"]
    assert abs(a - b) <= 1e-8 * abs(b)
    # the fused-movement variant runs and moves the bodies
    r = subprocess.run([BENCH, "150000", "2", "pagewise", "fused"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and json.loads(r.stdout.strip().splitlines()[-1])["fused_movement"] is True
