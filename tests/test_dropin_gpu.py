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
