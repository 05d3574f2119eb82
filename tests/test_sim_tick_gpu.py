"""GPU: BASELINE config C1 as the reference itself runs it — the reference's own ECSSimulator::tick (src/sim.cpp,
unmodified: Fluid -> Boundary -> BasicGravity -> RigidBodyCollision -> BarnesHut -> Rotation -> Movement -> Sleep,
sim.cpp:107-114) on its own Keplerian-disk scenario at 10 000 bodies, 100 ticks, once with every system the
reference's (oracle/_ref/sim_ref) and once with BarnesHutSystem and BoundarySystem replaced by the drop-in classes
(oracle/_ref/sim_dropin: same sim.cpp, same scenario code, same registry snapshot). Built where /root/reference exists
(oracle/Makefile target `sim`); the binaries travel to the GPU box. Besides parity this is the proof that sim.cpp and
i_scenario.hpp compile unchanged against the drop-in headers."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "sim_ref")
OURS = os.path.join(ROOT, "oracle", "_ref", "sim_dropin")


def run(exe, n, ticks, out):
    r = subprocess.run([exe, str(n), str(ticks), out], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-500:])
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    return rep, np.fromfile(out, dtype=np.float64).reshape(-1, 4)


def test_c1_hundred_ticks_of_the_reference_simulator_with_the_drop_in(tmp_path):
    if not (os.path.exists(REF) and os.path.exists(OURS)):
        pytest.skip("oracle/_ref/sim_ref / sim_dropin not built (need /root/reference at build time)")
    n, ticks = 10000, 100
    ref_rep, ref = run(REF, n, ticks, str(tmp_path / "ref.bin"))
    our_rep, got = run(OURS, n, ticks, str(tmp_path / "ours.bin"))
    assert ref.shape == got.shape == (n, 4)
    # 100 ticks of kick + drift: velocities within 1e-4 (norm-wise) of the reference's, positions within 1e-9 of the universe
    dv = np.hypot(got[:, 2] - ref[:, 2], got[:, 3] - ref[:, 3])
    vmag = np.hypot(ref[:, 2], ref[:, 3])
    assert np.sqrt(np.sum(dv ** 2) / np.sum(vmag ** 2)) <= 1e-4
    assert np.max(np.hypot(got[:, 0] - ref[:, 0], got[:, 1] - ref[:, 1])) / 6e9 <= 1e-9
    print(json.dumps({"c1_real_tick": {"reference": ref_rep, "drop_in": our_rep,
                                       "speedup": ref_rep["ms_per_tick"] / our_rep["ms_per_tick"]}}))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "c1_real_tick.json"), "w") as f:
            json.dump({"reference": ref_rep, "drop_in": our_rep, "speedup": ref_rep["ms_per_tick"] / our_rep["ms_per_tick"]}, f)
