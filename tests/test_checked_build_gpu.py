"""GPU: the parity and decomposition tests once more against liblpe_bh_checked.so — the same kernels compiled with
-DLPE_CHECKED, where every computed index into the big device arrays (sort scatter, gather, topology, aggregation,
record loads of both traversal kernels, frame stack, accept list, export and top-of-tree writes) is compared with the
array's extent before use and a violation fails the next synchronising call (bh_common.cuh: LPE_CHECK / lpe_idx).
compute-sanitizer is closed on this pool ("find a bad access with bounds checks and asserts of your own"); this is
that. The production library carries none of the checks."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, "little-physics-engine_b200", "liblpe_bh_checked.so")


@pytest.mark.gpu
def test_parity_and_decomposition_suites_pass_with_bounds_checks_on():
    if not os.path.exists(CHECKED):
        pytest.skip("liblpe_bh_checked.so not built")
    env = dict(os.environ, LPE_BH_LIB=CHECKED)
    probe = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, 'little-physics-engine_b200'); import lpe_bh; "
                            "print(lpe_bh.load_library().lpe_bh_version().decode())"], cwd=ROOT, env=env, capture_output=True, text=True)
    assert "CHECKED" in probe.stdout, probe.stdout + probe.stderr
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_parity_gpu.py", "tests/test_dd_gpu.py", "tests/test_boundary.py",
                        "-m", "gpu", "-x", "-q"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
