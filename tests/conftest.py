import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "little-physics-engine_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The built libraries are git-ignored. When a checkout is tested before __graft_entry__.build() ran, build them
    # here (nvcc cross-compiles without a GPU; nothing is built when they are present and newer than their sources).
    import shutil
    try:
        import lpe_bh
        if shutil.which("nvcc") or not os.path.exists(lpe_bh.LIB_PATH):
            lpe_bh.build_library()
    except Exception as e:   # the tests that need the library then fail loudly on their own
        print(f"conftest: liblpe_bh.so could not be built: {e}", file=sys.stderr)
    try:
        import oracle_py
        oracle_py.build_port()
    except Exception as e:
        print(f"conftest: oracle port could not be built: {e}", file=sys.stderr)


@pytest.fixture(scope="session")
def port():
    import oracle_py
    return oracle_py.PortLib()


@pytest.fixture(scope="session")
def reflib():
    import oracle_py
    if not oracle_py.RefLib.available():
        pytest.skip("oracle/_ref/libref_bh.so not built (needs /root/reference)")
    return oracle_py.RefLib()


@pytest.fixture(scope="session")
def bh():
    import lpe_bh
    ctx = lpe_bh.BarnesHut(0)   # raises without a CUDA device or without the built library: no fallback
    yield ctx
    ctx.close()


def golden_names():
    g = os.path.join(ROOT, "tests", "golden")
    return sorted(f[:-4] for f in os.listdir(g) if f.endswith(".npz"))


def load_golden(name):
    import numpy as np
    d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    U, eps, theta, thr, dt_kick, dt_drift, steps = d["cfg"]
    return d, dict(U=float(U), eps=float(eps), theta=float(theta), thr=float(thr), dt_kick=float(dt_kick),
                   dt_drift=float(dt_drift), steps=int(steps))
