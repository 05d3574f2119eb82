"""CPU: the C-ABI library loads and exports every symbol include/lpe_bh.h declares; host-only entry points work;
the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import lpe_bh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lpe_bh.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lpe_bh_\w+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ("lpe_bh_create", "lpe_bh_destroy", "lpe_bh_upload", "lpe_bh_step", "lpe_bh_download",
              "lpe_bh_update_host", "lpe_bh_dump_tree", "lpe_bh_get_stats", "lpe_bh_last_error",
              "lpe_bh_set_shard", "lpe_bh_step_begin", "lpe_bh_step_finish", "lpe_bh_workload", "lpe_bh_boundary",
              "lpe_bh_xchg_export", "lpe_bh_xchg_import", "lpe_bh_xchg_set_peer", "lpe_bh_xchg_p2p_ready", "lpe_bh_xchg_reset"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    lib = lpe_bh.load_library()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/lpe_bh.h but not exported: {missing}"


def test_struct_layouts_match_header():
    assert C.sizeof(lpe_bh.Params) == 7 * 8 + 6 * 4
    assert C.sizeof(lpe_bh.Stats) == 16 * 8 + 4 * 4 + 6 * 4 + 2 * 8
    assert C.sizeof(lpe_bh.TreeDump) == 10 * 8
    assert C.sizeof(lpe_bh.DeviceView) == 8 * 8
    assert C.sizeof(lpe_bh.BoundaryParams) == 4 * 8


def test_workloads_are_deterministic_and_in_bounds():
    U = float(2 ** 20)
    for kind in ("disk", "plummer", "two_galaxies"):
        a = lpe_bh.workload(kind, 5000, 42, U)
        b = lpe_bh.workload(kind, 5000, 42, U)
        c = lpe_bh.workload(kind, 5000, 43, U)
        assert all(np.array_equal(p, q) for p, q in zip(a, b))
        assert not np.array_equal(a[0], c[0])
        assert np.all((a[0] >= 0) & (a[0] < U) & (a[1] >= 0) & (a[1] < U))
        assert np.all(a[4] > 0)
    x, y, vx, vy, m = lpe_bh.workload("keplerian", 2000, 5, 6e9)
    assert m[0] == 1e36 and x[0] == 3e9 and y[0] == 3e9          # createCentralBody, keplerian_disk.cpp:45-53
    r = np.hypot(x[1:] - 3e9, y[1:] - 3e9)
    assert r.min() > 0.9e9 and r.max() < 2.6e9                    # inner 100 px, outer 240 px at 1e7 m/px
    assert np.all(m[1:] > 1e3)


def test_shard_helpers_partition_every_position_once():
    for n, R in ((1, 1), (5000, 2), (100_000, 8), (2048 * 7 + 3, 4)):
        chunk = lpe_bh.shard_chunk(n, R)
        assert chunk % lpe_bh.SHARD_BLOCK == 0 and chunk * R >= n
        pos = np.arange(0, n, 97)
        seen = set()
        for i in pos:
            r, s = lpe_bh.shard_owner(int(i), R)
            assert 0 <= r < R and 0 <= s < chunk
            assert (r, s) not in seen
            seen.add((r, s))


def test_no_gpu_means_loud_failure_not_fallback():
    lib = lpe_bh.load_library()
    if lib.lpe_bh_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        lpe_bh.BarnesHut(0)


def test_library_is_sm100a_with_packed_fp32_list_loops():
    """The shipped library holds sm_100a SASS and the traversal's list loops really are packed fp32
    (FADD2 / FMUL2 / FFMA2): a silent fall-back to scalar code or another arch would show here, without a GPU."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    lpe_bh.load_library()
    out = subprocess.run(["cuobjdump", "-sass", lpe_bh.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in out
    body = out[out.index("k_traverse2"):]
    for op in ("FADD2", "FMUL2", "FFMA2", "MUFU.RSQ"):
        assert op in body, op
