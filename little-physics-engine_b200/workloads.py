"""Deterministic synthetic body distributions (csrc/workloads.cpp) through libworkloads.so — host code only.

The same generators are also linked into liblpe_bh.so (lpe_bh_workload); this module exists so that a process that
must not load the CUDA product library (bench.py --impl reference) can still make the very same inputs.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libworkloads.so")
KINDS = {"disk": 0, "plummer": 1, "two_galaxies": 2, "keplerian": 3, "keplerian_counter": 4}
_lib = None


def workload(kind, n, seed, U):
    """Returns x, y, vx, vy, m (float64, creation order); SURVEY.md 8(d)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run little-physics-engine_b200/build.sh")
        _lib = C.CDLL(LIB_PATH)
    arrs = [np.empty(n, np.float64) for _ in range(5)]
    rc = _lib.lpe_bh_workload(C.c_int(KINDS[kind] if isinstance(kind, str) else kind), C.c_uint64(n), C.c_uint64(seed),
                              C.c_double(U), *[a.ctypes.data_as(C.c_void_p) for a in arrs])
    if rc:
        raise RuntimeError("lpe_bh_workload failed")
    return tuple(arrs)
