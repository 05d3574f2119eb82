"""ctypes bindings for the two CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module. The product path (little-physics-engine_b200) never does.

  RefLib    oracle/_ref/libref_bh.so   the reference's own sources (built here from /root/reference)
  PortLib   oracle/liboracle_bh.so     plain-C restatement, oracle/bh_oracle.c
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libref_bh.so")
PORT_SO = os.path.join(HERE, "liboracle_bh.so")

HAS_MASS, HAS_VELOCITY, BOUNDARY, LIQUID, ASLEEP = 1, 2, 4, 8, 16


class Params(C.Structure):
    _fields_ = [
        ("universe_size", C.c_double),
        ("softening", C.c_double),
        ("seconds_per_tick", C.c_double),
        ("time_acceleration", C.c_double),
        ("base_time_acceleration", C.c_double),
        ("time_scale", C.c_double),
        ("theta", C.c_double),
        ("small_mass_threshold", C.c_double),
        ("G", C.c_double),
        ("run_movement", C.c_int32),
        ("quirk", C.c_int32),
    ]


class Node(C.Structure):
    _fields_ = [
        ("mass", C.c_double), ("comx", C.c_double), ("comy", C.c_double),
        ("bx", C.c_double), ("by", C.c_double), ("bsize", C.c_double),
        ("single", C.c_int64), ("is_leaf", C.c_int32), ("all_small", C.c_int32),
    ]


NODE_DTYPE = np.dtype([
    ("mass", "<f8"), ("comx", "<f8"), ("comy", "<f8"), ("bx", "<f8"), ("by", "<f8"), ("bsize", "<f8"),
    ("single", "<i8"), ("is_leaf", "<i4"), ("all_small", "<i4"),
])


class Stats(C.Structure):
    _fields_ = [
        ("pool_nodes", C.c_uint64), ("nonempty_nodes", C.c_uint64), ("internal_nodes", C.c_uint64),
        ("accepted", C.c_uint64), ("visited", C.c_uint64), ("max_depth", C.c_int32),
        ("pool_overflow", C.c_int32), ("build_seconds", C.c_double), ("force_seconds", C.c_double),
        ("total_seconds", C.c_double), ("force_max", C.c_double), ("force_sum", C.c_double), ("force_count", C.c_uint64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


G_REAL = 6.674e-11  # SimulatorConstants::RealG, reference src/core/constants.cpp:8


def make_params(U, eps, theta=0.5, dt_kick=1.0 / 120, dt_drift=None, thr=0.0, run_movement=True, quirk=True,
                G=G_REAL):
    """dt_kick = SecondsPerTick*baseTimeAcceleration*timeScale; dt_drift = SecondsPerTick*TimeAcceleration."""
    if dt_drift is None:
        dt_drift = dt_kick
    p = Params()
    p.universe_size = U
    p.softening = eps
    p.seconds_per_tick = dt_kick
    p.time_acceleration = dt_drift / dt_kick if dt_kick != 0 else 0.0
    p.base_time_acceleration = 1.0
    p.time_scale = 1.0
    p.theta = theta
    p.small_mass_threshold = thr
    p.G = G
    p.run_movement = 1 if run_movement else 0
    p.quirk = 1 if quirk else 0
    return p


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def build_port(force=False):
    if force or not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(
            os.path.join(HERE, "bh_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return PORT_SO


def build_ref():
    """Builds oracle/_ref when /root/reference is present; otherwise keeps whatever was prebuilt."""
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    return REF_SO if os.path.exists(REF_SO) else None


class BoundaryParams(C.Structure):
    """orc_boundary_params (oracle_abi.h): BoundarySystem's configuration, margin in metres."""
    _fields_ = [("universe_size", C.c_double), ("margin", C.c_double), ("bounce_damping", C.c_double),
                ("max_speed", C.c_double)]


def _boundary(fn, U, margin, damping, max_speed, x, y, vx, vy, comp):
    out = [np.array(a, dtype=np.float64, copy=True) for a in (x, y, vx, vy)]
    comp = None if comp is None else np.ascontiguousarray(comp, dtype=np.uint8)
    bp = BoundaryParams(U, margin, damping, max_speed)
    rc = fn(C.byref(bp), C.c_uint64(len(out[0])), *[_ptr(a, C.c_double) for a in out],
            None if comp is None else _ptr(comp, C.c_uint8))
    if rc:
        raise RuntimeError("boundary failed")
    return dict(x=out[0], y=out[1], vx=out[2], vy=out[3])


class PortLib:
    kind = "port"

    def __init__(self):
        build_port()
        self.lib = C.CDLL(PORT_SO)
        self.lib.orc_bh_describe.restype = C.c_char_p

    def describe(self):
        return self.lib.orc_bh_describe().decode()

    def run(self, p, x, y, vx, vy, m, comp=None, rank=None, nsteps=1, threads=0, per_body=False):
        x, y, vx, vy, m = map(_f64, (x, y, vx, vy, m))
        n = len(x)
        comp = None if comp is None else np.ascontiguousarray(comp, dtype=np.uint8)
        rank = None if rank is None else np.ascontiguousarray(rank, dtype=np.uint32)
        ox, oy, ovx, ovy = (np.empty(n) for _ in range(4))
        acc = np.zeros(n, np.uint32) if per_body else None
        vis = np.zeros(n, np.uint32) if per_body else None
        st = Stats()
        rc = self.lib.orc_bh_run(
            C.byref(p), C.c_uint64(n), _ptr(x, C.c_double), _ptr(y, C.c_double), _ptr(vx, C.c_double),
            _ptr(vy, C.c_double), _ptr(m, C.c_double), _ptr(comp, C.c_uint8), _ptr(rank, C.c_uint32),
            C.c_int(nsteps), C.c_int(threads), _ptr(ox, C.c_double), _ptr(oy, C.c_double),
            _ptr(ovx, C.c_double), _ptr(ovy, C.c_double), _ptr(acc, C.c_uint32), _ptr(vis, C.c_uint32),
            C.byref(st))
        if rc:
            raise RuntimeError(f"orc_bh_run failed rc={rc}")
        out = dict(x=ox, y=oy, vx=ovx, vy=ovy, stats=st.as_dict())
        if per_body:
            out["accepted"] = acc
            out["visited"] = vis
        return out

    def boundary(self, U, x, y, vx, vy, comp=None, margin=15.0, damping=0.7, max_speed=1.0):
        """BoundarySystem::update restated (bh_oracle.c: orc_boundary)."""
        return _boundary(self.lib.orc_boundary, U, margin, damping, max_speed, x, y, vx, vy, comp)

    def tree(self, p, x, y, m, comp=None, rank=None):
        x, y, m = map(_f64, (x, y, m))
        n = len(x)
        comp = None if comp is None else np.ascontiguousarray(comp, dtype=np.uint8)
        rank = None if rank is None else np.ascontiguousarray(rank, dtype=np.uint32)
        cap = 4 * n + 16
        out = np.zeros(cap, NODE_DTYPE)
        cnt = C.c_uint64(0)
        st = Stats()
        rc = self.lib.orc_bh_tree(C.byref(p), C.c_uint64(n), _ptr(x, C.c_double), _ptr(y, C.c_double),
                                  _ptr(m, C.c_double), _ptr(comp, C.c_uint8), _ptr(rank, C.c_uint32),
                                  out.ctypes.data_as(C.c_void_p), C.c_uint64(cap), C.byref(cnt), C.byref(st))
        if rc:
            raise RuntimeError(f"orc_bh_tree failed rc={rc}")
        assert cnt.value <= cap
        return out[:cnt.value], st.as_dict()

    def direct(self, p, x, y, m, comp=None, first=0, count=None, threads=0):
        x, y, m = map(_f64, (x, y, m))
        n = len(x)
        count = n - first if count is None else count
        comp = None if comp is None else np.ascontiguousarray(comp, dtype=np.uint8)
        ax, ay = np.empty(count), np.empty(count)
        rc = self.lib.orc_direct_accel(C.byref(p), C.c_uint64(n), _ptr(x, C.c_double), _ptr(y, C.c_double),
                                       _ptr(m, C.c_double), _ptr(comp, C.c_uint8), C.c_uint64(first),
                                       C.c_uint64(count), C.c_int(threads), _ptr(ax, C.c_double),
                                       _ptr(ay, C.c_double))
        if rc:
            raise RuntimeError("orc_direct_accel failed")
        return ax, ay


class RefLib:
    kind = "reference"

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        self.lib = C.CDLL(REF_SO, mode=os.RTLD_LOCAL)
        self.lib.ref_bh_describe.restype = C.c_char_p

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def describe(self):
        return self.lib.ref_bh_describe().decode()

    def run(self, p, x, y, vx, vy, m, comp=None, nsteps=1, pool_nodes=0):
        x, y, vx, vy, m = map(_f64, (x, y, vx, vy, m))
        n = len(x)
        comp = None if comp is None else np.ascontiguousarray(comp, dtype=np.uint8)
        ox, oy, ovx, ovy = (np.empty(n) for _ in range(4))
        st = Stats()
        rc = self.lib.ref_bh_run(
            C.byref(p), C.c_uint64(n), _ptr(x, C.c_double), _ptr(y, C.c_double), _ptr(vx, C.c_double),
            _ptr(vy, C.c_double), _ptr(m, C.c_double), _ptr(comp, C.c_uint8), C.c_int(nsteps),
            C.c_uint64(pool_nodes), _ptr(ox, C.c_double), _ptr(oy, C.c_double), _ptr(ovx, C.c_double),
            _ptr(ovy, C.c_double), C.byref(st))
        if rc:
            raise RuntimeError(f"ref_bh_run failed rc={rc} (2 = node pool grew: reference defect D1)")
        return dict(x=ox, y=oy, vx=ovx, vy=ovy, stats=st.as_dict())

    def boundary(self, U, x, y, vx, vy, comp=None, margin=15.0, damping=0.7, max_speed=1.0):
        """The reference's own BoundarySystem::update (ref_harness.cpp: ref_boundary)."""
        return _boundary(self.lib.ref_boundary, U, margin, damping, max_speed, x, y, vx, vy, comp)

    def view_rank(self, n, comp=None):
        comp = None if comp is None else np.ascontiguousarray(comp, dtype=np.uint8)
        rank = np.empty(n, np.uint32)
        self.lib.ref_bh_view_rank(C.c_uint64(n), _ptr(comp, C.c_uint8), _ptr(rank, C.c_uint32))
        return rank

    def tree(self, p, x, y, m, comp=None, pool_nodes=0):
        x, y, m = map(_f64, (x, y, m))
        n = len(x)
        comp = None if comp is None else np.ascontiguousarray(comp, dtype=np.uint8)
        cap = 4 * n + 16
        out = np.zeros(cap, NODE_DTYPE)
        cnt = C.c_uint64(0)
        st = Stats()
        rc = self.lib.ref_bh_tree(C.byref(p), C.c_uint64(n), _ptr(x, C.c_double), _ptr(y, C.c_double),
                                  _ptr(m, C.c_double), _ptr(comp, C.c_uint8), C.c_uint64(pool_nodes),
                                  out.ctypes.data_as(C.c_void_p), C.c_uint64(cap), C.byref(cnt), C.byref(st))
        if rc:
            raise RuntimeError(f"ref_bh_tree failed rc={rc}")
        assert cnt.value <= cap
        return out[:cnt.value], st.as_dict()
