"""ctypes binding of liblpe_bh.so (include/lpe_bh.h) — the reference-facing host API in Python form.

Thin by design: every method is one C-ABI call. There is no CPU fallback; constructing a
`BarnesHut` without a CUDA device raises. Used by tests/, bench.py and __graft_entry__.py.

Reference interface mirrored (names and meaning of the knobs):
  Systems::BarnesHutConfig{theta, smallMassThreshold}    include/systems/barnes_hut.hpp:31-46
  SharedSystemConfig{UniverseSizeMeters, GravitationalSoftener, SecondsPerTick, TimeAcceleration}
                                                         include/systems/shared_system_config.hpp:10-21
  Components::SimulatorState{baseTimeAcceleration, timeScale}   include/entities/sim_components.hpp:4-11
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# LPE_BH_LIB selects another build of the same library (tests: liblpe_bh_checked.so, the kernels' bounds checks on)
LIB_PATH = os.environ.get("LPE_BH_LIB") or os.path.join(HERE, "liblpe_bh.so")

HAS_MASS, HAS_VELOCITY, BOUNDARY, LIQUID, ASLEEP = 1, 2, 4, 8, 16
PREC_FAST, PREC_STRICT = 0, 1
KEYS_AUTO, KEYS_MORTON, KEYS_HILBERT = 0, 1, 2
SHARD_BLOCK = 2048
G_REAL = 6.674e-11  # SimulatorConstants::RealG, reference src/core/constants.cpp:8

WORKLOADS = {"disk": 0, "plummer": 1, "two_galaxies": 2, "keplerian": 3, "keplerian_counter": 4}


class Params(C.Structure):
    _fields_ = [
        ("universe_size", C.c_double), ("softening", C.c_double), ("theta", C.c_double),
        ("small_mass_threshold", C.c_double), ("G", C.c_double), ("dt_kick", C.c_double),
        ("dt_drift", C.c_double), ("quirk_mode", C.c_int32), ("precision", C.c_int32),
        ("do_drift", C.c_int32), ("max_depth", C.c_int32), ("key_order", C.c_int32), ("reserved", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_bodies", C.c_uint64), ("n_in_tree", C.c_uint64), ("n_terminals", C.c_uint64),
        ("n_nodes", C.c_uint64), ("interactions", C.c_uint64), ("visits", C.c_uint64), ("warp_visits", C.c_uint64), ("overflow_chunks", C.c_uint64), ("t2_kinds", C.c_uint64 * 8),
        ("depth", C.c_int32), ("sort_passes", C.c_int32), ("hilbert", C.c_int32), ("pad_", C.c_int32),
        ("ms_keygen", C.c_float),
        ("ms_sort", C.c_float), ("ms_build", C.c_float), ("ms_traverse", C.c_float), ("ms_total", C.c_float),
        ("pad2_", C.c_float), ("force_max", C.c_double), ("force_sum", C.c_double),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["t2_kinds"] = list(d["t2_kinds"])
        return d


class BoundaryParams(C.Structure):
    """lpe_bh_boundary_params: Systems::BoundaryConfig (include/systems/boundary.hpp:27-36), margin in metres."""
    _fields_ = [("universe_size", C.c_double), ("margin", C.c_double), ("bounce_damping", C.c_double),
                ("max_speed", C.c_double)]


class TreeDump(C.Structure):
    _fields_ = [
        ("sorted_keys", C.c_void_p), ("sorted_index", C.c_void_p), ("node_level", C.c_void_p),
        ("node_key", C.c_void_p), ("node_skip", C.c_void_p), ("node_first", C.c_void_p),
        ("node_count", C.c_void_p), ("node_mass", C.c_void_p), ("node_comx", C.c_void_p),
        ("node_comy", C.c_void_p),
    ]


class DeviceView(C.Structure):
    _fields_ = [
        ("body", C.c_void_p), ("vel", C.c_void_p), ("orig", C.c_void_p), ("xchg_send", C.c_void_p),
        ("xchg_recv", C.c_void_p), ("n", C.c_uint64), ("xchg_chunk", C.c_uint64), ("key_ordered", C.c_int32),
        ("pad_", C.c_int32),
    ]


class DDStats(C.Structure):
    """lpe_bh_dd_stats"""
    _fields_ = [
        ("capacity", C.c_uint64), ("n_live", C.c_uint64), ("n_in_tree", C.c_uint64), ("n_terminals", C.c_uint64),
        ("n_cells", C.c_uint64), ("n_roots", C.c_uint64), ("exported_blocks", C.c_uint64 * 8),
        ("interactions", C.c_uint64), ("work_cost", C.c_uint64), ("overflow_chunks", C.c_uint64),
        ("import_blocks", C.c_uint32), ("fault", C.c_uint32), ("rank", C.c_int32), ("nranks", C.c_int32),
        ("depth", C.c_int32), ("export_rounds", C.c_int32),
        ("ms_keygen", C.c_float), ("ms_wait_a", C.c_float), ("ms_sort", C.c_float), ("ms_build", C.c_float),
        ("ms_export", C.c_float), ("ms_wait_b", C.c_float), ("ms_top", C.c_float), ("ms_traverse", C.c_float),
        ("ms_total", C.c_float), ("pad2_", C.c_float),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["exported_blocks"] = list(d["exported_blocks"])
        return d


def make_params(U, eps, theta=0.5, dt_kick=1.0 / 120, dt_drift=None, thr=0.0, quirk=True, precision=PREC_FAST,
                do_drift=True, max_depth=0, G=G_REAL, key_order=KEYS_AUTO):
    p = Params()
    p.universe_size = U
    p.softening = eps
    p.theta = theta
    p.small_mass_threshold = thr
    p.G = G
    p.dt_kick = dt_kick
    p.dt_drift = dt_kick if dt_drift is None else dt_drift
    p.quirk_mode = 1 if quirk else 0
    p.precision = precision
    p.do_drift = 1 if do_drift else 0
    p.max_depth = max_depth
    p.key_order = key_order
    p.reserved = 0
    return p


def pinned_array(n, dtype=np.float64):
    """A page-locked host array (lpe_bh_alloc_pinned): asynchronous copies at full PCIe rate, and host ticks on such arrays are
    replayed as one CUDA graph. Keep the array referenced while in use; it is freed with the process."""
    lib = load_library()
    nbytes = int(n) * np.dtype(dtype).itemsize
    p = lib.lpe_bh_alloc_pinned(C.c_uint64(max(nbytes, 8)))
    if not p:
        raise MemoryError("lpe_bh_alloc_pinned failed")
    buf = (C.c_char * max(nbytes, 8)).from_address(p)
    return np.frombuffer(buf, dtype=dtype, count=int(n))


def build_library(force=False):
    """Compile liblpe_bh.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))]
    srcs.append(os.path.join(HERE, "..", "include", "lpe_bh.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    subprocess.check_call(["bash", os.path.join(HERE, "build.sh")])
    return LIB_PATH


_lib = None


def load_library():
    """Load the CUDA library. Fails loudly when it has not been built: there is nothing to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run little-physics-engine_b200/build.sh (no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    lib.lpe_bh_version.restype = C.c_char_p
    lib.lpe_bh_last_error.restype = C.c_char_p
    lib.lpe_bh_last_error.argtypes = [C.c_void_p]
    lib.lpe_bh_launch_count.restype = C.c_uint64
    lib.lpe_bh_launch_count.argtypes = [C.c_void_p]
    lib.lpe_bh_graph_replays.restype = C.c_uint64
    lib.lpe_bh_graph_replays.argtypes = [C.c_void_p]
    lib.lpe_bh_alloc_pinned.restype = C.c_void_p
    lib.lpe_bh_alloc_pinned.argtypes = [C.c_uint64]
    lib.lpe_bh_free_pinned.argtypes = [C.c_void_p]
    lib.lpe_bh_shard_chunk.restype = C.c_uint64
    lib.lpe_bh_shard_chunk.argtypes = [C.c_uint64, C.c_int]
    lib.lpe_bh_dd_window.restype = C.c_void_p
    lib.lpe_bh_dd_window.argtypes = [C.c_void_p]
    lib.lpe_bh_cell_key.restype = C.c_uint64
    lib.lpe_bh_cell_key.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int]
    lib.lpe_bh_key_cell.restype = None
    lib.lpe_bh_key_cell.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def workload(kind, n, seed, U):
    """Deterministic synthetic bodies; returns x, y, vx, vy, m."""
    lib = load_library()
    arrs = [np.empty(n, np.float64) for _ in range(5)]
    rc = lib.lpe_bh_workload(C.c_int(WORKLOADS[kind] if isinstance(kind, str) else kind), C.c_uint64(n),
                             C.c_uint64(seed), C.c_double(U), *[_dp(a) for a in arrs])
    if rc:
        raise RuntimeError("lpe_bh_workload failed")
    return tuple(arrs)


def shard_chunk(n, nranks):
    return int(load_library().lpe_bh_shard_chunk(C.c_uint64(n), C.c_int(nranks)))


def shard_owner(pos, nranks):
    r, s = C.c_int(0), C.c_uint64(0)
    load_library().lpe_bh_shard_owner(C.c_uint64(pos), C.c_int(nranks), C.byref(r), C.byref(s))
    return r.value, s.value


def cell_key(ix, iy, level, hilbert=True):
    """Sort key of cell (ix, iy) of a quadtree level (host helper, no GPU)."""
    return int(load_library().lpe_bh_cell_key(ix, iy, level, 1 if hilbert else 0))


def key_cell(key, level, hilbert=True):
    x, y = C.c_uint32(0), C.c_uint32(0)
    load_library().lpe_bh_key_cell(C.c_uint64(key), level, 1 if hilbert else 0, C.byref(x), C.byref(y))
    return x.value, y.value


class BarnesHut:
    """One device context = one Systems::BarnesHutSystem instance with device-resident bodies."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.lpe_bh_create(C.c_int(device), C.byref(self.h))
        if rc:
            raise RuntimeError("lpe_bh_create: " + self.lib.lpe_bh_last_error(None).decode())
        self.n = 0
        self._rank, self._nranks = 0, 1

    def close(self):
        if self.h:
            self.lib.lpe_bh_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc:
            raise RuntimeError(f"{what}: {self.lib.lpe_bh_last_error(self.h).decode()}")

    def set_stream(self, cuda_stream_ptr):
        self._chk(self.lib.lpe_bh_set_stream(self.h, C.c_void_p(cuda_stream_ptr)), "set_stream")

    def set_instrumentation(self, timing=False, counts=False, force_dfs=False, force_overflow=False, plain_launches=False,
                            warp_only=False):
        self._chk(self.lib.lpe_bh_set_instrumentation(
            self.h, C.c_int((1 if timing else 0) | (2 if counts else 0) | (4 if force_dfs else 0) |
                            (8 if force_overflow else 0) | (16 if plain_launches else 0) | (32 if warp_only else 0))),
                  "set_instrumentation")

    def upload(self, x, y, vx, vy, m, rank=None, comp=None):
        x, y, vx, vy, m = map(_f64, (x, y, vx, vy, m))
        rank = None if rank is None else np.ascontiguousarray(rank, dtype=np.uint32)
        comp = None if comp is None else np.ascontiguousarray(comp, dtype=np.uint8)
        self.n = len(x)
        self._keep = (x, y, vx, vy, m, rank, comp)
        self._chk(self.lib.lpe_bh_upload(self.h, C.c_uint64(self.n), _dp(x), _dp(y), _dp(vx), _dp(vy), _dp(m),
                                         _dp(rank), _dp(comp)), "upload")

    def generate(self, kind, n, seed, U):
        """Make the bodies on the device (lpe_bh_generate): kind "keplerian_counter"."""
        self.n = n
        self._chk(self.lib.lpe_bh_generate(self.h, C.c_int(WORKLOADS[kind] if isinstance(kind, str) else kind), C.c_uint64(n),
                                           C.c_uint64(seed), C.c_double(U)), "generate")

    def upload_positions(self, x, y):
        """Replace the positions of the resident bodies (creation order), e.g. after a host-side MovementSystem."""
        x, y = _f64(x), _f64(y)
        assert len(x) == self.n and len(y) == self.n
        self._chk(self.lib.lpe_bh_upload_positions(self.h, _dp(x), _dp(y)), "upload_positions")

    def upload_velocities(self, vx, vy):
        vx, vy = _f64(vx), _f64(vy)
        assert len(vx) == self.n and len(vy) == self.n
        self._chk(self.lib.lpe_bh_upload_velocities(self.h, _dp(vx), _dp(vy)), "upload_velocities")

    def upload_ptrs(self, n, x, y, vx, vy, m):
        """Raw host pointers (e.g. pinned torch tensors' data_ptr())."""
        self.n = n
        self._chk(self.lib.lpe_bh_upload(self.h, C.c_uint64(n), C.c_void_p(x), C.c_void_p(y), C.c_void_p(vx),
                                         C.c_void_p(vy), C.c_void_p(m), None, None), "upload")

    def boundary(self, universe_size, margin=15.0, bounce_damping=0.7, max_speed=1.0):
        """Systems::BoundarySystem::update as a device pass over the resident bodies (reference boundary.cpp:13-69)."""
        bp = BoundaryParams(universe_size, margin, bounce_damping, max_speed)
        self._chk(self.lib.lpe_bh_boundary(self.h, C.byref(bp)), "boundary")

    # ---- direct exchange over peer memory (include/lpe_bh.h) ----
    def xchg_export(self):
        buf = C.create_string_buffer(64)
        self._chk(self.lib.lpe_bh_xchg_export(self.h, buf), "xchg_export")
        return buf.raw

    def xchg_import(self, rank, handle):
        self._chk(self.lib.lpe_bh_xchg_import(self.h, C.c_int(rank), C.c_char_p(handle)), "xchg_import")

    def xchg_set_peer(self, rank, recv_ptr):
        self._chk(self.lib.lpe_bh_xchg_set_peer(self.h, C.c_int(rank), C.c_void_p(recv_ptr)), "xchg_set_peer")

    def xchg_p2p_ready(self):
        return bool(self.lib.lpe_bh_xchg_p2p_ready(self.h))

    def xchg_reset(self):
        self._chk(self.lib.lpe_bh_xchg_reset(self.h), "xchg_reset")

    def update_host_ptrs(self, params, n, x, y, vx, vy, m, rank=None, comp=None):
        """lpe_bh_update_host on raw host pointers (pinned memory makes the uploads overlap the step)."""
        self.n = n
        vp = lambda a: None if a is None else C.c_void_p(a)
        self._chk(self.lib.lpe_bh_update_host(self.h, C.byref(params), C.c_uint64(n), vp(x), vp(y), vp(vx), vp(vy),
                                              vp(m), vp(rank), vp(comp)), "update_host")

    def step(self, params, nsteps=1):
        self._chk(self.lib.lpe_bh_step(self.h, C.byref(params), C.c_int(nsteps)), "step")

    def synchronize(self):
        self._chk(self.lib.lpe_bh_synchronize(self.h), "synchronize")

    def synchronize_quiet(self):
        if self.h:
            self.lib.lpe_bh_synchronize(self.h)

    def download(self):
        out = [np.empty(self.n, np.float64) for _ in range(4)]
        self._chk(self.lib.lpe_bh_download(self.h, *[_dp(a) for a in out]), "download")
        return dict(x=out[0], y=out[1], vx=out[2], vy=out[3])

    def download_ptrs(self, x, y, vx, vy):
        self._chk(self.lib.lpe_bh_download(self.h, C.c_void_p(x), C.c_void_p(y), C.c_void_p(vx), C.c_void_p(vy)),
                  "download")

    def update_host(self, params, x, y, vx, vy, m, rank=None, comp=None):
        """The drop-in call: upload, one step, download; arrays updated in place (must be float64 contiguous)."""
        for a in (x, y, vx, vy, m):
            assert a.dtype == np.float64 and a.flags.c_contiguous
        rank = None if rank is None else np.ascontiguousarray(rank, dtype=np.uint32)
        comp = None if comp is None else np.ascontiguousarray(comp, dtype=np.uint8)
        self.n = len(x)
        self._chk(self.lib.lpe_bh_update_host(self.h, C.byref(params), C.c_uint64(self.n), _dp(x), _dp(y), _dp(vx),
                                              _dp(vy), _dp(m), _dp(rank), _dp(comp)), "update_host")

    def stats(self):
        s = Stats()
        self._chk(self.lib.lpe_bh_get_stats(self.h, C.byref(s)), "get_stats")
        return s.as_dict()

    def dump_tree(self):
        st = self.stats()
        n, nn = st["n_bodies"], st["n_nodes"]
        out = dict(
            sorted_keys=np.zeros(n, np.uint64), sorted_index=np.zeros(n, np.uint32),
            node_level=np.zeros(nn, np.int32), node_key=np.zeros(nn, np.uint64), node_skip=np.zeros(nn, np.uint32),
            node_first=np.zeros(nn, np.uint32), node_count=np.zeros(nn, np.uint32), node_mass=np.zeros(nn),
            node_comx=np.zeros(nn), node_comy=np.zeros(nn))
        d = TreeDump(**{k: v.ctypes.data for k, v in out.items()})
        self._chk(self.lib.lpe_bh_dump_tree(self.h, C.byref(d)), "dump_tree")
        out["stats"] = st
        return out

    def counts(self):
        acc, vis = np.zeros(self.n, np.uint32), np.zeros(self.n, np.uint32)
        self._chk(self.lib.lpe_bh_get_counts(self.h, _dp(acc), _dp(vis)), "get_counts")
        return acc, vis

    def direct_accel(self, params, first=0, count=None):
        count = self.n - first if count is None else count
        ax, ay = np.empty(count), np.empty(count)
        self._chk(self.lib.lpe_bh_direct_accel(self.h, C.byref(params), C.c_uint64(first), C.c_uint64(count), _dp(ax),
                                               _dp(ay)), "direct_accel")
        return ax, ay

    def max_source_mass(self):
        v = C.c_double(0.0)
        self._chk(self.lib.lpe_bh_max_source_mass(self.h, C.byref(v)), "max_source_mass")
        return v.value

    def launch_count(self):
        return int(self.lib.lpe_bh_launch_count(self.h))

    def graph_replays(self):
        """Steps / ticks so far that were replays of a captured CUDA graph."""
        return int(self.lib.lpe_bh_graph_replays(self.h))

    def fma_peak_tflops(self):
        t = C.c_double(0.0)
        self._chk(self.lib.lpe_bh_fma_peak(self.h, C.byref(t)), "fma_peak")
        return t.value

    # ---- multi-GPU ----
    def set_shard(self, rank, nranks):
        self._chk(self.lib.lpe_bh_set_shard(self.h, C.c_int(rank), C.c_int(nranks)), "set_shard")
        self._rank, self._nranks = rank, nranks

    def step_begin(self, params):
        self._chk(self.lib.lpe_bh_step_begin(self.h, C.byref(params)), "step_begin")

    def step_finish(self):
        self._chk(self.lib.lpe_bh_step_finish(self.h), "step_finish")

    def xchg_read_send(self):
        buf = np.empty(4 * shard_chunk(self.n, self._nranks), np.float64)
        self._chk(self.lib.lpe_bh_xchg_read_send(self.h, _dp(buf)), "xchg_read_send")
        return buf

    def xchg_write_recv(self, src_rank, buf):
        buf = np.ascontiguousarray(buf, dtype=np.float64)
        self._chk(self.lib.lpe_bh_xchg_write_recv(self.h, C.c_int(src_rank), _dp(buf)), "xchg_write_recv")

    def device_view(self):
        v = DeviceView()
        self._chk(self.lib.lpe_bh_get_device_view(self.h, C.byref(v)), "get_device_view")
        return v

    # ---- multi-GPU, domain-decomposed (include/lpe_bh.h) ----
    def dd_init(self, rank, nranks, capacity, import_blocks=0):
        self._chk(self.lib.lpe_bh_dd_init(self.h, C.c_int(rank), C.c_int(nranks), C.c_uint64(capacity),
                                          C.c_uint32(import_blocks)), "dd_init")
        self._rank, self._nranks, self._capacity = rank, nranks, max(int(capacity), 1024)

    def dd_window(self):
        return self.lib.lpe_bh_dd_window(self.h)

    def dd_set_peer(self, rank, window, peer_device=-1):
        self._chk(self.lib.lpe_bh_dd_set_peer(self.h, C.c_int(rank), C.c_void_p(window), C.c_int(peer_device)), "dd_set_peer")

    def dd_export(self):
        buf = C.create_string_buffer(64)
        self._chk(self.lib.lpe_bh_dd_export(self.h, buf), "dd_export")
        return buf.raw

    def dd_import(self, rank, handle):
        self._chk(self.lib.lpe_bh_dd_import(self.h, C.c_int(rank), C.c_char_p(handle)), "dd_import")

    def dd_ready(self):
        return bool(self.lib.lpe_bh_dd_ready(self.h))

    def dd_upload(self, params, x, y, vx, vy, m, rank=None, comp=None):
        """Every rank is given the whole input and keeps the bodies of its own key range."""
        x, y, vx, vy, m = map(_f64, (x, y, vx, vy, m))
        rank = None if rank is None else np.ascontiguousarray(rank, dtype=np.uint32)
        comp = None if comp is None else np.ascontiguousarray(comp, dtype=np.uint8)
        self.n = len(x)
        self._chk(self.lib.lpe_bh_dd_upload(self.h, C.byref(params), C.c_uint64(self.n), _dp(x), _dp(y), _dp(vx), _dp(vy),
                                            _dp(m), _dp(rank), _dp(comp)), "dd_upload")

    def dd_phase(self, params, phase):
        self._chk(self.lib.lpe_bh_dd_phase(self.h, C.byref(params), C.c_int(phase)), "dd_phase")

    def dd_step(self, params, nsteps=1):
        self._chk(self.lib.lpe_bh_dd_step(self.h, C.byref(params), C.c_int(nsteps)), "dd_step")

    def dd_download(self, counts=False):
        cap = self._capacity
        idx = np.empty(cap, np.uint32)
        out = [np.empty(cap, np.float64) for _ in range(4)]
        acc = np.empty(cap, np.uint32) if counts else None
        n = C.c_uint64(0)
        self._chk(self.lib.lpe_bh_dd_download(self.h, C.byref(n), _dp(idx), *[_dp(a) for a in out], _dp(acc)), "dd_download")
        k = n.value
        d = dict(index=idx[:k], x=out[0][:k], y=out[1][:k], vx=out[2][:k], vy=out[3][:k])
        if counts:
            d["accepted"] = acc[:k]
        return d

    def dd_stats(self):
        s = DDStats()
        self._chk(self.lib.lpe_bh_dd_get_stats(self.h, C.byref(s)), "dd_get_stats")
        return s.as_dict()

    def dd_chunk_costs(self):
        cap = self._capacity // 32 + 2
        keys, cost = np.empty(cap, np.uint64), np.empty(cap, np.uint32)
        n = C.c_uint64(0)
        self._chk(self.lib.lpe_bh_dd_chunk_costs(self.h, C.byref(n), _dp(keys), _dp(cost)), "dd_chunk_costs")
        return keys[:n.value].copy(), cost[:n.value].copy()

    def dd_get_splitters(self):
        a = (C.c_uint64 * (self._nranks + 1))()
        self._chk(self.lib.lpe_bh_dd_get_splitters(self.h, a), "dd_get_splitters")
        return [int(v) for v in a]

    def dd_set_splitters(self, split30):
        a = (C.c_uint64 * (self._nranks + 1))(*[int(v) for v in split30])
        self._chk(self.lib.lpe_bh_dd_set_splitters(self.h, a), "dd_set_splitters")


class DDGroup:
    """All ranks of a domain-decomposed run driven by ONE host thread (SURVEY.md 8(b) "Threading").

    devices = one CUDA device index per rank. Distinct devices: every rank's phases are queued on its own GPU and the
    in-stream flag barriers make the GPUs wait for each other (peer access between the devices). The same device for
    every rank (tests on a one-GPU box): the phases are run rank after rank with a host synchronisation in between,
    since kernels of one GPU must not wait for each other.
    """

    def __init__(self, devices, capacity, import_blocks=0):
        self.devices = list(devices)
        self.R = len(self.devices)
        self.lockstep = len(set(self.devices)) == self.R
        if not self.lockstep and len(set(self.devices)) != 1:
            raise ValueError("either one device per rank or the same device for all ranks")
        self.ranks = [BarnesHut(d) for d in self.devices]
        for r, c in enumerate(self.ranks):
            c.dd_init(r, self.R, capacity, import_blocks)
        wins = [c.dd_window() for c in self.ranks]
        for c in self.ranks:
            for r, w in enumerate(wins):
                c.dd_set_peer(r, w, self.devices[r] if self.lockstep else -1)
        self.n = 0

    def close(self):
        for c in self.ranks:
            c.synchronize_quiet()
        for c in self.ranks:
            c.close()

    def upload(self, params, x, y, vx, vy, m, rank=None, comp=None):
        self.n = len(x)
        for c in self.ranks:
            c.dd_upload(params, x, y, vx, vy, m, rank=rank, comp=comp)

    def step(self, params, nsteps=1):
        for _ in range(nsteps):
            if self.lockstep:
                for c in self.ranks:
                    c.dd_step(params, 1)
            else:
                for ph in range(3):
                    for c in self.ranks:
                        c.dd_phase(params, ph)
                    for c in self.ranks:
                        c.synchronize()

    def download(self, counts=False):
        """State of all ranks assembled in creation order."""
        out = {k: np.full(self.n, np.nan) for k in ("x", "y", "vx", "vy")}
        if counts:
            out["accepted"] = np.zeros(self.n, np.uint32)
        seen = np.zeros(self.n, np.int32)
        for c in self.ranks:
            d = c.dd_download(counts=counts)
            i = d["index"]
            seen[i] += 1
            for k in out:
                out[k][i] = d[k]
        if not np.all(seen == 1):
            raise RuntimeError(f"domain decomposition lost or duplicated bodies: {np.count_nonzero(seen != 1)} of {self.n}")
        return out

    def stats(self):
        return [c.dd_stats() for c in self.ranks]

    def rebalance(self, beta=240.0):
        """New splitters from the last step's per-chunk traversal cost (+ beta per chunk for sort and build)."""
        new = balanced_splitters([c.dd_chunk_costs() for c in self.ranks], self.R, beta)
        for c in self.ranks:
            c.dd_set_splitters(new)
        return new


def balanced_splitters(per_rank, nranks, beta=240.0, scale=None):
    """Splitters (depth-30 keys) that give every rank the same share of sum(cost + beta) over the 32-body chunks.
    per_rank = [(first_key30, cost)] in rank order: the ranks' key ranges are disjoint and ascending, so the
    concatenation is globally key-ordered. scale = optional per-rank correction factors (measured time of the rank /
    mean over ranks): what the cost model missed on a rank (deeper tree, more cells per body) is charged to its chunks.
    Pure host arithmetic, the same on every rank."""
    keys = np.concatenate([k for k, _ in per_rank])
    scale = [1.0] * len(per_rank) if scale is None else scale
    w = np.concatenate([(c.astype(np.float64) + float(beta)) * float(f) for (_, c), f in zip(per_rank, scale)])
    top = 1 << 60
    split = [0]
    if len(keys):
        cum = np.cumsum(w)
        for r in range(1, nranks):
            i = int(np.searchsorted(cum, cum[-1] * r / nranks))
            k = int(keys[min(i, len(keys) - 1)])
            split.append(max(k, split[-1]))
    else:
        split += [top * r // nranks for r in range(1, nranks)]
    split.append(top)
    return split
