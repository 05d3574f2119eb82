#!/usr/bin/env python
"""bench.py — Barnes-Hut body-steps/s (theta = 0.5) on B200, the metric of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c1|c2|c3|c4|c5] [--precision fast|strict] [--mgpu dd|replicated]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (one rank per GPU)

A "step" is one full pass of the hot path over all bodies: space-filling-curve keys, radix sort, quadtree build +
aggregation, theta traversal, kick and drift. One JSON line is printed by rank 0.

Workload: C3 (16 M-body Plummer sphere) at EVERY N, so that the driver's 1 -> 8 GPU series is one problem (strong
scaling). At N = 1 the line also carries a "c2" object: the 1 M-body disk (BASELINE config 2) measured the same way.

  value      whole-job body-steps/s with the bodies resident in HBM (device-timed, CUDA events on the launching
             stream, max over ranks); L2 is flushed between timed steps.
  e2e        the same metric through the call the ECS drop-in makes (lpe_bh_update_host): pinned host buffers ->
             upload -> step -> download, every step, copies inside the timed region (N = 1).
  roofline   the dominant kernel (k_traverse2): FP32-pipe bound, algorithmic flops = 20 per accepted
             interaction, against the FP32 FMA peak measured on this GPU by lpe_bh_fma_peak.
  roofline_hbm  the HBM-bound phases (keygen + sort + build) against MEASURED_PEAKS.json's copy bandwidth.
  cpu_baseline  the reference's own barnes_hut.cpp + movement.cpp (oracle/_ref, compiled unmodified) on the host.
  parity_check  N > 1: one extra counted step of the decomposed run against the unsharded step on rank 0's GPU.

N > 1 (--mgpu dd, default): domain decomposition — every rank owns a key range, sorts and builds only its bodies,
exchanges the top of the tree and the locally essential cells through NVLink peer memory (include/lpe_bh.h,
lpe_bh_dd_*); --mgpu replicated keeps round 1's replicated-tree variant for comparison.
--impl reference times the reference's CPU code (single-threaded) on bounded samples; that process
never loads the CUDA library.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "little-physics-engine_b200"))

U = float(2 ** 20)       # power-of-two universe, eps = U/2^14, G = RealG, dt_K = dt_D = 1/120
EPS = U / 2 ** 14
THETA = 0.5
DT = 1.0 / 120.0
WORKLOADS = {
    # C1: the reference's own Keplerian-disk scenario configuration (keplerian_disk.cpp:16-29) at 10 k bodies
    "c1": dict(kind="keplerian", n=10_000, seed=5, U=6e9, eps=2e7, thr=1e3, dt_drift=6.756e-3,
               name="C1: 10k-body Keplerian disk (reference scenario config: U=6e9, eps=2e7, threshold 1e3), theta=0.5"),
    "c2": dict(kind="disk", n=1_000_000, seed=42, name="C2: 1M-body uniform disk, theta=0.5, U=2^20, eps=U/2^14"),
    "c3": dict(kind="plummer", n=16_000_000, seed=43, name="C3: 16M-body Plummer sphere, theta=0.5, U=2^20, eps=U/2^14"),
    "c4": dict(kind="two_galaxies", n=4_000_000, seed=44, name="C4: 4M-body two-galaxy collision, theta=0.5, U=2^20, eps=U/2^14"),
    "c5": dict(kind="disk", n=1_000_000, seed=42, name="C5: theta sweep 0.3-1.0 on the 1M-body disk + direct O(N^2) at 256k bodies"),
}
FLOPS_PER_INTERACTION = 20.0
# dram__bytes_read.sum + dram__bytes_write.sum of one k_traverse2 launch, from the committed ncu --set full captures
NCU_TRAFFIC = {"c2": (154.6e6, "profiles/r02_ncu_traverse2_c2.txt (ncu --set full, one launch: dram read 127.9 MB + write 26.7 MB)"),
               "c3": (2043.0e6, "profiles/r02_ncu_traverse2_c3.txt (ncu --set full, one launch: dram read 1298.8 MB + write 744.2 MB)")}
# SURVEY.md §8(d) algorithmic bytes of the HBM phases: state 40 B read + 32 B written, key + index 12 B, 24 B per sort
# pass, sorted gather 24 B, node arrays 85 B  ->  193 + 24 * passes (4 passes for the 33-bit keys of depth 16)
def hbm_bytes_per_body(passes):
    return 193.0 + 24.0 * passes


def wl_params(wl):
    return dict(U=wl.get("U", U), eps=wl.get("eps", EPS), thr=wl.get("thr", 0.0), dt_kick=DT, dt_drift=wl.get("dt_drift", DT))


def config_for(wl, world, precision, mgpu):
    """Identical in both arms (the driver compares the two config objects)."""
    p = wl_params(wl)
    return {"workload": wl["name"], "bodies": wl["n"], "theta": THETA, "softening": p["eps"], "universe": p["U"],
            "quirk_mode": "reference", "precision": precision, "n_gpus": world,
            "l2": "256 MB buffer written between timed steps (L2 flush); per-step CUDA events",
            "parallelism": "single GPU" if world == 1 else f"{world} GPUs, {mgpu}"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.t = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.t.append(time.time())
            self.rows.append([c.strip() for c in line.split(",")])

    def wait_first(self, timeout=5.0):
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sel = [r for r, t in zip(self.rows, self.t) if t0 - 0.03 <= t <= t1 + 0.03] or self.rows
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in sel:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def reference_sample(wl, steps, warmup):
    """The reference's own CPU implementation on this box's host cores, on a bounded sample of the workload.

    Sample: every `sub`-th body of the workload (all of it up to 2 M bodies; about 1 M bodies of C3 / C4, where one CPU
    step of the whole thing would take minutes and a 5 GB node pool); per step the full tree is built over those
    bodies and the force loop runs over every `stride`-th of them — done by giving only those a Velocity component, so
    it is still the unmodified code path (barnes_hut.cpp:89 iterates view<Position,Velocity,Mass>).
    Reported separately: what was MEASURED (seconds per sample step) and what is derived from it (a full step's time
    = build + force time scaled to all targets; the reference gets slower per body as N grows, so running it on a
    subsample flatters the reference, not us)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    import workloads
    p = wl_params(wl)
    n_full = wl["n"]
    x, y, vx, vy, m = workloads.workload(wl["kind"], n_full, wl["seed"], p["U"])
    sub = max(1, n_full // 1_000_000) if n_full > 2_000_000 else 1
    if sub > 1:
        x, y, vx, vy, m = (np.ascontiguousarray(a[::sub]) for a in (x, y, vx, vy, m))
    n = len(x)
    po = O.make_params(p["U"], p["eps"], theta=THETA, thr=p["thr"], dt_kick=p["dt_kick"], dt_drift=p["dt_drift"])
    stride = max(1, n // 50_000)          # ~50k targets per sample: ~1-2 s of force work per step
    comp = np.full(n, O.HAS_MASS, np.uint8)
    comp[::stride] |= O.HAS_VELOCITY
    ntargets = int(np.count_nonzero(comp & O.HAS_VELOCITY))
    if O.RefLib.available():
        lib, kind = O.RefLib(), "reference"
        run = lambda: lib.run(po, x, y, vx, vy, m, comp=comp, nsteps=1, pool_nodes=4 * n + 4096)["stats"]["total_seconds"]
        build = lambda: lib.tree(po, x, y, m, pool_nodes=4 * n + 4096)[1]["build_seconds"]
    else:
        lib, kind = O.PortLib(), "port"
        run = lambda: lib.run(po, x, y, vx, vy, m, comp=comp, nsteps=1, threads=1)["stats"]["total_seconds"]
        build = lambda: lib.tree(po, x, y, m)[1]["build_seconds"]
    t_build = build()
    measured, full = [], []
    for s in range(warmup + steps):
        t = run()
        if s >= warmup:
            measured.append(t)
            full.append(t_build + max(t - t_build, 0.0) * (n / ntargets))
    t_meas, t_full = float(np.mean(measured)), float(np.mean(full))
    value = n / t_full
    return {
        "value": value, "unit": "body-steps/s", "cores": 1, "kind": kind, "host_cores_available": os.cpu_count(),
        "measured_seconds_per_sample_step": t_meas, "build_seconds": t_build, "sample_bodies": n, "sample_targets": ntargets,
        "extrapolation": f"full step of the {n}-body sample = build + (sample step - build) x {n / ntargets:.1f} "
                         f"= {t_full:.3f} s; value = {n} bodies / that",
        "sample": (f"every {sub}th body of the workload ({n} bodies): per step a full tree build over them + the force "
                   f"loop over every {stride}th ({ntargets} targets); {lib.describe()}"),
    }, t_meas


def reference_arm(args, wl, rank, world):
    if rank != 0:
        return
    base, t_meas = reference_sample(wl, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "barnes_hut_body_steps_per_sec", "value": base["value"], "unit": "body-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_meas,       # MEASURED: one bounded sample step (see cpu_baseline.extrapolation for `value`)
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_for(wl, max(args.gpus, 1), args.precision, args.mgpu),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------ our arm
class CudaArray:
    """Expose a raw device pointer of the library to torch (zero copy) for the NCCL allgather."""

    def __init__(self, ptr, nelem):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def log(rank, msg):
    print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)


def make_gpu_params(lpe_bh, wl, precision="fast", theta=THETA, quirk=True):
    p = wl_params(wl)
    return lpe_bh.make_params(p["U"], p["eps"], theta=theta, thr=p["thr"], dt_kick=p["dt_kick"], dt_drift=p["dt_drift"],
                              quirk=quirk, precision=lpe_bh.PREC_STRICT if precision == "strict" else lpe_bh.PREC_FAST)


def measure_single(torch, lpe_bh, bh, stream, wl, bodies, params, steps, warmup, flush, precision):
    """Device-resident single-GPU steps: per-step CUDA events, L2 flushed in between. Returns a dict of measurements."""
    x, y, vx, vy, m = bodies
    n = len(x)
    bh.set_instrumentation(counts=True)
    bh.upload(x, y, vx, vy, m)
    bh.step(params, 1)
    st = bh.stats()
    interactions, passes = st["interactions"], st["sort_passes"]
    # phase table: a few steps with the library's own events between the phases (plain launches)
    bh.set_instrumentation(timing=True)
    bh.upload(x, y, vx, vy, m)
    phases = []
    for it in range(3 + min(steps, 5)):
        flush.zero_()
        bh.step(params, 1)
        if it >= 3:
            s = bh.stats()
            phases.append((s["ms_keygen"], s["ms_sort"], s["ms_build"], s["ms_traverse"]))
    # timed region: the step as a user runs it — no instrumentation, so the library replays its captured CUDA graph of the
    # step (captured the second time a step with the same parameters / buffers comes up: 5 steps settle that)
    bh.set_instrumentation()
    bh.upload(x, y, vx, vy, m)
    for _ in range(max(warmup, 6)):
        bh.step(params, 1)
    torch.cuda.synchronize()
    launches0 = bh.launch_count()
    step_ms = []
    for _ in range(steps):
        flush.zero_()                       # evict L2 between timed steps (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        bh.step(params, 1)
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    launches = bh.launch_count() - launches0
    ph = np.mean(np.array(phases), axis=0)
    ms = float(np.mean(step_ms))
    return dict(n=n, ms_per_step=ms, value=n / (ms * 1e-3), interactions=interactions, passes=passes,
                phases={"keygen": float(ph[0]), "sort": float(ph[1]), "build": float(ph[2]), "traverse": float(ph[3])},
                launches=int(launches), kernel="k_traverse2 (two-phase traversal)" if precision == "fast" else "k_traverse<STRICT> (depth-first, fp64)")


def measure_e2e(torch, bh, bodies, params, steps):
    """Through host buffers, as Systems::BarnesHutSystem::update pays it: lpe_bh_update_host on pinned SoA arrays."""
    x, y, vx, vy, m = bodies
    n = len(x)
    hx, hy, hvx, hvy, hm = (torch.from_numpy(a.copy()).pin_memory() for a in (x, y, vx, vy, m))
    ptrs = [t.data_ptr() for t in (hx, hy, hvx, hvy, hm)]
    e2e_steps = max(3, min(steps, 20))
    bh.set_instrumentation()              # no phase events in the host-clocked loop
    for it in range(5 + e2e_steps):         # (5 warm-up ticks: the tick's CUDA graph is captured on the 3rd / 4th)
        if it == 5:
            torch.cuda.synchronize()
            te0 = time.perf_counter()
        bh.update_host_ptrs(params, n, *ptrs)   # synchronises; the result lands in the pinned host arrays
    e2e_ms = (time.perf_counter() - te0) * 1e3 / e2e_steps
    return {"value": n / (e2e_ms * 1e-3), "unit": "body-steps/s", "ms_per_step": e2e_ms,
            "h2d_bytes_per_step": 40 * n, "d2h_bytes_per_step": 32 * n,
            "path": "pinned host SoA -> lpe_bh_update_host (one CUDA graph: uploads on a copy stream in the order the step needs "
                    "them, the tree walk beside the velocity upload, kick + drift in creation order, x/y/vx/vy downloaded; "
                    "host wall clock incl. sync)"}


def e2e_registry(n, ticks):
    """Wall clock of Systems::BarnesHutSystem::update(entt::registry&) itself — the drop-in class on a real EnTT
    registry (host/dropin_bench.cpp, compiled against the reference's headers where they exist; the binary travels)."""
    exe = os.path.join(ROOT, "little-physics-engine_b200", "host", "_build", "dropin_bench")
    if not os.path.exists(exe):
        return {"value": None, "note": "host/_build/dropin_bench not built (needs the reference headers at build time)"}
    out = {}
    for mode in ("pagewise", "per_entity"):
        r = subprocess.run([exe, str(n), str(ticks), mode], capture_output=True, text=True, timeout=900)
        if r.returncode != 0 or not r.stdout.strip():
            return {"value": None, "note": f"dropin_bench failed: {r.stderr[-300:]}"}
        out[mode] = json.loads(r.stdout.strip().splitlines()[-1])
    pw = out["pagewise"]
    return {"value": pw["body_steps_per_s"], "unit": "body-steps/s", "ms_per_step": pw["ms_per_update"],
            "h2d_bytes_per_step": pw["h2d_bytes_per_tick"], "d2h_bytes_per_step": pw["d2h_bytes_per_tick"],
            "staging_path_taken": pw["staging_path_taken"],
            "ms_per_step_entity_by_entity_staging": out["per_entity"]["ms_per_update"],
            "path": "entt::registry -> Systems::BarnesHutSystem::update (pool pages copied into page-locked {x,y} buffers by "
                    "4 worker threads -> lpe_bh_tick_begin / _mass / _finish, each pool staged while the device works on the one "
                    "before, the pools' alignment checked beside the tree walk -> velocities copied back into the pool pages); "
                    "host wall clock"}


def nvlink_counters(index):
    """Sum of the NVLink data counters of one GPU (nvidia-smi nvlink -gt d), bytes; None when unavailable."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        return None
    tx = rx = 0
    seen = False
    for line in out.splitlines():
        parts = line.replace(":", " ").split()
        if "Tx" in parts and "KiB" in parts:
            tx += int(parts[parts.index("KiB") - 1]) * 1024; seen = True
        if "Rx" in parts and "KiB" in parts:
            rx += int(parts[parts.index("KiB") - 1]) * 1024; seen = True
    return {"tx": tx, "rx": rx} if seen else None


def rooflines(meas, fma_peak, peaks, peak_kind, workload_key, world=1):
    n, trav_ms = meas["n"], meas["phases"]["traverse"]
    flops = FLOPS_PER_INTERACTION * meas["interactions"] / max(world, 1)
    achieved = flops / (trav_ms * 1e-3) / 1e12
    bpb = hbm_bytes_per_body(meas["passes"])
    hbm_ms = meas["phases"]["keygen"] + meas["phases"]["sort"] + meas["phases"]["build"]
    hbm_ach = bpb * n / max(world, 1) / (hbm_ms * 1e-3) / 1e9
    traffic, tsrc = NCU_TRAFFIC.get(workload_key, (None, None))
    t_roof = 1e3 * (bpb * n / max(world, 1) / (peaks["hbm_gbs"] * 1e9) + (flops / (fma_peak * 1e12) if fma_peak else 0.0))
    return {
        "roofline": {"bound": "fp32_fma", "kernel": meas["kernel"], "achieved": achieved, "peak": fma_peak,
                     "unit": "TFLOP/s", "frac": achieved / fma_peak if fma_peak else None, "traffic": traffic,
                     "traffic_source": tsrc,
                     "peak_source": "lpe_bh_fma_peak measured in this run (MEASURED_PEAKS.json has no FP32 figure)",
                     "algorithmic": f"{FLOPS_PER_INTERACTION:.0f} flop x {meas['interactions']} accepted interactions / step"
                                    + (f" / {world} ranks" if world > 1 else "")},
        "roofline_hbm": {"bound": "hbm", "kernels": "k_keygen + k_sort_* + build kernels", "achieved": hbm_ach,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / peaks["hbm_gbs"],
                         "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                         "algorithmic": f"{bpb:.0f} B/body ({meas['passes']} sort passes) x {n} bodies"
                                        + (f" / {world} ranks" if world > 1 else "")},
        # T_roof = N*B_alg/BW_HBM + N*F_alg/P_FP32 (phases are sequential), against the measured step
        "roofline_step": {"t_roof_ms": t_roof, "t_measured_ms": meas["ms_per_step"], "frac": t_roof / meas["ms_per_step"]},
    }


def cpu_baseline_n1(wl):
    base, _ = reference_sample(wl, 1, 0)
    return base


def c5_lines(torch, lpe_bh, bh, stream, wl, bodies, flush):
    """theta sweep on the 1 M-body disk and the direct O(N^2) sum on its first 262 144 bodies."""
    x, y, vx, vy, m = bodies
    n = len(x)
    sweep = []
    for theta in (0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0):
        p = make_gpu_params(lpe_bh, wl, theta=theta)
        r = measure_single(torch, lpe_bh, bh, stream, wl, bodies, p, 5, 3, flush, "fast")
        sweep.append({"theta": theta, "ms_per_step": r["ms_per_step"], "value": r["value"],
                      "interactions_per_body": r["interactions"] / n, "traverse_ms": r["phases"]["traverse"]})
    nd = 262_144
    sub = tuple(a[:nd] for a in bodies)
    pq = make_gpu_params(lpe_bh, wl, quirk=False)
    bh.set_instrumentation()
    bh.upload(*sub)
    bh.direct_accel(pq, 0, 1024)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ax, ay = bh.direct_accel(pq, 0, nd)
    direct_s = time.perf_counter() - t0
    # Barnes-Hut on the same bodies, textbook tree (no first-occupant double count), against the direct sum
    one = lpe_bh.make_params(U, EPS, theta=THETA, dt_kick=1.0, dt_drift=1.0, quirk=False, do_drift=False)
    bh.upload(sub[0], sub[1], np.zeros(nd), np.zeros(nd), sub[4])
    bh.step(one, 1)
    got = bh.download()
    G = 6.674e-11
    mag = np.hypot(ax, ay)
    err = np.hypot(got["vx"] - ax, got["vy"] - ay) / np.maximum(mag, 1e-3 * np.median(mag))
    return {"theta_sweep": sweep,
            "direct": {"bodies": nd, "seconds": direct_s, "pair_interactions_per_s": nd * (nd - 1) / direct_s,
                       "value": nd / direct_s, "note": "fp64 direct sum (lpe_bh_direct_accel), host wall clock incl. download",
                       "bh_textbook_vs_direct_median_rel_err": float(np.median(err)),
                       "bh_textbook_vs_direct_p99_rel_err": float(np.percentile(err, 99))}}


def our_arm(args, wl, key, rank, world, local_rank):
    import torch
    import lpe_bh
    log(rank, f"start world={world} local_rank={local_rank} workload={wl['name']}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    # NCCL prints its version banner on stdout; the contract is ONE JSON line there, so everything but the final
    # print goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = wl["n"]
    wp = wl_params(wl)
    bodies = lpe_bh.workload(wl["kind"], n, wl["seed"], wp["U"])
    log(rank, "workload generated")
    params = make_gpu_params(lpe_bh, wl, args.precision)
    bh = lpe_bh.BarnesHut(local_rank)
    # a real (non-default) stream shared by torch and the library, so torch.cuda.Event brackets the library's kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    bh.set_stream(stream.cuda_stream)
    peaks, peak_kind = measured_peaks()
    warmup = max(args.warmup, 3)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    fma_peak = bh.fma_peak_tflops()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.wait_first()
    t_busy0 = time.time()
    line = None

    if world == 1:
        meas = measure_single(torch, lpe_bh, bh, stream, wl, bodies, params, args.steps, warmup, flush, args.precision)
        clocks = sampler.stop(t_busy0, time.time()) if sampler else None
        if clocks is not None:
            clocks["window"] = "warm-up + device-timed region"
        e2e = measure_e2e(torch, bh, bodies, params, args.steps)
        log(rank, "timed region done")
        line = {
            "metric": "barnes_hut_body_steps_per_sec", "value": meas["value"], "unit": "body-steps/s", "n_gpus": 1,
            "steps": args.steps, "warmup": warmup, "ms_per_step": meas["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 state, f32 interaction math" if args.precision == "fast" else "f64", "data": "synthetic",
            "config": config_for(wl, 1, args.precision, args.mgpu), "phases_ms": meas["phases"],
            "interactions_per_body": meas["interactions"] / n,
            **rooflines(meas, fma_peak, peaks, peak_kind, key),
            "e2e": e2e, "gpu_launches": meas["launches"], "clocks": clocks,
            "launch_mode": "timed steps replay the library's captured CUDA graph of a step (one submission per step, gpu_launches "
                           "counts the kernels inside); phases_ms come from separate steps with events between the phases",
        }
        if key == "c5":
            line.update(c5_lines(torch, lpe_bh, bh, stream, wl, bodies, flush))
        if key == "c2" and not args.bodies:
            line["e2e_registry"] = e2e_registry(n, 10)
        if args.workload is None and not args.bodies:
            # the 1 M-body configuration of BASELINE.json in the same run
            w2 = dict(WORKLOADS["c2"])
            b2 = lpe_bh.workload(w2["kind"], w2["n"], w2["seed"], U)
            p2 = make_gpu_params(lpe_bh, w2, args.precision)
            m2 = measure_single(torch, lpe_bh, bh, stream, w2, b2, p2, args.steps, warmup, flush, args.precision)
            c2 = {"config": config_for(w2, 1, args.precision, args.mgpu), "value": m2["value"], "unit": "body-steps/s",
                  "ms_per_step": m2["ms_per_step"], "phases_ms": m2["phases"], "interactions_per_body": m2["interactions"] / w2["n"],
                  **rooflines(m2, fma_peak, peaks, peak_kind, "c2"), "e2e": measure_e2e(torch, bh, b2, p2, args.steps),
                  "e2e_registry": e2e_registry(w2["n"], 10), "gpu_launches": m2["launches"]}
            if not args.no_cpu_baseline:
                c2["cpu_baseline"] = cpu_baseline_n1(w2)
            line["c2"] = c2
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_n1(wl)
    elif args.mgpu == "dd":
        line = dd_arm(args, torch, dist, lpe_bh, bh, stream, wl, key, bodies, params, rank, world, local_rank, warmup, flush,
                      fma_peak, peaks, peak_kind, sampler, t_busy0)
    else:
        line = replicated_arm(args, torch, dist, lpe_bh, bh, stream, wl, key, bodies, params, rank, world, warmup, flush,
                              fma_peak, peaks, peak_kind, sampler, t_busy0)
    if rank == 0:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        log(rank, "result printed")
    bh.close()
    if dist:
        dist.destroy_process_group()


def dd_arm(args, torch, dist, lpe_bh, bh, stream, wl, key, bodies, params, rank, world, local_rank, warmup, flush,
           fma_peak, peaks, peak_kind, sampler, t_busy0):
    """Domain-decomposed run: one process per GPU, windows exchanged once as CUDA IPC handles, then lockstep steps whose
    only synchronisation is the library's own in-stream flag barriers."""
    x, y, vx, vy, m = bodies
    n = len(x)
    capacity = int(n / world * 1.5) + 65536
    bh.dd_init(rank, world, capacity)
    handles = [None] * world
    dist.all_gather_object(handles, bh.dd_export())
    for r, hnd in enumerate(handles):
        if r != rank:
            bh.dd_import(r, hnd)
    bh.dd_upload(params, x, y, vx, vy, m)
    torch.cuda.synchronize()
    dist.barrier()
    log(rank, "bodies distributed")

    def gather_costs():
        mine = bh.dd_chunk_costs()
        allc = [None] * world
        dist.all_gather_object(allc, mine)
        return allc

    def rebalance():
        st = bh.dd_stats()
        local = st["ms_keygen"] + st["ms_sort"] + st["ms_build"] + st["ms_export"] + st["ms_top"]
        allc = gather_costs()
        mean_cost = float(np.mean(np.concatenate([c for _, c in allc]))) if sum(len(c) for _, c in allc) else 1.0
        mine = torch.tensor([local, st["ms_traverse"]], device="cuda", dtype=torch.float64)
        allt = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allt, mine)
        tl = np.array([float(t[0].item()) for t in allt]); tt = np.array([float(t[1].item()) for t in allt])
        beta = mean_cost * float(tl.sum()) / max(float(tt.sum()), 1e-9)     # sort + build share of a chunk, in list entries
        # what the cost model misses on a rank (deeper tree, more cells per body) is charged to its chunks: measured busy
        # time of the rank per unit of model weight it holds now
        busy = tl + tt
        held = np.array([float(np.sum(c.astype(np.float64) + beta)) for _, c in allc])
        rho = busy / np.maximum(held, 1e-9)
        scale = rho / max(float(rho.mean()), 1e-30)
        new = lpe_bh.balanced_splitters(allc, world, beta, scale)
        bh.dd_set_splitters(new)
        return beta

    # settle: a few steps, re-balance on the measured per-chunk cost, a few more (the bodies move to their new owners)
    bh.set_instrumentation(timing=True)
    balance_log = []
    for it in range(4):
        bh.dd_step(params, 2)
        bh.synchronize()
        balance_log.append({"traverse_ms": bh.dd_stats()["ms_traverse"], "n_live": bh.dd_stats()["n_live"]})
        beta = rebalance()
        dist.barrier()
    for _ in range(warmup):
        bh.dd_step(params, 1)
    bh.synchronize()
    # phase table: a few steps with the library's own events between the phases (plain launches)
    phases = []
    for _ in range(min(args.steps, 5)):
        flush.zero_()
        bh.dd_step(params, 1)
        s = bh.dd_stats()
        phases.append([s[k] for k in ("ms_keygen", "ms_wait_a", "ms_sort", "ms_build", "ms_export", "ms_wait_b", "ms_top", "ms_traverse")])
    # timed region: no instrumentation, so every rank replays its captured CUDA graph of the step (two buffer parities:
    # captured on the 3rd / 4th step with the same splitters, 6 steps settle that)
    bh.set_instrumentation()
    for _ in range(6):
        bh.dd_step(params, 1)
    bh.synchronize()
    torch.cuda.synchronize()
    log(rank, "warm-up done")
    launches0 = bh.launch_count()
    dist.barrier()
    torch.cuda.synchronize()
    nv0 = nvlink_counters(local_rank) if rank == 0 else None
    dist.barrier()   # (rank 0's nvidia-smi call takes tens of ms: nobody starts the timed steps before it is back)
    torch.cuda.synchronize()
    step_ms = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        bh.dd_step(params, 1)
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    torch.cuda.synchronize()
    dist.barrier()
    nv1 = nvlink_counters(local_rank) if rank == 0 else None
    nvlink = None
    if nv0 and nv1:   # rank 0's GPU: what the step's own kernels moved over NVLink (exports, roots, migrants, mail)
        nvlink = {"gpu": local_rank, "tx_bytes_per_step": (nv1["tx"] - nv0["tx"]) / args.steps,
                  "rx_bytes_per_step": (nv1["rx"] - nv0["rx"]) / args.steps,
                  "source": "nvidia-smi nvlink -gt d before / after the timed steps (includes the L2-flush-free idle gaps; "
                            "no collective runs in between)"}
    launches = bh.launch_count() - launches0
    t = torch.tensor([float(np.sum(step_ms))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    ph = np.mean(np.array(phases), axis=0)
    allph = [None] * world
    dist.all_gather_object(allph, [float(v) for v in ph])
    st = bh.dd_stats()
    allst = [None] * world
    dist.all_gather_object(allst, {"n_live": st["n_live"], "n_cells": st["n_cells"], "exported_blocks": sum(st["exported_blocks"]),
                                   "n_roots": st["n_roots"]})
    clocks = sampler.stop(t_busy0, time.time()) if sampler else None
    if clocks is not None:
        clocks["window"] = "settling + warm-up + device-timed region"

    # ---- parity: one counted step of the decomposed run against the unsharded step on rank 0's GPU ----
    parity = dd_parity_check(torch, dist, lpe_bh, bh, bodies, params, rank, world, local_rank, capacity)
    log(rank, "timed region done")
    if rank != 0:
        return None
    names = ("keygen", "wait_migrants", "sort", "build", "export", "wait_export", "top", "traverse")
    arr = np.array(allph)
    imax = arr.max(axis=0)
    interactions = parity["interactions"]
    meas = dict(n=n, ms_per_step=ms_per_step, interactions=interactions, passes=bh.stats()["sort_passes"],
                phases={"keygen": float(imax[0]), "sort": float(imax[2]), "build": float(imax[3] + imax[4]),
                        "traverse": float(imax[7])}, kernel="k_traverse2 (two-phase traversal)")
    return {
        "metric": "barnes_hut_body_steps_per_sec", "value": n / (ms_per_step * 1e-3), "unit": "body-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64 state, f32 interaction math", "data": "synthetic",
        "config": config_for(wl, world, args.precision, args.mgpu),
        "parallelism_detail": "domain decomposition by key range: per-rank sort + build, published roots + locally essential "
                              "child blocks stored into peer memory (CUDA IPC over NVLink) by the step's own kernels, two "
                              "in-stream flag barriers per step, no collective in the data path",
        "phases_ms_max_over_ranks": dict(zip(names, (float(v) for v in imax))),
        "phases_ms_per_rank": {nm: [float(v) for v in arr[:, i]] for i, nm in enumerate(names)},
        "per_rank": allst, "balance": {"beta_list_entries_per_chunk": beta, "settling": balance_log},
        "interactions_per_body": interactions / n, "parity_check": parity, "nvlink": nvlink,
        **rooflines(meas, fma_peak, peaks, peak_kind, key, world),
        "e2e": {"value": None, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "path": "multi-GPU runs keep the bodies resident on their owners; the host round trip is measured at N=1"},
        "gpu_launches": int(launches), "clocks": clocks,
        "launch_mode": "timed steps replay each rank's captured CUDA graph of the whole decomposed step (three phases, two "
                       "in-stream barriers; gpu_launches counts rank 0's kernels inside); the phase tables come from "
                       "separate steps with events between the phases",
    }


def dd_parity_check(torch, dist, lpe_bh, bh, bodies, params, rank, world, local_rank, capacity):
    """Re-distribute the ORIGINAL bodies, run one counted step on all ranks, gather on rank 0 and compare with one
    unsharded step of the same bodies on rank 0's GPU: identical per-body accepted-interaction counts (every theta
    decision), velocities within fp32 summation noise."""
    x, y, vx, vy, m = bodies
    n = len(x)
    bh.set_instrumentation(counts=True)
    bh.dd_upload(params, x, y, vx, vy, m)
    torch.cuda.synchronize()
    dist.barrier()
    bh.dd_step(params, 1)
    d = bh.dd_download(counts=True)
    nl = torch.tensor([len(d["index"])], device="cuda", dtype=torch.int64)
    counts = [torch.zeros_like(nl) for _ in range(world)]
    dist.all_gather(counts, nl)
    pad = lambda a, dt: torch.from_numpy(np.concatenate([a, np.zeros(capacity - len(a), a.dtype)]).astype(dt)).cuda()
    mine = {"index": pad(d["index"], np.int64), "vx": pad(d["vx"], np.float64), "vy": pad(d["vy"], np.float64),
            "acc": pad(d["accepted"], np.int64)}
    got = {}
    for k, t in mine.items():
        parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, parts, dst=0)
        if rank == 0:
            got[k] = np.concatenate([p[:int(c.item())].cpu().numpy() for p, c in zip(parts, counts)])
    bh.set_instrumentation(timing=True)
    out = None
    if rank == 0:
        one = lpe_bh.BarnesHut(local_rank)
        one.set_instrumentation(counts=True)
        one.upload(x, y, vx, vy, m)
        one.step(params, 1)
        ref = one.download()
        racc, _ = one.counts()
        interactions = one.stats()["interactions"]
        one.close()
        idx = got["index"]
        seen = np.bincount(idx, minlength=n)
        dvx, dvy = got["vx"] - vx[idx], got["vy"] - vy[idx]
        rvx, rvy = ref["vx"][idx] - vx[idx], ref["vy"][idx] - vy[idx]
        mag = np.hypot(rvx, rvy)
        err = np.hypot(dvx - rvx, dvy - rvy) / np.maximum(mag, 1e-3 * np.median(mag))
        out = {"what": "one counted step of the decomposed run vs the unsharded step on one GPU, all bodies",
               "bodies": n, "every_body_owned_once": bool(np.all(seen == 1)),
               "accepted_counts_identical": bool(np.array_equal(got["acc"], racc[idx])),
               "max_rel_dv_diff": float(err.max()), "interactions": int(interactions)}
    dist.barrier()
    return out


def replicated_arm(args, torch, dist, lpe_bh, bh, stream, wl, key, bodies, params, rank, world, warmup, flush, fma_peak,
                   peaks, peak_kind, sampler, t_busy0):
    """Round 1's variant: every rank holds all bodies and builds the same tree, traversal by block-cyclic key slices,
    new state stored into every rank's receive buffer by the traversal kernel itself."""
    x, y, vx, vy, m = bodies
    n = len(x)
    bh.set_instrumentation(counts=True)
    bh.upload(x, y, vx, vy, m)
    bh.step(params, 1)
    interactions = bh.stats()["interactions"]
    bh.set_shard(rank, world)
    bh.set_instrumentation(timing=True)
    bh.upload(x, y, vx, vy, m)
    view = bh.device_view()
    send = torch.as_tensor(CudaArray(view.xchg_send, 4 * view.xchg_chunk), device="cuda")
    recv = torch.as_tensor(CudaArray(view.xchg_recv, 4 * view.xchg_chunk * world), device="cuda")
    ok = 0
    if not args.no_p2p and world <= 8:
        mine = None
        try:
            mine = bh.xchg_export()
        except Exception as e:
            log(rank, f"direct exchange: export failed: {e}")
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        if all(h is not None for h in handles):
            try:
                for r, hnd in enumerate(handles):
                    if r != rank:
                        bh.xchg_import(r, hnd)
                ok = 1 if bh.xchg_p2p_ready() else 0
            except Exception as e:
                log(rank, f"direct exchange unavailable: {e}")
    flag = torch.tensor([ok], device="cuda", dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    p2p = bool(flag.item())
    if not p2p:
        bh.xchg_reset()
    token = torch.zeros(1, device="cuda", dtype=torch.int32)

    def one_step():
        bh.step_begin(params)
        if p2p:
            dist.all_reduce(token)      # stream-ordered barrier: every rank's stores have landed
        else:
            dist.all_gather_into_tensor(recv, send)
        bh.step_finish()

    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize()
    launches0 = bh.launch_count()
    dist.barrier()
    torch.cuda.synchronize()
    step_ms, phases = [], []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        one_step()
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        s = bh.stats()
        phases.append((s["ms_keygen"], s["ms_sort"], s["ms_build"], s["ms_traverse"]))
    torch.cuda.synchronize()
    dist.barrier()
    launches = bh.launch_count() - launches0
    t = torch.tensor([float(np.sum(step_ms))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    ph = np.mean(np.array(phases), axis=0)
    clocks = sampler.stop(t_busy0, time.time()) if sampler else None
    if rank != 0:
        return None
    meas = dict(n=n, ms_per_step=ms_per_step, interactions=interactions, passes=bh.stats()["sort_passes"],
                phases={"keygen": float(ph[0]) * world, "sort": float(ph[1]) * world, "build": float(ph[2]) * world,
                        "traverse": float(ph[3])}, kernel="k_traverse2 (two-phase traversal)")
    return {
        "metric": "barnes_hut_body_steps_per_sec", "value": n / (ms_per_step * 1e-3), "unit": "body-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64 state, f32 interaction math", "data": "synthetic",
        "config": config_for(wl, world, args.precision, args.mgpu),
        "parallelism_detail": "replicated tree, block-cyclic key-slice traversal, " +
                              ("(x,y,vx,vy) stored into every rank's receive buffer by the traversal kernel (NVLink peer memory)"
                               if p2p else "NCCL allgather of (x,y,vx,vy)"),
        "phases_ms": {"keygen": float(ph[0]), "sort": float(ph[1]), "build": float(ph[2]), "traverse": float(ph[3])},
        "interactions_per_body": interactions / n,
        **rooflines(meas, fma_peak, peaks, peak_kind, key, world),
        "e2e": {"value": None, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "path": "multi-GPU runs keep bodies resident; host round trip measured at N=1 only"},
        "gpu_launches": int(launches), "clocks": clocks,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fast", choices=["fast", "strict"])
    ap.add_argument("--mgpu", default="dd", choices=["dd", "replicated"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bodies", type=int, default=None,
                    help="TESTS ONLY: shrink the workload to this many bodies; the line is then marked reduced and is not a bench value")
    ap.add_argument("--no-p2p", action="store_true", help="--mgpu replicated: NCCL allgather instead of the fused peer-memory exchange")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run on this node
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    # ONE workload at every N (the driver's scaling series must be one problem): C3, the sharded 16 M-body configuration
    key = args.workload or "c3"
    wl = dict(WORKLOADS[key])
    if args.bodies:
        wl["n"] = int(args.bodies)
        wl["name"] = f"REDUCED to {wl['n']} bodies (contract test, not a bench value): " + wl["name"]
    if args.impl == "reference":
        reference_arm(args, wl, rank, world)
    else:
        our_arm(args, wl, key, rank, world, local_rank)


if __name__ == "__main__":
    main()
