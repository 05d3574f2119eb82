#!/usr/bin/env python
"""bench.py — Barnes-Hut body-steps/s (theta = 0.5) on B200, the metric of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (one rank per GPU, NCCL)

A "step" is one full pass of the hot path over all bodies: space-filling-curve keys, radix sort, quadtree build + aggregation,
theta traversal, kick and drift. One JSON line is printed by rank 0.

  value      whole-job body-steps/s with the bodies resident in HBM (device-timed, CUDA events on the launching
             stream, max over ranks); L2 is flushed between timed steps.
  e2e        the same metric through the call the ECS drop-in makes (lpe_bh_update_host semantics): pinned host
             buffers -> upload -> step -> download, every step, copies inside the timed region.
  roofline   the dominant kernel (k_traverse): FP32-pipe bound, algorithmic flops = 20 per
             accepted interaction, against the FP32 FMA peak measured on this GPU by lpe_bh_fma_peak.
  roofline_hbm  the HBM-bound phases (keygen + sort + build) against MEASURED_PEAKS.json's copy bandwidth.
  cpu_baseline  the reference's own barnes_hut.cpp + movement.cpp (oracle/_ref, compiled unmodified) on the host.

--impl reference times that CPU code (all it can use: it is single-threaded) on bounded samples.
The oracle is only ever the baseline / checker here; the measured product path is the CUDA library.
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
    "c2": dict(kind="disk", n=1_000_000, seed=42, name="C2: 1M-body uniform disk, theta=0.5, U=2^20, eps=U/2^14"),
    "c3": dict(kind="plummer", n=16_000_000, seed=43, name="C3: 16M-body Plummer sphere, theta=0.5, U=2^20, eps=U/2^14"),
    "c4": dict(kind="two_galaxies", n=4_000_000, seed=44, name="C4: 4M-body two-galaxy collision, theta=0.5, U=2^20, eps=U/2^14"),
}
FLOPS_PER_INTERACTION = 20.0
# dram__bytes_read.sum + dram__bytes_write.sum of one k_traverse2 launch, from the committed ncu --set full capture
# (profiles/r01_ncu_traverse2_c2_final.txt: 146.32 MB read + 32.57 MB written); the kernel is not DRAM-bound, the
# figure only shows that nothing is re-read (the records alone are 52 MB, the bodies 64 MB).
NCU_TRAFFIC_BYTES = {"c2": 178.9e6}
HBM_BYTES_PER_BODY = 340.0     # keygen + sort + gather + node arrays, 64-bit keys


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
def reference_arm(args, wl, rank, world):
    """The reference's own CPU implementation on this box's host cores. Bounded sample per step: the full tree is
    built over all N bodies (every body is a source) and the force loop runs over every `stride`-th body — done by
    giving only those bodies a Velocity component, so it is still the unmodified reference code path
    (barnes_hut.cpp:89 iterates view<Position,Velocity,Mass>). A step's time is scaled to all N targets:
    t_step = t_build + (t_sample - t_build) * stride, with t_build from a build-only call."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    import lpe_bh
    n_full = wl["n"]
    x, y, vx, vy, m = lpe_bh.workload(wl["kind"], n_full, wl["seed"], U)
    # Workloads above 2 M bodies (C3: 16 M would be minutes and 5 GB of node pool per CPU step): the reference runs on
    # every k-th body of the same workload (about 1 M bodies); its per-body throughput on that sample is reported
    # (the reference gets slower per body as N grows, so this flatters the reference, not us).
    sub = max(1, n_full // 1_000_000) if n_full > 2_000_000 else 1
    if sub > 1:
        x, y, vx, vy, m = (np.ascontiguousarray(a[::sub]) for a in (x, y, vx, vy, m))
    n = len(x)
    p = O.make_params(U, EPS, theta=THETA, dt_kick=DT, dt_drift=DT)
    stride = max(1, n // 50_000)          # ~50k targets per sample: ~1-2 s of force work per step
    comp = np.full(n, O.HAS_MASS, np.uint8)
    comp[::stride] |= O.HAS_VELOCITY
    ntargets = int(np.count_nonzero(comp & O.HAS_VELOCITY))
    if O.RefLib.available():
        lib, kind = O.RefLib(), "reference"
        run = lambda: lib.run(p, x, y, vx, vy, m, comp=comp, nsteps=1, pool_nodes=4 * n + 4096)["stats"]["total_seconds"]
        build = lambda: lib.tree(p, x, y, m, pool_nodes=4 * n + 4096)[1]["build_seconds"]
    else:
        lib, kind = O.PortLib(), "port"
        run = lambda: lib.run(p, x, y, vx, vy, m, comp=comp, nsteps=1, threads=1)["stats"]["total_seconds"]
        build = lambda: lib.tree(p, x, y, m)[1]["build_seconds"]
    t_build = build()
    times = []
    for s in range(args.warmup + args.steps):
        t = run()
        if s >= args.warmup:
            times.append(t_build + max(t - t_build, 0.0) * (n / ntargets))
    ms = 1e3 * float(np.mean(times))
    value = n / (ms * 1e-3)
    ms = ms * (n_full / n)                # time of one step of the full workload at the measured per-body rate
    sample = (f"per step: full tree build over all {n} bodies + force loop over every {stride}th body "
              f"({ntargets} targets), scaled to {n} targets; {lib.describe()}")
    if sub > 1:
        sample = (f"every {sub}th body of the {n_full}-body workload ({n} bodies); " + sample +
                  f"; ms_per_step = {n_full} bodies at the measured per-body rate")
    line = {
        "impl": "reference", "metric": "barnes_hut_body_steps_per_sec", "value": value, "unit": "body-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "bodies": n_full, "theta": THETA},
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_n1(wl):
    """Rank 0, N=1 only: ONE full, unsampled step of the reference on the same bodies (about 20 s for C2)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    import lpe_bh
    n = min(wl["n"], 1_000_000)            # C3/C4: the first 1M bodies of the same distribution (bounded sample)
    x, y, vx, vy, m = lpe_bh.workload(wl["kind"], wl["n"], wl["seed"], U)
    x, y, vx, vy, m = (a[:n] for a in (x, y, vx, vy, m))
    p = O.make_params(U, EPS, theta=THETA, dt_kick=DT, dt_drift=DT)
    if O.RefLib.available():
        lib, kind = O.RefLib(), "reference"
        t = lib.run(p, x, y, vx, vy, m, nsteps=1, pool_nodes=4 * n + 4096)["stats"]["total_seconds"]
    else:
        lib, kind = O.PortLib(), "port"
        t = lib.run(p, x, y, vx, vy, m, nsteps=1, threads=1)["stats"]["total_seconds"]
    return {"value": n / t, "unit": "body-steps/s", "cores": 1, "kind": kind, "seconds_per_step": t,
            "sample": f"one full step (tree build + all {n} targets + movement) of {n} bodies of the workload; "
                      f"{lib.describe()}; host has {os.cpu_count()} cores, the reference can use 1"}


# ------------------------------------------------------------------------------------------------------ our arm
class CudaArray:
    """Expose a raw device pointer of the library to torch (zero copy) for the NCCL allgather."""

    def __init__(self, ptr, nelem):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def log(rank, msg):
    print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)


def our_arm(args, wl, rank, world, local_rank):
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
    x, y, vx, vy, m = lpe_bh.workload(wl["kind"], n, wl["seed"], U)
    log(rank, "workload generated")
    params = lpe_bh.make_params(U, EPS, theta=THETA, dt_kick=DT, dt_drift=DT)
    bh = lpe_bh.BarnesHut(local_rank)
    # a real (non-default) stream shared by torch and the library, so torch.cuda.Event brackets the library's kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    bh.set_stream(stream.cuda_stream)
    peaks, peak_kind = measured_peaks()
    warmup = max(args.warmup, 3)

    # one counting step on a scratch copy of the bodies: interactions per step for the flop roofline
    bh.set_instrumentation(counts=True)
    bh.upload(x, y, vx, vy, m)
    bh.step(params, 1)
    interactions = bh.stats()["interactions"]
    fma_peak = bh.fma_peak_tflops()

    send = recv = token = None
    p2p = False
    if world > 1:
        bh.set_shard(rank, world)
    bh.set_instrumentation(timing=True)
    bh.upload(x, y, vx, vy, m)
    if world > 1:
        view = bh.device_view()
        send = torch.as_tensor(CudaArray(view.xchg_send, 4 * view.xchg_chunk), device="cuda")
        recv = torch.as_tensor(CudaArray(view.xchg_recv, 4 * view.xchg_chunk * world), device="cuda")
        # Direct exchange: every rank opens every other rank's receive buffer (CUDA IPC over NVLink peer memory) and
        # the traversal kernel stores its results there itself. All ranks must agree, else the NCCL allgather is used.
        ok = 0
        if not args.no_p2p and world <= 8:
            mine = None
            try:
                mine = bh.xchg_export()
            except Exception as e:
                log(rank, f"direct exchange: export failed: {e}")
            handles = [None] * world
            dist.all_gather_object(handles, mine)          # every rank takes part, whatever happened above
            if all(h is not None for h in handles):
                try:
                    for r, hnd in enumerate(handles):
                        if r != rank:
                            bh.xchg_import(r, hnd)
                    ok = 1 if bh.xchg_p2p_ready() else 0
                except Exception as e:   # IPC not permitted on this box: fall back to the collective
                    log(rank, f"direct exchange unavailable: {e}")
        flag = torch.tensor([ok], device="cuda", dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        p2p = bool(flag.item())
        if not p2p:
            bh.xchg_reset()      # some rank could not open every handle: every rank goes back to the allgather
        token = torch.zeros(1, device="cuda", dtype=torch.int32)

    xe = [torch.cuda.Event(enable_timing=True) for _ in range(3)]   # multi-GPU: allgather / scatter split

    def one_step():
        if world == 1:
            bh.step(params, 1)
        else:
            bh.step_begin(params)
            xe[0].record(stream)
            if p2p:
                dist.all_reduce(token)      # stream-ordered barrier: every rank's stores have landed
            else:
                dist.all_gather_into_tensor(recv, send)
            xe[1].record(stream)
            bh.step_finish()
            xe[2].record(stream)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.wait_first()
    t_busy0 = time.time()
    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize()
    log(rank, "warm-up done")

    launches0 = bh.launch_count()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    step_ms, phases = [], []
    for _ in range(args.steps):
        flush.zero_()                       # evict L2 between timed steps (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        one_step()
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        st = bh.stats()
        phases.append((st["ms_keygen"], st["ms_sort"], st["ms_build"], st["ms_traverse"]) +
                      ((xe[0].elapsed_time(xe[1]), xe[1].elapsed_time(xe[2])) if world > 1 else (0.0, 0.0)))
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t_wall1 = time.time()
    launches = bh.launch_count() - launches0
    total_ms = float(np.sum(step_ms))
    if dist:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n / (ms_per_step * 1e-3)
    ph = np.mean(np.array(phases), axis=0)
    per_rank_traverse = None
    if dist:   # load balance of the block-cyclic slices (C4 is the stress case): every rank's mean traversal time
        mine = torch.tensor([float(ph[3])], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank_traverse = [float(t.item()) for t in allr]

    # clocks / throttle reasons sampled by nvidia-smi every 20 ms from the first warm-up step to the end of the
    # device-timed region (which alone is only tens of milliseconds long). The sampler stops before the end-to-end
    # loop: that one is host wall clock over ~55 CUDA API calls per tick, and a driver query every 20 ms stalls them
    # (2.0 -> 2.3 ms per tick at 1 M bodies).
    clocks = sampler.stop(t_busy0, time.time()) if sampler else None
    if clocks is not None:
        clocks["window"] = "warm-up + device-timed region"

    # ---- end to end through host buffers (what Systems::BarnesHutSystem::update pays), N=1 path of the C ABI ----
    e2e = None
    if world == 1:
        hx, hy, hvx, hvy, hm = (torch.from_numpy(a.copy()).pin_memory() for a in (x, y, vx, vy, m))
        ptrs = [t.data_ptr() for t in (hx, hy, hvx, hvy, hm)]
        e2e_steps = max(3, min(args.steps, 20))
        bh.set_instrumentation()              # no phase events in the host-clocked loop
        for it in range(2 + e2e_steps):
            if it == 2:
                torch.cuda.synchronize()
                te0 = time.perf_counter()
            bh.update_host_ptrs(params, n, *ptrs)   # synchronises; the result lands in the pinned host arrays
        e2e_ms = (time.perf_counter() - te0) * 1e3 / e2e_steps
        e2e = {"value": n / (e2e_ms * 1e-3), "unit": "body-steps/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": 40 * n, "d2h_bytes_per_step": 32 * n,
               "path": "pinned host SoA -> lpe_bh_update_host (uploads on a copy stream behind the step, kick + drift, "
                       "x/y/vx/vy downloaded; host wall clock incl. sync)"}
    else:
        e2e = {"value": None, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "path": "multi-GPU runs keep bodies resident; host round trip measured at N=1 only"}

    log(rank, "timed region done")
    if rank == 0:
        trav_ms = float(ph[3])
        flops = FLOPS_PER_INTERACTION * interactions / max(world, 1)    # this rank's share of the targets
        achieved = flops / (trav_ms * 1e-3) / 1e12
        hbm_ms = float(ph[0] + ph[1] + ph[2])
        hbm_ach = HBM_BYTES_PER_BODY * n / (hbm_ms * 1e-3) / 1e9
        line = {
            "metric": "barnes_hut_body_steps_per_sec", "value": value, "unit": "body-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64 state, f32 interaction math", "data": "synthetic",
            "config": {"workload": wl["name"], "bodies": n, "theta": THETA, "softening": EPS, "universe": U,
                       "quirk_mode": "reference", "precision": "fast",
                       "l2": "256 MB buffer written between timed steps (L2 flush); per-step CUDA events",
                       "parallelism": "single GPU" if world == 1 else
                       f"{world} GPUs: replicated tree, block-cyclic key-slice traversal, " +
                       ("(x,y,vx,vy) stored into every rank's receive buffer by the traversal kernel itself (NVLink peer "
                        "memory, CUDA IPC) + one-element allreduce as barrier" if p2p else "NCCL allgather of (x,y,vx,vy)")},
            "phases_ms": {"keygen": float(ph[0]), "sort": float(ph[1]), "build": float(ph[2]), "traverse": trav_ms,
                          "allgather": float(ph[4]), "scatter": float(ph[5])},
            "per_rank_traverse_ms": per_rank_traverse,
            "interactions_per_body": interactions / n,
            "roofline": {"bound": "fp32_fma", "kernel": "k_traverse2 (two-phase traversal)", "achieved": achieved,
                         "peak": fma_peak, "unit": "TFLOP/s", "frac": achieved / fma_peak if fma_peak else None,
                         "traffic": NCU_TRAFFIC_BYTES.get(args.workload or ("c2" if world == 1 else "c3")),
                         "peak_source": "lpe_bh_fma_peak measured in this run (MEASURED_PEAKS.json has no FP32 figure)",
                         "algorithmic": f"{FLOPS_PER_INTERACTION:.0f} flop x {interactions} accepted interactions / launch"},
            "roofline_hbm": {"bound": "hbm", "kernels": "k_keygen + k_sort_* + build kernels", "achieved": hbm_ach,
                             "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / peaks["hbm_gbs"],
                             "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                             "algorithmic": f"{HBM_BYTES_PER_BODY:.0f} B/body x {n} bodies"},
            # T_roof = N*B_alg/BW_HBM + N*F_alg/P_FP32 (phases are sequential; the replicated build
            # is counted once per rank, the traversal divides by the ranks), against the measured step
            "roofline_step": (lambda t_roof: {"t_roof_ms": t_roof, "t_measured_ms": ms_per_step, "frac": t_roof / ms_per_step})(
                1e3 * (HBM_BYTES_PER_BODY * n / (peaks["hbm_gbs"] * 1e9) +
                       (flops / (fma_peak * 1e12) if fma_peak else 0.0))),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_n1(wl)
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        log(rank, "result printed")
    bh.close()
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bodies", type=int, default=None,
                    help="TESTS ONLY: shrink the workload to this many bodies; the line is then marked reduced and is not a bench value")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL allgather instead of the fused peer-memory exchange")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run on this node
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    # N=1: the configuration the metric is quoted on that the reference can also run (C2); N>1: the sharded 16M config
    wl = dict(WORKLOADS[args.workload or ("c2" if world == 1 else "c3")])
    if args.bodies:
        wl["n"] = int(args.bodies)
        wl["name"] = f"REDUCED to {wl['n']} bodies (contract test, not a bench value): " + wl["name"]
    if args.impl == "reference":
        reference_arm(args, wl, rank, world)
    else:
        our_arm(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
