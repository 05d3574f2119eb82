"""Multi-GPU check of the domain-decomposed step (include/lpe_bh.h, lpe_bh_dd_*) on real peers.

  torchrun --nproc-per-node N scripts/check_dd.py [n_bodies] [steps]     one process per GPU, CUDA IPC windows
  python scripts/check_dd.py --one-process N [n_bodies] [steps]          one host thread drives N GPUs (peer access)

Both use the library's own in-stream flag barriers. Checked against an unsharded run of the same bodies on one GPU:
FAST precision — every body owned exactly once, per-body accepted-interaction counts identical, velocities within fp32
summation noise; STRICT precision — positions and velocities bit for bit. `steps` steps with a drift long enough that
bodies change owner.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "little-physics-engine_b200"))
import lpe_bh  # noqa: E402


def bodies(n):
    U = 2.0 ** 20
    x, y, vx, vy, m = lpe_bh.workload("plummer", n, 7, U)
    rng = np.random.default_rng(11)
    vx = rng.normal(0, 2e3, n); vy = rng.normal(0, 2e3, n)     # fast enough to cross key-range borders in a few steps
    return U, (x, y, vx, vy, m)


def params(U, precision):
    return lpe_bh.make_params(U, U / 2 ** 14, theta=0.5, dt_kick=1e5, dt_drift=0.05, precision=precision)   # (a long kick: dv stays resolvable next to |v| ~ 2e3)


def reference(device, p, b, steps, counts):
    one = lpe_bh.BarnesHut(device)
    one.set_instrumentation(counts=counts)
    one.upload(*b)
    one.step(p, steps)
    out = one.download()
    if counts:
        out["accepted"], _ = one.counts()
    one.close()
    return out


def compare(tag, got, ref, b, precision):
    x, y, vx, vy, m = b
    idx = got["index"]
    ok = True
    if precision == lpe_bh.PREC_STRICT:
        bad = [k for k in ("x", "y", "vx", "vy") if not np.array_equal(got[k], ref[k][idx])]
        ok = not bad
        msg = "bit for bit" if ok else f"MISMATCH in {bad}"
    else:
        acc_ok = np.array_equal(got["accepted"], ref["accepted"][idx])
        rvx, rvy = ref["vx"][idx] - vx[idx], ref["vy"][idx] - vy[idx]
        mag = np.hypot(rvx, rvy)
        err = np.hypot(got["vx"] - vx[idx] - rvx, got["vy"] - vy[idx] - rvy) / np.maximum(mag, 1e-3 * np.median(mag))
        ok = acc_ok and err.max() <= 5e-5
        msg = f"accepted counts {'identical' if acc_ok else 'DIFFER'}, max rel dv diff {err.max():.2e}"
    print(f"[{tag}] {'OK' if ok else 'FAIL'}: {msg}", flush=True)
    return ok


def one_process(ndev, n, steps):
    U, b = bodies(n)
    fails = 0
    for precision in (lpe_bh.PREC_FAST, lpe_bh.PREC_STRICT):
        p = params(U, precision)
        fast = precision == lpe_bh.PREC_FAST
        ref = reference(0, p, b, steps, fast)
        g = lpe_bh.DDGroup(list(range(ndev)), int(n / ndev * 1.5) + 4096)
        for c in g.ranks:
            c.set_instrumentation(counts=fast)
        g.upload(p, *b)
        g.step(p, steps)
        seen = np.zeros(n, np.int32)
        for r, c in enumerate(g.ranks):
            d = c.dd_download(counts=fast)
            seen[d["index"]] += 1
            fails += not compare(f"one process, {ndev} GPUs, rank {r}, {'FAST' if fast else 'STRICT'}", d, ref, b, precision)
        fails += not np.all(seen == 1)
        print(f"moved bodies per rank: {[s['n_live'] for s in g.stats()]}", flush=True)
        g.close()
    sys.exit(1 if fails else 0)


def multi_process(n, steps):
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    U, b = bodies(n)
    fails = 0
    for precision in (lpe_bh.PREC_FAST, lpe_bh.PREC_STRICT):
        p = params(U, precision)
        fast = precision == lpe_bh.PREC_FAST
        ref = reference(local, p, b, steps, fast)
        bh = lpe_bh.BarnesHut(local)
        bh.set_instrumentation(counts=fast)
        bh.dd_init(rank, world, int(n / world * 1.5) + 4096)
        handles = [None] * world
        dist.all_gather_object(handles, bh.dd_export())
        for r, h in enumerate(handles):
            if r != rank:
                bh.dd_import(r, h)
        bh.dd_upload(p, *b)
        dist.barrier()
        # half of the steps, a re-balance on the measured costs, the other half
        bh.dd_step(p, max(1, steps // 2))
        allc = [None] * world
        dist.all_gather_object(allc, bh.dd_chunk_costs())
        bh.dd_set_splitters(lpe_bh.balanced_splitters(allc, world))
        dist.barrier()
        bh.dd_step(p, steps - max(1, steps // 2))
        d = bh.dd_download(counts=fast)
        fails += not compare(f"rank {rank} of {world}, {'FAST' if fast else 'STRICT'}", d, ref, b, precision)
        if rank == 0:
            print(f"[rank 0] {bh.graph_replays()} of {steps} steps were replays of the rank's captured CUDA graph", flush=True)
        owned = torch.tensor([len(d["index"])], device="cuda")
        dist.all_reduce(owned)
        fails += int(owned.item()) != n
        dist.barrier()
        bh.close()
    t = torch.tensor([fails], device="cuda")
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(1 if t.item() else 0)


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "--one-process":
        ndev = int(args[1])
        rest = args[2:]
        one_process(ndev, int(rest[0]) if rest else 300_000, int(rest[1]) if len(rest) > 1 else 4)
    else:
        multi_process(int(args[0]) if args else 300_000, int(args[1]) if len(args) > 1 else 4)
