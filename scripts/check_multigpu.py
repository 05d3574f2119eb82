"""Multi-process check of the sharded step (run under torchrun, one rank per GPU).

Every rank runs `steps` sharded steps — once with the fused peer-memory exchange, once with the NCCL allgather — and
compares the downloaded state bit for bit with an unsharded run of the same bodies on its own GPU.
usage: python -m torch.distributed.run --nproc-per-node N scripts/check_multigpu.py [n_bodies] [steps]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "little-physics-engine_b200"))
import lpe_bh  # noqa: E402
from bench import CudaArray  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    U = 2.0 ** 20
    x, y, vx, vy, m = lpe_bh.workload("plummer", n, 7, U)
    p = lpe_bh.make_params(U, U / 2 ** 14, theta=0.5, dt_kick=1 / 120, dt_drift=1 / 120)
    stream = torch.cuda.Stream()
    one = lpe_bh.BarnesHut(local)
    one.upload(x, y, vx, vy, m)
    one.step(p, steps)
    ref = one.download()
    one.close()
    failures = 0
    for mode in ("p2p", "nccl"):
        bh = lpe_bh.BarnesHut(local)
        bh.set_stream(stream.cuda_stream)
        bh.set_shard(rank, world)
        bh.upload(x, y, vx, vy, m)
        view = bh.device_view()
        send = torch.as_tensor(CudaArray(view.xchg_send, 4 * view.xchg_chunk), device="cuda")
        recv = torch.as_tensor(CudaArray(view.xchg_recv, 4 * view.xchg_chunk * world), device="cuda")
        token = torch.zeros(1, device="cuda", dtype=torch.int32)
        if mode == "p2p":
            handles = [None] * world
            dist.all_gather_object(handles, bh.xchg_export())
            for r, h in enumerate(handles):
                if r != rank:
                    bh.xchg_import(r, h)
            assert bh.xchg_p2p_ready()
        with torch.cuda.stream(stream):
            for _ in range(steps):
                bh.step_begin(p)
                if mode == "p2p":
                    dist.all_reduce(token)
                else:
                    dist.all_gather_into_tensor(recv, send)
                bh.step_finish()
        got = bh.download()
        bad = [k for k in ("x", "y", "vx", "vy") if not np.array_equal(got[k], ref[k])]
        if bad:
            failures += 1
            print(f"[rank {rank}] {mode}: MISMATCH in {bad}", flush=True)
        else:
            print(f"[rank {rank}] {mode}: {steps} sharded steps of {n} bodies on {world} GPUs == unsharded, bit for bit",
                  flush=True)
        dist.barrier()
        bh.close()
    t = torch.tensor([failures], device="cuda")
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(1 if t.item() else 0)


if __name__ == "__main__":
    main()
