"""CPU, world_size 2 over gloo: the multi-GPU step's host logic — block-cyclic ownership of Morton-sorted bodies,
the packed exchange layout (lpe_bh_shard_owner / lpe_bh_shard_chunk) and the allgather — with the oracle standing
in for the GPU force phase. Every rank must end with the single-process result, bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _morton_order(x, y, U, D=12):
    h = U / (1 << D)
    ix = np.clip((x / h).astype(np.int64), 0, (1 << D) - 1)
    iy = np.clip((y / h).astype(np.int64), 0, (1 << D) - 1)
    key = np.zeros(len(x), np.int64)
    for b in range(D):
        key |= ((ix >> b) & 1) << (2 * b)
        key |= ((iy >> b) & 1) << (2 * b + 1)
    return np.argsort(key, kind="stable")


def _worker(rank, world, port, n, q):
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "little-physics-engine_b200")):
        sys.path.insert(0, p)
    import oracle_py as O
    import lpe_bh
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        U = 1024.0
        x, y = rng.random(n) * U, rng.random(n) * U
        vx, vy = rng.standard_normal(n), rng.standard_normal(n)
        m = 1e6 * (0.5 + rng.random(n))
        p = O.make_params(U, U / 2 ** 12, dt_kick=1 / 120, dt_drift=0.005)
        full = O.PortLib().run(p, x, y, vx, vy, m, threads=1)          # what one process would produce
        order = _morton_order(x, y, U)                                 # sorted position -> body
        chunk = lpe_bh.shard_chunk(n, world)
        send = torch.zeros(chunk, 4, dtype=torch.float64)
        mine = 0
        for pos, b in enumerate(order):
            r, slot = lpe_bh.shard_owner(pos, world)
            if r == rank:                                               # this rank "computed" body b
                send[slot] = torch.tensor([full["x"][b], full["y"][b], full["vx"][b], full["vy"][b]])
                mine += 1
        recv = torch.zeros(world * chunk, 4, dtype=torch.float64)
        dist.all_gather_into_tensor(recv, send)
        out = np.zeros((n, 4))
        for pos, b in enumerate(order):
            r, slot = lpe_bh.shard_owner(pos, world)
            out[b] = recv[r * chunk + slot].numpy()
        ok = (np.array_equal(out[:, 0], full["x"]) and np.array_equal(out[:, 1], full["y"]) and
              np.array_equal(out[:, 2], full["vx"]) and np.array_equal(out[:, 3], full["vy"]))
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([mine]))
        q.put((rank, bool(ok), [int(c) for c in counts], chunk))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [5000, 2048 * 3 + 17])
def test_block_cyclic_exchange_world2(n):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, counts, chunk in res:
        assert ok, f"rank {rank} did not reconstruct the single-process state"
        assert sum(counts) == n                      # every body owned exactly once
        assert max(counts) <= chunk
        assert max(counts) - min(counts) <= 2048     # balanced to within one block
