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


def _balance_worker(rank, world, port, q):
    """Host logic of the domain-decomposed run's load balancer (lpe_bh.balanced_splitters): every rank contributes its
    per-chunk (first key, cost) arrays, gathers everybody's, and must arrive at the SAME splitters — the kernels of all
    ranks then agree on who owns which key (no GPU involved)."""
    sys.path.insert(0, os.path.join(ROOT, "little-physics-engine_b200"))
    import lpe_bh
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        nchunks = 3000 + 500 * rank
        lo, hi = rank * (1 << 59), (rank + 1) * (1 << 59)               # disjoint ascending key ranges (depth-30 keys)
        keys = np.sort(rng.integers(lo, hi, nchunks, dtype=np.int64)).astype(np.uint64)
        cost = (200 + 400 * rng.random(nchunks) * (1 + 3 * rank)).astype(np.uint32)   # rank 1's chunks are dearer
        allc = [None] * world
        dist.all_gather_object(allc, (keys, cost))
        split = lpe_bh.balanced_splitters(allc, world, beta=100.0)
        scaled = lpe_bh.balanced_splitters(allc, world, beta=100.0, scale=[1.0, 2.0][:world])
        allk = np.concatenate([k for k, _ in allc]); allw = np.concatenate([c for _, c in allc]).astype(np.float64) + 100.0
        share = [float(allw[(allk >= split[r]) & (allk < split[r + 1])].sum()) for r in range(world)]
        q.put((rank, split, scaled, share))
    finally:
        dist.destroy_process_group()


def test_balanced_splitters_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_balance_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, s0, sc0, share0), (_, s1, sc1, share1) = res
    assert s0 == s1 and sc0 == sc1                       # every rank computes the same splitters
    assert s0[0] == 0 and s0[-1] == 1 << 60 and all(a <= b for a, b in zip(s0, s0[1:]))
    assert abs(share0[0] - share0[1]) <= 0.01 * sum(share0)   # equal shares of cost + beta (to one chunk)
    assert sc0[1] > s0[1]                                # charging rank 1's chunks double moves the border into its range
