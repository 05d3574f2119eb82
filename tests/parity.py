"""Shared parity helpers for the tests: compare the CUDA path (through the C ABI) with the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "little-physics-engine_b200"))

import oracle_py  # noqa: E402
import lpe_bh  # noqa: E402


def deinterleave(key):
    """Morton key -> (ix, iy): digit bit0 = x half, bit1 = y half (reference getQuadrant, barnes_hut.hpp:121-131)."""
    key = np.asarray(key, dtype=np.uint64)

    def compact(v):
        v = v & np.uint64(0x5555555555555555)
        v = (v | (v >> np.uint64(1))) & np.uint64(0x3333333333333333)
        v = (v | (v >> np.uint64(2))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v >> np.uint64(4))) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v >> np.uint64(8))) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v >> np.uint64(16))) & np.uint64(0x00000000FFFFFFFF)
        return v

    return compact(key), compact(key >> np.uint64(1))


def hilbert_to_xy(h, D):
    """Hilbert index -> (ix, iy) at depth D (inverse of hilbert_index in csrc/bh_build.cuh), vectorised."""
    t = np.asarray(h, dtype=np.uint64).copy()
    x = np.zeros(len(t), np.int64)
    y = np.zeros(len(t), np.int64)
    for b in range(D):
        s = 1 << b
        rx = ((t >> np.uint64(1)) & np.uint64(1)).astype(np.int64)
        ry = ((t ^ rx.astype(np.uint64)) & np.uint64(1)).astype(np.int64)
        flip = (ry == 0) & (rx == 1)
        x = np.where(flip, s - 1 - x, x)
        y = np.where(flip, s - 1 - y, y)
        swap = ry == 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        x = x + s * rx
        y = y + s * ry
        t = t >> np.uint64(2)
    return x.astype(np.uint64), y.astype(np.uint64)


def keys_to_xy(keys, D, hilbert):
    """Depth-D cell coordinates of sort keys, whichever curve they are on."""
    return hilbert_to_xy(keys, D) if hilbert else deinterleave(keys)


def oracle_cells(nodes, U):
    """Index the oracle's dumped nodes by (level, ix, iy)."""
    level = np.rint(np.log2(U / nodes["bsize"])).astype(np.int64)
    ix = np.rint(nodes["bx"] / nodes["bsize"]).astype(np.int64)
    iy = np.rint(nodes["by"] / nodes["bsize"]).astype(np.int64)
    cells = {}
    nchild = {}
    for k in range(len(nodes)):
        key = (int(level[k]), int(ix[k]), int(iy[k]))
        cells[key] = k
        if level[k] > 0:
            pk = (int(level[k]) - 1, int(ix[k]) >> 1, int(iy[k]) >> 1)
            nchild[pk] = nchild.get(pk, 0) + 1
    return cells, nchild


def compare_tree(dump, nodes, U, rtol=1e-12):
    """GPU path-compressed tree (lpe_bh_dump_tree) vs the oracle's full quadtree dump. Returns a report dict.

    Every GPU branching cell must be an internal oracle node at the same (level, ix, iy) with the same mass / COM;
    every GPU leaf must be the oracle leaf holding the same body; the number of branching cells must equal the
    number of oracle internal nodes with >= 2 non-empty children above the depth bound.
    """
    D = dump["stats"]["depth"]
    cells, nchild = oracle_cells(nodes, U)
    lv = dump["node_level"]
    ix, iy = keys_to_xy(dump["node_key"], D, dump["stats"]["hilbert"])
    worst = 0.0
    n_branch = n_leaf = n_aggr = 0
    for i in range(len(lv)):
        L = int(lv[i])
        if L >= 0:
            sh = D - L
            key = (L, int(ix[i]) >> sh, int(iy[i]) >> sh)
            n_branch += 1
        elif L == -2:
            key = (D, int(ix[i]), int(iy[i]))
            n_aggr += 1
        else:
            n_leaf += 1
            key = None
        if key is not None:
            assert key in cells, f"GPU node {i} level {L} has no oracle cell {key}"
            o = nodes[cells[key]]
            assert o["is_leaf"] == 0, f"oracle cell {key} is a leaf but GPU node {i} is internal"
            assert o["single"] == dump["node_first"][i], (
                f"first occupant differs at {key}: oracle {o['single']} gpu {dump['node_first'][i]}")
        else:
            # the oracle leaf that holds this body: find via its position is expensive; use the 'single' index map
            o = None
        if o is not None:
            for a, b in ((o["mass"], dump["node_mass"][i]), (o["comx"], dump["node_comx"][i]),
                         (o["comy"], dump["node_comy"][i])):
                err = abs(a - b) / max(abs(a), 1e-300)
                worst = max(worst, err)
    # leaves: every in-tree body appears exactly once as a leaf or under an aggregated terminal
    leaf_bodies = dump["node_first"][lv == -1]
    single_of_leaf = {int(n["single"]): k for k, n in enumerate(nodes) if n["is_leaf"] == 1}
    for i in np.nonzero(lv == -1)[0]:
        b = int(dump["node_first"][i])
        if b in single_of_leaf:
            o = nodes[single_of_leaf[b]]
            assert o["mass"] == dump["node_mass"][i] and o["comx"] == dump["node_comx"][i] and o["comy"] == dump["node_comy"][i]
        else:
            # the oracle subdivided below the depth bound D: the body must then sit below a depth-D cell alone
            raise AssertionError(f"GPU leaf body {b} is not an oracle leaf")
    expect_branch = sum(1 for (L, _, _), c in nchild.items() if c >= 2 and L < D)
    assert worst <= rtol, f"aggregate mismatch {worst:.3e} > {rtol}"
    assert n_branch == expect_branch, f"branching cells: gpu {n_branch} oracle {expect_branch}"
    assert len(np.unique(leaf_bodies)) == len(leaf_bodies)
    return dict(worst_rel=worst, branching=n_branch, leaves=n_leaf, aggregated=n_aggr, depth=D)


def check_preorder(dump):
    """Structural invariants of the pre-order array: skip pointers nest, leaves skip to the next node."""
    skip = dump["node_skip"].astype(np.int64)
    lv = dump["node_level"]
    n = len(skip)
    assert np.all(skip > np.arange(n)) and np.all(skip <= n)
    assert np.all(skip[lv < 0] == np.nonzero(lv < 0)[0] + 1)
    if n:
        assert skip[0] == n
    # nesting: a child's subtree ends inside its parent's
    stack = []
    for i in range(n):
        while stack and stack[-1] <= i:
            stack.pop()
        if stack:
            assert skip[i] <= stack[-1], f"node {i} overruns its ancestor"
        stack.append(int(skip[i]))


def rel_err(test, ref, floor_frac=1e-3):
    """Per-body relative error of a 2-vector field, SURVEY.md §8(d): a body whose |ref| is below floor_frac of the
    median |ref| (a near-perfect cancellation of a few hundred terms) is judged against the median instead."""
    tx, ty = test
    rx, ry = ref
    mag = np.hypot(rx, ry)
    med = np.median(mag[mag > 0]) if np.any(mag > 0) else 1.0
    den = np.where(mag >= floor_frac * med, mag, med)
    err = np.hypot(tx - rx, ty - ry) / den
    norm = np.sqrt(np.sum((tx - rx) ** 2 + (ty - ry) ** 2) / max(np.sum(rx ** 2 + ry ** 2), 1e-300))
    return dict(max=float(err.max()) if len(err) else 0.0, median=float(np.median(err)) if len(err) else 0.0,
                p999=float(np.quantile(err, 0.999)) if len(err) else 0.0, norm=float(norm))


def gen_uniform(n, U, seed, mass_lo=0.5e6, mass_hi=1.5e6):
    rng = np.random.default_rng(seed)
    x = rng.random(n) * U
    y = rng.random(n) * U
    m = mass_lo + (mass_hi - mass_lo) * rng.random(n)
    vx = rng.standard_normal(n)
    vy = rng.standard_normal(n)
    return x, y, vx, vy, m
