// lpe_bh_dd.inl — host side of the domain-decomposed step (kernels: bh_dd.cuh). Included by lpe_bh.cu.
//
// One context = one rank = one GPU. Every rank allocates the same peer-visible WINDOW (header with mailboxes and the
// migrant counter, root tables, both sets of state buffers, the record array with one import region per sender) and
// learns every other rank's window: a raw device pointer when the ranks share a process (lpe_bh_dd_set_peer), a CUDA
// IPC handle when there is one process per GPU (lpe_bh_dd_export / lpe_bh_dd_import). From then on a step is three
// phases on the rank's stream with an in-stream flag barrier after the first two; all ranks must step in lockstep.

namespace {

struct DDLayout {
    size_t hdr, roots, body[2], vel[2], orig[2], rec, xrec, total;
    unsigned int nblocks, importBase;
};
DDLayout dd_layout(uint64_t S, int R, unsigned int icap) {
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    DDLayout L{};
    size_t o = 0;
    L.hdr = o; o = al(o + sizeof(DDHeader));
    L.roots = o; o = al(o + sizeof(DDRoot) * (size_t)DD_MAXROOTS * LPE_MAX_P2P);
    for (int b = 0; b < 2; ++b) {
        L.body[b] = o; o = al(o + sizeof(Body) * S);
        L.vel[b] = o; o = al(o + sizeof(double2) * S);
        L.orig[b] = o; o = al(o + sizeof(unsigned int) * S);
    }
    L.importBase = DD_TOPCAP + (unsigned int)S + 8u;
    L.nblocks = L.importBase + (unsigned int)R * icap;
    L.rec = o; o = al(o + sizeof(TravRec) * 4 * (size_t)L.nblocks);
    L.xrec = o; o = al(o + sizeof(double4) * 4 * (size_t)L.nblocks);
    L.total = o;
    return L;
}

void dd_close_peers(lpe_bh_ctx* c) {
    for (int r = 0; r < LPE_MAX_P2P; ++r) {
        if (c->dd_peer_opened[r]) cudaIpcCloseMemHandle(c->dd_peer_opened[r]);
        c->dd_peer_opened[r] = nullptr;
        c->dd_peer_win[r] = nullptr;
    }
}

void dd_release(lpe_bh_ctx* c) {
    if (!c->dd && !c->dd_win) return;
    dd_close_peers(c);
    if (c->dd_win) cudaFree(c->dd_win);
    c->dd_win = nullptr;
    c->dd_win_bytes = 0;
    c->dd = false;
    c->dd_R = 1; c->dd_rank = 0;
    c->body = c->body2 = nullptr; c->vel = c->vel2 = nullptr; c->orig = c->orig2 = nullptr;
    c->rec = nullptr; c->dd_xrec = nullptr;
    c->dd_dom = nullptr; c->dd_myroots = nullptr; c->dd_queue = nullptr; c->dd_oob = nullptr; c->dd_payload = nullptr;
    c->dd_chunk_cost = nullptr; c->dd_top = DDTop{};
    c->dd_dom_depth = -1;
    c->n = 0;
}

bool dd_ready(const lpe_bh_ctx* c) {
    if (!c->dd) return false;
    for (int r = 0; r < c->dd_R; ++r)
        if (!c->dd_peer_win[r]) return false;
    return true;
}

DDPeers dd_peers(const lpe_bh_ctx* c, int parity) {
    const DDLayout L = dd_layout(c->cap, c->dd_R, c->dd_icap);
    DDPeers P{};
    for (int r = 0; r < c->dd_R; ++r) {
        char* w = c->dd_peer_win[r];
        P.hdr[r] = reinterpret_cast<DDHeader*>(w + L.hdr);
        P.body[r] = reinterpret_cast<Body*>(w + L.body[parity]);
        P.vel[r] = reinterpret_cast<double2*>(w + L.vel[parity]);
        P.orig[r] = reinterpret_cast<unsigned int*>(w + L.orig[parity]);
        P.roots[r] = reinterpret_cast<DDRoot*>(w + L.roots);
        P.rec[r] = reinterpret_cast<TravRec*>(w + L.rec);
        P.xrec[r] = reinterpret_cast<double4*>(w + L.xrec);
    }
    return P;
}
DDHeader* dd_hdr(const lpe_bh_ctx* c) { return reinterpret_cast<DDHeader*>(c->dd_win); }

// ---- cell <-> key on the host (must agree with the device's hilbert_index / Morton interleave) -----------------------
unsigned int compact_bits(unsigned long long v) {
    v &= 0x5555555555555555ull;
    v = (v | (v >> 1)) & 0x3333333333333333ull;
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FFull;
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFFull;
    v = (v | (v >> 16)) & 0x00000000FFFFFFFFull;
    return (unsigned int)v;
}
unsigned long long spread_bits_host(unsigned int v) {
    unsigned long long x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}
void key_to_cell(unsigned long long key, int level, int hilbert, unsigned int& x, unsigned int& y) {
    if (!hilbert) {
        x = compact_bits(key);
        y = compact_bits(key >> 1);
        return;
    }
    x = y = 0;
    unsigned long long t = key;
    for (int b = 0; b < level; ++b) {
        const unsigned int s = 1u << b;
        const unsigned int rx = 1u & (unsigned int)(t >> 1);
        const unsigned int ry = 1u & ((unsigned int)t ^ rx);
        if (ry == 0u) {
            if (rx == 1u) { x = s - 1u - x; y = s - 1u - y; }
            const unsigned int tmp = x; x = y; y = tmp;
        }
        x += s * rx;
        y += s * ry;
        t >>= 2;
    }
}
unsigned long long cell_to_key(unsigned int x, unsigned int y, int level, int hilbert) {
    return hilbert ? hilbert_index(x, y, level) : (spread_bits_host(x) | (spread_bits_host(y) << 1));
}

// depth-D splitters from the depth-30 ones (a depth-D cell is never split: round up to the next cell)
void dd_split_at_depth(const lpe_bh_ctx* c, int D, unsigned long long* out /* R + 1 */) {
    const int sh = 2 * (LPE_MAX_DEPTH - D);
    out[0] = 0ull;
    for (int r = 1; r < c->dd_R; ++r) {
        const unsigned long long k30 = c->dd_split30[r];
        unsigned long long kd = sh ? ((k30 + ((1ull << sh) - 1ull)) >> sh) : k30;
        if (kd < out[r - 1]) kd = out[r - 1];
        out[r] = kd;
    }
    out[c->dd_R] = 1ull << (2 * D);
    for (int r = 1; r < c->dd_R; ++r)
        if (out[r] > out[c->dd_R]) out[r] = out[c->dd_R];
}

DDSplit dd_split(const lpe_bh_ctx* c, int D) {
    unsigned long long kd[LPE_MAX_P2P + 1];
    dd_split_at_depth(c, D, kd);
    DDSplit sp{};
    for (int r = 0; r < c->dd_R; ++r) sp.k[r] = kd[r];
    for (int r = c->dd_R; r <= LPE_MAX_P2P; ++r) sp.k[r] = ~0ull;   // bodies outside the tree belong to the last rank
    sp.me = c->dd_rank;
    sp.R = c->dd_R;
    return sp;
}

// every rank's key range as maximal aligned quadrants, with their boxes: the domain a cell is tested against
int dd_build_domain(lpe_bh_ctx* c, const StepConst& k) {
    if (c->dd_dom_depth == k.D) return 0;
    c->dd_dom_host.resize(sizeof(DDDomain));
    DDDomain* H = reinterpret_cast<DDDomain*>(c->dd_dom_host.data());
    std::memset(H, 0, sizeof(DDDomain));
    unsigned long long kd[LPE_MAX_P2P + 1];
    dd_split_at_depth(c, k.D, kd);
    for (int r = 0; r < c->dd_R; ++r) {
        unsigned long long a = kd[r];
        const unsigned long long e = kd[r + 1];
        int nq = 0;
        double bx0 = 1e300, by0 = 1e300, bx1 = -1e300, by1 = -1e300;
        while (a < e) {
            int j = 0;
            while (j < k.D && (a & ((1ull << (2 * (j + 1))) - 1ull)) == 0ull && a + (1ull << (2 * (j + 1))) <= e) ++j;
            if (nq >= DD_MAXQ) return fail(c, "domain decomposition: too many quadrants in one key range (internal error)");
            DDQuad& q = H->q[r][nq++];
            q.key = a;
            q.level = k.D - j;
            unsigned int ix, iy;
            key_to_cell(a >> (2 * j), q.level, k.hilbert, ix, iy);
            // the bounds keygen itself uses: cell i of the finest level spans [fl(i*h), fl((i+1)*h))
            q.x0 = (double)((unsigned long long)ix << j) * k.h;
            q.x1 = (double)(((unsigned long long)ix + 1ull) << j) * k.h;
            q.y0 = (double)((unsigned long long)iy << j) * k.h;
            q.y1 = (double)(((unsigned long long)iy + 1ull) << j) * k.h;
            bx0 = std::min(bx0, q.x0); by0 = std::min(by0, q.y0); bx1 = std::max(bx1, q.x1); by1 = std::max(by1, q.y1);
            a += 1ull << (2 * j);
        }
        H->nq[r] = nq;
        H->box[r][0] = bx0; H->box[r][1] = by0; H->box[r][2] = bx1; H->box[r][3] = by1;
        // tree of boxes over the quadrants (scaled units, floats rounded outward)
        auto down = [](double v) { float f = (float)v; return ((double)f > v) ? std::nextafterf(f, -INFINITY) : f; };
        auto up = [](double v) { float f = (float)v; return ((double)f < v) ? std::nextafterf(f, INFINITY) : f; };
        float4* B = H->bvh[r];
        for (int j = 0; j < DD_BVH_LEAVES; ++j) {
            if (j < nq) {
                const DDQuad& q = H->q[r][j];
                B[DD_BVH_LEAVES + j] = make_float4(down(q.x0 * k.invS), down(q.y0 * k.invS), up(q.x1 * k.invS), up(q.y1 * k.invS));
            } else {
                B[DD_BVH_LEAVES + j] = make_float4(1.f, 1.f, 0.f, 0.f);   // empty
            }
        }
        for (int n = DD_BVH_LEAVES - 1; n >= 1; --n) {
            const float4 l = B[2 * n], rr = B[2 * n + 1];
            const bool le = l.x > l.z, re = rr.x > rr.z;
            B[n] = le ? rr : re ? l : make_float4(std::min(l.x, rr.x), std::min(l.y, rr.y), std::max(l.z, rr.z), std::max(l.w, rr.w));
        }
        B[0] = make_float4(1.f, 1.f, 0.f, 0.f);
    }
    // (pageable source: the copy is staged before the call returns, so the host table may be rebuilt right away)
    CU_TRY(c, cudaMemcpyAsync(c->dd_dom, H, sizeof(DDDomain), cudaMemcpyHostToDevice, c->stream));
    c->dd_dom_depth = k.D;
    return 0;
}

__global__ void k_dd_payload(int point, const unsigned long long* __restrict__ oob, const Scal* __restrict__ s,
                             double* __restrict__ payload) {
    if (threadIdx.x != 0) return;
    double* p = payload + 6 * point;
    if (point == 0) {   // box of the targets outside the tree that this rank held at the start of the step
        const bool any = oob[0] != ~0ull;
        p[0] = any ? ordered_value(oob[0]) : 1.0;
        p[1] = any ? ordered_value(oob[1]) : 1.0;
        p[2] = any ? ordered_value(oob[2]) : 0.0;
        p[3] = any ? ordered_value(oob[3]) : 0.0;
        p[4] = p[5] = 0.0;
    } else {
        p[0] = (double)s->n_live; p[1] = (double)s->n_in;
        p[2] = p[3] = p[4] = p[5] = 0.0;
    }
}

int dd_make_const(lpe_bh_ctx* c, const lpe_bh_params& p, StepConst& k) {
    if (!c->dd) return fail(c, "context is not in domain-decomposed mode (lpe_bh_dd_init + lpe_bh_dd_upload first)");
    if (!dd_ready(c)) return fail(c, "domain decomposition: not every rank's window is known (lpe_bh_dd_import / lpe_bh_dd_set_peer)");
    if (c->dd_hilbert < 0) return fail(c, "domain decomposition: no bodies uploaded (lpe_bh_dd_upload)");
    if (make_const(c, p, k)) return 1;
    if (k.hilbert != c->dd_hilbert) return fail(c, "domain decomposition: key_order differs from the one the bodies were distributed with");
    if (p.universe_size != c->dd_U) return fail(c, "domain decomposition: universe_size differs from the one the bodies were distributed with");
    k.dd = 1;
    k.blockBase = DD_TOPCAP;
    k.recSlots = 4u * dd_layout(c->cap, c->dd_R, c->dd_icap).nblocks;
    k.shard_rank = 0;
    k.shard_n = 1;
    k.n = (int)c->cap;
    return 0;
}

int dd_phase(lpe_bh_ctx* c, const lpe_bh_params& p, int phase) {
    StepConst k;
    if (dd_make_const(c, p, k)) return 1;
    cudaStream_t st = c->stream;
    const int S = (int)c->cap;
    const bool timing = c->instr & 1;
    const DDSplit sp = dd_split(c, k.D);
    DDHeader* hdr = dd_hdr(c);
    if (phase == 0) {
        if (dd_build_domain(c, k)) return 1;
        ++c->dd_epoch;
        k_dd_next_step<<<1, 1, 0, st>>>(c->dd_step_dev);
        if (timing) cudaEventRecord(c->dd_ev[0], st);
        if (step_prologue(c, S)) return 1;
        CU_TRY(c, cudaMemsetAsync(c->dd_oob, 0xFF, 16, st));
        CU_TRY(c, cudaMemsetAsync(c->dd_oob + 2, 0, 16 + sizeof(unsigned int) * LPE_MAX_P2P * DD_EXPORT_ROUNDS, st));   // + the exporter's round flags
        const DDPeers peers = dd_peers(c, c->dd_cur);
        k_dd_keygen<<<cdiv(S, 256), 256, 0, st>>>(k, sp, c->body, c->vel, c->orig, c->keys[0], c->vals[0], c->scal, hdr,
                                                  peers, c->dd_oob);
        k_dd_payload<<<1, 32, 0, st>>>(0, c->dd_oob, c->scal, c->dd_payload);
        if (timing) cudaEventRecord(c->dd_ev[1], st);
        k_dd_signal<<<1, 32, 0, st>>>(0, c->dd_step_dev, c->dd_rank, c->dd_R, peers, c->dd_payload);
        c->launches += 3;
    } else if (phase == 1) {
        k_dd_wait<<<1, 32, 0, st>>>(0, c->dd_step_dev, c->dd_R, hdr);
        if (timing) cudaEventRecord(c->dd_ev[2], st);
        k_dd_keygen_inbox<<<64, 256, 0, st>>>(k, sp, c->body, c->keys[0], c->vals[0], c->scal, hdr);
        if (step_sort(c, k, S, &c->scal->n_sort)) return 1;
        if (timing) cudaEventRecord(c->dd_ev[3], st);
        if (step_build(c, k, S, &c->scal->n_sort)) return 1;
        c->dd_cur ^= 1;
        if (timing) cudaEventRecord(c->dd_ev[4], st);
        const DDLayout L = dd_layout(c->cap, c->dd_R, c->dd_icap);
        const DDPeers peers = dd_peers(c, c->dd_cur);
        DDRootsIn ri{c->tkey, c->tfirst, c->mask, c->P, c->agg, c->body};
        k_dd_roots<<<1, 256, 0, st>>>(k, sp, c->dd_dom, ri, c->dd_myroots, c->scal, hdr);
        DDExportArgs ea{c->dd_myroots, c->child, c->rec, c->agg, c->body, c->dd_queue, c->dd_pushed, c->dd_icap, L.importBase};
        {   // one cluster of DD_EXPORT_CLUSTER blocks per destination
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(DD_EXPORT_CLUSTER, c->dd_R);
            cfg.blockDim = dim3(DD_EXPORT_THREADS);
            cfg.dynamicSmemBytes = 0;
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = DD_EXPORT_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            const DDDomain* domc = c->dd_dom;
            CU_TRY(c, cudaLaunchKernelEx(&cfg, k_dd_export, k, sp, domc, ea, peers, c->scal, hdr));
        }
        if (c->dd_R > 1) k_dd_export_x<<<dim3(16, c->dd_R), 256, 0, st>>>(k, c->dd_rank, ea, peers, c->scal);
        k_dd_payload<<<1, 32, 0, st>>>(1, c->dd_oob, c->scal, c->dd_payload);
        if (timing) cudaEventRecord(c->dd_ev[5], st);
        k_dd_signal<<<1, 32, 0, st>>>(1, c->dd_step_dev, c->dd_rank, c->dd_R, peers, c->dd_payload + 6);
        c->launches += 7;
    } else if (phase == 2) {
        k_dd_wait<<<1, 32, 0, st>>>(1, c->dd_step_dev, c->dd_R, hdr);
        if (timing) cudaEventRecord(c->dd_ev[6], st);
        const DDLayout L = dd_layout(c->cap, c->dd_R, c->dd_icap);
        k_dd_top<<<1, 1024, DD_TOP_SMEM_BYTES, st>>>(k, c->dd_rank, c->dd_R, hdr, reinterpret_cast<const DDRoot*>(c->dd_win + L.roots), c->dd_top,
                                     c->rec, c->dd_xrec, c->selfslot, c->scal);
        if (timing) cudaEventRecord(c->dd_ev[7], st);
        if (step_traverse(c, k, p, S, false)) return 1;
        if (timing) cudaEventRecord(c->dd_ev[8], st);
        c->launches += 2;
        c->last_c = k;
        c->have_step = true;
        c->last.depth = k.D;
        c->last.hilbert = k.hilbert;
    } else {
        return fail(c, "phase must be 0, 1 or 2");
    }
    CU_TRY(c, cudaGetLastError());
    return 0;
}

// faults raised on the device since the last check (synchronise first)
int dd_check_fault(lpe_bh_ctx* c) {
    if (!c->dd || !c->dd_win) return 0;
    unsigned int f = 0, ovf = 0;
    CU_TRY(c, cudaMemcpy(&f, &dd_hdr(c)->fault, sizeof(f), cudaMemcpyDeviceToHost));
    (void)ovf;
    if (!f) return 0;
    CU_TRY(c, cudaMemset(&dd_hdr(c)->fault, 0, sizeof(unsigned int)));
    std::string m = "domain decomposition fault:";
    if (f & 1u) m += " more migrants than free slots (raise the capacity or rebalance);";
    if (f & 2u) m += " more exported child blocks than import capacity (raise import_blocks);";
    if (f & 4u) m += " root table overflow;";
    if (f & 8u) m += " barrier timed out (a peer did not reach it);";
    if (f & 16u) m += " a migrant arrived outside its owner's key range (ranks out of lockstep?);";
    if (f & 32u) m += " traversal frontier overflow (not supported in this mode);";
    return fail(c, m + " results of the last step are invalid");
}

}  // namespace

extern "C" {

int lpe_bh_dd_init(lpe_bh_ctx* c, int rank, int nranks, uint64_t capacity, uint32_t import_blocks) {
    if (!c) return 1;
    if (nranks < 1 || nranks > LPE_MAX_P2P || rank < 0 || rank >= nranks) return fail(c, "domain decomposition: 1..8 ranks, 0 <= rank < nranks");
    if (capacity == 0 || capacity > LPE_MAX_BODIES) return fail(c, "domain decomposition: capacity out of range");
    if (c->shard_n > 1) return fail(c, "context is in replicated-tree sharded mode (lpe_bh_set_shard): reset it to 1 rank first");
    DevGuard _dg(c->device);
    if (ensure_capacity(c, capacity, true)) return 1;
    const uint64_t S = c->cap;
    c->dd_icap = import_blocks ? import_blocks : (unsigned int)std::max<uint64_t>(65536, S / 8);
    c->dd_R = nranks;
    c->dd_rank = rank;
    const DDLayout L = dd_layout(S, nranks, c->dd_icap);
    void* w = nullptr;
    cudaError_t e = cudaMalloc(&w, L.total);
    if (e != cudaSuccess) {
        free_all(c);
        return fail(c, std::string("cudaMalloc (window): ") + cudaGetErrorString(e));
    }
    c->dd_win = static_cast<char*>(w);
    c->dd_win_bytes = L.total;
    c->dd = true;
    c->dd_cur = 0;
    c->dd_epoch = 0;
    c->dd_hilbert = -1;
    c->dd_dom_depth = -1;
    cudaMemsetAsync(c->dd_win, 0, L.body[0], c->stream);   // header + root tables
    c->body = reinterpret_cast<Body*>(c->dd_win + L.body[0]);   c->body2 = reinterpret_cast<Body*>(c->dd_win + L.body[1]);
    c->vel = reinterpret_cast<double2*>(c->dd_win + L.vel[0]);  c->vel2 = reinterpret_cast<double2*>(c->dd_win + L.vel[1]);
    c->orig = reinterpret_cast<unsigned int*>(c->dd_win + L.orig[0]);
    c->orig2 = reinterpret_cast<unsigned int*>(c->dd_win + L.orig[1]);
    c->rec = reinterpret_cast<TravRec*>(c->dd_win + L.rec);
    c->dd_xrec = reinterpret_cast<double4*>(c->dd_win + L.xrec);
    int rc = 0;
    rc |= dalloc(c, c->dd_dom, 1) | dalloc(c, c->dd_myroots, DD_MAXROOTS) | dalloc(c, c->dd_queue, (size_t)nranks * c->dd_icap) |
          dalloc(c, c->dd_oob, 4 + LPE_MAX_P2P * DD_EXPORT_ROUNDS / 2 + 2) | dalloc(c, c->dd_payload, 12) | dalloc(c, c->dd_chunk_cost, S / 32 + 64) |
          dalloc(c, c->dd_step_dev, 1);
    if (!rc) cudaMemsetAsync(c->dd_step_dev, 0, sizeof(unsigned long long), c->stream);
    // (set on every init: the attribute is kept per device)
    if (cudaFuncSetAttribute(k_dd_top, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DD_TOP_SMEM_BYTES) != cudaSuccess) rc = 1;
    c->dd_pushed = reinterpret_cast<unsigned int*>(c->dd_oob + 4);
    rc |= dalloc(c, c->dd_top.mask, DD_TOPROOTS) | dalloc(c, c->dd_top.P, DD_TOPROOTS + 1) |
          dalloc(c, c->dd_top.wstart, DD_TOPROOTS) | dalloc(c, c->dd_top.child, 4 * DD_TOPROOTS) |
          dalloc(c, c->dd_top.cellLevel, DD_TOPROOTS) | dalloc(c, c->dd_top.agg, DD_TOPROOTS) |
          dalloc(c, c->dd_top.delta, DD_TOPROOTS);
    if (rc) {
        free_all(c);
        dd_release(c);
        return 1;
    }
    for (auto& ev : c->dd_ev)
        if (!ev) cudaEventCreate(&ev);
    c->dd_peer_win[rank] = c->dd_win;
    k_dd_poison<<<1, 32, 0, c->stream>>>(c->rec);
    c->n = 0;
    c->have_step = false;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    CU_TRY(c, cudaGetLastError());
    return 0;
}

int lpe_bh_dd_export(lpe_bh_ctx* c, void* handle64) {
    if (!c || !handle64) return 1;
    if (!c->dd) return fail(c, "context is not in domain-decomposed mode");
    DevGuard _dg(c->device);
    cudaIpcMemHandle_t hnd;
    CU_TRY(c, cudaIpcGetMemHandle(&hnd, c->dd_win));
    std::memcpy(handle64, &hnd, sizeof(hnd));
    return 0;
}

int lpe_bh_dd_import(lpe_bh_ctx* c, int rank, const void* handle64) {
    if (!c || !handle64) return 1;
    if (!c->dd) return fail(c, "context is not in domain-decomposed mode");
    if (rank < 0 || rank >= c->dd_R || rank == c->dd_rank) return fail(c, "bad peer rank");
    DevGuard _dg(c->device);
    cudaIpcMemHandle_t hnd;
    std::memcpy(&hnd, handle64, sizeof(hnd));
    void* ptr = nullptr;
    CU_TRY(c, cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
    if (c->dd_peer_opened[rank]) cudaIpcCloseMemHandle(c->dd_peer_opened[rank]);
    c->dd_peer_opened[rank] = ptr;
    c->dd_peer_win[rank] = static_cast<char*>(ptr);
    return 0;
}

void* lpe_bh_dd_window(lpe_bh_ctx* c) { return (c && c->dd) ? c->dd_win : nullptr; }

int lpe_bh_dd_set_peer(lpe_bh_ctx* c, int rank, void* window, int peer_device) {
    if (!c) return 1;
    if (!c->dd) return fail(c, "context is not in domain-decomposed mode");
    if (rank < 0 || rank >= c->dd_R || !window) return fail(c, "bad peer rank / window");
    if (peer_device >= 0 && peer_device != c->device) {   // ranks of one process on different devices: direct peer access
        DevGuard _dg(c->device);
        int can = 0;
        CU_TRY(c, cudaDeviceCanAccessPeer(&can, c->device, peer_device));
        if (!can) return fail(c, "no peer access between the two devices");
        const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(c, cudaGetErrorString(e));
        cudaGetLastError();
    }
    c->dd_peer_win[rank] = static_cast<char*>(window);
    return 0;
}

int lpe_bh_dd_ready(const lpe_bh_ctx* c) { return (c && dd_ready(c)) ? 1 : 0; }

// Every rank is handed the WHOLE input (creation order) and keeps the bodies of its own key range. The splitters come
// from a strided sample (equal body counts; lpe_bh_dd_set_splitters re-balances later), identical on every rank.
int lpe_bh_dd_upload(lpe_bh_ctx* c, const lpe_bh_params* p, uint64_t n, const double* x, const double* y, const double* vx,
                     const double* vy, const double* m, const uint32_t* rank, const uint8_t* comp) {
    if (!c || !p) return 1;
    if (!c->dd) return fail(c, "context is not in domain-decomposed mode (lpe_bh_dd_init first)");
    if (n > LPE_MAX_BODIES * (uint64_t)LPE_MAX_P2P) return fail(c, "too many bodies");
    if (n && (!x || !y || !m)) return fail(c, "x, y and m are required");
    DevGuard _dg(c->device);
    cudaStream_t st = c->stream;
    CU_TRY(c, cudaStreamSynchronize(st));
    const uint64_t S = c->cap;
    c->n = S;
    c->have_step = false;
    c->dd_n_total = n;
    c->dd_U = p->universe_size;
    c->dd_dom_depth = -1;
    c->dd_cur = 0;
    {   // back to buffer set 0
        const DDLayout L = dd_layout(S, c->dd_R, c->dd_icap);
        c->body = reinterpret_cast<Body*>(c->dd_win + L.body[0]);   c->body2 = reinterpret_cast<Body*>(c->dd_win + L.body[1]);
        c->vel = reinterpret_cast<double2*>(c->dd_win + L.vel[0]);  c->vel2 = reinterpret_cast<double2*>(c->dd_win + L.vel[1]);
        c->orig = reinterpret_cast<unsigned int*>(c->dd_win + L.orig[0]);
        c->orig2 = reinterpret_cast<unsigned int*>(c->dd_win + L.orig[1]);
    }
    lpe_bh_params p30 = *p;
    p30.max_depth = LPE_MAX_DEPTH;
    StepConst k;
    if (make_const(c, p30, k)) return 1;
    c->dd_hilbert = k.hilbert;
    // ---- splitters: quantiles of the keys of every stride-th body ----
    const uint64_t stride = std::max<uint64_t>(1, n / 262144);
    const uint64_t ns = n ? (n + stride - 1) / stride : 0;
    std::vector<unsigned long long> skeys(ns);
    if (ns) {
        std::vector<double> sx(ns), sy(ns);
        std::vector<unsigned char> sc(ns);
        for (uint64_t i = 0; i < ns; ++i) { sx[i] = x[i * stride]; sy[i] = y[i * stride]; sc[i] = comp ? comp[i * stride] : 3; }
        for (uint64_t off = 0; off < ns; off += S) {
            const uint64_t cnt = std::min<uint64_t>(S, ns - off);
            CU_TRY(c, cudaMemcpyAsync(c->tmp, sx.data() + off, 8 * cnt, cudaMemcpyHostToDevice, st));
            CU_TRY(c, cudaMemcpyAsync(c->tmp + S, sy.data() + off, 8 * cnt, cudaMemcpyHostToDevice, st));
            CU_TRY(c, cudaMemcpyAsync(c->comp_in, sc.data() + off, cnt, cudaMemcpyHostToDevice, st));
            k_dd_sample_keys<<<cdiv((long long)cnt, 256), 256, 0, st>>>(k, (int)cnt, c->tmp, c->tmp + S, c->comp_in, c->keys[0]);
            CU_TRY(c, cudaMemcpyAsync(skeys.data() + off, c->keys[0], 8 * cnt, cudaMemcpyDeviceToHost, st));
            CU_TRY(c, cudaStreamSynchronize(st));
        }
    }
    std::vector<unsigned long long> in;
    in.reserve(ns);
    for (unsigned long long v : skeys)
        if (v < (1ull << (2 * LPE_MAX_DEPTH))) in.push_back(v);
    std::sort(in.begin(), in.end());
    c->dd_split30[0] = 0ull;
    for (int r = 1; r < c->dd_R; ++r)
        c->dd_split30[r] = in.empty() ? ((1ull << (2 * LPE_MAX_DEPTH)) / (unsigned long long)c->dd_R) * (unsigned long long)r
                                      : in[(size_t)((unsigned long long)r * in.size() / (unsigned long long)c->dd_R)];
    c->dd_split30[c->dd_R] = 1ull << (2 * LPE_MAX_DEPTH);
    // ---- selection: chunks of the input, a scan of "mine" flags whose sink writes the kept bodies ----
    DDSplit sp{};
    for (int r = 0; r < c->dd_R; ++r) sp.k[r] = c->dd_split30[r];
    for (int r = c->dd_R; r <= LPE_MAX_P2P; ++r) sp.k[r] = ~0ull;
    sp.me = c->dd_rank; sp.R = c->dd_R;
    CU_TRY(c, cudaMemsetAsync(c->scal, 0, sizeof(Scal), st));
    CU_TRY(c, cudaMemsetAsync(c->totals, 0, sizeof(unsigned int) * (512 * SORT_MAX_PASSES + 16), st));
    CU_TRY(c, cudaMemsetAsync(c->lbstatus, 0, sizeof(unsigned long long) *
                                  ((size_t)cdiv((long long)c->cap, SORT_TILE) * (256 * (SORT_MAX_PASSES - 1) + 512) +
                                   2 * ((size_t)cdiv((long long)c->cap + 1, SCAN_TILE) + 2)), st));
    CU_TRY(c, cudaMemsetAsync(c->epoch_dev, 0, sizeof(unsigned int), st));
    c->epoch = 0u;
    unsigned int kept = 0;
    unsigned int* total = c->totals + 512 * SORT_MAX_PASSES + 12;   // a free word of the per-step scratch
    unsigned int* ticket = c->totals + 512 * SORT_MAX_PASSES + SORT_MAX_PASSES + 1;
    unsigned int* sfault = c->totals + 512 * SORT_MAX_PASSES + SORT_MAX_PASSES;
    unsigned long long* scanStatus = c->lbstatus + (size_t)cdiv((long long)c->cap, SORT_TILE) * (256 * (SORT_MAX_PASSES - 1) + 512);
    for (uint64_t off = 0; off < n; off += S) {
        const uint64_t cnt = std::min<uint64_t>(S, n - off);
        double* t = c->tmp;
        CU_TRY(c, cudaMemcpyAsync(t, x + off, 8 * cnt, cudaMemcpyHostToDevice, st));
        CU_TRY(c, cudaMemcpyAsync(t + S, y + off, 8 * cnt, cudaMemcpyHostToDevice, st));
        CU_TRY(c, cudaMemcpyAsync(t + 2 * S, m + off, 8 * cnt, cudaMemcpyHostToDevice, st));
        if (vx) CU_TRY(c, cudaMemcpyAsync(t + 3 * S, vx + off, 8 * cnt, cudaMemcpyHostToDevice, st));
        if (vy) CU_TRY(c, cudaMemcpyAsync(t + 4 * S, vy + off, 8 * cnt, cudaMemcpyHostToDevice, st));
        if (rank) CU_TRY(c, cudaMemcpyAsync(c->rank_in, rank + off, 4 * cnt, cudaMemcpyHostToDevice, st));
        if (comp) CU_TRY(c, cudaMemcpyAsync(c->comp_in, comp + off, cnt, cudaMemcpyHostToDevice, st));
        k_dd_max_mass<<<cdiv((long long)cnt, 256), 256, 0, st>>>((int)cnt, t + 2 * S, comp ? c->comp_in : nullptr, c->scal);
        DDSelIn in{t, t + S, vx ? t + 3 * S : nullptr, vy ? t + 4 * S : nullptr, t + 2 * S, rank ? c->rank_in : nullptr,
                   comp ? c->comp_in : nullptr, (unsigned int)off, (unsigned int)n};
        ++c->epoch;
        k_epoch_next<<<1, 1, 0, st>>>(c->epoch_dev);
        CU_TRY(c, cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st));
        k_scan_chained<<<cdiv((long long)cnt + 1, SCAN_TILE), SCAN_THREADS, 0, st>>>(
            DDSelLoad{k, sp, in}, DDSelSink{in, c->body, c->vel, c->orig, kept, (unsigned int)S, total, (int)cnt}, (int)cnt,
            scanStatus, c->epoch_dev, ticket, sfault);
        unsigned int got = 0;
        CU_TRY(c, cudaMemcpyAsync(&got, total, sizeof(got), cudaMemcpyDeviceToHost, st));
        CU_TRY(c, cudaStreamSynchronize(st));
        kept += got;
        if (kept > S) return fail(c, "domain decomposition: this rank's share of the bodies exceeds its capacity");
    }
    DDHeader h{};
    h.n_live = kept;
    CU_TRY(c, cudaMemcpyAsync(dd_hdr(c), &h, sizeof(DDHeader), cudaMemcpyHostToDevice, st));
    CU_TRY(c, cudaStreamSynchronize(st));
    CU_TRY(c, cudaGetLastError());
    c->orig_valid = true;
    c->dd_epoch = 0;
    CU_TRY(c, cudaMemset(c->dd_step_dev, 0, sizeof(unsigned long long)));
    drop_graphs(c);   // (new splitters, new bodies)
    return 0;
}

int lpe_bh_dd_phase(lpe_bh_ctx* c, const lpe_bh_params* p, int phase) {
    if (!c || !p) return 1;
    DevGuard _dg(c->device);
    return dd_phase(c, *p, phase);
}

// nsteps full steps; the in-stream barriers make every rank wait for the others, so ALL ranks must make the same call
// (one process per GPU, or one host thread looping over its contexts: the launches are asynchronous)
int lpe_bh_dd_step(lpe_bh_ctx* c, const lpe_bh_params* p, int nsteps) {
    if (!c || !p) return 1;
    DevGuard _dg(c->device);
    for (int s = 0; s < nsteps; ++s) {
        // the whole step — three phases, two in-stream barriers — is one CUDA graph when nothing it depends on changed
        // (parameters, splitters, buffer parity); the domain table is brought up to date outside the graph
        StepConst k;
        if (dd_make_const(c, *p, k)) return 1;
        if (dd_build_domain(c, k)) return 1;
        std::string key("dd");
        key_add(key, *p); key_add(key, c->cap); key_add(key, c->instr); key_add(key, c->force_dfs); key_add(key, c->dd_cur);
        key_add(key, c->dd_R); key_add(key, c->dd_rank); key_add(key, c->body); key_add(key, c->stream);
        for (int r = 0; r <= c->dd_R; ++r) key_add(key, c->dd_split30[r]);
        for (int r = 0; r < c->dd_R; ++r) key_add(key, c->dd_peer_win[r]);
        auto body = [&]() -> int {
            for (int ph = 0; ph < 3; ++ph)
                if (dd_phase(c, *p, ph)) return 1;
            return 0;
        };
        if (run_graphed(c, key, body, true)) return 1;
    }
    return 0;
}

// the bodies this rank owns, with their creation indices; arrays of at least `capacity` elements; synchronises
int lpe_bh_dd_download(lpe_bh_ctx* c, uint64_t* n_out, uint32_t* index, double* x, double* y, double* vx, double* vy,
                       uint32_t* accepted) {
    if (!c || !n_out) return 1;
    if (!c->dd) return fail(c, "context is not in domain-decomposed mode");
    DevGuard _dg(c->device);
    cudaStream_t st = c->stream;
    CU_TRY(c, cudaStreamSynchronize(st));
    if (dd_check_fault(c)) return 1;
    if (fetch_fault(c)) return 1;
    unsigned int nl = 0;
    CU_TRY(c, cudaMemcpy(&nl, &dd_hdr(c)->n_live, sizeof(nl), cudaMemcpyDeviceToHost));
    *n_out = nl;
    if (nl) {
        const size_t bytes = sizeof(double) * nl;
        const int g = cdiv((long long)nl, 256);
        double *t0 = c->tmp, *t1 = c->tmp + c->cap, *t2 = c->tmp + 2 * c->cap, *t3 = c->tmp + 3 * c->cap;
        if (x || y) {
            k_get_pos<<<g, 256, 0, st>>>((int)nl, c->body, t0, t1, nullptr);
            if (x) CU_TRY(c, cudaMemcpyAsync(x, t0, bytes, cudaMemcpyDeviceToHost, st));
            if (y) CU_TRY(c, cudaMemcpyAsync(y, t1, bytes, cudaMemcpyDeviceToHost, st));
        }
        if (vx || vy) {
            k_unpack2<<<g, 256, 0, st>>>((int)nl, c->vel, t2, t3, nullptr);
            if (vx) CU_TRY(c, cudaMemcpyAsync(vx, t2, bytes, cudaMemcpyDeviceToHost, st));
            if (vy) CU_TRY(c, cudaMemcpyAsync(vy, t3, bytes, cudaMemcpyDeviceToHost, st));
        }
        if (index) CU_TRY(c, cudaMemcpyAsync(index, c->orig, 4 * (size_t)nl, cudaMemcpyDeviceToHost, st));
        if (accepted) {
            if (!(c->instr & 2)) return fail(c, "enable instrumentation bit1 before the step");
            CU_TRY(c, cudaMemcpyAsync(accepted, c->cntAcc, 4 * (size_t)nl, cudaMemcpyDeviceToHost, st));
        }
    }
    CU_TRY(c, cudaStreamSynchronize(st));
    CU_TRY(c, cudaGetLastError());
    return check_fault(c);
}

int lpe_bh_dd_get_stats(lpe_bh_ctx* c, lpe_bh_dd_stats* o) {
    if (!c || !o) return 1;
    if (!c->dd) return fail(c, "context is not in domain-decomposed mode");
    DevGuard _dg(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    std::memset(o, 0, sizeof(*o));
    o->capacity = c->cap;
    o->import_blocks = c->dd_icap;
    o->rank = c->dd_rank; o->nranks = c->dd_R;
    DDHeader h;
    CU_TRY(c, cudaMemcpy(&h, dd_hdr(c), sizeof(h), cudaMemcpyDeviceToHost));
    o->n_live = h.n_live;
    o->fault = h.fault;
    for (int r = 0; r < c->dd_R; ++r) o->n_roots += h.root_count[r];
    if (c->have_step) {
        Scal s;
        CU_TRY(c, cudaMemcpy(&s, c->scal, sizeof(s), cudaMemcpyDeviceToHost));
        o->n_in_tree = s.n_in; o->n_terminals = s.n_term; o->n_cells = s.n_internal;
        o->interactions = s.interactions; o->work_cost = s.work_cost; o->overflow_chunks = s.ovf_count;
        for (int r = 0; r < c->dd_R; ++r) o->exported_blocks[r] = s.exp_count[r];
        o->depth = c->last_c.D;
        o->export_rounds = (int32_t)s.dd_rounds;
        if (c->instr & 1) {
            cudaEventElapsedTime(&o->ms_keygen, c->dd_ev[0], c->dd_ev[1]);
            cudaEventElapsedTime(&o->ms_wait_a, c->dd_ev[1], c->dd_ev[2]);
            cudaEventElapsedTime(&o->ms_sort, c->dd_ev[2], c->dd_ev[3]);
            cudaEventElapsedTime(&o->ms_build, c->dd_ev[3], c->dd_ev[4]);
            cudaEventElapsedTime(&o->ms_export, c->dd_ev[4], c->dd_ev[5]);
            cudaEventElapsedTime(&o->ms_wait_b, c->dd_ev[5], c->dd_ev[6]);
            cudaEventElapsedTime(&o->ms_top, c->dd_ev[6], c->dd_ev[7]);
            cudaEventElapsedTime(&o->ms_traverse, c->dd_ev[7], c->dd_ev[8]);
            cudaEventElapsedTime(&o->ms_total, c->dd_ev[0], c->dd_ev[8]);
        }
    }
    return 0;
}

// Load-balance input: for every 32-body chunk of this rank's sorted bodies, the depth-30 key of its first body and the
// number of list entries its warp evaluated in the last step's traversal. The caller concatenates the ranks' arrays
// (they are globally key-ordered), weighs them and picks new splitters. Synchronises.
int lpe_bh_dd_chunk_costs(lpe_bh_ctx* c, uint64_t* n_chunks, uint64_t* first_key30, uint32_t* cost) {
    if (!c || !n_chunks) return 1;
    if (!c->dd || !c->have_step) return fail(c, "no domain-decomposed step has run");
    DevGuard _dg(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    unsigned int nl = 0;
    CU_TRY(c, cudaMemcpy(&nl, &dd_hdr(c)->n_live, sizeof(nl), cudaMemcpyDeviceToHost));
    const uint64_t nc = ((uint64_t)nl + 31) / 32;
    *n_chunks = nc;
    if (!nc) return 0;
    if (cost) CU_TRY(c, cudaMemcpy(cost, c->dd_chunk_cost, 4 * nc, cudaMemcpyDeviceToHost));
    if (first_key30) {
        const int sh = 2 * (LPE_MAX_DEPTH - c->last_c.D);
        const uint64_t top = 1ull << (2 * LPE_MAX_DEPTH);
        if (c->last_c.k32) {   // every 32nd sorted key and payload (strided copies: 4 bytes out of every 128)
            std::vector<uint32_t> k32(nc), v32(nc);
            CU_TRY(c, cudaMemcpy2D(k32.data(), 4, c->keys[c->sorted_sel], 128, 4, nc, cudaMemcpyDeviceToHost));
            CU_TRY(c, cudaMemcpy2D(v32.data(), 4, c->vals[c->sorted_sel], 128, 4, nc, cudaMemcpyDeviceToHost));
            for (uint64_t i = 0; i < nc; ++i)
                first_key30[i] = (v32[i] & LPE_VAL_OUT) ? top : ((uint64_t)k32[i] << sh);   // bodies outside the tree sort last
        } else {
            CU_TRY(c, cudaMemcpy2D(first_key30, 8, c->keys[c->sorted_sel], 256, 8, nc, cudaMemcpyDeviceToHost));
            for (uint64_t i = 0; i < nc; ++i) {
                const uint64_t k = first_key30[i];
                first_key30[i] = (k >> (2 * c->last_c.D)) ? top : (k << sh);
            }
        }
    }
    return 0;
}

int lpe_bh_dd_get_splitters(lpe_bh_ctx* c, uint64_t* split30) {
    if (!c || !split30) return 1;
    if (!c->dd) return fail(c, "context is not in domain-decomposed mode");
    for (int r = 0; r <= c->dd_R; ++r) split30[r] = c->dd_split30[r];
    return 0;
}

// New splitters (depth-30 keys, R + 1 values, non-decreasing, first 0): bodies that now belong to another rank move
// there in the next step's phase A. Every rank must set the same values before the same step.
int lpe_bh_dd_set_splitters(lpe_bh_ctx* c, const uint64_t* split30) {
    if (!c || !split30) return 1;
    if (!c->dd) return fail(c, "context is not in domain-decomposed mode");
    if (split30[0] != 0) return fail(c, "splitters must start at key 0");
    for (int r = 0; r < c->dd_R; ++r)
        if (split30[r] > split30[r + 1]) return fail(c, "splitters must be non-decreasing");
    for (int r = 0; r <= c->dd_R; ++r) c->dd_split30[r] = split30[r];
    c->dd_split30[c->dd_R] = 1ull << (2 * LPE_MAX_DEPTH);
    c->dd_dom_depth = -1;
    return 0;
}

// host helpers for tests (no GPU): sort key of a cell and back, the rule the device uses
uint64_t lpe_bh_cell_key(uint32_t ix, uint32_t iy, int level, int hilbert) { return cell_to_key(ix, iy, level, hilbert); }
void lpe_bh_key_cell(uint64_t key, int level, int hilbert, uint32_t* ix, uint32_t* iy) {
    unsigned int x, y;
    key_to_cell(key, level, hilbert, x, y);
    if (ix) *ix = x;
    if (iy) *iy = y;
}

}  // extern "C"
