/*
 * bh_oracle.c — plain-C restatement of the reference's Barnes-Hut step.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_abi.h): the product path (the CUDA library
 * behind include/lpe_bh.h) never links, loads or calls this file.
 *
 * PARITY PINNING: this port is checked bit-for-bit against oracle/_ref/libref_bh.so
 * (the reference's own barnes_hut.cpp/movement.cpp compiled unmodified) by
 * tests/test_oracle_pinning.py, and against golden vectors generated from that
 * library (tests/golden/, made by tests/golden/make_golden.py). The reference
 * itself ships no tests or golden vectors.
 *
 * Every function cites the reference lines it follows. Arithmetic is kept in the
 * reference's expression order; build with -ffp-contract=off so no FMA is formed.
 * Bodies are addressed by creation index; node "pointers" are pool indices, which
 * also removes the reference's pool-growth use-after-free (defect D1).
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>

#include "oracle_abi.h"

#define NIL (-1)
#define MAX_DEPTH 1000 /* the reference recurses without bound on coincident points (defect D3) */

typedef struct {
    /* QuadTreeNode, include/systems/barnes_hut.hpp:81-107, as SoA */
    double *M, *cx, *cy, *bx, *by, *bs;
    uint8_t *leaf, *small;
    int64_t *single;
    int64_t *child; /* 4 per node: nw, ne, sw, se */
    int64_t cap, next;
    /* body arrays (borrowed) */
    const double *x, *y, *m;
    double U, eps, theta, thr, G;
    int quirk;
    int failed;
} tree_t;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int tree_reserve(tree_t* t, int64_t cap) {
    if (cap <= t->cap) return 0;
#define GROW(p, T, k) do { T* q = (T*)realloc(t->p, sizeof(T) * (size_t)(cap) * (k)); if (!q) return 1; t->p = q; } while (0)
    GROW(M, double, 1); GROW(cx, double, 1); GROW(cy, double, 1);
    GROW(bx, double, 1); GROW(by, double, 1); GROW(bs, double, 1);
    GROW(leaf, uint8_t, 1); GROW(small, uint8_t, 1);
    GROW(single, int64_t, 1); GROW(child, int64_t, 4);
#undef GROW
    t->cap = cap;
    return 0;
}

static void tree_free(tree_t* t) {
    free(t->M); free(t->cx); free(t->cy); free(t->bx); free(t->by); free(t->bs);
    free(t->leaf); free(t->small); free(t->single); free(t->child);
    memset(t, 0, sizeof(*t));
}

/* allocateNode, barnes_hut.cpp:38-48 (index-based, so growth is safe) + QuadTreeNode() defaults, barnes_hut.hpp:83-87 */
static int64_t alloc_node(tree_t* t) {
    if (t->next >= t->cap) {
        if (tree_reserve(t, t->cap ? t->cap * 2 : 1024)) { t->failed = 1; return NIL; }
    }
    int64_t k = t->next++;
    t->M[k] = 0.0; t->cx[k] = 0.0; t->cy[k] = 0.0;
    t->bx[k] = 0.0; t->by[k] = 0.0; t->bs[k] = 0.0;
    t->leaf[k] = 1; t->small[k] = 1; t->single[k] = NIL;
    t->child[4 * k + 0] = t->child[4 * k + 1] = t->child[4 * k + 2] = t->child[4 * k + 3] = NIL;
    return k;
}

/* QuadTreeNode::contains, barnes_hut.hpp:112-115 */
static int contains(const tree_t* t, int64_t k, double x, double y) {
    return (x >= t->bx[k] && x < t->bx[k] + t->bs[k] && y >= t->by[k] && y < t->by[k] + t->bs[k]);
}

/* QuadTreeNode::getQuadrant, barnes_hut.hpp:121-131: 0=NW 1=NE 2=SW 3=SE */
static int quadrant(const tree_t* t, int64_t k, double x, double y) {
    double midX = t->bx[k] + t->bs[k] * 0.5;
    double midY = t->by[k] + t->bs[k] * 0.5;
    if (x < midX) return (y < midY) ? 0 : 2;
    return (y < midY) ? 1 : 3;
}

/* subdivide, barnes_hut.cpp:199-238 */
static void subdivide(tree_t* t, int64_t k) {
    t->leaf[k] = 0;
    double half = t->bs[k] * 0.5;
    double x = t->bx[k], y = t->by[k];
    for (int q = 0; q < 4; ++q) {
        int64_t c = alloc_node(t);
        if (c == NIL) return;
        t->child[4 * k + q] = c;
        t->bx[c] = (q & 1) ? x + half : x;
        t->by[c] = (q & 2) ? y + half : y;
        t->bs[c] = half;
    }
}

/* insertParticle, barnes_hut.cpp:133-197 */
static void insert(tree_t* t, int64_t k, int64_t e, double px, double py, double mass, int depth) {
    if (k == NIL || t->failed) return;
    if (depth > MAX_DEPTH) { t->failed = 2; return; }
    if (!contains(t, k, px, py)) return;                                  /* :139 */
    if (t->M[k] == 0.0) {                                                 /* :144-154 */
        t->M[k] = mass; t->cx[k] = px; t->cy[k] = py; t->single[k] = e;
        if (mass >= t->thr) t->small[k] = 0;
        return;
    }
    if (t->leaf[k]) {                                                     /* :157-169 */
        int64_t old = t->single[k];
        double ox = t->x[old], oy = t->y[old], om = t->m[old];
        subdivide(t, k);
        if (t->failed) return;
        if (t->quirk) {
            /* the old occupant is re-inserted THROUGH this node, which already holds its mass:
             * the first-occupant double count */
            insert(t, k, old, ox, oy, om, depth);
        } else {
            /* textbook variant (not the reference): hand the old occupant straight to its child */
            insert(t, t->child[4 * k + quadrant(t, k, ox, oy)], old, ox, oy, om, depth + 1);
        }
        insert(t, k, e, px, py, mass, depth);
    } else {                                                              /* :170-196 */
        double newTotal = t->M[k] + mass;
        t->cx[k] = (t->cx[k] * t->M[k] + px * mass) / newTotal;
        t->cy[k] = (t->cy[k] * t->M[k] + py * mass) / newTotal;
        t->M[k] = newTotal;
        if (mass >= t->thr) t->small[k] = 0;
        int q = quadrant(t, k, px, py);
        insert(t, t->child[4 * k + q], e, px, py, mass, depth + 1);
    }
}

/* view<Position,Mass>(exclude<Boundary>) iteration order, barnes_hut.cpp:117: EnTT walks the
 * leading pool back to front, i.e. newest entity first, unless `rank` says otherwise. */
static int64_t* insertion_order(uint64_t n, const uint8_t* comp, const uint32_t* rank, uint64_t* count) {
    int64_t* ord = (int64_t*)malloc(sizeof(int64_t) * (n ? n : 1));
    if (!ord) return NULL;
    uint64_t k = 0;
    if (!rank) {
        for (int64_t i = (int64_t)n - 1; i >= 0; --i) {
            uint8_t c = comp ? comp[i] : (ORC_HAS_MASS | ORC_HAS_VELOCITY);
            if ((c & ORC_HAS_MASS) && !(c & ORC_BOUNDARY)) ord[k++] = i;
        }
    } else {
        /* counting placement by rank (ranks are a permutation of 0..k-1 over the in-view bodies) */
        for (uint64_t i = 0; i < n; ++i) ord[i] = NIL;
        for (uint64_t i = 0; i < n; ++i) {
            uint8_t c = comp ? comp[i] : (ORC_HAS_MASS | ORC_HAS_VELOCITY);
            if ((c & ORC_HAS_MASS) && !(c & ORC_BOUNDARY) && rank[i] < n) { ord[rank[i]] = (int64_t)i; }
        }
        for (uint64_t i = 0; i < n; ++i) if (ord[i] != NIL) ord[k++] = ord[i];
    }
    *count = k;
    return ord;
}

/* buildTree, barnes_hut.cpp:101-131 */
static int build(tree_t* t, const orc_params* p, uint64_t n, const double* x, const double* y, const double* m,
                 const uint8_t* comp, const uint32_t* rank) {
    t->x = x; t->y = y; t->m = m;
    t->U = p->universe_size; t->eps = p->softening; t->theta = p->theta;
    t->thr = p->small_mass_threshold; t->G = p->G; t->quirk = p->quirk;
    t->next = 0; t->failed = 0;
    if (tree_reserve(t, (int64_t)(3 * n + 1024))) return 1;
    int64_t root = alloc_node(t);
    t->bx[root] = 0.0; t->by[root] = 0.0; t->bs[root] = t->U;             /* :110-112 */
    uint64_t cnt = 0;
    int64_t* ord = insertion_order(n, comp, rank, &cnt);
    if (!ord) return 1;
    for (uint64_t k = 0; k < cnt; ++k) {
        int64_t i = ord[k];
        if (x[i] >= 0.0 && x[i] < t->U && y[i] >= 0.0 && y[i] < t->U)    /* :123-124 */
            insert(t, root, i, x[i], y[i], m[i], 0);
        if (t->failed) break;
    }
    free(ord);
    return t->failed;
}

typedef struct { uint64_t accepted, visited; double fmax, fsum; } counts_t;

/* ---- tiny pthread parallel-for (dynamic chunks); threads <= 1 runs inline ---- */
typedef void (*chunk_fn)(void* ctx, int64_t lo, int64_t hi, int tid);
typedef struct { chunk_fn fn; void* ctx; int64_t n, chunk; volatile int64_t next; int tid; pthread_mutex_t* mu; } pf_shared;
typedef struct { pf_shared* sh; int tid; } pf_arg;
static void* pf_worker(void* a_) {
    pf_arg* a = (pf_arg*)a_;
    pf_shared* sh = a->sh;
    for (;;) {
        int64_t lo = __atomic_fetch_add(&sh->next, sh->chunk, __ATOMIC_RELAXED);
        if (lo >= sh->n) break;
        int64_t hi = lo + sh->chunk < sh->n ? lo + sh->chunk : sh->n;
        sh->fn(sh->ctx, lo, hi, a->tid);
    }
    return NULL;
}
static void parallel_for(int64_t n, int64_t chunk, int threads, chunk_fn fn, void* ctx) {
    if (threads <= 1) { fn(ctx, 0, n, 0); return; }
    if (threads > 256) threads = 256;
    pf_shared sh; sh.fn = fn; sh.ctx = ctx; sh.n = n; sh.chunk = chunk; sh.next = 0;
    pthread_t th[256]; pf_arg args[256];
    for (int t = 0; t < threads; ++t) { args[t].sh = &sh; args[t].tid = t; pthread_create(&th[t], NULL, pf_worker, &args[t]); }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
}

/* calculateForce, barnes_hut.cpp:240-294. vel is updated in place, node by node, like the reference. */
static void force(const tree_t* t, int64_t k, int64_t e, double px, double py, double* vx, double* vy,
                  double mass, double dt, counts_t* c) {
    if (k == NIL || t->M[k] == 0.0) return;                               /* :248 */
    if (t->small[k] && (t->thr > 0.0)) return;                            /* :253 */
    c->visited++;
    double dx = t->cx[k] - px;
    double dy = t->cy[k] - py;
    double distSq = dx * dx + dy * dy + t->eps * t->eps;                  /* :261 */
    double dist = sqrt(distSq);
    double sizeSq = t->bs[k] * t->bs[k];
    double thetaSq = t->theta * t->theta;
    int useApprox = t->leaf[k] || (sizeSq / distSq < thetaSq);            /* :269 */
    if (useApprox) {
        if (t->leaf[k] && t->single[k] == e) return;                      /* :272 */
        double f = t->G * t->M[k] * mass / distSq;                        /* :277 */
        double invDistMass = f / (mass * dist);                           /* :280 */
        double accX = dx * invDistMass;
        double accY = dy * invDistMass;
        *vx += accX * dt;                                                 /* :285-286 */
        *vy += accY * dt;
        c->accepted++;
        if (f > c->fmax) c->fmax = f;                                     /* :278 DebugStats::updateForce(force) */
        c->fsum += f;
    } else {
        for (int q = 0; q < 4; ++q)                                       /* :289-292 nw, ne, sw, se */
            force(t, t->child[4 * k + q], e, px, py, vx, vy, mass, dt, c);
    }
}

static void fill_stats(const tree_t* t, orc_stats* st) {
    if (!st) return;
    st->pool_nodes = (uint64_t)t->next;
    uint64_t ne = 0, in = 0; int md = 0;
    for (int64_t k = 0; k < t->next; ++k) {
        if (t->M[k] != 0.0) {
            ++ne;
            int d = (int)lround(log2(t->U / t->bs[k]));
            if (d > md) md = d;
        }
        if (!t->leaf[k]) ++in;
    }
    st->nonempty_nodes = ne; st->internal_nodes = in; st->max_depth = md;
}

static int is_target(uint8_t c) { return (c & ORC_HAS_MASS) && (c & ORC_HAS_VELOCITY) && !(c & ORC_BOUNDARY); }
static int is_mover(uint8_t c) { return (c & ORC_HAS_VELOCITY) && !(c & ORC_BOUNDARY) && !(c & ORC_LIQUID); }

/* the target loop of BarnesHutSystem::update, barnes_hut.cpp:89-98 (targets are independent) */
typedef struct {
    const tree_t* t; const uint8_t* comp; const double *ox, *oy; double *ovx, *ovy; const double* m; double dt;
    uint32_t *acc_per_body, *vis_per_body; counts_t tot[256];
} force_job;
static void force_chunk(void* ctx, int64_t lo, int64_t hi, int tid) {
    force_job* j = (force_job*)ctx;
    uint64_t acc_loc = 0, vis_loc = 0;
    double fmax_loc = 0.0, fsum_loc = 0.0;
    for (int64_t i = lo; i < hi; ++i) {
        uint8_t c = j->comp ? j->comp[i] : (ORC_HAS_MASS | ORC_HAS_VELOCITY);
        if (!is_target(c)) continue;
        counts_t cn = {0, 0, 0.0, 0.0};
        double vxx = j->ovx[i], vyy = j->ovy[i];
        force(j->t, 0, i, j->ox[i], j->oy[i], &vxx, &vyy, j->m[i], j->dt, &cn);
        j->ovx[i] = vxx; j->ovy[i] = vyy;
        acc_loc += cn.accepted; vis_loc += cn.visited;
        if (cn.fmax > fmax_loc) fmax_loc = cn.fmax;
        fsum_loc += cn.fsum;
        if (j->acc_per_body) j->acc_per_body[i] = (uint32_t)cn.accepted;
        if (j->vis_per_body) j->vis_per_body[i] = (uint32_t)cn.visited;
    }
    j->tot[tid].accepted += acc_loc; j->tot[tid].visited += vis_loc;
    if (fmax_loc > j->tot[tid].fmax) j->tot[tid].fmax = fmax_loc;
    j->tot[tid].fsum += fsum_loc;
}

/* BarnesHutSystem::update (barnes_hut.cpp:50-99) then, if run_movement, MovementSystem::update
 * (movement.cpp:13-39), nsteps times. Arrays are updated in place in (ox,oy,ovx,ovy). */
int orc_bh_run(const orc_params* p, uint64_t n, const double* x, const double* y, const double* vx,
               const double* vy, const double* m, const uint8_t* comp, const uint32_t* rank, int nsteps,
               int threads, double* ox, double* oy, double* ovx, double* ovy, uint32_t* acc_per_body,
               uint32_t* vis_per_body, orc_stats* st) {
    if (!p || !x || !y || !m || !ox || !oy || !ovx || !ovy) return 1;
    if (st) memset(st, 0, sizeof(*st));
    memcpy(ox, x, sizeof(double) * n); memcpy(oy, y, sizeof(double) * n);
    if (vx) memcpy(ovx, vx, sizeof(double) * n); else memset(ovx, 0, sizeof(double) * n);
    if (vy) memcpy(ovy, vy, sizeof(double) * n); else memset(ovy, 0, sizeof(double) * n);
    tree_t t; memset(&t, 0, sizeof(t));
    double tb = 0.0, tf = 0.0, t00 = now_s();
    uint64_t acc_tot = 0, vis_tot = 0;
    double fmax_tot = 0.0, fsum_tot = 0.0;
    int rc = 0;
    for (int s = 0; s < nsteps && !rc; ++s) {
        /* early exit, barnes_hut.cpp:55-71 */
        if (p->small_mass_threshold > 0.0) {
            int skip = 1;
            for (uint64_t i = 0; i < n; ++i) {
                uint8_t c = comp ? comp[i] : (ORC_HAS_MASS | ORC_HAS_VELOCITY);
                if ((c & ORC_HAS_MASS) && !(c & ORC_BOUNDARY) && m[i] >= p->small_mass_threshold) { skip = 0; break; }
            }
            if (skip) goto movement;
        }
        {
            double t0 = now_s();
            rc = build(&t, p, n, ox, oy, m, comp, rank);
            double t1 = now_s();
            tb += t1 - t0;
            if (rc) break;
            double dt = p->seconds_per_tick * p->base_time_acceleration * p->time_scale; /* :284 */
            force_job job = {&t, comp, ox, oy, ovx, ovy, m, dt, acc_per_body, vis_per_body, {{0, 0, 0.0, 0.0}}};
            parallel_for((int64_t)n, 256, threads, force_chunk, &job);
            for (int k = 0; k < 256; ++k) {
                acc_tot += job.tot[k].accepted; vis_tot += job.tot[k].visited;
                if (job.tot[k].fmax > fmax_tot) fmax_tot = job.tot[k].fmax;
                fsum_tot += job.tot[k].fsum;
            }
            tf += now_s() - t1;
        }
    movement:
        if (p->run_movement) {
            double dt = p->seconds_per_tick * p->time_acceleration;       /* movement.cpp:17 */
            for (uint64_t i = 0; i < n; ++i) {
                uint8_t c = comp ? comp[i] : (ORC_HAS_MASS | ORC_HAS_VELOCITY);
                if (!is_mover(c)) continue;
                ox[i] += ovx[i] * dt;                                     /* movement.cpp:32-33 */
                oy[i] += ovy[i] * dt;
            }
        }
    }
    if (st) {
        fill_stats(&t, st);
        st->accepted = acc_tot; st->visited = vis_tot;
        st->force_max = fmax_tot; st->force_sum = fsum_tot; st->force_count = acc_tot;
        st->build_seconds = tb; st->force_seconds = tf; st->total_seconds = now_s() - t00;
    }
    tree_free(&t);
    return rc;
}

/* Build once and dump every non-empty node in pool (allocation) order — same order as ref_bh_tree. */
int orc_bh_tree(const orc_params* p, uint64_t n, const double* x, const double* y, const double* m,
                const uint8_t* comp, const uint32_t* rank, orc_node* out, uint64_t cap, uint64_t* count,
                orc_stats* st) {
    if (!p || !x || !y || !m) return 1;
    if (st) memset(st, 0, sizeof(*st));
    tree_t t; memset(&t, 0, sizeof(t));
    double t0 = now_s();
    int rc = build(&t, p, n, x, y, m, comp, rank);
    if (st) { fill_stats(&t, st); st->build_seconds = now_s() - t0; }
    uint64_t k = 0;
    for (int64_t i = 0; i < t.next && !rc; ++i) {
        if (t.M[i] == 0.0) continue;
        if (out && k < cap) {
            orc_node* o = &out[k];
            o->mass = t.M[i]; o->comx = t.cx[i]; o->comy = t.cy[i];
            o->bx = t.bx[i]; o->by = t.by[i]; o->bsize = t.bs[i];
            o->is_leaf = t.leaf[i]; o->all_small = t.small[i]; o->single = t.single[i];
        }
        ++k;
    }
    if (count) *count = k;
    tree_free(&t);
    return rc;
}

typedef struct {
    const orc_params* p; uint64_t n; const double *x, *y, *m; const uint8_t* comp; uint64_t first; double *ax, *ay;
} direct_job;
static void direct_chunk(void* ctx, int64_t lo, int64_t hi, int tid) {
    (void)tid;
    direct_job* d = (direct_job*)ctx;
    const double U = d->p->universe_size, e2 = d->p->softening * d->p->softening, G = d->p->G;
    const double *x = d->x, *y = d->y, *m = d->m;
    for (int64_t k = lo; k < hi; ++k) {
        uint64_t i = d->first + (uint64_t)k;
        double sx = 0.0, sy = 0.0;
        for (uint64_t j = 0; j < d->n; ++j) {
            if (j == i) continue;
            uint8_t c = d->comp ? d->comp[j] : (ORC_HAS_MASS | ORC_HAS_VELOCITY);
            if (!(c & ORC_HAS_MASS) || (c & ORC_BOUNDARY)) continue;
            if (!(x[j] >= 0.0 && x[j] < U && y[j] >= 0.0 && y[j] < U)) continue;
            double dx = x[j] - x[i], dy = y[j] - y[i];
            double d2 = dx * dx + dy * dy + e2;
            double f = G * m[j] / (d2 * sqrt(d2));
            sx += dx * f; sy += dy * f;
        }
        d->ax[k] = sx; d->ay[k] = sy;
    }
}

/* Direct O(N^2) Plummer-softened sum with the reference's force law (barnes_hut.cpp:257-282):
 * a_i = sum_{j != i, j source in [0,U)^2} G m_j d / (|d|^2 + eps^2)^{3/2}. Accuracy cross-check only. */
int orc_direct_accel(const orc_params* p, uint64_t n, const double* x, const double* y, const double* m,
                     const uint8_t* comp, uint64_t first_target, uint64_t n_targets, int threads, double* ax,
                     double* ay) {
    if (!p || !x || !y || !m || !ax || !ay) return 1;
    direct_job job = {p, n, x, y, m, comp, first_target, ax, ay};
    parallel_for((int64_t)n_targets, 64, threads, direct_chunk, &job);
    return 0;
}

const char* orc_bh_describe(void) {
    return "plain-C port of little-physics-engine src/systems/barnes_hut.cpp + movement.cpp (oracle/bh_oracle.c)";
}

/* ---------------------------------------------------------------------------------------------------------
 * BoundarySystem::update, reference src/systems/boundary.cpp:13-69 — clamp to [margin, U - margin], reflect the
 * velocity component with damping, and after a bounce cap the speed at maxSpeed. View = Position + Velocity
 * (boundary.cpp:22), asleep entities skipped (boundary.cpp:29-31).
 * --------------------------------------------------------------------------------------------------------- */
int orc_boundary(const orc_boundary_params* p, uint64_t n, double* x, double* y, double* vx, double* vy,
                 const uint8_t* comp) {
    if (!p || (n && (!x || !y || !vx || !vy))) return 1;
    const double marginM = p->margin, universeSizeM = p->universe_size;
    const double bounceDamping = p->bounce_damping, maxSpeed = p->max_speed;
    for (uint64_t i = 0; i < n; ++i) {
        const unsigned cm = comp ? comp[i] : (ORC_HAS_MASS | ORC_HAS_VELOCITY);
        if (!(cm & ORC_HAS_VELOCITY) || (cm & ORC_ASLEEP)) continue;
        int bounced = 0;
        if (x[i] < marginM) {                                   /* boundary.cpp:36-40 */
            x[i] = marginM;
            vx[i] = fabs(vx[i]) * bounceDamping;
            bounced = 1;
        } else if (x[i] > universeSizeM - marginM) {            /* boundary.cpp:42-46 */
            x[i] = universeSizeM - marginM;
            vx[i] = -fabs(vx[i]) * bounceDamping;
            bounced = 1;
        }
        if (y[i] < marginM) {                                   /* boundary.cpp:49-53 */
            y[i] = marginM;
            vy[i] = fabs(vy[i]) * bounceDamping;
            bounced = 1;
        } else if (y[i] > universeSizeM - marginM) {            /* boundary.cpp:55-59 */
            y[i] = universeSizeM - marginM;
            vy[i] = -fabs(vy[i]) * bounceDamping;
            bounced = 1;
        }
        if (bounced) {                                          /* boundary.cpp:62-68 */
            const double speed = sqrt(vx[i] * vx[i] + vy[i] * vy[i]);
            if (speed > maxSpeed) {
                vx[i] = (vx[i] / speed) * maxSpeed;
                vy[i] = (vy[i] / speed) * maxSpeed;
            }
        }
    }
    return 0;
}

