/**
 * @file barnes_hut.cpp — ECS side of the drop-in BarnesHutSystem (see barnes_hut.hpp here).
 *
 * Mirrors the control flow of the reference's update() (src/systems/barnes_hut.cpp:50-99):
 *   early exit when no mass reaches smallMassThreshold (:55-71), SimulatorState lookup with the same
 *   warning (:75-80), then "build tree + force loop" — which here is one call into liblpe_bh.so.
 */
#include "systems/barnes_hut.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <iostream>
#include <mutex>
#include <thread>
#include <vector>

#include "core/constants.hpp"
#include "core/profile.hpp"
#include "lpe_bh.h"

namespace Systems {

static_assert(sizeof(Components::Position) == 2 * sizeof(double), "Position is an {x, y} record");
static_assert(sizeof(Components::Velocity) == 2 * sizeof(double), "Velocity is an {x, y} record");
static_assert(sizeof(Components::Mass) == sizeof(double), "Mass is one double");

namespace {

/** A handful of persistent worker threads: run(f) calls f(k, K) for k = 0..K-1 (k = 0 on the caller) and returns when
 *  all are done. The registry is only ever touched from inside run(), i.e. while update() owns it. */
class Workers {
public:
    explicit Workers(int extra) {
        for (int i = 0; i < extra; ++i) threads_.emplace_back([this, i] { loop(i + 1); });
    }
    ~Workers() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            ++epoch_;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int size() const { return (int)threads_.size() + 1; }
    void run(const std::function<void(int, int)>& f) {
        const int K = size();
        if (K == 1) { f(0, 1); return; }
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = &f;
            pending_ = K - 1;
            ++epoch_;
        }
        cv_.notify_all();
        f(0, K);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(int k) {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int, int)>* job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
                job = job_;
            }
            (*job)(k, size());
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)>* job_ = nullptr;
    int pending_ = 0;
    unsigned long long epoch_ = 0;
    bool stop_ = false;
};

template <class T>
struct PinnedArray {   // page-locked host memory (full PCIe rate, asynchronous copies); grows geometrically
    T* p = nullptr;
    std::size_t cap = 0;
    bool reserve(std::size_t n) {
        if (n <= cap) return true;
        const std::size_t want = std::max(n, cap + cap / 2);
        T* q = static_cast<T*>(lpe_bh_alloc_pinned(want * sizeof(T)));
        if (!q) return false;
        if (p) lpe_bh_free_pinned(p);
        p = q;
        cap = want;
        return true;
    }
    ~PinnedArray() { if (p) lpe_bh_free_pinned(p); }
};

constexpr std::size_t kPage = ENTT_PACKED_PAGE;   // elements per pool page (entt.hpp:60)

}  // namespace

struct BarnesHutSystem::Staging {
    PinnedArray<double> pos, vel, m;     // {x, y} records, {vx, vy} records, masses — in staging order
    PinnedArray<std::uint8_t> comp;
    PinnedArray<std::uint32_t> rank;
    std::vector<entt::entity> entities;  // entity-by-entity path only
    std::unique_ptr<Workers> workers;
    int workerCount = -1;
    bool reserve(std::size_t n) {
        return pos.reserve(2 * n) && vel.reserve(2 * n) && m.reserve(n) && comp.reserve(n) && rank.reserve(n);
    }
};

BarnesHutSystem::BarnesHutSystem() : st_(new Staging()) {}

BarnesHutSystem::~BarnesHutSystem() {
    st_.reset();   // page-locked buffers go before the context
    if (ctx_) lpe_bh_destroy(ctx_);
}

bool BarnesHutSystem::ensureContext() {
    if (ctx_) return true;
    if (contextFailed_) return false;
    if (lpe_bh_create(options_.device, &ctx_) != 0) {
        std::cerr << "[BarnesHut] Warning: cannot open CUDA device " << options_.device << ": "
                  << lpe_bh_last_error(nullptr) << ". Skipping update.\n";
        contextFailed_ = true;
        ctx_ = nullptr;
        return false;
    }
    return true;
}

// Pool pages as they are. Valid when Position, Mass and Velocity hold exactly the same entities in the same packed
// order and nothing is a Boundary: then buildTree's view (barnes_hut.cpp:117) and the force loop's view (:89) both
// visit packed index n-1, n-2, ..., 0 (EnTT iterates the leading pool back to front), i.e. the body at packed index
// i has insertion rank n-1-i — the C ABI's default when no rank array is given — and every body has every component.
// This is the cheap part of the test and the copy of the positions, the first thing the device needs; whether the three
// pools really hold the same entities in the same order is checked page by page while the velocities are copied
// (update() below), i.e. while the device already builds and walks the tree: a tick that fails there is started over on
// the entity-by-entity path, its queued device work is simply superseded.
bool BarnesHutSystem::stagePagewise(entt::registry& registry, std::size_t& n) {
    auto& ps = registry.storage<Components::Position>();
    auto& ms = registry.storage<Components::Mass>();
    auto& vs = registry.storage<Components::Velocity>();
    n = ms.size();
    if (n == 0 || ps.size() != n || vs.size() != n) return false;
    if (!registry.storage<Components::Boundary>().empty()) return false;
    if (options_.fuseMovement) {   // the fused drift needs per-entity phases (movement.cpp:25-29): any liquid -> slow path
        for (const auto& ph : registry.storage<Components::ParticlePhase>())
            if (ph.phase == Components::Phase::Liquid) return false;
    }
    if (!st_->reserve(n)) return false;
    auto** ppages = ps.raw();
    const std::size_t pages = (n + kPage - 1) / kPage;
    double* pos = st_->pos.p;
    st_->workers->run([&](int k, int K) {
        for (std::size_t pg = pages * k / K; pg < pages * (k + 1) / K; ++pg) {
            const std::size_t a = pg * kPage, cnt = std::min(kPage, n - a);
            std::memcpy(pos + 2 * a, ppages[pg], cnt * sizeof(Components::Position));
        }
    });
    return true;
}

// The reference's own views, entity by entity, in the iteration order of buildTree's view (barnes_hut.cpp:117): that
// order IS the insertion order, which decides each cell's first occupant. Every target of the force
// loop (view<Position,Velocity,Mass>, :89) is also in this view, so one pass stages sources and targets.
std::size_t BarnesHutSystem::stagePerEntity(entt::registry& registry) {
    auto insertView = registry.view<Components::Position, Components::Mass>(entt::exclude<Components::Boundary>);
    auto& ents = st_->entities;
    ents.clear();
    for (auto entity : insertView) ents.push_back(entity);
    const std::size_t nMass = ents.size();
    if (options_.fuseMovement) {
        // MovementSystem also moves entities without Mass (movement.cpp:20): they ride along as pure movers
        auto moveView = registry.view<Components::Position, Components::Velocity>(entt::exclude<Components::Boundary>);
        for (auto entity : moveView)
            if (!registry.all_of<Components::Mass>(entity)) ents.push_back(entity);
    }
    const std::size_t n = ents.size();
    if (n == 0 || !st_->reserve(n)) return 0;
    for (std::size_t i = 0; i < n; ++i) {
        const auto entity = ents[i];
        const auto& pos = registry.get<Components::Position>(entity);
        std::uint8_t comp = 0;
        double mass = 0.0, vx = 0.0, vy = 0.0;
        if (i < nMass) {
            comp |= LPE_HAS_MASS;
            mass = registry.get<Components::Mass>(entity).value;
        }
        if (const auto* vel = registry.try_get<Components::Velocity>(entity)) {
            comp |= LPE_HAS_VELOCITY;
            vx = vel->x;
            vy = vel->y;
        }
        if (options_.fuseMovement) {
            if (const auto* ph = registry.try_get<Components::ParticlePhase>(entity))
                if (ph->phase == Components::Phase::Liquid) comp |= LPE_LIQUID;   // movement.cpp:25-29
        }
        st_->pos.p[2 * i] = pos.x; st_->pos.p[2 * i + 1] = pos.y;
        st_->vel.p[2 * i] = vx;    st_->vel.p[2 * i + 1] = vy;
        st_->m.p[i] = mass;
        st_->comp.p[i] = comp;
        st_->rank.p[i] = static_cast<std::uint32_t>(i);   // position in the view's iteration (massless movers: unused)
    }
    return n;
}

void BarnesHutSystem::update(entt::registry& registry) {
    PROFILE_SCOPE("BarnesHutSystem");

    // Early exit, reference barnes_hut.cpp:55-71: skip when every non-boundary mass is below the threshold.
    if (specificConfig.smallMassThreshold > 0.0) {
        bool shouldSkip = true;
        auto massCheckView = registry.view<Components::Mass>(entt::exclude<Components::Boundary>);
        for (auto entity : massCheckView) {
            if (massCheckView.get<Components::Mass>(entity).value >= specificConfig.smallMassThreshold) {
                shouldSkip = false;
                break;
            }
        }
        if (shouldSkip) return;
    }

    // reference barnes_hut.cpp:75-80
    auto stateView = registry.view<Components::SimulatorState>();
    if (stateView.empty()) {
        std::cerr << "[BarnesHut] Warning: No SimulatorState found. Skipping update.\n";
        return;
    }
    const auto& simState = stateView.get<Components::SimulatorState>(stateView.front());

    if (!ensureContext()) return;
    const int wantWorkers = std::max(0, options_.stagingThreads);
    if (!st_->workers || st_->workerCount != wantWorkers) {
        st_->workers.reset(new Workers(wantWorkers > 0 ? wantWorkers - 1 : 0));
        st_->workerCount = wantWorkers;
    }

    // LPE_DROPIN_TRACE=1: where the host time of one update goes (stderr, one line per update)
    static const bool trace = std::getenv("LPE_DROPIN_TRACE") != nullptr;
    double tms[8] = {};
    int tk = 0;
    const auto tstart = std::chrono::steady_clock::now();
    auto mark = [&]() {
        if (trace && tk < 8) tms[tk++] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tstart).count();
    };

    std::size_t n = 0;
    bool pagewise = options_.pagewiseStaging && stagePagewise(registry, n);
    if (!pagewise) n = stagePerEntity(registry);
    lastPath_ = pagewise ? 1 : 0;
    if (n == 0) return;

    lpe_bh_params p{};
    p.universe_size = sysConfig.UniverseSizeMeters;
    p.softening = sysConfig.GravitationalSoftener;
    p.theta = specificConfig.theta;
    p.small_mass_threshold = specificConfig.smallMassThreshold;
    p.G = SimulatorConstants::RealG;
    p.dt_kick = sysConfig.SecondsPerTick * simState.baseTimeAcceleration * simState.timeScale;  // barnes_hut.cpp:284
    p.dt_drift = sysConfig.SecondsPerTick * sysConfig.TimeAcceleration;                           // movement.cpp:17
    p.quirk_mode = options_.referenceQuirk ? 1 : 0;
    p.precision = options_.strictFp64 ? LPE_PREC_STRICT : LPE_PREC_FAST;
    p.do_drift = options_.fuseMovement ? 1 : 0;
    p.max_depth = 0;

    // Three calls: each queues the device work its array unlocks, so the next pool is copied while the GPU runs
    // (positions -> keys + sort | masses -> tree build | velocities -> traversal + kick).
    auto failed = [&]() {
        std::cerr << "[BarnesHut] Warning: device step failed: " << lpe_bh_last_error(ctx_) << ". Skipping update.\n";
    };
    lpe_bh_set_instrumentation(ctx_, options_.collectForceStats ? 2 : 0);
    mark();
    if (lpe_bh_tick_begin(ctx_, &p, n, st_->pos.p, pagewise ? nullptr : st_->comp.p) != 0) { failed(); return; }
    mark();
    std::size_t pages = (n + kPage - 1) / kPage;
    if (pagewise) {
        auto** mpages = registry.storage<Components::Mass>().raw();
        st_->workers->run([&](int k, int K) {
            for (std::size_t pg = pages * k / K; pg < pages * (k + 1) / K; ++pg) {
                const std::size_t a = pg * kPage, cnt = std::min(kPage, n - a);
                std::memcpy(st_->m.p + a, mpages[pg], cnt * sizeof(Components::Mass));
            }
        });
    }
    mark();
    if (lpe_bh_tick_mass(ctx_, st_->m.p, pagewise ? nullptr : st_->rank.p) != 0) { failed(); return; }
    mark();
    if (pagewise) {
        // the velocities, and — while the device builds the tree and walks it — the check that the three pools hold the
        // same entities in the same packed order (entities come and go: checked every tick)
        auto& vsr = registry.storage<Components::Velocity>();
        auto** vpages = vsr.raw();
        const auto* pe = registry.storage<Components::Position>().data();
        const auto* me = registry.storage<Components::Mass>().data();
        const auto* ve = vsr.data();
        std::atomic<bool> aligned{true};
        st_->workers->run([&](int k, int K) {
            for (std::size_t pg = pages * k / K; pg < pages * (k + 1) / K; ++pg) {
                const std::size_t a = pg * kPage, cnt = std::min(kPage, n - a);
                if (std::memcmp(pe + a, me + a, cnt * sizeof(entt::entity)) != 0 ||
                    std::memcmp(pe + a, ve + a, cnt * sizeof(entt::entity)) != 0) {
                    aligned.store(false, std::memory_order_relaxed);
                    return;
                }
                std::memcpy(st_->vel.p + 2 * a, vpages[pg], cnt * sizeof(Components::Velocity));
            }
        });
        if (!aligned.load()) {
            // start the tick over, entity by entity: lpe_bh_tick_begin supersedes everything queued so far (the masses
            // staged above went to the wrong bodies, nothing of that tick is kept)
            pagewise = false;
            lastPath_ = 0;
            n = stagePerEntity(registry);
            if (n == 0) return;
            if (lpe_bh_tick_begin(ctx_, &p, n, st_->pos.p, st_->comp.p) != 0) { failed(); return; }
            if (lpe_bh_tick_mass(ctx_, st_->m.p, st_->rank.p) != 0) { failed(); return; }
        }
    }
    mark();
    if (lpe_bh_tick_finish(ctx_, st_->pos.p, st_->vel.p) != 0) { failed(); return; }
    mark();
    if (options_.collectForceStats) {
        lpe_bh_stats s{};
        if (lpe_bh_get_stats(ctx_, &s) == 0) {
            forceStats_.maxForce = s.force_max;
            forceStats_.totalForce = s.force_sum;
            forceStats_.count = s.interactions;
        }
    }

    // The reference mutates Velocity in place through the view reference (barnes_hut.cpp:285-286): no signals.
    if (pagewise) {
        auto** vpages = registry.storage<Components::Velocity>().raw();
        auto** ppages = registry.storage<Components::Position>().raw();
        const bool drift = options_.fuseMovement;
        st_->workers->run([&](int k, int K) {
            for (std::size_t pg = pages * k / K; pg < pages * (k + 1) / K; ++pg) {
                const std::size_t a = pg * kPage, cnt = std::min(kPage, n - a);
                std::memcpy(vpages[pg], st_->vel.p + 2 * a, cnt * sizeof(Components::Velocity));
                if (drift) std::memcpy(ppages[pg], st_->pos.p + 2 * a, cnt * sizeof(Components::Position));
            }
        });
        mark();
        if (trace)
            std::fprintf(stderr, "[BarnesHut] update trace (ms): positions staged %.3f | tick_begin queued %.3f | masses staged %.3f | "
                                 "tick_mass queued %.3f | velocities staged %.3f | tick_finish returned %.3f | velocities back in the pool %.3f\n",
                         tms[0], tms[1], tms[2], tms[3], tms[4], tms[5], tms[6]);
        return;
    }
    for (std::size_t i = 0; i < n; ++i) {
        if (!(st_->comp.p[i] & LPE_HAS_VELOCITY)) continue;
        auto& vel = registry.get<Components::Velocity>(st_->entities[i]);
        vel.x = st_->vel.p[2 * i];
        vel.y = st_->vel.p[2 * i + 1];
        if (options_.fuseMovement) {
            auto& pos = registry.get<Components::Position>(st_->entities[i]);
            pos.x = st_->pos.p[2 * i];
            pos.y = st_->pos.p[2 * i + 1];
        }
    }
}

}  // namespace Systems
