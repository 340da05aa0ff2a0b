// nb_engine.cu -- engine lifetime, capacity management and step sequencing behind the nb_* ABI.
#include <stdarg.h>
#include <string.h>

#include "nb_engine.h"

namespace nb {

static char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

void fatal(const char* what, const char* file, int line) {
    // The reference panics across the FFI boundary; nb_* return void, so there is nobody to tell.
    fprintf(stderr, "nbody_b200: fatal: %s (%s:%d)\n", what, file, line);
    fflush(stderr);
    abort();
}

Engine& engine() {
    static Engine* e = new Engine();  // intentionally leaked: no static-destruction-order games with CUDA
    return *e;
}

int local_begin(const Engine& e) {
    const long long b = static_cast<long long>(e.rank) * static_cast<long long>(e.lay.L);
    return b < e.n ? static_cast<int>(b) : e.n;
}
int local_count(const Engine& e) {
    const long long b = static_cast<long long>(e.rank) * static_cast<long long>(e.lay.L);
    long long c = static_cast<long long>(e.n) - b;
    if (c < 0) c = 0;
    if (c > static_cast<long long>(e.lay.L)) c = static_cast<long long>(e.lay.L);
    return static_cast<int>(c);
}

int try_init(Engine& e, int device) {
    if (e.inited) {
        if (device == e.device) return 0;
        set_error("already bound to device %d", e.device);
        return -2;
    }
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        set_error("no CUDA device: %s (libnbody_b200 has no CPU fallback)", cudaGetErrorString(err));
        (void)cudaGetLastError();
        return -1;
    }
    if (device < 0 || device >= count) {
        set_error("device %d out of range (0..%d)", device, count - 1);
        return -1;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        set_error("cudaGetDeviceProperties failed");
        return -1;
    }
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libnbody_b200 carries sm_100a code only", device, prop.major, prop.minor);
        return -1;
    }
    NB_CUDA(cudaSetDevice(device));
    e.device = device;
    e.num_sms = prop.multiProcessorCount;
    NB_CUDA(cudaStreamCreateWithFlags(&e.own_stream, cudaStreamNonBlocking));
    e.stream = e.own_stream;
    e.inited = true;
    return 0;
}

void ensure_init(Engine& e) {
    if (e.inited) {
        NB_CUDA(cudaSetDevice(e.device));
        return;
    }
    int dev = 0;
    if (const char* s = getenv("NB_DEVICE")) dev = atoi(s);
    if (const char* s = getenv("NB_PEER_TIMEOUT_MS")) e.peer_timeout_ns = atoll(s) <= 0 ? 0ull : static_cast<uint64_t>(atoll(s)) * 1000000ull;
    if (try_init(e, dev) != 0) fatal(g_err, __FILE__, __LINE__);
}

// Size the arena and scratch for n global bodies.  In the sharded case the arena was allocated (and
// exported to the peers) by nbx_dist_init for max_particles and cannot move.
void ensure_capacity(Engine& e, int n) {
    if (e.dist) {
        if (n > e.max_particles) {
            set_error("n=%d exceeds nbx_dist_init max_particles=%d", n, e.max_particles);
            fatal(g_err, __FILE__, __LINE__);
        }
        // Balanced shards: L follows the size of the set (ceil(n / world), 1024-aligned), not the capacity the arena
        // was allocated for -- otherwise a set much smaller than max_particles would sit on the first ranks only.
        // All ranks derive the same layout from the same n; the flags at offset 0 never move.
        size_t L = ((static_cast<size_t>(n) + e.world - 1) / e.world + kShardAlign - 1) / kShardAlign * kShardAlign;
        if (L == 0) L = kShardAlign;
        if (L > e.L_cap) L = e.L_cap;
        if (L != e.lay.L) {
            NB_CUDA(cudaStreamSynchronize(e.stream));
            e.lay.set(L);
        }
    } else {
        size_t L = (static_cast<size_t>(n) + kShardAlign - 1) / kShardAlign * kShardAlign;
        if (L == 0) L = kShardAlign;
        ArenaLayout want;
        want.set(L);
        if (want.bytes > e.arena_cap_bytes || e.arena.base == nullptr) {
            NB_CUDA(cudaStreamSynchronize(e.stream));
            if (e.arena.base) NB_CUDA(cudaFree(e.arena.base));
            NB_CUDA(cudaMalloc(&e.arena.base, want.bytes));
            e.arena_cap_bytes = want.bytes;
        }
        // the layout always matches the current set exactly, so no work is spent on stale padding
        e.lay = want;
        e.peer[0] = e.arena;
    }
    const size_t L = e.lay.L;
    if (L > e.force_cap) {
        if (e.force) NB_CUDA(cudaFree(e.force));
        NB_CUDA(cudaMalloc(&e.force, L * sizeof(float2)));
        e.force_cap = L;
    }
}

// ---- phase timing -------------------------------------------------------------------------------
PhaseScope::PhaseScope(Engine& e_, int id_) : e(e_), id(id_) {
    if (e.phase_timing && e.ev_slot < Engine::kPhaseRing && e.ev_sub[id][e.ev_slot] < Engine::kPhaseSub) {
        sub = e.ev_sub[id][e.ev_slot]++;
        cudaEvent_t* pr = e.ev[id][e.ev_slot][sub];
        if (!pr[0]) {
            NB_CUDA(cudaEventCreate(&pr[0]));
            NB_CUDA(cudaEventCreate(&pr[1]));
        }
        NB_CUDA(cudaEventRecord(pr[0], e.stream));
    }
}
PhaseScope::~PhaseScope() {
    if (sub >= 0) {
        NB_CUDA(cudaEventRecord(e.ev[id][e.ev_slot][sub][1], e.stream));
        if (e.ev_count[id] < e.ev_slot + 1) e.ev_count[id] = e.ev_slot + 1;
    }
}
// average ms per step of every phase since the previous call (sub-intervals of a step are summed); resets the ring
void collect_phase_times(Engine& e) {
    NB_CUDA(cudaStreamSynchronize(e.stream));
    for (int p = 0; p < NBX_NUM_PHASES; p++) {
        double sum = 0.0, subsum[Engine::kPhaseSub] = {};
        int cnt = 0;
        for (int k = 0; k < e.ev_count[p]; k++) {
            bool any = false;
            for (int u = 0; u < e.ev_sub[p][k]; u++) {
                float ms = 0.f;
                if (e.ev[p][k][u][0] && cudaEventElapsedTime(&ms, e.ev[p][k][u][0], e.ev[p][k][u][1]) == cudaSuccess) {
                    sum += ms;
                    subsum[u] += ms;
                    any = true;
                }
                (void)cudaGetLastError();
            }
            if (any) cnt++;
        }
        e.phase_ms[p] = cnt ? static_cast<float>(sum / cnt) : 0.f;
        for (int u = 0; u < Engine::kPhaseSub; u++) e.phase_sub_ms[p][u] = cnt ? static_cast<float>(subsum[u] / cnt) : 0.f;
        e.ev_count[p] = 0;
        for (int k = 0; k < Engine::kPhaseRing; k++) e.ev_sub[p][k] = 0;
    }
    e.ev_slot = 0;
}

// ---- step sequencing ------------------------------------------------------------------------------

static void fill_local(Engine& e, AllPairsArgs& a) {
    a.xi = e.arena.x(e.lay, e.cur);
    a.yi = e.arena.y(e.lay, e.cur);
    a.mi = e.arena.m(e.lay);
}

// rs-src/nbody.rs:106-162
void step_brute_force(Engine& e, float dt) {
    if (e.n == 0) return;
    e.kdk_theta = 0.f;
    AllPairsArgs a{};
    allpairs_plan(e, a);
    fill_local(e, a);
    {
        PhaseScope ps(e, 7);
        dist_fill_segments(e, a);
    }
    if (e.mode == NBX_MODE_FAST) {
        {
            PhaseScope ps(e, 0);
            launch_allpairs_fast(e, a);
        }
        PhaseScope ps(e, 1);
        launch_integrate_fast(e, a.partial, a.nseg * a.slices_per_seg, dt, false);
    } else {
        {
            PhaseScope ps(e, 0);
            launch_allpairs_exact(e, a, e.force);
        }
        PhaseScope ps(e, 1);
        launch_integrate_exact(e, e.force, dt, false);
    }
    e.step_count++;
    dist_signal_step_done(e);
    e.ctr.steps++;
    if (e.phase_timing) e.ev_slot++;
}

// Leapfrog (KDK) keeps v at the half step between steps.  Before anything reads the state, the owed closing half kick
// v += (dt/2) a(x) is applied with one force evaluation by the method of the last step; stepping then continues from a
// synchronised (x, v), which is algebraically the same kick-drift-kick sequence.
void kdk_close(Engine& e) {
    if (e.integrator != NBX_INTEGRATOR_LEAPFROG_KDK || e.kdk_pending == 0.f || e.n == 0) return;
    const float h = e.kdk_pending;
    e.kdk_pending = 0.f;
    if (e.kdk_theta == 0.f) accelerations_local(e);
    else bh_accelerations(e, e.kdk_theta, e.force);
    launch_kick(e, e.force, h);
}

// accelerations of the local shard into e.force (as a_i), no state change
void accelerations_local(Engine& e) {
    AllPairsArgs a{};
    allpairs_plan(e, a);
    fill_local(e, a);
    dist_fill_segments(e, a);
    if (e.mode == NBX_MODE_FAST) {
        launch_allpairs_fast(e, a);
        launch_accel_from_partial(e, a.partial, a.nseg * a.slices_per_seg, e.force);
    } else {
        launch_allpairs_exact(e, a, e.force);
        launch_accel_from_force(e, e.force, e.force);
    }
    // a read-only collective still counts as an epoch for the cross-rank protocol
    e.step_count++;
    dist_signal_step_done(e);
}

}  // namespace nb
