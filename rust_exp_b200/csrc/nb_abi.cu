// nb_abi.cu -- the extern "C" surface of libnbody_b200 (include/nbody_b200.h).
//
// The implementations carry the b200_ prefix; nb_alias.cu adds the unprefixed nb_* names as one-line forwarders in a
// SEPARATE object, so that a host module which itself defines nb_* (the Rust shim that replaces rs-src/nbody.rs inside
// the rust_exp static archive) can link libnbody_b200.a without duplicate symbols -- the linker then never pulls
// nb_alias.o out of the archive.
//
// The six reference symbols keep the exact names / argument order / C types that
// hs-src/RustNBodyExperiment.hs:101-106 imports from rs-src/nbody.rs; each takes the engine mutex for
// the whole call, like the reference's PARTICLES.lock() (rs-src/nbody.rs:36,43,82,117,380,517).
// No C++ exception crosses this boundary: everything below is noexcept-by-construction (abort on error).
#include <string.h>

#include "nb_engine.h"

using namespace nb;

#define NB_LOCK() std::lock_guard<std::mutex> lock__(engine().mu)

static void replace_set_begin(Engine& e, int n) {
    ensure_init(e);
    if (n < 0) n = 0;
    if (e.dist && e.world > 1) dist_wait_all(e, e.step_count);  // peers may still be reading our arena
    e.n = n;
    e.set_gen++;
    e.kdk_pending = 0.f;
    e.mirror_mass_valid = false;
    ensure_capacity(e, n);
}
static void replace_set_end(Engine& e) {
    if (e.dist && e.world > 1) {
        e.step_count++;
        dist_signal_step_done(e);
    }
}

extern "C" {

// rs-src/nbody.rs:34-37
int32_t b200_nb_num_particles(void) {
    NB_LOCK();
    return engine().n;
}

// rs-src/nbody.rs:39-64
void b200_nb_random_disk(int32_t num_particles) {
    NB_LOCK();
    Engine& e = engine();
    replace_set_begin(e, num_particles);
    generate_random_disk(e, e.n);
    replace_set_end(e);
}

// rs-src/nbody.rs:73-104
void b200_nb_stable_orbits(int32_t num_particles, float rmin, float rmax) {
    NB_LOCK();
    Engine& e = engine();
    // the reference always pushes the sun, then num_particles-1 planets (rs-src/nbody.rs:93-95)
    replace_set_begin(e, num_particles < 1 ? 1 : num_particles);
    generate_stable_orbits(e, e.n, rmin, rmax);
    replace_set_end(e);
}

// Step calls return only when the GPU has finished the step, like the reference's synchronous CPU call: its host
// puts a wall-clock timer around the step (hs-src/RustNBodyExperiment.hs:55-57, timeIt) and would otherwise display
// launch latency.  nbx_set_async(1) (or NB_ASYNC_STEPS=1 in the environment) makes steps return after enqueueing;
// nb_get_particles / nb_draw / nbx_synchronize still order after them.
static int g_async = -1;
static bool sync_steps() {
    if (g_async < 0) { const char* s = getenv("NB_ASYNC_STEPS"); g_async = (s && atoi(s) != 0) ? 1 : 0; }
    return g_async == 0;
}

// rs-src/nbody.rs:106-162
void b200_nb_step_brute_force(float dt) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    step_brute_force(e, dt);
    if (sync_steps()) NB_CUDA(cudaStreamSynchronize(e.stream));
}

// rs-src/nbody.rs:186-480 -- theta first
void b200_nb_step_barnes_hut(float theta, float dt, int32_t nthreads) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    if (theta == 0.0f) {  // rs-src/nbody.rs:197-200 (before nthreads is ever used)
        step_brute_force(e, dt);
    } else {
        // rs-src/nbody.rs:424-428: the division by nthreads sits inside the closure mapped over (0..nthreads); for
        // nthreads <= 0 that range is empty -- the reference builds the tree and returns without moving any body.
        if (nthreads <= 0) return;
        bh_step(e, theta, dt);
    }
    if (sync_steps()) NB_CUDA(cudaStreamSynchronize(e.stream));
}

// rs-src/nbody.rs:482-583
void b200_nb_draw(int32_t w, int32_t h, uint32_t* fb) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    kdk_close(e);   // the velocity tails are drawn from synchronised velocities
    draw_to_host(e, w, h, fb);
}

// ---- required additions ---------------------------------------------------------------------------
void b200_nb_set_particles(const float* aos5, int32_t n) {
    NB_LOCK();
    Engine& e = engine();
    replace_set_begin(e, n);
    state_upload_aos(e, aos5, e.n);
    replace_set_end(e);
}

void b200_nb_get_particles(float* aos5_out, int32_t n) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    kdk_close(e);
    if (e.dist && e.world > 1) dist_wait_all(e, e.step_count);
    state_download_aos(e, aos5_out, n);
    if (e.dist && e.world > 1) {
        e.step_count++;
        dist_signal_step_done(e);
    }
}

// ---- extension surface ----------------------------------------------------------------------------
int32_t nbx_get_particles_local(float* aos5_out_full, int32_t n) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    kdk_close(e);
    state_download_local(e, aos5_out_full, n);
    if (e.bh) bh_poll(e);
    return 0;
}

int32_t nbx_init(int32_t device) {
    NB_LOCK();
    return try_init(engine(), device);
}

void nbx_shutdown(void) {
    NB_LOCK();
    Engine& e = engine();
    if (!e.inited) return;
    cudaSetDevice(e.device);
    cudaStreamSynchronize(e.stream);
    dist_shutdown(e);
    bh_shutdown(e);
    x3_shutdown(e);
    draw_shutdown();
    if (e.arena.base) cudaFree(e.arena.base);
    if (e.bh_arena) cudaFree(e.bh_arena);
    e.bh_arena = nullptr;
    if (e.mirror) cudaFree(e.mirror);
    if (e.partial) cudaFree(e.partial);
    if (e.force) cudaFree(e.force);
    if (e.stage_dev) cudaFree(e.stage_dev);
    for (int p = 0; p < NBX_NUM_PHASES; p++)
        for (int k = 0; k < Engine::kPhaseRing; k++) {
            e.ev_sub[p][k] = 0;
            for (int u = 0; u < Engine::kPhaseSub; u++)
                for (int j = 0; j < 2; j++)
                    if (e.ev[p][k][u][j]) {
                        cudaEventDestroy(e.ev[p][k][u][j]);
                        e.ev[p][k][u][j] = nullptr;
                    }
        }
    if (e.own_stream) cudaStreamDestroy(e.own_stream);
    (void)cudaGetLastError();
    // reset to a fresh engine (keep the mutex)
    e.inited = false;
    e.arena.base = nullptr; e.mirror = nullptr; e.partial = nullptr; e.force = nullptr;
    e.stage_dev = nullptr; e.own_stream = nullptr; e.stream = nullptr;
    e.mirror_cap = e.partial_cap = e.force_cap = e.stage_dev_cap = 0;
    e.lay = ArenaLayout(); e.L_cap = 0; e.arena_cap_bytes = 0; e.n = 0; e.dist = false; e.rank = 0; e.world = 1; e.peers_mapped = false;
    e.cur = 0; e.step_count = 0; e.mode = NBX_MODE_FAST; e.tune = Tuning(); e.ctr = nbx_counters{};
    e.integrator = NBX_INTEGRATOR_EULER; e.kdk_pending = 0.f; e.kdk_theta = 0.f; e.square_aabb = 0;
    e.transport = NBX_TRANSPORT_P2P_DIRECT; e.max_particles = 0; e.mirror_mass_valid = false; e.bh_partition = 0;
    e.bh_count = false; e.phase_timing = false; e.ev_slot = 0; e.bh_lay = BhArenaLayout();
    for (int p = 0; p < NBX_NUM_PHASES; p++) e.ev_count[p] = 0;
    for (int g = 0; g < kMaxRanks; g++) { e.peer[g].base = nullptr; e.bh_peer[g] = nullptr; }
}

int32_t nbx_set_async(int32_t enable) {
    NB_LOCK();
    g_async = enable != 0 ? 1 : 0;
    return 0;
}
int32_t nbx_set_peer_timeout_ms(int32_t ms) {
    NB_LOCK();
    engine().peer_timeout_ns = ms <= 0 ? 0ull : static_cast<uint64_t>(ms) * 1000000ull;
    return 0;
}

int32_t nbx_set_integrator(int32_t integrator) {
    NB_LOCK();
    Engine& e = engine();
    if (integrator != NBX_INTEGRATOR_EULER && integrator != NBX_INTEGRATOR_LEAPFROG_KDK) {
        set_error("unknown integrator %d", integrator);
        return -1;
    }
    if (e.inited) kdk_close(e);   // leave the old scheme with synchronised velocities
    e.integrator = integrator;
    e.kdk_pending = 0.f;
    return 0;
}
int32_t nbx_set_square_aabb(int32_t enable) {
    NB_LOCK();
    engine().square_aabb = enable != 0;
    return 0;
}

const char* nbx_last_error(void) { return last_error(); }
const char* nbx_version(void) { return "nbody_b200 0.1 (sm_100a)"; }

int32_t nbx_set_mode(int32_t mode) {
    NB_LOCK();
    if (mode != NBX_MODE_FAST && mode != NBX_MODE_EXACT) {
        set_error("unknown mode %d", mode);
        return -1;
    }
    if (engine().inited) kdk_close(engine());
    engine().mode = mode;
    return 0;
}
int32_t nbx_get_mode(void) {
    NB_LOCK();
    return engine().mode;
}

int32_t nbx_set_stream(void* cuda_stream) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    NB_CUDA(cudaStreamSynchronize(e.stream));
    e.stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e.own_stream;
    return 0;
}

int32_t nbx_synchronize(void) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    cudaError_t err = cudaStreamSynchronize(e.stream);
    if (err != cudaSuccess) {
        set_error("cudaStreamSynchronize: %s", cudaGetErrorString(err));
        return -1;
    }
    bh_poll(e);
    return 0;
}

void nbx_seed(uint64_t seed) {
    NB_LOCK();
    engine().seed = seed;
}

// the Plummer set of configurations C2-C5 generated on the device (no reference counterpart)
void nbx_plummer(int32_t num_particles, float a_scale, float mass_per_body) {
    NB_LOCK();
    Engine& e = engine();
    if (!(a_scale > 0.f) || !(mass_per_body > 0.f)) fatal("nbx_plummer: scale and mass per body must be positive", __FILE__, __LINE__);
    replace_set_begin(e, num_particles < 0 ? 0 : num_particles);
    generate_plummer(e, e.n, a_scale, mass_per_body);
    replace_set_end(e);
}

int32_t nbx_tune(int32_t bodies_per_thread, int32_t target_waves, int32_t ctas_per_sm) {
    NB_LOCK();
    Engine& e = engine();
    if (!(bodies_per_thread == 0 || bodies_per_thread == 1 || bodies_per_thread == 2 || bodies_per_thread == 4)) {
        set_error("bodies_per_thread must be 0, 1, 2 or 4");
        return -1;
    }
    e.tune.bodies_per_thread = bodies_per_thread;
    e.tune.target_waves = target_waves;
    e.tune.ctas_per_sm = ctas_per_sm;
    if (const char* sh = getenv("NB_SHARE_RCP")) e.tune.share_rcp = atoi(sh);   // experiment knob
    return 0;
}

void nbx_get_counters(nbx_counters* out) {
    NB_LOCK();
    if (engine().inited) bh_poll(engine());
    *out = engine().ctr;
}
void nbx_reset_counters(void) {
    NB_LOCK();
    if (engine().inited) bh_poll(engine());
    engine().ctr = nbx_counters{};
}
int32_t nbx_bh_count_interactions(int32_t enable) {
    NB_LOCK();
    engine().bh_count = enable != 0;
    return 0;
}

int32_t nbx_bh_pop_histogram(uint64_t* out33, int32_t reset) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    bh_pop_histogram(e, out33, reset != 0);
    return 0;
}

int32_t nbx_bh_flatten(float* out9, int32_t cap) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    return bh_flatten(e, out9, cap);
}

int32_t nbx_bh_partition(int32_t parts) {
    NB_LOCK();
    if (parts < 0 || parts > kMaxRanks) {
        set_error("parts must be 0..%d", kMaxRanks);
        return -1;
    }
    engine().bh_partition = parts;
    return 0;
}

int32_t nbx_phase_timing(int32_t enable) {
    NB_LOCK();
    Engine& e = engine();
    e.phase_timing = enable != 0;
    e.ev_slot = 0;
    for (int p = 0; p < NBX_NUM_PHASES; p++) {
        e.ev_count[p] = 0;
        for (int k = 0; k < Engine::kPhaseRing; k++) e.ev_sub[p][k] = 0;
    }
    return 0;
}
int32_t nbx_get_phase_ms(float* out8) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    collect_phase_times(e);
    memcpy(out8, e.phase_ms, sizeof(float) * NBX_NUM_PHASES);
    return 0;
}

int32_t nbx_get_phase_sub_ms(int32_t phase, float* out4) {
    NB_LOCK();
    if (phase < 0 || phase >= NBX_NUM_PHASES) return -1;
    memcpy(out4, engine().phase_sub_ms[phase], sizeof(float) * 4);   // filled by the last nbx_get_phase_ms
    return 0;
}

int32_t nbx_accelerations(float* axy_out, int32_t n) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    if (e.n == 0 || n <= 0) return 0;
    accelerations_local(e);
    // each rank returns its own rows at their global position; other rows are left untouched
    const int b = local_begin(e), c0 = local_count(e);
    int c = c0;
    if (b + c > n) c = n - b;
    if (c > 0)
        NB_CUDA(cudaMemcpyAsync(axy_out + 2 * static_cast<size_t>(b), e.force, sizeof(float2) * c, cudaMemcpyDeviceToHost,
                                e.stream));
    NB_CUDA(cudaStreamSynchronize(e.stream));
    return 0;
}

int32_t nbx_bh_accelerations(float theta, float* axy_out, int32_t n) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    if (e.n == 0 || n <= 0) return 0;
    bh_accelerations(e, theta, e.force);
    const int b = local_begin(e);
    int c = local_count(e);
    if (b + c > n) c = n - b;
    if (c > 0)
        NB_CUDA(cudaMemcpyAsync(axy_out + 2 * static_cast<size_t>(b), e.force, sizeof(float2) * c, cudaMemcpyDeviceToHost,
                                e.stream));
    NB_CUDA(cudaStreamSynchronize(e.stream));
    return 0;
}

int32_t nbx3_num_particles(void) {
    NB_LOCK();
    return engine().inited ? x3_num(engine()) : 0;
}
int32_t nbx3_set_particles(const float* aos7, int32_t n) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    x3_set(e, aos7, n);
    return 0;
}
int32_t nbx3_get_particles(float* aos7_out, int32_t n) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    x3_get(e, aos7_out, n);
    return 0;
}
int32_t nbx3_configure(int32_t law, float eps2) {
    NB_LOCK();
    if ((law != NBX3_LAW_NEWTON && law != NBX3_LAW_REF) || !(eps2 > 0.0f)) {
        set_error("nbx3_configure: law must be NBX3_LAW_NEWTON or NBX3_LAW_REF and eps2 > 0");
        return -1;
    }
    Engine& e = engine();
    ensure_init(e);
    x3_config(e, law, eps2);
    return 0;
}
int32_t nbx3_set_sharded(int32_t on) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    x3_set_sharded(e, on);
    return 0;
}
int32_t nbx3_step_all_pairs(float dt) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    x3_step(e, dt, 1, nullptr);
    return 0;
}
int32_t nbx3_accelerations(float* axyz_out, int32_t n) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    if (n < x3_num(e)) {
        set_error("nbx3_accelerations: output holds %d bodies, the set has %d", n, x3_num(e));
        return -1;
    }
    x3_step(e, 0.0f, 0, axyz_out);
    return 0;
}

int32_t nbx_dist_init(int32_t rank, int32_t world, int32_t max_particles) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    return dist_init(e, rank, world, max_particles);
}
int32_t nbx_dist_handle_bytes(void) { return 128; }
int32_t nbx_dist_export(void* handle_out) {
    NB_LOCK();
    return dist_export(engine(), handle_out);
}
int32_t nbx_dist_import(const void* all_handles, int32_t world) {
    NB_LOCK();
    return dist_import(engine(), all_handles, world);
}
int32_t nbx_dist_nccl_unique_id(void* id_out128) { return dist_nccl_unique_id(id_out128); }
int32_t nbx_dist_nccl_init(const void* id128) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    return dist_nccl_init(e, id128);
}
int32_t nbx_dist_set_transport(int32_t transport) {
    NB_LOCK();
    if (transport < 0 || transport > 2) {
        set_error("unknown transport %d", transport);
        return -1;
    }
    engine().transport = transport;
    return 0;
}
float nbx_dist_sync_test(int32_t iters) {
    NB_LOCK();
    Engine& e = engine();
    ensure_init(e);
    return bh_sync_test(e, iters);
}
int32_t nbx_dist_local_range(int32_t* begin, int32_t* count) {
    NB_LOCK();
    Engine& e = engine();
    *begin = local_begin(e);
    *count = local_count(e);
    return 0;
}

}  // extern "C"
