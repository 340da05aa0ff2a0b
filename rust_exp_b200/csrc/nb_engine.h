// nb_engine.h -- host-side engine behind the nb_* C ABI: device-resident particle state, step
// sequencing, multi-GPU wiring.  Mirrors the role of the process-global PARTICLES vector of the
// reference (rs-src/nbody.rs:28-32); one Engine per process, guarded by one mutex (the reference holds
// its mutex for the whole of every nb_* call).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "../../include/nbody_b200.h"
#include "nb_common.cuh"

namespace nb {

constexpr int kMaxRanks = 8;
constexpr int kShardAlign = 1024;  // shard length granularity (bodies); j tiles divide it
constexpr int kFlagSlots = 64;

// Symmetric arena: the part of a rank's state that peers may read over NVLink.  One cudaMalloc, one
// IPC handle.  All offsets are identical on every rank.  L = shard_len (bodies per rank, padded).
//   x[2][L]  y[2][L]   ping-pong position buffers (step s reads buffer s&1, writes (s+1)&1)
//   m[L]               masses (constant during stepping; 0 for padding slots)
//   vx[L] vy[L]        velocities of the local shard (read by peers only for nb_get_particles)
//   flags[kFlagSlots]  uint32 step counters written by peers (flags[r] = steps rank r has completed); at offset 0
struct ArenaLayout {
    size_t L = 0;
    size_t off_x[2] = {0, 0}, off_y[2] = {0, 0}, off_m = 0, off_vx = 0, off_vy = 0, off_flags = 0, bytes = 0;
    // flags come FIRST: their address must not move when the shard length follows the size of the set
    void set(size_t shard_len) {
        L = shard_len;
        size_t o = 0, f = shard_len * sizeof(float);
        off_flags = o; o += kFlagSlots * sizeof(uint32_t);
        off_x[0] = o; o += f;
        off_x[1] = o; o += f;
        off_y[0] = o; o += f;
        off_y[1] = o; o += f;
        off_m = o; o += f;
        off_vx = o; o += f;
        off_vy = o; o += f;
        bytes = (o + 255) & ~size_t(255);
    }
};

// Second symmetric allocation, used by the domain-partitioned Barnes-Hut step (nb_bh.cu).  Everything a PEER touches
// lives here; offsets are identical on every rank.
//   flags[4][64]        epoch flags of the step's four cross-rank ordering points (boxes / bodies / trees / walks)
//   aabb[G][4]          aabb[s] = rank s's local bounding box (ordered-int encoded), written by rank s
//   count_in[G]         count_in[s] = bodies rank s delivered into inbox region s this step, written by rank s
//   cellwork[2][cells]  walk cost per cut-level cell measured by this rank's walk (double-buffered by step parity)
//   celltab[cells]      the cut-level cell table: every rank stores the entries of the cells it owns into EVERY rank's copy
//   workpub[cells]      last step's walk cost per cell, stored here by whichever rank walked the cell (same broadcast)
//   in_key/in_rec[G][R] inbox: region s receives the (key-sorted) bodies rank s owns by index that fall into THIS rank's
//                       cells -- peer stores over NVLink (the all-to-all-v of the step)
//   nblk / ncblk        this rank's subtree forest (block-SoA records + child indices), walked in place by the peers
//   acc[L_cap]          accelerations of the bodies this rank owns by index, scattered here by whoever walked them
constexpr int kBhCutLevel = 5;
constexpr int kBhNumCells = 1 << (2 * kBhCutLevel);   // 1024 cells at the cut level
constexpr int kBhFlagRows = 5;   // boxes / bodies / trees / walks + one row for nbx_dist_sync_test
struct BhArenaLayout {
    int world = 1;
    size_t R = 0;            // inbox region capacity (bodies) per source rank = shard capacity
    size_t cap_bodies = 0;   // most bodies one part may hold: the whole set (any imbalance is legal, only slow)
    int cap_blocks = 0;
    size_t off_flags = 0, off_aabb = 0, off_count_in = 0, off_cellwork = 0, off_celltab = 0, off_workpub = 0, off_in_key = 0, off_in_rec = 0,
           off_nblk = 0, off_ncblk = 0, off_acc = 0, bytes = 0;
    void set(size_t shard_cap, int world_, size_t max_particles) {
        world = world_;
        R = shard_cap;
        cap_bodies = max_particles;
        size_t cb = 2 * max_particles + 4096;
        if (cb > (size_t(1) << 22) - 1) cb = (size_t(1) << 22) - 1;
        cap_blocks = static_cast<int>(cb);
        size_t o = 0;
        auto take = [&](size_t bytes_) { size_t at = o; o = (o + bytes_ + 255) & ~size_t(255); return at; };
        off_flags = take(kBhFlagRows * 64 * sizeof(uint32_t));
        off_aabb = take(kMaxRanks * 4 * sizeof(int));
        off_count_in = take(kMaxRanks * sizeof(int));
        off_cellwork = take(2 * kBhNumCells * sizeof(unsigned));
        off_celltab = take(kBhNumCells * 64);
        off_workpub = take(kBhNumCells * sizeof(unsigned));
        off_in_key = take(static_cast<size_t>(world) * R * 8);
        off_in_rec = take(static_cast<size_t>(world) * R * 16);
        off_nblk = take(cb * 64);
        off_ncblk = take(cb * 16);
        off_acc = take(shard_cap * 8);
        bytes = o;
    }
};

struct ArenaView {
    char* base = nullptr;
    float* x(const ArenaLayout& l, int b) const { return reinterpret_cast<float*>(base + l.off_x[b]); }
    float* y(const ArenaLayout& l, int b) const { return reinterpret_cast<float*>(base + l.off_y[b]); }
    float* m(const ArenaLayout& l) const { return reinterpret_cast<float*>(base + l.off_m); }
    float* vx(const ArenaLayout& l) const { return reinterpret_cast<float*>(base + l.off_vx); }
    float* vy(const ArenaLayout& l) const { return reinterpret_cast<float*>(base + l.off_vy); }
    uint32_t* flags(const ArenaLayout& l) const { return reinterpret_cast<uint32_t*>(base + l.off_flags); }
};

// j-side source of one rank's shard for the all-pairs kernels.
struct JSeg {
    const float* x;
    const float* y;
    const float* m;
};

struct AllPairsArgs {
    JSeg seg[kMaxRanks];       // seg[g] = positions/masses of global bodies [g*L, (g+1)*L)
    int nseg;                  // world size
    int seg_len;               // L
    int n_total;               // real bodies in the global set (exact mode loops j < n_total)
    const float* xi;           // local shard positions (current buffer)
    const float* yi;
    const float* mi;
    int i_global_begin;        // global index of local body 0
    int n_local;               // real local bodies
    // fast path decomposition
    int slice_len;             // j bodies per work item (multiple of the j tile)
    int slices_per_seg;
    float2* partial;           // [nseg*slices_per_seg][L] partial accelerations
    // cross-rank ordering (P2P_DIRECT): wait until flags[g] >= wait_step before touching seg g != me
    const uint32_t* flags;
    uint32_t wait_step;
    int my_rank;
    unsigned long long timeout_ns;   // bound on the cross-rank waits (0 = none)
    // P2P_DIRECT, staged: the first 2*(nseg-1) CTAs first copy one half of one REMOTE segment each from the peer's
    // HBM (src[]) into the local mirror, once per step, then join the work loop; every j tile is then streamed from
    // local memory (seg[] points at the mirror for remote segments).  NVLink volume = the all-gather's, not x n_itiles.
    int stage;                       // 0: seg[] are read in place (peer memory for remote segments)
    JSeg src[kMaxRanks];             // where the stagers read segment g (peer arena)
    JSeg dst[kMaxRanks];             // where they write it (local mirror) == seg[g] for remote g
    int stage_mass;                  // masses are constant between set calls: staged on the first step only
    uint32_t* staged;                // staged[g] += 1 per finished half
    uint32_t stage_want;             // producers read segment g once staged[g] >= stage_want
};

struct Tuning {
    int bodies_per_thread = 0;  // 0 = auto
    int target_waves = 0;
    int ctas_per_sm = 0;
    int share_rcp = -1;         // -1 = library default
};

struct Engine {
    std::mutex mu;
    bool inited = false;
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int mode = NBX_MODE_FAST;
    uint64_t seed = 0x5eed5eedULL;
    Tuning tune;

    // global set
    int n = 0;
    // sharding
    int rank = 0, world = 1;
    bool dist = false;
    int transport = NBX_TRANSPORT_P2P_DIRECT;
    int max_particles = 0;  // arena capacity in the dist case
    size_t L_cap = 0;       // dist: shard capacity (bodies per rank) the arena was allocated and exported for
    ArenaLayout lay;        // layout for the CURRENT set: L = balanced shard length, <= L_cap when sharded
    size_t arena_cap_bytes = 0;
    ArenaView arena;                 // own
    ArenaView peer[kMaxRanks];       // peer[rank] == arena
    bool peers_mapped = false;
    BhArenaLayout bh_lay;
    char* bh_arena = nullptr;        // own (sharded mode only)
    char* bh_peer[kMaxRanks] = {};   // bh_peer[rank] == bh_arena
    int bh_partition = 0;            // 0 auto, 1 force the replicated tree, >1 virtual parts on one GPU
    uint64_t peer_timeout_ns = 30000000000ull;   // device-side waits on peers trap after this long (0 = never)
    int cur = 0;                     // current position buffer
    uint32_t step_count = 0;         // steps completed (also the cross-rank flag value)
    // opt-in integrator (rs-src/nbody.rs:449 TODO): leapfrog kick-drift-kick, FAST mode.  Velocities are kept at the half
    // step between steps; kdk_pending = the half kick (dt/2) still owed, applied before any read-back (kdk_close).
    int integrator = NBX_INTEGRATOR_EULER;
    float kdk_pending = 0.f;
    float kdk_theta = 0.f;           // force method of the last step (0 = all-pairs), used for the closing kick
    int square_aabb = 0;             // opt-in (rs-src/nbody.rs:400-407): square up the root box of the quadtree
    uint64_t set_gen = 0;            // bumped whenever the particle set is replaced (set / generators)
    // full mirror of all shards (GATHER / NCCL transports and Barnes-Hut): x,y,m of G*L bodies
    float* mirror = nullptr;
    size_t mirror_cap = 0;
    bool mirror_mass_valid = false;  // masses are constant between set/generate calls: gathered once
    uint32_t* stage_flags = nullptr; // [kMaxRanks] halves staged per remote segment (P2P_DIRECT staged)
    uint32_t stage_epoch = 0;
    // scratch
    float2* partial = nullptr;
    size_t partial_cap = 0;  // in float2
    float2* force = nullptr; // exact-mode forces / accelerations [L]
    size_t force_cap = 0;
    float* stage_dev = nullptr;
    size_t stage_dev_cap = 0;

    // counters / timing
    nbx_counters ctr{};
    bool bh_count = false;
    bool phase_timing = false;
    // ring of event pairs per phase: one slot per step since the last nbx_get_phase_ms, so the
    // caller gets the AVERAGE device time per step of every phase over its timed region
    // a phase may be entered several times per step (the cross-rank waits are): kPhaseSub sub-intervals per slot
    static constexpr int kPhaseRing = 128;
    static constexpr int kPhaseSub = 4;
    cudaEvent_t ev[NBX_NUM_PHASES][kPhaseRing][kPhaseSub][2] = {};
    unsigned char ev_sub[NBX_NUM_PHASES][kPhaseRing] = {};
    int ev_count[NBX_NUM_PHASES] = {};
    int ev_slot = 0;  // step index within the ring
    float phase_ms[NBX_NUM_PHASES] = {};
    float phase_sub_ms[NBX_NUM_PHASES][4] = {};   // the same, per sub-interval of a step (entry order)

    // Barnes-Hut workspace, opaque (nb_bh.cu)
    void* bh = nullptr;
    // 3-D extension state, opaque (nb_3d.cu)
    void* ext3 = nullptr;
};

Engine& engine();
const char* last_error();
int try_init(Engine& e, int device);
void ensure_init(Engine& e);
void step_brute_force(Engine& e, float dt);
void accelerations_local(Engine& e);   // a_i of the local shard -> e.force
void ensure_capacity(Engine& e, int n);   // (re)allocate arena/scratch for n global bodies
int local_begin(const Engine& e);
int local_count(const Engine& e);

// phase timing helpers
struct PhaseScope {
    Engine& e;
    int id;
    int sub = -1;
    PhaseScope(Engine& e_, int id_);
    ~PhaseScope();
};
void collect_phase_times(Engine& e);

// nb_state.cu
void state_upload_aos(Engine& e, const float* aos5, int n);
void state_download_aos(Engine& e, float* aos5, int n);
void state_download_local(Engine& e, float* aos5_full, int n);
void launch_integrate_fast(Engine& e, const float2* partial, int nslices, float dt, bool kill);
void launch_integrate_exact(Engine& e, const float2* force, float dt, bool kill);
void launch_kick(Engine& e, const float2* acc, float h);
inline float kick_dt(const Engine& e, float dt) { return e.integrator == NBX_INTEGRATOR_LEAPFROG_KDK ? e.kdk_pending + 0.5f * dt : dt; }
void kdk_close(Engine& e);   // nb_engine.cu: apply the owed half kick (one force evaluation) so that v is synchronised with x
void launch_fill_zero_f32(Engine& e, float* p, size_t n);
void generate_random_disk(Engine& e, int n);
void generate_stable_orbits(Engine& e, int n, float rmin, float rmax);
void generate_plummer(Engine& e, int n, float a_scale, float mass_per_body);

// nb_allpairs.cu
void allpairs_slices(const Engine& e, int n_local, int L, int nseg, int* slice_len, int* per_seg);
void allpairs_plan(Engine& e, AllPairsArgs& a);  // fills decomposition fields, grows scratch
void launch_allpairs_fast(Engine& e, const AllPairsArgs& a);
void launch_allpairs_exact(Engine& e, const AllPairsArgs& a, float2* force_out);
void launch_accel_from_partial(Engine& e, const float2* partial, int nslices, float2* out);
void launch_accel_from_force(Engine& e, const float2* force, float2* out);

// nb_dist.cu
void dist_fill_segments(Engine& e, AllPairsArgs& a);   // seg[] according to transport; may enqueue gather
void dist_signal_step_done(Engine& e);                 // publish step_count to peers
void dist_wait_all(Engine& e, uint32_t step);          // device-side wait for all peers
void dist_gather_mirror(Engine& e, int buf);           // mirror <- all shards (own kernel or NCCL)
void dist_shutdown(Engine& e);
void dist_require_peers(const Engine& e);               // fatal unless the peers' arenas are mapped
int dist_init(Engine& e, int rank, int world, int max_particles);
int dist_export(Engine& e, void* out64);
int dist_import(Engine& e, const void* all, int world);
int dist_nccl_unique_id(void* out128);
bool dist_nccl_ready();
void dist_nccl_allgather_inplace(Engine& e, float* const* bufs, int nbufs, size_t count);
int dist_nccl_init(Engine& e, const void* id128);

// nb_bh.cu
void bh_step(Engine& e, float theta, float dt);
void bh_accelerations(Engine& e, float theta, float2* out);
void bh_shutdown(Engine& e);
int bh_flatten(Engine& e, float* out9, int cap);   // FAST tree of the last BH call, oracle flatten format
void bh_poll(Engine& e);
float bh_sync_test(Engine& e, int iters);   // average ms of one cross-rank ordering point (diagnostic)
void bh_pop_histogram(Engine& e, uint64_t* out33, bool reset);   // fold the last Barnes-Hut step's status/counters in (synchronises)

// nb_3d.cu
void x3_set(Engine& e, const float* aos7, int n);
void x3_get(Engine& e, float* aos7, int n);
int x3_num(Engine& e);
void x3_config(Engine& e, int law, float eps2);
void x3_set_sharded(Engine& e, int on);
void x3_step(Engine& e, float dt, int update, float* acc_host);
void x3_shutdown(Engine& e);

// nb_draw.cu
void draw_to_host(Engine& e, int w, int h, uint32_t* fb);
void draw_shutdown();

}  // namespace nb
