/*
 * nbody_b200.h -- C ABI of libnbody_b200.so, the B200-native replacement for the N-body hot path of
 * blitzcode/rust-exp (rs-src/nbody.rs).
 *
 * The six nb_* symbols below are exactly the `#[no_mangle] pub extern fn` surface that
 * hs-src/RustNBodyExperiment.hs:101-106 binds with `foreign import ccall`; names, argument order,
 * C types and (void) error behaviour are the reference's.  State is a hidden process-global particle
 * set that survives between calls (rs-src/nbody.rs:28-32), here resident in GPU HBM.
 *
 * Two further nb_* symbols are REQUIRED ADDITIONS (the reference has no state getter/setter and seeds
 * its initial conditions from an unseeded RNG, so parity cannot be tested without them).
 *
 * Everything prefixed nbx_ is an extension surface (device selection, bit-exact mode, stream
 * injection, multi-GPU wiring, counters).  None of it is needed for single-GPU drop-in use.
 *
 * Error behaviour: the reference panics across the FFI boundary (rs-src/nbody.rs:231,246,267,...);
 * the nb_* functions return void, so on an unrecoverable CUDA error they print to stderr and abort().
 * nbx_* functions return 0 on success and a negative code on failure (message in nbx_last_error()).
 * There is NO CPU fallback: every entry point that computes requires a CUDA device.
 */
#ifndef NBODY_B200_H
#define NBODY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Reference surface (drop-in).
 * ---------------------------------------------------------------------------------------------- */

/* rs-src/nbody.rs:34-37.  Number of particles in the global set (all ranks' particles when sharded). */
int32_t nb_num_particles(void);

/* rs-src/nbody.rs:39-64.  Replace the set with n particles uniformly placed in a disk of radius 23,
 * v in [-3.5,3.5)^2, m in [0.1,1.5).  Distribution-level match only (the reference RNG is unseeded);
 * generated on the device from a counter-based RNG seeded by nbx_seed(). */
void nb_random_disk(int32_t num_particles);

/* rs-src/nbody.rs:73-104.  Sun (m=1000, at rest at the origin) + n-1 unit-mass planets on circular
 * orbits with r in [rmin,rmax), speed sqrt(1000) independent of r (2-D log-potential gravity). */
void nb_stable_orbits(int32_t num_particles, float rmin, float rmax);

/* rs-src/nbody.rs:106-162.  One O(N^2) step: forces from OLD positions, then v += (dt*F)/m,
 * p += dt*v.  Returns when the GPU has finished the step (the reference's call is synchronous and its host times it
 * with a wall clock, hs-src/RustNBodyExperiment.hs:55-57); see nbx_set_async for the enqueue-only variant. */
void nb_step_brute_force(float dt);

/* rs-src/nbody.rs:186-480.  NOTE argument order: theta first.  theta == 0.0 is the brute-force step
 * (and then no velocity kill, rs-src/nbody.rs:197-200).  Otherwise: tight non-square AABB, quadtree
 * over old positions, per-body tree force with opening test (x2-x1)/d < theta, Euler, then
 * v = 0 for bodies with |px| > 55 or |py| > 55 (rs-src/nbody.rs:466-471).
 * nthreads is the reference's CPU thread-count hint (UI range 1..16): accepted and ignored, except
 * that nthreads <= 0 moves no body, like the reference: its per-thread closure (which holds the division by
 * nthreads, rs-src/nbody.rs:424-428) is mapped over the empty range 0..nthreads, so the tree is built and dropped. */
void nb_step_barnes_hut(float theta, float dt, int32_t nthreads);

/* rs-src/nbody.rs:482-583.  Render into the caller-owned w*h ABGR8 framebuffer `fb` (HOST memory, a
 * mapped GL PBO in the reference host, hs-src/FrameBuffer.hs:127-135).  The image is composed on the
 * device and written to fb with one contiguous copy; fb is never read. */
void nb_draw(int32_t w, int32_t h, uint32_t *fb);

/* ------------------------------------------------------------------------------------------------
 * Required additions: inject / read back state.  AoS records of 5 floats
 * {px,py,vx,vy,m}, 20-byte stride, the field order of `struct Particle` (rs-src/nbody.rs:19-26).
 * Host pointers.  When sharded, every rank passes the same full array to nb_set_particles and
 * nb_get_particles returns the full, gathered state on every rank.
 * ---------------------------------------------------------------------------------------------- */
void nb_set_particles(const float *aos5, int32_t n);
void nb_get_particles(float *aos5_out, int32_t n);

/* ------------------------------------------------------------------------------------------------
 * Prefixed aliases of the eight symbols above -- the same functions under names that cannot collide.
 * The reference links ONE static archive (Cargo.toml:5-7 crate-type staticlib, rust-exp.cabal:46-47) whose
 * nbody module must keep exporting nb_*: the replacement module (rust-shim/nbody.rs) defines nb_* as
 * #[no_mangle] forwarders to these, and libnbody_b200.a is linked into the archive.  In the static library the
 * unprefixed names live in a separate object (nb_alias.o) that the linker only pulls in when nothing else
 * defines nb_*, so both integration styles link without duplicate symbols.
 * ---------------------------------------------------------------------------------------------- */
int32_t b200_nb_num_particles(void);
void b200_nb_random_disk(int32_t num_particles);
void b200_nb_stable_orbits(int32_t num_particles, float rmin, float rmax);
void b200_nb_step_brute_force(float dt);
void b200_nb_step_barnes_hut(float theta, float dt, int32_t nthreads);
void b200_nb_draw(int32_t w, int32_t h, uint32_t *fb);
void b200_nb_set_particles(const float *aos5, int32_t n);
void b200_nb_get_particles(float *aos5_out, int32_t n);

/* ------------------------------------------------------------------------------------------------
 * Extension surface.
 * ---------------------------------------------------------------------------------------------- */

/* Bind the library to a CUDA device (default: device 0 on first use).  Returns 0, or <0 if no usable
 * sm_100 device exists -- there is no CPU fallback. */
int32_t nbx_init(int32_t device);
void nbx_shutdown(void);
const char *nbx_last_error(void);
const char *nbx_version(void);

/* Arithmetic mode.  FAST (default): throughput kernels (packed-FP32 FMA + MUFU.RCP, tiled sums).
 * EXACT: every operation is a separately rounded IEEE binary32 op in the reference's order
 * (sequential ascending-j accumulation, (m1*m2)/(d2+EPS), true division, nested tree sums) and the
 * result is bit-identical to the CPU restatement of rs-src/nbody.rs. */
#define NBX_MODE_FAST 0
#define NBX_MODE_EXACT 1
int32_t nbx_set_mode(int32_t mode);
int32_t nbx_get_mode(void);

/* Opt-in integrator (the reference author's TODO at rs-src/nbody.rs:449, "Try different integration scheme, Verlet /
 * RK4 etc.").  EULER (default) is the reference's semi-implicit Euler.  LEAPFROG_KDK is kick-drift-kick leapfrog
 * (velocity Verlet): v += dt/2 a(x); x += dt v; v += dt/2 a(x') -- second order and time-reversible, one force
 * evaluation per step (consecutive half kicks are merged) plus one when the state is read back.  It changes results,
 * so it is never the default; it applies to FAST mode only (EXACT always integrates like the reference). */
#define NBX_INTEGRATOR_EULER 0
#define NBX_INTEGRATOR_LEAPFROG_KDK 1
int32_t nbx_set_integrator(int32_t integrator);

/* Opt-in tree shape (the reference author's TODO at rs-src/nbody.rs:400-407): square up the root box of the quadtree,
 * exactly as the commented-out code there does (the shorter side is extended from the min corner).  Default 0 = the
 * reference's tight, non-square box. */
int32_t nbx_set_square_aabb(int32_t enable);

/* Launch all work on a caller-owned cudaStream_t (e.g. the host framework's current stream) so the
 * caller's CUDA events bracket it.  NULL restores the library's own stream. */
int32_t nbx_set_stream(void *cuda_stream);
int32_t nbx_synchronize(void);

/* Step calls are synchronous by default (they return when the GPU has finished, like the reference's CPU call).
 * nbx_set_async(1) makes nb_step_* return after enqueueing on the library stream -- any nb_get_particles / nb_draw /
 * nbx_synchronize still orders after them.  NB_ASYNC_STEPS=1 in the environment selects the same at start-up. */
int32_t nbx_set_async(int32_t enable);

/* Seed for nb_random_disk / nb_stable_orbits. */
void nbx_seed(uint64_t seed);

/* Replaces the set with a Plummer model generated on the device (same counter-based RNG): scale radius a_scale,
 * truncated at 10 a_scale, projected to z = 0, equal masses, isotropic Gaussian velocities with the local dispersion
 * -- the synthetic cloud of BASELINE.json configurations 2-5.  The reference has no such generator (its two are
 * nb_random_disk and nb_stable_orbits); like those it aborts on invalid arguments. */
void nbx_plummer(int32_t num_particles, float a_scale, float mass_per_body);

/* Kernel-level tuning knobs (0 = library default): bodies per thread of the all-pairs kernel,
 * target waves of work items, resident CTAs per SM. */
int32_t nbx_tune(int32_t bodies_per_thread, int32_t target_waves, int32_t ctas_per_sm);

/* Counters since the last reset: kernels launched by this library, pair interactions evaluated by the
 * all-pairs path, Barnes-Hut interactions and visited nodes (the latter two only when counting is on). */
typedef struct nbx_counters {
    uint64_t kernel_launches;
    uint64_t allpairs_pairs;
    uint64_t bh_interactions;
    uint64_t bh_nodes_visited;
    uint64_t bh_nodes_built;
    uint64_t steps;
    /* lane efficiency of the warp-cooperative tree walk (counting mode only): stack entries popped by all warps,
     * and the sum over those pops of the number of lanes (bodies) that needed the entry.  Every pop is evaluated
     * by all 32 lanes, so bh_pop_lanes / (32 * bh_pops) is the fraction of useful lane work. */
    uint64_t bh_pops;
    uint64_t bh_pop_lanes;
    /* gauges of the most recent Barnes-Hut step whose status was read back: bodies in this rank's domain part
     * (partitioned multi-GPU step; n otherwise) and the key levels its sort ordered on */
    uint64_t bh_part_bodies;
    uint64_t bh_sort_levels;
} nbx_counters;
void nbx_get_counters(nbx_counters *out);
void nbx_reset_counters(void);
int32_t nbx_bh_count_interactions(int32_t enable);
/* Histogram (33 bins, HOST array) of the number of lanes per popped stack entry since the previous call with
 * reset != 0; filled only while counting is enabled. */
int32_t nbx_bh_pop_histogram(uint64_t *out33, int32_t reset);

/* Parity aid: the quadtree of the most recent FAST Barnes-Hut call in DFS pre-order (children UL,UR,LL,LR),
 * 9 floats per node {x1,y1,x2,y2, px,py,m, has_children, depth} -- the layout the CPU oracle dumps -- into a HOST
 * array holding `cap` nodes.  Returns the number of nodes of the tree (may exceed cap), <0 on error. */
int32_t nbx_bh_flatten(float *out9, int32_t cap);

/* Domain partitioning of the FAST Barnes-Hut step: 0 = automatic (one part per GPU when
 * sharded over several GPUs and the set has >= 65,536 bodies; one tree otherwise), 1 = always one replicated
 * tree, 2..8 on a single GPU = that many "virtual ranks" (exercises the partitioned code path for tests). */
int32_t nbx_bh_partition(int32_t parts);

/* Device time (ms, CUDA events on the library stream) of the phases of the most recent step:
 * [0] all-pairs / traversal kernel  [1] integrate  [2] aabb  [3] keys  [4] sort  [5] tree build
 * [6] centre-of-mass  [7] cross-rank wait/gather.  Only recorded when enabled (adds event overhead). */
#define NBX_NUM_PHASES 8
int32_t nbx_phase_timing(int32_t enable);
int32_t nbx_get_phase_ms(float *out8);
/* A phase may be entered several times per step (the cross-rank ordering points are: boxes, bodies, trees, walks): the
 * per-entry averages behind the last nbx_get_phase_ms call, in entry order (4 floats). */
int32_t nbx_get_phase_sub_ms(int32_t phase, float *out4);

/* Accelerations only (no state update) of the current set with the current mode, a_i = F_i / m_i in
 * the reference's units, written to a HOST array of 2*n floats.  Used by sampled parity checks. */
int32_t nbx_accelerations(float *axy_out, int32_t n);
/* Same through the Barnes-Hut tree (build + traversal with opening angle theta, no state update). */
int32_t nbx_bh_accelerations(float theta, float *axy_out, int32_t n);

/* ---- 3-D extension (NOT part of the drop-in surface) ------------------------------------
 * A separate particle set of 7-float records {px,py,pz,vx,vy,vz,m} and an all-pairs step with a selectable
 * pair law.  LAW_NEWTON: a_i = sum m_j d/(|d|^2+eps2)^(3/2) (rsqrt-based, 20 flop/pair -- the kernel the
 * project brief describes).  LAW_REF: the reference's un-normalised law m_j d/(|d|^2+eps2) lifted to 3-D; with
 * z == 0 it reproduces the 2-D FAST kernel bit for bit.  Integrator as in the reference
 * (rs-src/nbody.rs:153-160).  Default: LAW_NEWTON, eps2 = 1e-4.
 * Multi-GPU: once the process group is wired WITH the NCCL communicator (nbx_dist_init + nbx_dist_nccl_init), the rows
 * shard by index -- every rank keeps all positions, evaluates and integrates its own rows and one in-place
 * ncclAllGather per coordinate exchanges the new positions (all nbx3_* calls are then collective).  The result equals
 * the single-GPU result bit for bit.  nbx3_set_sharded(0) makes every rank compute every row instead (default 1). */
#define NBX3_LAW_NEWTON 0
#define NBX3_LAW_REF 1
int32_t nbx3_num_particles(void);
int32_t nbx3_set_particles(const float *aos7, int32_t n);
int32_t nbx3_get_particles(float *aos7_out, int32_t n);
int32_t nbx3_configure(int32_t law, float eps2);
int32_t nbx3_set_sharded(int32_t on);
int32_t nbx3_step_all_pairs(float dt);
int32_t nbx3_accelerations(float *axyz_out, int32_t n);

/* ---- multi-GPU (one process per GPU; particles shard by index) -------------
 * Wiring is done by the host framework's process group (torch.distributed in this repo):
 *   1. every rank: nbx_dist_init(rank, world, max_particles) -> allocates the symmetric arena
 *   2. every rank: nbx_dist_export(handle)  ; all-gather the nbx_dist_handle_bytes()-byte handles ; nbx_dist_import(all)
 *   3. (NCCL transport only) rank 0: nbx_dist_nccl_unique_id ; broadcast ; all: nbx_dist_nccl_init
 * After that nb_set_particles / nb_step_* / nb_get_particles are collective calls.
 */
#define NBX_TRANSPORT_P2P_DIRECT 0 /* all-pairs kernel TMA-loads j tiles straight from peer HBM over NVLink */
#define NBX_TRANSPORT_P2P_GATHER 1 /* pull kernel copies peer shards into local HBM, then local kernel */
#define NBX_TRANSPORT_NCCL 2       /* ncclAllGather baseline */
int32_t nbx_dist_init(int32_t rank, int32_t world, int32_t max_particles);
int32_t nbx_dist_handle_bytes(void);
int32_t nbx_dist_export(void *handle_out);
int32_t nbx_dist_import(const void *all_handles, int32_t world);
int32_t nbx_dist_nccl_unique_id(void *id_out128);
int32_t nbx_dist_nccl_init(const void *id128);
int32_t nbx_dist_set_transport(int32_t transport);
/* Device-side waits on a peer rank give up (message + trap -> abort) after this long; 0 = wait for ever.
 * Default 30 s (NB_PEER_TIMEOUT_MS in the environment overrides): long enough for host-side skew between ranks
 * (initial-condition loading, rank-0-only I/O), short enough that a dead peer does not hang the GPU. */
int32_t nbx_set_peer_timeout_ms(int32_t ms);
int32_t nbx_dist_local_range(int32_t *begin, int32_t *count);
/* Diagnostic (collective): average device time in ms of one cross-rank ordering point -- a 32-thread kernel that stores
 * an epoch flag into every peer's arena and waits for every peer's -- over `iters` back-to-back launches. */
float nbx_dist_sync_test(int32_t iters);
/* Owner-only read-back: fills rows [begin, begin+count) (this rank's shard, see nbx_dist_local_range) of the
 * caller's FULL n-row AoS array and leaves the other rows untouched -- count*20 bytes of device->host traffic
 * instead of n*20 on every rank.  Not a collective.  On one GPU it equals nb_get_particles. */
int32_t nbx_get_particles_local(float *aos5_out_full, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* NBODY_B200_H */
