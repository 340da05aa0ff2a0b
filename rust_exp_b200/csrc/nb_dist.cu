// nb_dist.cu -- multi-GPU wiring for the index-sharded particle set.
//
// One process per GPU.  Each rank owns the bodies [rank*L, (rank+1)*L) and keeps them in a "symmetric
// arena" (nb_engine.h) that it exports with a CUDA IPC handle; peers map it and read it over
// NVLink 5 / NVSwitch.  Three transports feed the j side of the all-pairs kernel:
//
//   P2P_DIRECT  the all-pairs kernel's producer warp bulk-TMA-loads j tiles straight from the peer's HBM,
//               tile by tile, so the "all-gather" is fused into the force kernel: no gathered copy
//               ever exists, and the transfer overlaps the math (local segment first, no wait).
//   P2P_GATHER  a pull kernel copies the peers' shards into a local mirror, then the kernel reads locally.
//   NCCL        ncclAllGather into the mirror (the library baseline the other two are measured against).
//
// Cross-rank ordering is device-side: flags[r] in every arena holds the number of collective epochs
// rank r has completed; an epoch's remote reads wait for flags[g] >= epoch.  Positions ping-pong
// between two buffers, so one wait per epoch covers both the RAW and the WAR hazard.
#include <dlfcn.h>
#include <string.h>

#include "nb_engine.h"

namespace nb {

// ---- minimal NCCL binding via dlopen (so the library loads on hosts without NCCL) ----------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7 };

struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclComm_t comm = nullptr;
};

static Nccl* nccl_load() {
    static Nccl* n = nullptr;
    if (n) return n;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("cannot dlopen libnccl.so.2: %s", dlerror());
        return nullptr;
    }
    Nccl* t = new Nccl();
    t->h = h;
    *(void**)&t->GetUniqueId = dlsym(h, "ncclGetUniqueId");
    *(void**)&t->CommInitRank = dlsym(h, "ncclCommInitRank");
    *(void**)&t->CommDestroy = dlsym(h, "ncclCommDestroy");
    *(void**)&t->AllGather = dlsym(h, "ncclAllGather");
    *(void**)&t->GroupStart = dlsym(h, "ncclGroupStart");
    *(void**)&t->GroupEnd = dlsym(h, "ncclGroupEnd");
    *(void**)&t->GetErrorString = dlsym(h, "ncclGetErrorString");
    if (!t->GetUniqueId || !t->CommInitRank || !t->AllGather || !t->GroupStart || !t->GroupEnd) {
        set_error("libnccl is missing required symbols");
        delete t;
        return nullptr;
    }
    n = t;
    return n;
}

// ---- device-side epoch flags -------------------------------------------------------------------
struct PeerFlagPtrs {
    uint32_t* p[kMaxRanks];
};

__global__ void signal_kernel(PeerFlagPtrs peers, int world, int my_rank, uint32_t value) {
    const int g = threadIdx.x;
    if (g < world) {
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(peers.p[g] + my_rank) = value;
        __threadfence_system();
    }
}

__global__ void wait_all_kernel(const uint32_t* flags, int world, uint32_t want, unsigned long long timeout_ns) {
    const int g = threadIdx.x;
    if (g < world) wait_epoch(flags + g, want, timeout_ns, g, "epoch");
}

// mirror[g*L ..] <- peer g's current x / y / m, 16-byte vector loads over NVLink
struct GatherSrc {
    const float4* x[kMaxRanks];
    const float4* y[kMaxRanks];
    const float4* m[kMaxRanks];
};
__global__ void gather_kernel(GatherSrc src, int world, int L4, float4* __restrict__ mx, float4* __restrict__ my,
                              float4* __restrict__ mm, int with_mass) {
    const int g = blockIdx.y;
    if (g >= world) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L4; i += gridDim.x * blockDim.x) {
        mx[static_cast<size_t>(g) * L4 + i] = src.x[g][i];
        my[static_cast<size_t>(g) * L4 + i] = src.y[g][i];
        if (with_mass) mm[static_cast<size_t>(g) * L4 + i] = src.m[g][i];
    }
}

static float* mirror_x(Engine& e) { return e.mirror; }
static float* mirror_y(Engine& e) { return e.mirror + static_cast<size_t>(e.world) * e.lay.L; }
static float* mirror_m(Engine& e) { return e.mirror + 2 * static_cast<size_t>(e.world) * e.lay.L; }

static void ensure_mirror(Engine& e) {
    const size_t need = 3 * static_cast<size_t>(e.world) * e.lay.L;
    if (need > e.mirror_cap) {
        if (e.mirror) NB_CUDA(cudaFree(e.mirror));
        NB_CUDA(cudaMalloc(&e.mirror, need * sizeof(float)));
        e.mirror_cap = need;
        e.mirror_mass_valid = false;
    }
}

void dist_require_peers(const Engine& e) {
    if (e.dist && e.world > 1 && !e.peers_mapped) fatal("nbx_dist_import was not called (peer arenas are not mapped)", __FILE__, __LINE__);
}

void dist_wait_all(Engine& e, uint32_t epoch) {
    if (!e.dist || e.world == 1) return;
    dist_require_peers(e);
    wait_all_kernel<<<1, 32, 0, e.stream>>>(e.arena.flags(e.lay), e.world, epoch, e.peer_timeout_ns);
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
}

void dist_signal_step_done(Engine& e) {
    if (!e.dist || e.world == 1) return;
    dist_require_peers(e);
    PeerFlagPtrs p{};
    for (int g = 0; g < e.world; g++) p.p[g] = e.peer[g].flags(e.lay);
    signal_kernel<<<1, 32, 0, e.stream>>>(p, e.world, e.rank, e.step_count);
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
}

// Copy every rank's current positions (+ masses) into the local mirror.
void dist_gather_mirror(Engine& e, int buf) {
    ensure_mirror(e);
    const int L = static_cast<int>(e.lay.L);
    if (e.dist && e.world > 1 && e.transport == NBX_TRANSPORT_NCCL) {
        Nccl* n = nccl_load();
        if (!n || !n->comm) fatal("NCCL transport selected but nbx_dist_nccl_init was not called", __FILE__, __LINE__);
        // three all-gathers (x, y, m) of L floats per rank, grouped into one launch
        n->GroupStart();
        n->AllGather(e.arena.x(e.lay, buf), mirror_x(e), L, ncclFloat32, n->comm, e.stream);
        n->AllGather(e.arena.y(e.lay, buf), mirror_y(e), L, ncclFloat32, n->comm, e.stream);
        if (!e.mirror_mass_valid) n->AllGather(e.arena.m(e.lay), mirror_m(e), L, ncclFloat32, n->comm, e.stream);
        ncclResult_t r = n->GroupEnd();
        if (r != 0) fatal(n->GetErrorString ? n->GetErrorString(r) : "ncclAllGather failed", __FILE__, __LINE__);
        e.mirror_mass_valid = true;
        return;
    }
    dist_require_peers(e);
    dist_wait_all(e, e.step_count);
    GatherSrc s{};
    for (int g = 0; g < e.world; g++) {
        const ArenaView& av = (g == e.rank) ? e.arena : e.peer[g];
        s.x[g] = reinterpret_cast<const float4*>(av.x(e.lay, buf));
        s.y[g] = reinterpret_cast<const float4*>(av.y(e.lay, buf));
        s.m[g] = reinterpret_cast<const float4*>(av.m(e.lay));
    }
    const int L4 = L / 4;
    dim3 grid(static_cast<unsigned>(min(e.num_sms * 2, (L4 + 255) / 256)), static_cast<unsigned>(e.world));
    gather_kernel<<<grid, 256, 0, e.stream>>>(s, e.world, L4, reinterpret_cast<float4*>(mirror_x(e)),
                                             reinterpret_cast<float4*>(mirror_y(e)),
                                             reinterpret_cast<float4*>(mirror_m(e)), e.mirror_mass_valid ? 0 : 1);
    e.mirror_mass_valid = true;
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
}

// Point the all-pairs kernel's j segments at the right memory for the selected transport.
void dist_fill_segments(Engine& e, AllPairsArgs& a) {
    const int buf = e.cur;
    a.flags = nullptr;
    a.wait_step = e.step_count;
    if (!e.dist || e.world == 1) {
        a.seg[0] = JSeg{e.arena.x(e.lay, buf), e.arena.y(e.lay, buf), e.arena.m(e.lay)};
        return;
    }
    if (e.transport == NBX_TRANSPORT_P2P_DIRECT) {
        dist_require_peers(e);
        for (int g = 0; g < e.world; g++) {
            const ArenaView& av = (g == e.rank) ? e.arena : e.peer[g];
            a.seg[g] = JSeg{av.x(e.lay, buf), av.y(e.lay, buf), av.m(e.lay)};
        }
        if (e.mode == NBX_MODE_FAST) {
            a.flags = e.arena.flags(e.lay);  // the kernel's producer warp (or its stagers) wait per segment
            static const bool stage = [] { const char* v = getenv("NB_P2P_STAGE"); return v ? atoi(v) != 0 : true; }();
            // A rank without local bodies launches no force kernel (no items), so it must not count a staging epoch either:
            // its stage counters would fall behind stage_want for good and the next staged step would wait for ever
            // (found by tools/dist_check.py on 8 GPUs with a 20,000-body set: the last rank's shard is empty).
            if (stage && local_count(e) > 0) {
                // staged: each remote segment crosses NVLink once per step into the local mirror (2 stager CTAs per
                // segment, inside the force kernel), then every j tile streams from local HBM
                ensure_mirror(e);
                if (!e.stage_flags) {
                    NB_CUDA(cudaMalloc(&e.stage_flags, kMaxRanks * sizeof(uint32_t)));
                    NB_CUDA(cudaMemsetAsync(e.stage_flags, 0, kMaxRanks * sizeof(uint32_t), e.stream));
                    e.stage_epoch = 0;
                }
                e.stage_epoch++;
                a.stage = 1;
                a.staged = e.stage_flags;
                a.stage_want = 2u * e.stage_epoch;
                a.stage_mass = e.mirror_mass_valid ? 0 : 1;
                const size_t L = e.lay.L;
                for (int g = 0; g < e.world; g++) {
                    a.src[g] = a.seg[g];
                    a.dst[g] = JSeg{mirror_x(e) + g * L, mirror_y(e) + g * L, mirror_m(e) + g * L};
                    if (g != e.rank) a.seg[g] = a.dst[g];
                }
                e.mirror_mass_valid = true;
            }
        } else {
            dist_wait_all(e, e.step_count);
        }
        return;
    }
    // P2P_GATHER / NCCL: gather first, then everything is local
    dist_gather_mirror(e, buf);
    const size_t L = e.lay.L;
    for (int g = 0; g < e.world; g++) {
        a.seg[g] = JSeg{mirror_x(e) + g * L, mirror_y(e) + g * L, mirror_m(e) + g * L};
    }
    if (e.transport == NBX_TRANSPORT_NCCL) {
        // integrate overwrites v in place and the other position buffer; peers only read our arena through
        // get_particles, which is epoch-ordered, so no extra wait is needed here.
    }
}

void dist_shutdown(Engine& e) {
    if (e.stage_flags) { cudaFree(e.stage_flags); e.stage_flags = nullptr; e.stage_epoch = 0; }
    if (e.peers_mapped) {
        for (int g = 0; g < e.world; g++) {
            if (g != e.rank && e.peer[g].base) cudaIpcCloseMemHandle(e.peer[g].base);
            if (g != e.rank && e.bh_peer[g]) cudaIpcCloseMemHandle(e.bh_peer[g]);
            e.peer[g].base = nullptr;
            e.bh_peer[g] = nullptr;
        }
        e.peers_mapped = false;
    }
    Nccl* n = nccl_load();
    if (n && n->comm && n->CommDestroy) {
        n->CommDestroy(n->comm);
        n->comm = nullptr;
    }
    (void)cudaGetLastError();
}

// ---- C ABI pieces that live here -----------------------------------------------------------------
int dist_init(Engine& e, int rank, int world, int max_particles) {
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) {
        set_error("bad rank/world %d/%d (max %d ranks)", rank, world, kMaxRanks);
        return -1;
    }
    if (max_particles < 1) {
        set_error("max_particles must be >= 1");
        return -1;
    }
    if (e.dist) {
        set_error("nbx_dist_init called twice");
        return -2;
    }
    NB_CUDA(cudaStreamSynchronize(e.stream));
    if (e.arena.base) {
        NB_CUDA(cudaFree(e.arena.base));
        e.arena.base = nullptr;
    }
    const size_t per = (static_cast<size_t>(max_particles) + world - 1) / world;
    const size_t L = (per + kShardAlign - 1) / kShardAlign * kShardAlign;
    e.lay.set(L);
    e.L_cap = L;
    NB_CUDA(cudaMalloc(&e.arena.base, e.lay.bytes));
    NB_CUDA(cudaMemset(e.arena.base, 0, e.lay.bytes));
    if (world > 1) {
        e.bh_lay.set(L, world, static_cast<size_t>(max_particles));
        NB_CUDA(cudaMalloc(&e.bh_arena, e.bh_lay.bytes));
        NB_CUDA(cudaMemset(e.bh_arena, 0, e.bh_lay.bytes));
    }
    e.rank = rank;
    e.world = world;
    e.max_particles = max_particles;
    e.dist = true;
    e.n = 0;
    e.step_count = 0;
    e.cur = 0;
    for (int g = 0; g < kMaxRanks; g++) { e.peer[g].base = nullptr; e.bh_peer[g] = nullptr; }
    e.peer[rank] = e.arena;
    e.bh_peer[rank] = e.bh_arena;
    e.peers_mapped = (world == 1);
    return 0;
}

int dist_export(Engine& e, void* out) {
    if (!e.dist) {
        set_error("nbx_dist_init first");
        return -1;
    }
    cudaIpcMemHandle_t h;
    cudaError_t err = cudaIpcGetMemHandle(&h, e.arena.base);
    if (err != cudaSuccess) {
        set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(err));
        return -1;
    }
    static_assert(sizeof(h) == 64, "IPC handle size");
    memcpy(out, &h, sizeof(h));
    memset(static_cast<char*>(out) + 64, 0, 64);
    if (e.bh_arena) {
        err = cudaIpcGetMemHandle(&h, e.bh_arena);
        if (err != cudaSuccess) {
            set_error("cudaIpcGetMemHandle(bh arena): %s", cudaGetErrorString(err));
            return -1;
        }
        memcpy(static_cast<char*>(out) + 64, &h, sizeof(h));
    }
    return 0;
}

int dist_import(Engine& e, const void* all, int world) {
    if (!e.dist || world != e.world) {
        set_error("nbx_dist_import: world mismatch");
        return -1;
    }
    for (int g = 0; g < world; g++) {
        if (g == e.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(all) + 128 * g, 64);
        void* p = nullptr;
        cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess) {
            set_error("cudaIpcOpenMemHandle(rank %d): %s", g, cudaGetErrorString(err));
            (void)cudaGetLastError();
            return -1;
        }
        e.peer[g].base = static_cast<char*>(p);
        memcpy(&h, static_cast<const char*>(all) + 128 * g + 64, 64);
        p = nullptr;
        err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess) {
            set_error("cudaIpcOpenMemHandle(rank %d, bh arena): %s", g, cudaGetErrorString(err));
            (void)cudaGetLastError();
            return -1;
        }
        e.bh_peer[g] = static_cast<char*>(p);
    }
    e.peers_mapped = true;
    return 0;
}

bool dist_nccl_ready() {
    Nccl* n = nccl_load();
    return n && n->comm;
}

// In-place ncclAllGather of `nbufs` arrays on the library stream: rank r contributes bufs[k][r * count .. (r + 1) * count).
void dist_nccl_allgather_inplace(Engine& e, float* const* bufs, int nbufs, size_t count) {
    Nccl* n = nccl_load();
    if (!n || !n->comm) fatal("this path needs the NCCL communicator (nbx_dist_nccl_init was not called)", __FILE__, __LINE__);
    n->GroupStart();
    for (int k = 0; k < nbufs; k++) n->AllGather(bufs[k] + static_cast<size_t>(e.rank) * count, bufs[k], count, ncclFloat32, n->comm, e.stream);
    const ncclResult_t r = n->GroupEnd();
    if (r != 0) fatal(n->GetErrorString ? n->GetErrorString(r) : "ncclAllGather failed", __FILE__, __LINE__);
}

int dist_nccl_unique_id(void* out128) {
    Nccl* n = nccl_load();
    if (!n) return -1;
    ncclUniqueId id;
    ncclResult_t r = n->GetUniqueId(&id);
    if (r != 0) {
        set_error("ncclGetUniqueId failed: %d", r);
        return -1;
    }
    memcpy(out128, &id, 128);
    return 0;
}

int dist_nccl_init(Engine& e, const void* id128) {
    Nccl* n = nccl_load();
    if (!n) return -1;
    if (!e.dist) {
        set_error("nbx_dist_init first");
        return -1;
    }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = n->CommInitRank(&n->comm, e.world, id, e.rank);
    if (r != 0) {
        set_error("ncclCommInitRank failed: %s", n->GetErrorString ? n->GetErrorString(r) : "?");
        return -1;
    }
    return 0;
}

}  // namespace nb
