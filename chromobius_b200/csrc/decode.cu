// CUDA kernels (sm_100a) and the C ABI of the B200 decoder. See include/chromobius_b200.h for the boundary and
// decode_core.h for the per-shot algorithm. No PyTorch, no CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/chromobius_b200.h"
#include "decode_core.h"
#include "host_compile.h"

using namespace cb200;

namespace {

thread_local std::string g_error;

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
#define CUDA_CHECK(expr)                                                                                       \
    do {                                                                                                       \
        cudaError_t e_ = (expr);                                                                               \
        if (e_ != cudaSuccess)                                                                                 \
            throw CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(e_) + " (" __FILE__ ":" +  \
                            std::to_string(__LINE__) + ")");                                                   \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// Device groups: GROUP == 32 -> one warp per shot (4 shots per 128-thread block); otherwise one block of GROUP threads
// per shot. Warp-level primitives serve the ball growth (one warp per ball).
// ---------------------------------------------------------------------------------------------------------------
template <int GROUP>
struct DevGroup {
    // Stateless: everything derives from threadIdx / GROUP, so passing the group by reference through non-inlined
    // functions costs no local memory and n / lanes / nwarps are compile-time constants.
    __device__ __forceinline__ int tid() const { return GROUP == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x; }
    __device__ __forceinline__ constexpr int n() const { return GROUP; }
    __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31); }
    __device__ __forceinline__ constexpr int lanes() const { return 32; }
    __device__ __forceinline__ int warp() const { return GROUP == 32 ? 0 : (int)(threadIdx.x >> 5); }
    __device__ __forceinline__ constexpr int nwarps() const { return GROUP == 32 ? 1 : GROUP / 32; }
    __device__ __forceinline__ void sync() const {
        if (GROUP == 32) __syncwarp();
        else __syncthreads();
    }
    __device__ __forceinline__ void wsync() const { __syncwarp(); }
    __device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(0xFFFFFFFFu, p); }
    __device__ __forceinline__ uint32_t shfl32(uint32_t v, int src) const { return __shfl_sync(0xFFFFFFFFu, v, src); }
    __device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) const {
        return (uint64_t)__shfl_sync(0xFFFFFFFFu, (unsigned long long)v, src);
    }
    // 64-bit warp minimum from two 32-bit hardware reductions (REDUX): high words first, then the low words of the
    // lanes that tie on the high word.
    __device__ __forceinline__ uint64_t min64(uint64_t v) const {
#ifdef CB_NO_REDUX
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            uint64_t w = (uint64_t)__shfl_xor_sync(0xFFFFFFFFu, (unsigned long long)v, o);
            v = w < v ? w : v;
        }
        return v;
#endif
        const uint32_t hi = (uint32_t)(v >> 32);
        const uint32_t mh = __reduce_min_sync(0xFFFFFFFFu, hi);
        const uint32_t lo = hi == mh ? (uint32_t)v : 0xFFFFFFFFu;
        const uint32_t ml = __reduce_min_sync(0xFFFFFFFFu, lo);
        return ((uint64_t)mh << 32) | ml;
    }
};

constexpr int WARP_MODE_BLOCK = 128;  // threads per block when GROUP == 32

template <int GROUP>
__device__ __forceinline__ void group_ids(uint32_t &group, uint32_t &group_in_block) {
    if (GROUP == 32) {
        group_in_block = threadIdx.x / 32;
        group = blockIdx.x * (blockDim.x / 32) + group_in_block;
    } else {
        group_in_block = 0;
        group = blockIdx.x;
    }
}

// The Shot objects of a block: one per warp, in (static) shared memory -- see struct Shot.
template <int GROUP>
__device__ __forceinline__ Shot &block_shot() {
    constexpr int WARPS = (GROUP == 32 ? WARP_MODE_BLOCK : GROUP) / 32;
    __shared__ __align__(16) uint8_t shot_mem[WARPS * ((sizeof(Shot) + 15) / 16 * 16)];
    return *(Shot *)(shot_mem + (threadIdx.x >> 5) * ((sizeof(Shot) + 15) / 16 * 16));
}
// bind_shot for this warp's copy: lane 0 writes, the warp waits.
template <int GROUP>
__device__ __forceinline__ void bind_warp_shot(Shot &s, const DecodeParams &p, const ScratchConfig &cfg, uint8_t *arena,
                                               uint8_t *shared, const DevGroup<GROUP> &g) {
    __syncwarp();
    if (g.lane() == 0) bind_shot(s, p, cfg, arena, shared, g);
    __syncwarp();
}

extern __shared__ __align__(16) uint8_t cb_smem[];

// K0: establishes the persistent-clean invariants of every group's arena.
template <int GROUP>
__global__ void __launch_bounds__(GROUP == 32 ? WARP_MODE_BLOCK : GROUP) k_init_arena(DecodeParams p, uint32_t num_groups) {
    uint32_t group, gib;
    DevGroup<GROUP> g;
    group_ids<GROUP>(group, gib);
    if (group >= num_groups) return;
    Shot &s = block_shot<GROUP>();
    bind_warp_shot<GROUP>(s, p, p.cfg, p.arena + (uint64_t)group * p.arena_stride, cb_smem, g);
    init_group_arena(s, g);
}

// K1: bit-packed ingest statistics. One warp per shot: masked popcount of the row with coalesced loads.
// Emits per-shot detection-event counts, the batch maximum and the batch total.
__global__ void __launch_bounds__(256) k_count_events(const uint8_t *__restrict__ dets, uint64_t stride, uint32_t num_shots,
                                                      const uint8_t *__restrict__ mask, uint32_t num_det_bytes,
                                                      uint32_t *__restrict__ counts, uint32_t *__restrict__ max_count,
                                                      unsigned long long *__restrict__ total) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    uint32_t local_max = 0;
    unsigned long long local_total = 0;
    for (uint32_t shot = warp; shot < num_shots; shot += nwarps) {
        const uint8_t *row = dets + (uint64_t)shot * stride;
        uint32_t c = 0;
        // head bytes up to 16-byte alignment, 16-byte body, tail bytes
        uint32_t head = (uint32_t)((16 - ((uintptr_t)row & 15)) & 15);
        if (head > num_det_bytes) head = num_det_bytes;
        uint32_t body = (num_det_bytes - head) / 16;
        for (uint32_t k = lane; k < head; k += 32) c += __popc((uint32_t)(row[k] & mask[k]));
        const uint4 *row16 = (const uint4 *)(row + head);
        for (uint32_t k = lane; k < body; k += 32) {
            uint4 v = row16[k];
            const uint8_t *m = mask + head + 16 * k;  // mask is small and L1/L2 resident; unaligned -> byte loads
            uint32_t m0 = m[0] | (m[1] << 8) | (m[2] << 16) | ((uint32_t)m[3] << 24);
            uint32_t m1 = m[4] | (m[5] << 8) | (m[6] << 16) | ((uint32_t)m[7] << 24);
            uint32_t m2 = m[8] | (m[9] << 8) | (m[10] << 16) | ((uint32_t)m[11] << 24);
            uint32_t m3 = m[12] | (m[13] << 8) | (m[14] << 16) | ((uint32_t)m[15] << 24);
            c += __popc(v.x & m0) + __popc(v.y & m1) + __popc(v.z & m2) + __popc(v.w & m3);
        }
        for (uint32_t k = head + 16 * body + lane; k < num_det_bytes; k += 32) c += __popc((uint32_t)(row[k] & mask[k]));
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
        if (lane == 0) {
            counts[shot] = c;
            local_max = max(local_max, c);
            local_total += c;
        }
    }
    if (lane == 0 && local_total) {
        atomicMax(max_count, local_max);
        atomicAdd(total, local_total);
    }
}

// K2: the decode kernel. Persistent groups pull shots from a queue.
#ifndef CB_MIN_BLOCKS_256
#define CB_MIN_BLOCKS_256 3
#endif
// PROF: launches of a call with phase profiling enabled keep the cycle counters (PhaseTimer); all others compile them out.
template <int GROUP, bool PROF = false>
__global__ void __launch_bounds__(GROUP == 32 ? WARP_MODE_BLOCK : GROUP, GROUP == 512 ? 2 : (GROUP == 256 ? CB_MIN_BLOCKS_256 : (GROUP == 192 ? 4 : 6)))
    k_decode(DecodeParams p, uint32_t num_groups, uint32_t smem_per_group) {
    uint32_t group, gib;
    DevGroup<GROUP> g;
    typedef typename std::conditional<PROF, PhaseTimer, NullTimer>::type Timer;
    group_ids<GROUP>(group, gib);
    if (group >= num_groups) return;
    Shot &s = block_shot<GROUP>();
    bind_warp_shot<GROUP>(s, p, p.cfg, p.arena + (uint64_t)group * p.arena_stride, cb_smem + (size_t)gib * smem_per_group, g);
    init_group_shared(s, g);
    const uint32_t total = p.shot_list ? p.list_len : p.num_shots;
#ifdef CB_TAIL_PROF
    // diagnostic build: slot 21 = first group's start (globaltimer ns), 20 = sum of group finish offsets, 22 = last finish
    unsigned long long tail_t0 = 0;
    if (p.phase_cycles && g.tid() == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        atomicCAS(&p.phase_cycles[21], 0ull, now);
    }
#endif
    while (true) {
        if (g.tid() == 0) s.hdr->bcast = atomicAdd(p.work_counter, 1u);
        g.sync();
        uint32_t k = s.hdr->bcast;
        g.sync();
        if (k >= total) break;
        uint32_t shot = p.shot_list ? p.shot_list[k] : k;
        if (p.ready) {
            // streaming input: wait until the copy stream has delivered this shot's row
            if (g.tid() == 0) {
                while (*(const volatile uint32_t *)p.ready <= shot) __nanosleep(500);
                __threadfence();
            }
            g.sync();
        }
        decode_shot<DevGroup<GROUP>, Timer>(s, p, shot, g);
        if (p.event_total && g.tid() == 0) {
            const uint32_t ev = s.word_rank[s.nwords];  // detection events of the shot (set by ingest_shot, also on overflow)
            atomicAdd(p.event_total, (unsigned long long)ev);
            atomicMax(p.event_max, ev);
        }
        // Out of scratch: again, in a spare arena with the next level's scratch, if one is free (decode_shot ended with a
        // barrier after thread 0 stored the status). Otherwise the host's retry launch takes the shot.
        if (p.spare_arena && p.status_out[shot] == SHOT_SCRATCH_OVERFLOW) {
            if (g.tid() == 0) {
                uint32_t slot = NO_INDEX;
                for (uint32_t c = 0; c < p.num_spares && slot == NO_INDEX; c++)
                    if (atomicCAS(&p.spare_claim[c], 0u, 1u) == 0u) slot = c;
                s.hdr->bcast = slot;
            }
            g.sync();
            const uint32_t slot = s.hdr->bcast;
            g.sync();
            if (slot != NO_INDEX) {
                uint8_t *const shared = cb_smem + (size_t)gib * smem_per_group;
                bind_warp_shot<GROUP>(s, p, p.spare_cfg, p.spare_arena + (uint64_t)slot * p.spare_stride, shared, g);
                g.sync();
                decode_shot<DevGroup<GROUP>, Timer>(s, p, shot, g);
                if (g.tid() == 0) {
                    // still too small: tell the host that level 1 has been tried, so its retry starts at level 2
                    if (p.status_out[shot] == SHOT_SCRATCH_OVERFLOW) p.status_out[shot] = SHOT_SCRATCH_OVERFLOW_SPARE;
                    __threadfence();
                    atomicExch(&p.spare_claim[slot], 0u);
                    atomicAdd(p.spare_retries, 1u);
                }
                g.sync();  // thread 0 read the status through this binding
                bind_warp_shot<GROUP>(s, p, p.cfg, p.arena + (uint64_t)group * p.arena_stride, shared, g);
                g.sync();
            }
        }
    }
#ifdef CB_TAIL_PROF
    if (p.phase_cycles && g.tid() == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        tail_t0 = *(volatile unsigned long long *)&p.phase_cycles[21];
        atomicAdd(&p.phase_cycles[20], now - tail_t0);
        atomicMax(&p.phase_cycles[22], now - tail_t0);
    }
#endif
}

// K3: u64 masks -> little-endian observable bytes, integer weights -> float32 (decode/pymatcher.cc:40).
__global__ void k_finish(const uint64_t *__restrict__ obs64, const int64_t *__restrict__ wint, uint32_t num_shots,
                         uint32_t obs_bytes, double inv_norm, uint8_t *__restrict__ obs_out, float *__restrict__ w_out,
                         int64_t *__restrict__ wint_out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_shots) return;
    uint64_t m = obs64[i];
    for (uint32_t b = 0; b < obs_bytes; b++) obs_out[(uint64_t)i * obs_bytes + b] = (uint8_t)(m >> (8 * b));
    if (w_out) w_out[i] = (float)((double)wint[i] * inv_norm);
    if (wint_out) wint_out[i] = wint[i];
}

template <class T>
struct DeviceBuf {
    T *ptr = nullptr;
    size_t cap = 0;
    void ensure(size_t n) {
        if (n <= cap) return;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        CUDA_CHECK(cudaMalloc(&ptr, std::max<size_t>(n, 1) * sizeof(T)));
        cap = n;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

}  // namespace

struct cb200_decoder {
    std::shared_ptr<const HostTables> host_ptr;  // shared by the per-device replicas of one cb200_multi
    const HostTables &host_ref() const { return *host_ptr; }
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // streaming input: chunks of a large host batch (decode_impl)
    std::vector<uint32_t> ready_values;      // host source of the "shots resident" words the copy stream publishes
    cudaEvent_t ev[6] = {};
    // device tables (one allocation, see upload_tables)
    DeviceBuf<uint8_t> d_tables;
    size_t l2_window_bytes = 0;
    size_t l2_persist_want = 0;          // this decoder's share of the device's persisting-L2 set-aside (0: none)
    uint32_t tune_groups_per_sm = 0;
    TablesView dview{};
    uint64_t table_bytes = 0;
    // per-call buffers (grow only)
    DeviceBuf<uint8_t> d_dets, d_obs_bytes;
    struct ArenaState {
        DeviceBuf<uint8_t> buf;
        ScratchConfig cfg{};
        uint32_t group = 0, groups = 0;
        uint64_t stride = 0;
    } arenas[4];  // one cached arena per escalation level, so retry passes do not evict the main layout
    DeviceBuf<uint64_t> d_obs64;
    DeviceBuf<int64_t> d_wint, d_wint_out, d_edges;
    DeviceBuf<uint32_t> d_status, d_rounds, d_counts, d_scalars, d_list, d_spare_claim;
    DeviceBuf<unsigned long long> d_phase;
    bool profile_phases = false;
    unsigned long long phase_cycles[24] = {};
    DeviceBuf<float> d_wfloat;
    // arena layout currently initialised on the device
    // tuning
    uint32_t tune_group = 0, tune_ht_log2 = 0, tune_pool_bytes = 32u << 10;
};

namespace {

// The persisting-L2 set-aside is a per-DEVICE limit, while several decoders can live on one GPU (the sinter service
// caches one per model; a multi-GPU handle may list a device twice): it is kept at the largest request among the live
// decoders of that device and handed back (limit 0, persisting lines reset) when the last of them is freed.
struct L2SetAside {
    std::mutex mu;
    std::map<int, std::multiset<size_t>> wants;  // device -> requests of the live decoders
};
L2SetAside &l2_registry() {
    static L2SetAside r;
    return r;
}
bool l2_set_aside_register(int device, size_t want) {
    L2SetAside &r = l2_registry();
    std::lock_guard<std::mutex> lock(r.mu);
    auto &set = r.wants[device];
    set.insert(want);
    return cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, *set.rbegin()) == cudaSuccess;
}
void l2_set_aside_release(int device, size_t want) {
    L2SetAside &r = l2_registry();
    std::lock_guard<std::mutex> lock(r.mu);
    auto it = r.wants.find(device);
    if (it == r.wants.end()) return;
    auto pos = it->second.find(want);
    if (pos != it->second.end()) it->second.erase(pos);
    if (it->second.empty()) {
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
        r.wants.erase(it);
    } else {
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, *it->second.rbegin());
    }
    cudaGetLastError();
}

// All read-only tables live in ONE allocation so that a single L2 access-policy window can pin them: the matcher
// graph is re-read by every ball of every shot, while the per-shot scratch streams through L2 behind it.
void upload_tables(cb200_decoder *d) {
    const HostTables &h = d->host_ref();
    size_t total = 0;
    auto reserve = [&](size_t bytes) {
        size_t off = total;
        total += (bytes + 255) & ~size_t{255};
        return off;
    };
    // hot first: the window covers a prefix if the set-aside is smaller than the tables
    size_t o_near = reserve(h.near.size() * sizeof(NearEntry));
    size_t o_ell = reserve(h.ell_adj.size() * sizeof(AdjEntry));
    size_t o_row = reserve(h.row_ptr.size() * 4);
    size_t o_adj = reserve(h.adj.size() * sizeof(AdjEntry));
    size_t o_mask = reserve(h.ingest_mask.size());
    size_t o_info = reserve(h.det_info.size());
    size_t o_rep = reserve(h.rep_obs.size() * 8);
    size_t o_self = reserve(h.self_drag.size() * 4);
    size_t o_adrag = reserve(h.adj_drag.size() * 4);
    size_t o_drag = reserve(h.drag.size() * sizeof(DragEntry));
    d->d_tables.ensure(total + 256);
    uint8_t *base = d->d_tables.ptr;
    auto put = [&](size_t off, const void *src, size_t bytes) {
        if (bytes) CUDA_CHECK(cudaMemcpy(base + off, src, bytes, cudaMemcpyHostToDevice));
    };
    put(o_near, h.near.data(), h.near.size() * sizeof(NearEntry));
    put(o_ell, h.ell_adj.data(), h.ell_adj.size() * sizeof(AdjEntry));
    put(o_row, h.row_ptr.data(), h.row_ptr.size() * 4);
    put(o_adj, h.adj.data(), h.adj.size() * sizeof(AdjEntry));
    put(o_mask, h.ingest_mask.data(), h.ingest_mask.size());
    put(o_info, h.det_info.data(), h.det_info.size());
    put(o_rep, h.rep_obs.data(), h.rep_obs.size() * 8);
    put(o_self, h.self_drag.data(), h.self_drag.size() * 4);
    put(o_adrag, h.adj_drag.data(), h.adj_drag.size() * 4);
    put(o_drag, h.drag.data(), h.drag.size() * sizeof(DragEntry));
    d->table_bytes = total;
    TablesView v = h.view();
    v.near = h.near_stride ? (const NearEntry *)(base + o_near) : nullptr;
    v.ell_adj = (const AdjEntry *)(base + o_ell);
    v.row_ptr = (const uint32_t *)(base + o_row);
    v.adj = (const AdjEntry *)(base + o_adj);
    v.ingest_mask = base + o_mask;
    v.det_info = base + o_info;
    v.rep_obs = (const uint64_t *)(base + o_rep);
    v.self_drag = (const uint32_t *)(base + o_self);
    v.adj_drag = (const uint32_t *)(base + o_adrag);
    v.drag = (const DragEntry *)(base + o_drag);
    d->dview = v;
    // L2 persistence for the tables (best effort: older drivers / MIG slices may refuse).
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, d->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, d->device);
    // Only the head of the allocation is worth pinning: the first chunks of the near-table rows (the layout is chunk-major
    // and the near table comes first). The per-shot scratch of the shots in flight needs the rest of the L2.
    size_t hot = total;
    if (h.near_stride) hot = std::min<size_t>(total, (size_t)h.num_dets * 2 * 32 * sizeof(NearEntry) * 2);
    if (const char *env = getenv("CB200_L2_WINDOW_MB")) hot = std::min<size_t>(total, (size_t)strtoul(env, nullptr, 10) << 20);
    if (max_persist > 0 && max_window > 0 && hot > 0) {
        total = hot;
        size_t want = std::min<size_t>(total, (size_t)max_persist);
        d->l2_persist_want = want;
        if (l2_set_aside_register(d->device, want)) {
            cudaStreamAttrValue attr{};
            attr.accessPolicyWindow.base_ptr = base;
            attr.accessPolicyWindow.num_bytes = std::min<size_t>(total, (size_t)max_window);
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)attr.accessPolicyWindow.num_bytes);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(d->stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess)
                d->l2_window_bytes = attr.accessPolicyWindow.num_bytes;
        }
        cudaGetLastError();
    }
}

template <int GROUP>
void launch_init(const DecodeParams &p, uint32_t groups, cudaStream_t st) {
    uint32_t threads = GROUP == 32 ? WARP_MODE_BLOCK : GROUP;
    uint32_t per_block = GROUP == 32 ? WARP_MODE_BLOCK / 32 : 1;
    uint32_t blocks = (groups + per_block - 1) / per_block;
    k_init_arena<GROUP><<<blocks, threads, 0, st>>>(p, groups);
}
template <int GROUP>
void launch_decode(const DecodeParams &p, uint32_t groups, uint32_t smem_per_group, cudaStream_t st) {
    uint32_t threads = GROUP == 32 ? WARP_MODE_BLOCK : GROUP;
    uint32_t per_block = GROUP == 32 ? WARP_MODE_BLOCK / 32 : 1;
    uint32_t blocks = (groups + per_block - 1) / per_block;
    size_t smem = (size_t)smem_per_group * per_block;
    CUDA_CHECK(cudaFuncSetAttribute(k_decode<GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // Ask for the smallest shared-memory carve-out that still holds the resident blocks (occupancy is register limited):
    // the rest of the unified storage becomes L1, where the per-event scratch of the shots in flight lives.
    {
        int dev = 0, sm_bytes = 0, resident = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_bytes, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        // occupancy with the largest carve-out (the attribute of an earlier launch would otherwise cap the answer)
        cudaFuncSetAttribute(k_decode<GROUP>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_decode<GROUP>, (int)threads, smem);
        cudaFuncAttributes fa{};
        cudaFuncGetAttributes(&fa, k_decode<GROUP>);  // static shared memory: the warps' Shot objects
        int pct = -1;
        if (sm_bytes > 0 && resident > 0)
            pct = (int)(((size_t)resident * (smem + fa.sharedSizeBytes + 1024) * 100 + sm_bytes - 1) / sm_bytes);
        if (const char *env = getenv("CB200_CARVEOUT_PCT")) pct = atoi(env);  // experiments
        if (pct >= 0) cudaFuncSetAttribute(k_decode<GROUP>, cudaFuncAttributePreferredSharedMemoryCarveout, std::min(pct, 100));
        cudaGetLastError();
    }
    if (p.phase_cycles) {
        CUDA_CHECK(cudaFuncSetAttribute(k_decode<GROUP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_decode<GROUP, true><<<blocks, threads, smem, st>>>(p, groups, smem_per_group);
    } else {
        k_decode<GROUP><<<blocks, threads, smem, st>>>(p, groups, smem_per_group);
    }
}
template <int GROUP>
int blocks_per_sm(size_t smem_per_group) {
    uint32_t threads = GROUP == 32 ? WARP_MODE_BLOCK : GROUP;
    uint32_t per_block = GROUP == 32 ? WARP_MODE_BLOCK / 32 : 1;
    size_t smem = smem_per_group * per_block;
    int n = 0;
    cudaFuncSetAttribute(k_decode<GROUP>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);  // see launch_decode
    if (cudaFuncSetAttribute(k_decode<GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_decode<GROUP>, (int)threads, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *status_message(uint32_t st) {
    switch (st) {
        case SHOT_SCRATCH_OVERFLOW: return "decoder scratch memory exhausted even after enlarging it";
        case SHOT_NO_PERFECT_MATCHING: return "No perfect matching exists for the detection events.";
        case SHOT_EULER_FAILED: return "Graph didn't decompose into Euler tours.";
        case SHOT_DISCHARGE_FAILED:
            return "Failed to lift a flattened edge cycle from the matcher into an explanation of the detection events in "
                   "the cycle.\nThis error could be due to a coloring error in the model used to configure the decoder, or a "
                   "bug in the decoder.";
        case SHOT_ROUND_LIMIT: return "certified matching did not converge within the round limit";
        case SHOT_COMPONENT_TOO_LARGE: return "a matching component exceeded the supported size";
        default: return "unknown failure";
    }
}

// Chooses the group width and scratch configuration for a batch and (re)initialises the arena when the layout changes.
struct Plan {
    uint32_t group;
    uint32_t groups;
    uint32_t smem_per_group;
    ScratchConfig cfg;
};

Plan make_plan(cb200_decoder *d, uint64_t work_items, uint32_t max_det_events, double mean_det_events, uint32_t escalation) {
    Plan pl{};
    uint32_t group = d->tune_group;
    if (group == 0) group = mean_det_events >= 200 ? 256 : (mean_det_events >= 8 ? 128 : 32);
    if (group != 32 && group != 128 && group != 192 && group != 256 && group != 512) group = 32;
    pl.group = group;
    uint32_t full = 4;
    while ((1u << full) < 2 * std::max<uint32_t>(d->dview.num_nodes, 8)) full++;
    // With a near table the Dijkstra search is a rare fallback and runs in the global tables; shared memory then only
    // holds the shot's row and the blossom scratch pool, which leaves most of the SM's unified storage to the L1 cache
    // (the per-event arrays of the shots in flight live there). Without a near table every ball is a search: shared tables.
    const bool near = d->host_ref().near_stride != 0;
    uint32_t sh = d->tune_ht_log2 ? d->tune_ht_log2 : (near ? 0 : 8);
    uint32_t gl = std::min<uint32_t>(full, 14 + 3 * escalation);
    pl.cfg.sh_ht_log2 = std::min(sh, gl);
    // Blossom scratch pool: 32 KB where shots hold hundreds of events (a round then often has one 32-vertex and several
    // small components side by side), 20 KB -- exactly one 32-vertex layout -- for small syndromes, whose components are
    // small and which gain more from the extra L1 (measured: d9 r36 92.7 -> 85.2 ms, superdense d5 69.5 -> 63.9 ms per
    // 300 k shots; d25 r100 is 3 % slower with 24 KB).
    pl.cfg.sh_pool = near ? (mean_det_events >= 200 ? d->tune_pool_bytes : std::min<uint32_t>(d->tune_pool_bytes, 20u << 10)) : 0;
    if (const char *env = getenv("CB200_SH_EVENTS")) pl.cfg.sh_events = (uint32_t)strtoul(env, nullptr, 10);  // experiments
    if (const char *env = getenv("CB200_POOL_KB")) pl.cfg.sh_pool = near ? (uint32_t)strtoul(env, nullptr, 10) << 10 : 0;  // experiments
    pl.cfg.sh_found = 16;
    if (getenv("CB200_NO_PLAIN_LIFT")) pl.cfg.pad |= 1u;  // experiments / tests: always the general lift
    pl.cfg.ht_log2 = gl;
    pl.cfg.found_cap = 256u << (2 * escalation);
    // Event capacity in coarse steps, and sticky: a batch that fits the layout the arena already has keeps it, so a
    // stream of batches with slightly different maxima does not re-allocate and re-initialise the arena per call.
    pl.cfg.max_events = (std::max<uint32_t>(2 * max_det_events, 2) + 127u) & ~127u;
    {
        const cb200_decoder::ArenaState &prev = d->arenas[std::min<uint32_t>(escalation, 3)];
        if (prev.group == group && prev.cfg.max_events >= pl.cfg.max_events && prev.cfg.max_events / 2 <= pl.cfg.max_events)
            pl.cfg.max_events = prev.cfg.max_events;
    }
    pl.cfg.dyn_bytes = ((uint64_t)(64 << 10) + (uint64_t)3072 * pl.cfg.max_events) << (3 * escalation);
    // tests: a first-level scratch this many times too small, so that the retry paths (spare arenas, retry launch) run
    if (escalation == 0)
        if (const char *env = getenv("CB200_TEST_SCRATCH_DIV"))
            pl.cfg.dyn_bytes = std::max<uint64_t>(4096, pl.cfg.dyn_bytes / std::max<uint64_t>(1, strtoull(env, nullptr, 10)));
    uint32_t nwarps = group == 32 ? 1 : group / 32;
    pl.smem_per_group = (uint32_t)shared_bytes_per_group(pl.cfg, d->dview.num_dets, nwarps);
    auto occupancy = [&](uint32_t smem) {
        switch (group) {
            case 32: return blocks_per_sm<32>(smem) * (WARP_MODE_BLOCK / 32);
            case 128: return blocks_per_sm<128>(smem);
            case 192: return blocks_per_sm<192>(smem);
            case 256: return blocks_per_sm<256>(smem);
            default: return blocks_per_sm<512>(smem);
        }
    };
    int per_sm = occupancy(pl.smem_per_group);
    if (per_sm <= 0) {
        // the row does not fit beside the ball tables: drop the shared tables, keep the row
        pl.cfg.sh_ht_log2 = 0;
        pl.smem_per_group = (uint32_t)shared_bytes_per_group(pl.cfg, d->dview.num_dets, nwarps);
        per_sm = occupancy(pl.smem_per_group);
        if (per_sm <= 0) throw std::invalid_argument("detector error model too large: one shot's row does not fit in shared memory");
    }
    if (d->tune_groups_per_sm && (int)d->tune_groups_per_sm < per_sm) per_sm = (int)d->tune_groups_per_sm;
    uint64_t want = (uint64_t)d->num_sms * (uint32_t)per_sm;
    pl.groups = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(want, work_items));
    return pl;
}

void ensure_arena(cb200_decoder *d, const Plan &pl, DecodeParams &p, uint64_t &launches, uint32_t level) {
    cb200_decoder::ArenaState &a = d->arenas[level];
    uint32_t nwarps = pl.group == 32 ? 1 : pl.group / 32;
    uint64_t stride = arena_bytes_per_group(pl.cfg, d->dview.num_nodes, nwarps);
    bool same = a.group == pl.group && a.stride == stride && a.groups >= pl.groups && memcmp(&a.cfg, &pl.cfg, sizeof(ScratchConfig)) == 0;
    p.cfg = pl.cfg;
    p.arena_stride = stride;
    if (!same) {
        a.buf.ensure(stride * pl.groups);
        p.arena = a.buf.ptr;
        if (pl.group == 32) launch_init<32>(p, pl.groups, d->stream);
        else if (pl.group == 128) launch_init<128>(p, pl.groups, d->stream);
        else if (pl.group == 192) launch_init<192>(p, pl.groups, d->stream);
        else if (pl.group == 256) launch_init<256>(p, pl.groups, d->stream);
        else launch_init<512>(p, pl.groups, d->stream);
        CUDA_CHECK(cudaGetLastError());
        launches++;
        a.group = pl.group;
        a.stride = stride;
        a.groups = pl.groups;
        a.cfg = pl.cfg;
    }
    p.arena = a.buf.ptr;
}

int decode_impl(cb200_decoder *d, const uint8_t *dets, uint64_t num_shots, uint64_t row_stride, uint8_t *obs_out,
                float *weight_out, int64_t *weight_int_out, uint32_t flags, cb200_batch_stats *stats, int64_t *edges_out,
                uint64_t edges_cap, uint64_t *num_edge_ints, uint64_t *obs_mask_out) {
    cb200_batch_stats st{};
    st.first_failed_shot = -1;
    st.num_shots = num_shots;
    if (stats) *stats = st;
    if (num_shots == 0) return CB200_OK;
    if (num_shots > 0x7FFFFFF0ull) throw std::invalid_argument("too many shots in one batch (limit 2^31 - 16)");
    const uint32_t nb = d->dview.num_det_bytes;
    const uint32_t ob = (d->dview.num_obs + 7) / 8;
    if (row_stride < nb) throw std::invalid_argument("row_stride is smaller than the number of detector bytes");
    CUDA_CHECK(cudaSetDevice(d->device));
    const bool dev_io = (flags & CB200_IO_DEVICE) != 0;
    const uint32_t n = (uint32_t)num_shots;
    cudaStream_t s = d->stream;
    uint64_t launches = 0;

    CUDA_CHECK(cudaEventRecord(d->ev[0], s));
    const uint8_t *d_dets = dets;
    uint64_t d_stride = row_stride;
    d->d_obs64.ensure(n);
    d->d_wint.ensure(n);
    d->d_status.ensure(n);
    d->d_rounds.ensure(n);
    d->d_counts.ensure(n);
    d->d_scalars.ensure(16);
    CUDA_CHECK(cudaMemsetAsync(d->d_scalars.ptr, 0, 16 * sizeof(uint32_t), s));
    // scalars: [0] max count, [2..3] total (u64) of k_count_events, [4] work counter, [5] spare retries,
    //          [6] ready shots (streaming input), [8..9] total (u64) and [10] max gathered by the decode kernel
    // Streaming input: a large host batch is copied in chunks on a second stream while the persistent decode kernel
    // already works on the chunks that have arrived. The arena is sized from the first chunk (plus head room; a shot
    // beyond it takes the scratch-overflow retry path), so the kernel can start after one chunk instead of after the
    // whole copy and a count pass over all of it.
    // Not when launches are serialised with their stream's other work (CUDA_LAUNCH_BLOCKING=1, a profiler or sanitizer
    // injected into the process): the kernel would wait for chunks whose copies cannot start before it ends.
    auto launches_serialised = [] {
        const char *b = getenv("CUDA_LAUNCH_BLOCKING");
        if (b && b[0] == '1') return true;
        for (const char *name : {"CUDA_INJECTION64_PATH", "NV_COMPUTE_PROFILER_PERFWORKS_DIR", "NV_NSIGHT_INJECTION_TRANSPORT_TYPE",
                                 "NV_SANITIZER_INJECTION_TRANSPORT_TYPE"})
            if (getenv(name)) return true;
        return false;
    };
    uint64_t chunk = 0;
    if (!dev_io && nb && !getenv("CB200_NO_PIPELINE") && !launches_serialised()) {
        chunk = std::max<uint64_t>(8192, ((uint64_t)96 << 20) / nb);
        if (const char *env = getenv("CB200_PIPELINE_CHUNK")) chunk = std::max<uint64_t>(1, strtoull(env, nullptr, 10));
        if ((uint64_t)n < 3 * chunk) chunk = 0;
    }
    const bool streaming = chunk != 0;
    const uint32_t counted = streaming ? (uint32_t)chunk : n;  // shots that k_count_events looks at
    if (!dev_io) {
        if (nb) {
            d->d_dets.ensure((size_t)n * nb);
            CUDA_CHECK(cudaMemcpy2DAsync(d->d_dets.ptr, nb, dets, row_stride, nb, counted, cudaMemcpyHostToDevice, s));
        }
        d_dets = d->d_dets.ptr;
        d_stride = nb;
    }
    CUDA_CHECK(cudaEventRecord(d->ev[1], s));
    {
        uint32_t blocks = std::min<uint32_t>((counted + 7) / 8, (uint32_t)d->num_sms * 8);
        k_count_events<<<blocks, 256, 0, s>>>(d_dets, d_stride, counted, d->dview.ingest_mask, nb, d->d_counts.ptr, d->d_scalars.ptr,
                                              (unsigned long long *)(d->d_scalars.ptr + 2));
        CUDA_CHECK(cudaGetLastError());
        launches++;
    }
    uint32_t scal[16];
    CUDA_CHECK(cudaMemcpyAsync(scal, d->d_scalars.ptr, sizeof(scal), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaEventRecord(d->ev[2], s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    uint32_t max_count = scal[0];
    uint64_t total_events;
    memcpy(&total_events, &scal[2], 8);
    const double mean_events = (double)total_events / (double)counted;
    if (streaming) max_count += std::max<uint32_t>(32, max_count / 8);  // head room for the chunks not seen yet
    st.num_detection_events = total_events;
    st.max_events_per_shot = 2 * max_count;

    DecodeParams p{};
    p.T = d->dview;
    p.dets = d_dets;
    p.stride = d_stride;
    p.num_shots = n;
    p.obs_out = d->d_obs64.ptr;
    p.weight_out = d->d_wint.ptr;
    p.status_out = d->d_status.ptr;
    p.rounds_out = d->d_rounds.ptr;
    p.work_counter = d->d_scalars.ptr + 4;
    p.event_total = (unsigned long long *)(d->d_scalars.ptr + 8);
    p.event_max = d->d_scalars.ptr + 10;
    if (d->profile_phases) {
        d->d_phase.ensure(24);
        CUDA_CHECK(cudaMemsetAsync(d->d_phase.ptr, 0, 192, s));
        p.phase_cycles = d->d_phase.ptr;
    }
    if (edges_out || num_edge_ints) {
        d->d_edges.ensure(edges_cap + 2);
        p.edge_dump = d->d_edges.ptr;
        p.edge_dump_cap = edges_cap;
        p.edge_dump_len = (uint64_t *)(d->d_edges.ptr + edges_cap);
    }

    std::vector<uint32_t> status(n);
    std::vector<uint32_t> retry, deferred;  // deferred: already failed in a level-1 spare arena, joins the list at level 2
    for (uint32_t esc = 0; esc < 4; esc++) {
        if (esc == 1 && retry.empty()) {
            retry.swap(deferred);
            continue;
        }
        uint64_t items = esc == 0 ? n : retry.size();
        Plan pl = make_plan(d, items, max_count, mean_events, esc);
        ensure_arena(d, pl, p, launches, esc);
        p.spare_arena = nullptr;
        p.num_spares = 0;
        if (esc == 0) {
            st.group_threads = pl.group;
            st.num_groups = pl.groups;
            p.shot_list = nullptr;
            p.list_len = 0;
            // Spare arenas of the next escalation level for in-kernel retries (DecodeParams::spare_arena). They need the
            // main plan's shared-memory layout; only worth it when the groups work through a queue of shots.
            constexpr uint32_t NUM_SPARES = 16;
            const bool no_spares = getenv("CB200_NO_SPARES") != nullptr;  // experiments / tests
            if (!no_spares && n > pl.groups) {
                Plan sp = make_plan(d, NUM_SPARES, streaming ? 2 * max_count : max_count, mean_events, 1);
                if (sp.group == pl.group && sp.smem_per_group == pl.smem_per_group && sp.cfg.sh_ht_log2 == pl.cfg.sh_ht_log2 &&
                    sp.cfg.sh_pool == pl.cfg.sh_pool && sp.cfg.sh_events == pl.cfg.sh_events && sp.cfg.sh_found == pl.cfg.sh_found &&
                    sp.cfg.max_events >= pl.cfg.max_events) {
                    DecodeParams q = p;
                    ensure_arena(d, sp, q, launches, 1);
                    d->d_spare_claim.ensure(NUM_SPARES);
                    CUDA_CHECK(cudaMemsetAsync(d->d_spare_claim.ptr, 0, NUM_SPARES * 4, s));
                    p.spare_arena = q.arena;
                    p.spare_stride = q.arena_stride;
                    p.spare_cfg = q.cfg;
                    p.num_spares = std::min<uint32_t>(NUM_SPARES, sp.groups);
                    p.spare_claim = d->d_spare_claim.ptr;
                    p.spare_retries = d->d_scalars.ptr + 5;
                }
            }
        } else {
            d->d_list.ensure(retry.size());
            CUDA_CHECK(cudaMemcpyAsync(d->d_list.ptr, retry.data(), retry.size() * 4, cudaMemcpyHostToDevice, s));
            p.shot_list = d->d_list.ptr;
            p.list_len = (uint32_t)retry.size();
            CUDA_CHECK(cudaMemsetAsync(d->d_scalars.ptr + 4, 0, 4, s));
        }
        p.ready = nullptr;
        if (esc == 0 && streaming) {
            // the first chunk is resident; the kernel starts now and the copy stream feeds it
            d->ready_values.assign(((uint64_t)n + chunk - 1) / chunk + 1, 0);
            d->ready_values[0] = counted;
            CUDA_CHECK(cudaMemcpyAsync(d->d_scalars.ptr + 6, &d->ready_values[0], 4, cudaMemcpyHostToDevice, s));
            p.ready = d->d_scalars.ptr + 6;
        }
        if (esc != 0) {
            p.event_total = nullptr;
            p.event_max = nullptr;
        }
        if (pl.group == 32) launch_decode<32>(p, pl.groups, pl.smem_per_group, s);
        else if (pl.group == 128) launch_decode<128>(p, pl.groups, pl.smem_per_group, s);
        else if (pl.group == 192) launch_decode<192>(p, pl.groups, pl.smem_per_group, s);
        else if (pl.group == 256) launch_decode<256>(p, pl.groups, pl.smem_per_group, s);
        else launch_decode<512>(p, pl.groups, pl.smem_per_group, s);
        CUDA_CHECK(cudaGetLastError());
        launches++;
        if (esc == 0) CUDA_CHECK(cudaEventRecord(d->ev[3], s));
        if (esc == 0 && streaming) {
            // remaining chunks on the copy stream; after each, the number of resident shots is published
            if (!d->copy_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
            uint64_t k = 1;
            for (uint64_t first = counted; first < n; first += chunk, k++) {
                const uint64_t rows = std::min<uint64_t>(chunk, n - first);
                CUDA_CHECK(cudaMemcpy2DAsync(d->d_dets.ptr + first * nb, nb, dets + first * row_stride, row_stride, nb, rows,
                                             cudaMemcpyHostToDevice, d->copy_stream));
                d->ready_values[k] = (uint32_t)(first + rows);
                CUDA_CHECK(cudaMemcpyAsync(d->d_scalars.ptr + 6, &d->ready_values[k], 4, cudaMemcpyHostToDevice, d->copy_stream));
            }
            CUDA_CHECK(cudaStreamSynchronize(d->copy_stream));
        }
        CUDA_CHECK(cudaMemcpyAsync(status.data(), d->d_status.ptr, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        uint32_t spare_retries = 0;
        if (p.spare_arena) CUDA_CHECK(cudaMemcpyAsync(&spare_retries, d->d_scalars.ptr + 5, 4, cudaMemcpyDeviceToHost, s));
        uint32_t gathered[4] = {0, 0, 0, 0};
        if (esc == 0) CUDA_CHECK(cudaMemcpyAsync(gathered, d->d_scalars.ptr + 8, sizeof(gathered), cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        if (esc == 0 && streaming) {
            // what the first chunk could only estimate: exact totals from the kernel, used by the statistics and by the
            // retry levels' plans
            memcpy(&total_events, &gathered[0], 8);
            max_count = std::max(max_count, gathered[2]);
            st.num_detection_events = total_events;
            st.max_events_per_shot = 2 * gathered[2];
        }
        st.num_retried_shots += spare_retries;
        retry.clear();
        for (uint32_t k = 0; k < n; k++) {
            if (status[k] == SHOT_SCRATCH_OVERFLOW) retry.push_back(k);
            else if (status[k] == SHOT_SCRATCH_OVERFLOW_SPARE) deferred.push_back(k);
        }
        if (esc >= 1) {
            retry.insert(retry.end(), deferred.begin(), deferred.end());
            deferred.clear();
        }
        if (retry.empty() && deferred.empty()) break;
        st.num_retried_shots += retry.size() + (esc == 0 ? deferred.size() : 0);
    }
    for (uint32_t k = 0; k < n; k++)
        if (status[k] == SHOT_SCRATCH_OVERFLOW_SPARE) status[k] = SHOT_SCRATCH_OVERFLOW;  // one public code for "out of scratch"
    for (uint32_t k = 0; k < n; k++)
        if (status[k]) {
            if (st.first_failed_shot < 0) {
                st.first_failed_shot = k;
                st.first_failed_status = status[k];
            }
            st.num_failed_shots++;
        }

    // outputs
    uint8_t *d_obs_bytes = obs_out;
    float *d_wf = weight_out;
    int64_t *d_wi = weight_int_out;
    if (!dev_io) {
        d->d_obs_bytes.ensure((size_t)n * std::max<uint32_t>(ob, 1));
        d_obs_bytes = d->d_obs_bytes.ptr;
        if (weight_out) {
            d->d_wfloat.ensure(n);
            d_wf = d->d_wfloat.ptr;
        }
        d_wi = nullptr;
    }
    k_finish<<<(n + 255) / 256, 256, 0, s>>>(d->d_obs64.ptr, d->d_wint.ptr, n, ob, 1.0 / d->host_ref().normalising_constant,
                                            d_obs_bytes, d_wf, d_wi);
    CUDA_CHECK(cudaGetLastError());
    launches++;
    CUDA_CHECK(cudaEventRecord(d->ev[4], s));
    if (!dev_io) {
        if (ob && obs_out) CUDA_CHECK(cudaMemcpyAsync(obs_out, d_obs_bytes, (size_t)n * ob, cudaMemcpyDeviceToHost, s));
        if (weight_out) CUDA_CHECK(cudaMemcpyAsync(weight_out, d_wf, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        if (weight_int_out) CUDA_CHECK(cudaMemcpyAsync(weight_int_out, d->d_wint.ptr, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    }
    if (obs_mask_out) CUDA_CHECK(cudaMemcpyAsync(obs_mask_out, d->d_obs64.ptr, 8, cudaMemcpyDeviceToHost, s));
    if (p.edge_dump) {
        uint64_t need = 0;
        CUDA_CHECK(cudaMemcpyAsync(&need, p.edge_dump_len, 8, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        if (num_edge_ints) *num_edge_ints = need;
        uint64_t take = std::min<uint64_t>(need, edges_cap);
        if (edges_out && take) CUDA_CHECK(cudaMemcpyAsync(edges_out, p.edge_dump, take * 8, cudaMemcpyDeviceToHost, s));
    }
    if (stats) {
        std::vector<uint32_t> rounds(n);
        CUDA_CHECK(cudaMemcpyAsync(rounds.data(), d->d_rounds.ptr, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        for (uint32_t r : rounds) st.matcher_rounds += r;
    }
    if (d->profile_phases) CUDA_CHECK(cudaMemcpyAsync(d->phase_cycles, d->d_phase.ptr, 192, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaEventRecord(d->ev[5], s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    st.kernel_launches = launches;
    CUDA_CHECK(cudaEventElapsedTime(&st.ms_h2d, d->ev[0], d->ev[1]));
    CUDA_CHECK(cudaEventElapsedTime(&st.ms_count, d->ev[1], d->ev[2]));
    CUDA_CHECK(cudaEventElapsedTime(&st.ms_decode, d->ev[2], d->ev[3]));
    CUDA_CHECK(cudaEventElapsedTime(&st.ms_d2h, d->ev[4], d->ev[5]));
    CUDA_CHECK(cudaEventElapsedTime(&st.ms_total, d->ev[0], d->ev[5]));
    if (stats) *stats = st;
    if (st.num_failed_shots) {
        g_error = std::string(status_message(st.first_failed_status)) + " (shot " + std::to_string(st.first_failed_shot) + ")";
        return CB200_ERR_DECODE;
    }
    return CB200_OK;
}

template <class F>
int guarded(F &&f) {
    try {
        return f();
    } catch (const std::invalid_argument &e) {
        g_error = e.what();
        return CB200_ERR_INVALID_ARGUMENT;
    } catch (const CudaError &e) {
        g_error = e.what();
        return CB200_ERR_CUDA;
    } catch (const std::exception &e) {
        g_error = e.what();
        return CB200_ERR_INVALID_ARGUMENT;
    }
}

}  // namespace

extern "C" {

const char *cb200_last_error(void) { return g_error.c_str(); }

// Uploads already compiled tables to `device` and returns a decoder bound to it.
static cb200_decoder *create_decoder(std::shared_ptr<const HostTables> tables, int device) {
    std::unique_ptr<cb200_decoder, void (*)(cb200_decoder *)> d(new cb200_decoder(), cb200_free);
    d->host_ptr = std::move(tables);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw CudaError(std::string("chromobius_b200 needs a CUDA device and has no CPU fallback: ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "no devices found"));
    if (device < 0 || device >= count) throw std::invalid_argument("device index out of range");
    d->device = device;
    CUDA_CHECK(cudaSetDevice(device));
    CUDA_CHECK(cudaDeviceGetAttribute(&d->num_sms, cudaDevAttrMultiProcessorCount, device));
    CUDA_CHECK(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    for (auto &e2 : d->ev) CUDA_CHECK(cudaEventCreate(&e2));
    upload_tables(d.get());
    return d.release();
}

int cb200_compile(const char *dem_text, int device, cb200_decoder **out) {
    if (!out) {
        g_error = "out must not be null";
        return CB200_ERR_INVALID_ARGUMENT;
    }
    *out = nullptr;
    return guarded([&]() {
        if (!dem_text) throw std::invalid_argument("dem_text must not be null");
        // host compile first: a malformed model is reported as such even where no GPU exists
        auto tables = std::make_shared<const HostTables>(compile_tables_cached(dem_text));
        *out = create_decoder(std::move(tables), device);
        return CB200_OK;
    });
}

void cb200_free(cb200_decoder *d) {
    if (!d) return;
    if (d->stream) {
        cudaSetDevice(d->device);
        cudaStreamSynchronize(d->stream);
    }
    if (d->l2_persist_want) l2_set_aside_release(d->device, d->l2_persist_want);
    d->d_tables.release();
    d->d_dets.release(); d->d_obs_bytes.release(); for (auto &a : d->arenas) a.buf.release(); d->d_obs64.release(); d->d_wint.release();
    d->d_wint_out.release(); d->d_edges.release(); d->d_status.release(); d->d_rounds.release(); d->d_counts.release();
    d->d_scalars.release(); d->d_list.release(); d->d_spare_claim.release(); d->d_wfloat.release(); d->d_phase.release();
    for (auto &e : d->ev)
        if (e) cudaEventDestroy(e);
    if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
}

uint64_t cb200_num_detectors(const cb200_decoder *d) { return d->host_ref().num_dets; }
uint64_t cb200_num_observables(const cb200_decoder *d) { return d->host_ref().num_obs; }
uint64_t cb200_num_detector_bytes(const cb200_decoder *d) { return (d->host_ref().num_dets + 7) / 8; }
uint64_t cb200_num_observable_bytes(const cb200_decoder *d) { return (d->host_ref().num_obs + 7) / 8; }
double cb200_normalising_constant(const cb200_decoder *d) { return d->host_ref().normalising_constant; }
uint64_t cb200_table_bytes(const cb200_decoder *d) { return d->table_bytes; }
uint64_t cb200_graph_half_edges(const cb200_decoder *d) { return d->host_ref().adj.size(); }

char *cb200_describe(const cb200_decoder *d) {
    std::string s = d->host_ref().describe();
    char *r = (char *)malloc(s.size() + 1);
    if (r) memcpy(r, s.c_str(), s.size() + 1);
    return r;
}
void cb200_free_string(char *s) { free(s); }

int cb200_set_tuning(cb200_decoder *d, uint32_t group_threads, uint32_t ht_log2) {
    if (group_threads != 0 && group_threads != 32 && group_threads != 128 && group_threads != 192 && group_threads != 256 &&
        group_threads != 512) {
        g_error = "group_threads must be 0, 32, 128, 192, 256 or 512";
        return CB200_ERR_INVALID_ARGUMENT;
    }
    if (ht_log2 != 0 && (ht_log2 < 4 || ht_log2 > 24)) {
        g_error = "ht_log2 must be 0 or within [4, 24]";
        return CB200_ERR_INVALID_ARGUMENT;
    }
    d->tune_group = group_threads;
    d->tune_ht_log2 = ht_log2;
    return CB200_OK;
}

int cb200_set_groups_per_sm(cb200_decoder *d, uint32_t groups_per_sm) {
    if (!d) return CB200_ERR_INVALID_ARGUMENT;
    d->tune_groups_per_sm = groups_per_sm;
    return CB200_OK;
}

int cb200_profile_phases(cb200_decoder *d, int enable, unsigned long long *cycles_out24) {
    if (!d) return CB200_ERR_INVALID_ARGUMENT;
    if (cycles_out24) memcpy(cycles_out24, d->phase_cycles, sizeof(d->phase_cycles));
    d->profile_phases = enable != 0;
    return CB200_OK;
}

int cb200_decode_batch(cb200_decoder *d, const uint8_t *dets, uint64_t num_shots, uint64_t row_stride, uint8_t *obs_out,
                       float *weight_out, int64_t *weight_int_out, uint32_t flags, cb200_batch_stats *stats) {
    return guarded([&]() {
        if (!d) throw std::invalid_argument("decoder must not be null");
        if (num_shots && !dets && cb200_num_detector_bytes(d)) throw std::invalid_argument("dets must not be null");
        if (num_shots && !obs_out && cb200_num_observable_bytes(d)) throw std::invalid_argument("obs_out must not be null");
        return decode_impl(d, dets, num_shots, row_stride, obs_out, weight_out, weight_int_out, flags, stats, nullptr, 0,
                           nullptr, nullptr);
    });
}

int cb200_decode_shot_edges(cb200_decoder *d, const uint8_t *dets_row, int64_t *edges_out, uint64_t edges_cap,
                            uint64_t *num_edge_ints, uint64_t *obs_mask_out, int64_t *weight_int_out) {
    return guarded([&]() {
        if (!d || !dets_row) throw std::invalid_argument("decoder and dets_row must not be null");
        std::vector<uint8_t> obs(std::max<uint64_t>(cb200_num_observable_bytes(d), 1));
        uint64_t need = 0;
        int rc = decode_impl(d, dets_row, 1, cb200_num_detector_bytes(d), obs.data(), nullptr, weight_int_out, CB200_IO_HOST,
                             nullptr, edges_out, edges_cap, &need, obs_mask_out);
        if (num_edge_ints) *num_edge_ints = need;
        return rc;
    });
}

// ---- one decoder replicated over several GPUs of the box (include/chromobius_b200.h) ----------------------------------
struct cb200_multi {
    std::vector<cb200_decoder *> replicas;
};

int cb200_compile_multi(const char *dem_text, const int *devices, int num_devices, cb200_multi **out) {
    if (!out) {
        g_error = "out must not be null";
        return CB200_ERR_INVALID_ARGUMENT;
    }
    *out = nullptr;
    return guarded([&]() {
        if (!dem_text) throw std::invalid_argument("dem_text must not be null");
        auto tables = std::make_shared<const HostTables>(compile_tables_cached(dem_text));
        std::vector<int> devs;
        if (devices && num_devices > 0) {
            devs.assign(devices, devices + num_devices);
        } else {
            int count = 0;
            if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
                throw CudaError("chromobius_b200 needs a CUDA device and has no CPU fallback: no devices found");
            for (int k = 0; k < count; k++) devs.push_back(k);
        }
        std::unique_ptr<cb200_multi> m(new cb200_multi());
        try {
            for (int dev : devs) m->replicas.push_back(create_decoder(tables, dev));
        } catch (...) {
            for (auto *r : m->replicas) cb200_free(r);
            throw;
        }
        *out = m.release();
        return CB200_OK;
    });
}

void cb200_multi_free(cb200_multi *m) {
    if (!m) return;
    for (auto *r : m->replicas) cb200_free(r);
    delete m;
}

int cb200_multi_num_devices(const cb200_multi *m) { return m ? (int)m->replicas.size() : 0; }

cb200_decoder *cb200_multi_replica(cb200_multi *m, int index) {
    return (m && index >= 0 && index < (int)m->replicas.size()) ? m->replicas[index] : nullptr;
}

int cb200_multi_decode_batch(cb200_multi *m, const uint8_t *dets, uint64_t num_shots, uint64_t row_stride, uint8_t *obs_out,
                             float *weight_out, int64_t *weight_int_out, cb200_batch_stats *stats) {
    if (!m || m->replicas.empty()) {
        g_error = "multi-GPU decoder must not be null";
        return CB200_ERR_INVALID_ARGUMENT;
    }
    const uint64_t G = m->replicas.size();
    const uint64_t ob = cb200_num_observable_bytes(m->replicas[0]);
    // contiguous row ranges, one host thread and one stream per device; results land in the caller's buffers
    std::vector<int> rc(G, CB200_OK);
    std::vector<std::string> err(G);
    std::vector<cb200_batch_stats> st(G);
    std::vector<uint64_t> first(G + 1, 0);
    for (uint64_t k = 0; k <= G; k++) first[k] = num_shots * k / G;
    std::vector<std::thread> threads;
    for (uint64_t k = 0; k < G; k++) {
        threads.emplace_back([&, k]() {
            const uint64_t a = first[k], n = first[k + 1] - first[k];
            st[k] = cb200_batch_stats{};
            st[k].first_failed_shot = -1;
            if (n == 0) return;
            rc[k] = cb200_decode_batch(m->replicas[k], dets ? dets + a * row_stride : nullptr, n, row_stride,
                                       obs_out ? obs_out + a * ob : nullptr, weight_out ? weight_out + a : nullptr,
                                       weight_int_out ? weight_int_out + a : nullptr, CB200_IO_HOST, &st[k]);
            if (rc[k] != CB200_OK) err[k] = cb200_last_error();  // thread-local in the worker: carried back by hand
        });
    }
    for (auto &t : threads) t.join();
    cb200_batch_stats total{};
    total.first_failed_shot = -1;
    int result = CB200_OK;
    for (uint64_t k = 0; k < G; k++) {
        total.num_shots += st[k].num_shots;
        total.num_detection_events += st[k].num_detection_events;
        total.num_retried_shots += st[k].num_retried_shots;
        total.num_failed_shots += st[k].num_failed_shots;
        if (total.first_failed_shot < 0 && st[k].first_failed_shot >= 0) {
            total.first_failed_shot = st[k].first_failed_shot + (int64_t)first[k];
            total.first_failed_status = st[k].first_failed_status;
        }
        total.max_events_per_shot = std::max(total.max_events_per_shot, st[k].max_events_per_shot);
        total.group_threads = st[k].group_threads ? st[k].group_threads : total.group_threads;
        total.num_groups += st[k].num_groups;
        total.kernel_launches += st[k].kernel_launches;
        total.ms_h2d = std::max(total.ms_h2d, st[k].ms_h2d);
        total.ms_count = std::max(total.ms_count, st[k].ms_count);
        total.ms_decode = std::max(total.ms_decode, st[k].ms_decode);
        total.ms_d2h = std::max(total.ms_d2h, st[k].ms_d2h);
        total.ms_total = std::max(total.ms_total, st[k].ms_total);
        total.matcher_rounds += st[k].matcher_rounds;
        if (rc[k] != CB200_OK && result == CB200_OK) {
            result = rc[k];
            g_error = err[k];
        }
    }
    if (stats) *stats = total;
    return result;
}

}  // extern "C"
