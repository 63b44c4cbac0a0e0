// Per-shot decode path of the B200 decoder, written once for the device (nvcc, sm_100a) and for a
// single-threaded host "simulation" build that tests/ uses to exercise the exact same logic without a GPU.
//
// Replaces, per shot (citations relative to /root/reference/src/chromobius):
//   stage 1  lift-in            decode/decoder.cc:231-247   -> ingest_shot()
//   stage 2  MWPM               decode/pymatcher.cc:27-42   -> match_shot() + paths_shot()  (see "Matcher" below)
//   stage 3  Euler tours        graph/euler_tours.h:89-112, graph/euler_tours.cc:21-97 -> lift_shot()
//   stage 4  discharge          decode/decoder.cc:80-203, 268-276                     -> lift_shot()
//
// Matcher. The reference delegates stage 2 to PyMatching's sparse blossom. A B200 wants independent work items, so
// the matcher here is a *certified local* exact MWPM:
//   1. every event grows a Dijkstra ball on the Moebius CSR until it has seen its nearest event(s)   [one warp per ball]
//   2. events that saw each other are merged into components (lock-free union-find)                  [thread per event]
//   3. each component is solved exactly by a dense primal-dual blossom on the seen pair distances    [thread per component]
//   4. the blossom duals certify the component against every UNSEEN pair: 2 y_a <= R_a for all a implies
//      y_a + y_b <= max(R_a, R_b) < d(a, b) for unseen (a, b), i.e. the duals stay feasible on the complete event graph,
//      so the union of the component matchings is a global minimum-weight perfect matching. Events whose dual is too
//      large for their ball (or that sit in an odd / unmatched component) re-grow their ball and the affected
//      components are re-solved. Typical shots need 2-5 rounds and components of 2-6 events.
//   5. matched pairs are expanded into graph edges along a canonical shortest path                   [one warp per pair]
// The total integer weight is exact (it is the optimum of the complete event graph); ties between equal-weight
// matchings are broken by this file's own deterministic rules, not PyMatching's.
//
// Cooperative execution model: a *group* (1 thread, 1 warp or 1 block) decodes one shot. Per-event / per-component
// loops are strided by the group's thread id, ball growth is strided by warp with the 32 lanes relaxing one adjacency
// row together; `g.sync()` separates phases. Ball tables (open-addressing hash + frontier) live in shared memory with
// a per-warp global-memory fallback. The host build runs with a group of one thread and one lane.
#ifndef CHROMOBIUS_B200_DECODE_CORE_H
#define CHROMOBIUS_B200_DECODE_CORE_H

#include <cstddef>
#include <cstdint>

#include "tables.h"

#ifdef __CUDACC__
#define CB_HD __host__ __device__ __forceinline__
// Large routines are kept out of line: the decode kernel is latency bound with many warps at different points of the
// code, and a fully inlined body (several hundred KB of SASS) thrashes the instruction cache.
#define CB_HDN __host__ __device__ __noinline__
#else
#define CB_HD inline
#define CB_HDN inline
#endif

// The sequential blossom code is latency bound and large; unrolled loops only make it miss the instruction cache.
#ifdef __CUDACC__
#define CB_LOOP _Pragma("unroll 1")
#define CB_UNROLL _Pragma("unroll")
#else
#define CB_LOOP
#define CB_UNROLL
#endif

namespace cb200 {

// Per-shot status codes (status_out). 0 = decoded.
enum ShotStatus : uint32_t {
    SHOT_OK = 0,
    SHOT_SCRATCH_OVERFLOW = 1,      // retryable with larger scratch (the host does that, on the GPU)
    SHOT_NO_PERFECT_MATCHING = 2,   // "No perfect matching exists for the detection events."
    SHOT_EULER_FAILED = 3,          // "Graph didn't decompose into Euler tours."   (euler_tours.cc:81)
    SHOT_DISCHARGE_FAILED = 4,      // "Failed to lift a flattened edge cycle ..."   (decoder.cc:176-202)
    SHOT_ROUND_LIMIT = 5,
    SHOT_COMPONENT_TOO_LARGE = 6,   // a matching component exceeded MAX_COMPONENT events
    SHOT_SCRATCH_OVERFLOW_SPARE = 7,  // out of scratch again in a spare (level-1) arena: the host retries from level 2
};

constexpr uint32_t MAX_COMPONENT = 2000;      // larger matching components are reported, not decoded
#ifndef CB_NARROW_COMPONENT
#define CB_NARROW_COMPONENT 250
#endif
constexpr uint32_t NARROW_COMPONENT = CB_NARROW_COMPONENT;  // up to here the blossom's edge matrix stores endpoints as u8, beyond as u16
constexpr uint32_t MAX_ROUNDS = 64;
// An open-ended growth request (odd component, exposed vertex) asks for this many events beyond the one it needs: one extra
// event per request means 20 % fewer matcher rounds at d25 r100 (4.49 -> 3.60 per shot) for slightly larger components.
#ifndef CB_GROW_EXTRA
#define CB_GROW_EXTRA 1
#endif
constexpr uint32_t GROW_EXTRA_EVENTS = CB_GROW_EXTRA;
// Radius of a ball that has seen everything reachable (an open-ended search that ran out of graph after its required event).
constexpr int64_t R_UNBOUNDED = INT64_MAX / 4;
constexpr uint32_t HT_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t FRONT_CLOSED = 0xFFFFFFFFu;
constexpr uint32_t SCAN_TMP = 1024;      // group sizes up to 1024 threads

struct NbEntry {
    uint32_t ev;   // event index within the shot
    uint32_t pos;  // position of the event's node in the near-table row of the list's owner (NO_INDEX: found by a search)
    int64_t d;     // exact shortest-path distance
};

// Sizes of the scratch of one group.
struct ScratchConfig {
    uint32_t sh_ht_log2;  // shared-memory ball table slots per warp = 1 << sh_ht_log2 (0: no shared tables)
    uint32_t sh_found;    // events a ball may report from the shared-memory attempt
    uint32_t ht_log2;     // global-memory fallback table slots per warp = 1 << ht_log2
    uint32_t found_cap;   // events a ball may report from the fallback attempt
    uint32_t max_events;  // matcher events (2 x detection events) per shot the arena is sized for
    uint32_t sh_pool;     // bytes of shared memory for blossom scratch (shares its space with the shared ball tables)
    uint64_t dyn_bytes;   // dynamic region (neighbour lists, blossom scratch, paths, Euler arrays)
    uint32_t sh_events;   // per-event arrays kept in shared memory instead of the arena: bit 0 uf + root, bit 1 pc + the
                          // flag bytes, bit 2 h (each only when max_events entries fit the launch's shared memory)
    uint32_t pad;         // experiment / test switches: bit 0 = never take the plain-cycle lift (lift_shot_plain)
};

// Launch-wide parameters.
struct DecodeParams {
    TablesView T;
    const uint8_t *dets;         // [num_shots][stride] bit packed detection events
    uint64_t stride;
    uint32_t num_shots;
    const uint32_t *shot_list;   // optional: indices of the shots to decode (retry passes); nullptr = 0..num_shots-1
    uint32_t list_len;
    uint64_t *obs_out;           // [num_shots]
    int64_t *weight_out;         // [num_shots] (nullable)
    uint32_t *status_out;        // [num_shots]
    uint32_t *rounds_out;        // [num_shots] (nullable) matcher rounds, for statistics
    // single-shot debug dump of the matched edges (nullable): interleaved node pairs, edge_dump_len = number of ints
    int64_t *edge_dump;
    uint64_t edge_dump_cap;
    uint64_t *edge_dump_len;
    // scratch
    uint8_t *arena;              // num_groups * arena_stride bytes; persistent parts pre-initialised by init_group_arena
    uint64_t arena_stride;
    ScratchConfig cfg;
    uint32_t *work_counter;      // persistent-kernel shot queue
    // A few spare arenas with larger scratch (the next escalation level's configuration, same shared-memory layout): a
    // group whose shot ran out of scratch claims one and decodes the shot again right away, while the other groups
    // carry on -- instead of a retry launch after the kernel, where the heaviest shots of the batch would run alone.
    uint8_t *spare_arena;        // num_spares * spare_stride bytes (nullptr: none)
    uint64_t spare_stride;
    ScratchConfig spare_cfg;
    uint32_t num_spares;
    uint32_t *spare_claim;       // [num_spares] 0 = free
    uint32_t *spare_retries;     // counter: shots decoded again in a spare arena
    // optional profiling: cycles spent by thread 0 of every group in
    // [0] ingest [1] first ball growth [2] component building [3] component solves [4] ball re-growth
    // [5] path extraction [6] Euler tours + discharge [7] shots
    // [8..11] blossom cycles for components of <= 8, <= 16, <= 32, > 32 events; [12..15] their counts
    // [16] closed-form component solves [17] per-warp blossoms [18] block-wide blossoms
    // [19] ball pops [20] cycles inside ball_grow [21] balls [22] nodes touched by balls [23] ball queries answered by
    // the near table
    unsigned long long *phase_cycles;
    // Streaming input (nullable): the host copies the batch in chunks while this kernel already decodes; *ready is the
    // number of leading shots whose rows have arrived (written by the copy stream after each chunk). A group that draws
    // a shot index at or beyond it waits.
    const uint32_t *ready;
    // Batch statistics gathered by the first-level launch (nullable): total and maximum detection events per shot.
    unsigned long long *event_total;
    uint32_t *event_max;
};

// ---------------------------------------------------------------------------------------------------------------
// Atomics / group abstraction
// ---------------------------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
CB_HD uint32_t cb_atomic_add(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
CB_HD uint64_t cb_atomic_add64(uint64_t *p, uint64_t v) { return atomicAdd((unsigned long long *)p, (unsigned long long)v); }
CB_HD uint64_t cb_atomic_xor64(uint64_t *p, uint64_t v) { return atomicXor((unsigned long long *)p, (unsigned long long)v); }
CB_HD uint32_t cb_atomic_cas(uint32_t *p, uint32_t expect, uint32_t v) { return atomicCAS(p, expect, v); }
CB_HD uint32_t cb_atomic_max(uint32_t *p, uint32_t v) { return atomicMax(p, v); }
CB_HD uint32_t cb_atomic_min(uint32_t *p, uint32_t v) { return atomicMin(p, v); }
CB_HD unsigned long long cb_clock() { return (unsigned long long)clock64(); }
CB_HD uint32_t cb_load_volatile(const uint32_t *p) { return *(const volatile uint32_t *)p; }
#else
CB_HD uint32_t cb_atomic_add(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
CB_HD uint64_t cb_atomic_add64(uint64_t *p, uint64_t v) { uint64_t o = *p; *p = o + v; return o; }
CB_HD uint64_t cb_atomic_xor64(uint64_t *p, uint64_t v) { uint64_t o = *p; *p = o ^ v; return o; }
CB_HD uint32_t cb_atomic_cas(uint32_t *p, uint32_t expect, uint32_t v) { uint32_t o = *p; if (o == expect) *p = v; return o; }
CB_HD uint32_t cb_atomic_max(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }
CB_HD uint32_t cb_atomic_min(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
CB_HD unsigned long long cb_clock() { return 0; }
CB_HD uint32_t cb_load_volatile(const uint32_t *p) { return *p; }
#endif

// Experimental sub-phase timers (compile with -DCB_XPROF): thread 0 of a group adds the cycles since the previous lap to
// phase_cycles[slot]; the slots 12..15 are borrowed from the blossom counters.
#ifdef CB_XPROF
#define CB_XLAP(s, g, t, slot) if ((s).prof && (g).tid == 0) { unsigned long long now_ = cb_clock(); cb_atomic_add64((uint64_t *)&(s).prof[slot], now_ - (t)); (t) = now_; }
#else
#define CB_XLAP(s, g, t, slot)
#endif

// Phase cycle counters of one thread (profiled calls). Unprofiled launches are compiled with NullTimer instead: the
// clock reads and thirteen 64-bit counters per thread cost 1.2 % of the decode kernel, and a run-time switch costs more
// than it saves (measured).
struct PhaseTimer {
    unsigned long long t[12];
    unsigned long long last;
    CB_HD void start() {
        for (int k = 0; k < 12; k++) t[k] = 0;
        last = cb_clock();
    }
    CB_HD void lap(int phase) {
        unsigned long long now = cb_clock();
        t[phase] += now - last;
        last = now;
    }
};
struct NullTimer {
    unsigned long long t[12];  // never read: DecodeParams::phase_cycles is null whenever this timer is compiled in
    CB_HD void start() {}
    CB_HD void lap(int) {}
};

// A group of one thread with one-lane "warps": the host simulation, and the device's thread-per-shot mode.
struct HostGroup {
    CB_HD int tid() const { return 0; }
    CB_HD int n() const { return 1; }
    CB_HD int lane() const { return 0; }
    CB_HD int lanes() const { return 1; }
    CB_HD int warp() const { return 0; }
    CB_HD int nwarps() const { return 1; }
    CB_HD void sync() const {}
    CB_HD void wsync() const {}
    CB_HD uint32_t ballot(bool p) const { return p ? 1u : 0u; }
    CB_HD uint32_t shfl32(uint32_t v, int) const { return v; }
    CB_HD uint64_t shfl64(uint64_t v, int) const { return v; }
    CB_HD uint64_t min64(uint64_t v) const { return v; }
    CB_HD uint64_t xor64(uint64_t v) const { return v; }
    CB_HD uint64_t add64(uint64_t v) const { return v; }
};

struct alignas(16) U4 {
    uint32_t x, y, z, w;
};

CB_HD uint64_t align16(uint64_t x) { return (x + 15) & ~uint64_t{15}; }
CB_HD uint32_t popc32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}
CB_HD uint32_t ctz32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return (uint32_t)(__ffs((int)v) - 1);
#else
    return (uint32_t)__builtin_ctz(v);
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// Memory of one group.
//   arena (global): [GroupHeader][scan_tmp][per-warp fallback ball tables]
//                   [per-event arrays][dynamic region]
//   shared        : [masked row words][word ranks][per-warp ball tables]
// Persistent-clean invariants (established by init_group_arena / init_group_shared, restored by every user):
//   every ball-table key == HT_EMPTY.
// ---------------------------------------------------------------------------------------------------------------
struct GroupHeader {
    uint32_t status;
    uint32_t ncomp;
    uint32_t mem_top;
    uint32_t ngrow;
    uint64_t dyn_top;     // bump pointer (bytes from dyn base)
    uint64_t weight;      // total matched weight
    uint64_t obs;         // XOR of the lifted cycles
    uint32_t scan_total;
    uint32_t bcast;
    uint32_t lift_ncomp;
    uint32_t lift_cyc_top;
    uint32_t nbig;
    uint32_t nhuge;
    uint32_t task_next;   // growth requests handed out so far (this round)
    uint32_t pad_hdr[3];
};

// Ball tables of one warp: open-addressing hash (node -> distance, frontier position) plus the frontier itself.
struct BallTables {
    uint32_t *key;     // [H]
    int64_t *dist;     // [H]
    uint32_t *fpos;    // [H] position in `front`, FRONT_CLOSED once settled
    uint32_t *used;    // [cap] slots in insertion order (O(touched) clean-up)
    uint64_t *front;   // [cap] open nodes as (distance << 24 | node): the minimum is the next node to settle, so the
                       // settling order is (distance, node id) -- independent of table size and lane count
    uint32_t *fslot;   // [cap] hash slot of every open node (parallel to front)
    NbEntry *found;    // [fcap]
    uint32_t H, cap, fcap;
    uint32_t nused, nfront;
    uint8_t *raw;      // start of the region (reused as blossom scratch between ball phases)
    uint64_t raw_bytes;
};

CB_HD uint64_t ball_table_bytes(uint32_t log2, uint32_t fcap) {
    if (log2 == 0) return 0;
    uint64_t H = uint64_t{1} << log2;
    uint64_t cap = H * 3 / 4;
    return align16(H * 4) + align16(H * 8) + align16(H * 4) + align16(cap * 4) * 2 + align16(cap * 8) +
           align16(uint64_t{fcap} * sizeof(NbEntry));
}
CB_HD void bind_ball_tables(BallTables &t, uint8_t *p, uint32_t log2, uint32_t fcap) {
    uint64_t H = uint64_t{1} << log2;
    uint64_t cap = H * 3 / 4;
    t.raw = p;
    t.raw_bytes = ball_table_bytes(log2, fcap);
    t.key = (uint32_t *)p; p += align16(H * 4);
    t.dist = (int64_t *)p; p += align16(H * 8);
    t.fpos = (uint32_t *)p; p += align16(H * 4);
    t.used = (uint32_t *)p; p += align16(cap * 4);
    t.front = (uint64_t *)p; p += align16(cap * 8);
    t.fslot = (uint32_t *)p; p += align16(cap * 4);
    t.found = (NbEntry *)p;
    t.H = log2 ? (uint32_t)H : 0;
    t.cap = (uint32_t)cap;
    t.fcap = fcap;
    t.nused = t.nfront = 0;
}

// Per-event arrays (all sized max_events), see struct Shot.
CB_HD uint64_t event_bytes(const ScratchConfig &c) {
    uint64_t m = c.max_events;
    // evdet(u32) R(i64) h(i64) nb_off nb_cnt grow_ev(u32) grow_to(i64) partner(i32) + 11 x u32 + 5 x u8
    return align16(m * 4) + align16(m * 8) * 2 + align16(m * 4) * 2 + align16(m * 4) + align16(m * 8) + align16(m * 4) +
           align16(m * 4) * 11 + align16(m) * 5;
}
CB_HD uint64_t shared_bytes_per_group(const ScratchConfig &c, uint32_t num_dets, uint32_t nwarps) {
    uint64_t nwords = (uint64_t{num_dets} + 31) / 32;
    uint64_t tables = nwarps * ball_table_bytes(c.sh_ht_log2, c.sh_found), pool = align16(c.sh_pool);
    uint64_t m = c.max_events, ev = 0;
    if (c.sh_events & 1) ev += align16(m * 4) * 2;
    if (c.sh_events & 2) ev += align16(m * 4) + align16(m) * 4;
    if (c.sh_events & 4) ev += align16(m * 8);
    return align16(nwords * 4) + align16((nwords + 1) * 4) + (tables > pool ? tables : pool) + ev;
}
CB_HD uint64_t arena_bytes_per_group(const ScratchConfig &c, uint32_t num_nodes, uint32_t nwarps) {
    return align16(sizeof(GroupHeader)) + align16(SCAN_TMP * 4) + nwarps * ball_table_bytes(c.ht_log2, c.found_cap) +
           event_bytes(c) + align16(c.dyn_bytes);
}

// Everything one group needs to decode one shot. One copy per WARP, in shared memory on the device (k_decode): the
// copies of a group are identical (pointers are computed with the same arithmetic) except for the ball tables and the
// small-allocation chunk, which belong to the warp. The object is passed by reference through the non-inlined stages; a
// per-thread copy would live in local memory (~650 B per thread, 160 KB per 256-thread group, re-read on every field
// access: measured as twice the global-load sectors of the whole kernel). After bind_shot the lanes only ever store
// lane-uniform values into it (M, the chunk state by lane 0, the ball-table counters).
struct Shot {
    TablesView T;
    ScratchConfig cfg;
    GroupHeader *hdr;
    uint32_t *scan_tmp;    // [SCAN_TMP]
    BallTables sh_ball;    // this warp's shared-memory tables (H == 0: none)
    uint8_t *sh_tables;    // all warps' shared-memory tables, contiguous: the blossom scratch pool between ball phases
    uint64_t sh_tables_bytes;
    BallTables gl_ball;    // this warp's global-memory fallback tables
    uint32_t *words;       // masked detection-event bits as little-endian u32 words (shared memory)
    uint32_t *word_rank;   // [nwords + 1] exclusive prefix popcount (shared memory)
    uint32_t nwords;
    uint32_t M;            // matcher events = 2 * detection events
    // per event
    uint32_t *evdet;       // [M/2] ascending detector ids
    int64_t *R;            // explored (open) ball radius
    int64_t *h;            // doubled dual of a solved event (h_a = 2 y_a)
    uint32_t *nb_off, *nb_cnt;
    uint32_t *grow_ev;
    int64_t *grow_to;
    int32_t *partner;
    uint32_t *uf, *root, *cnt, *start, *fillp, *members, *lidx, *comp_root, *path_off, *path_len, *blist;
    uint32_t *pc;          // solved events: smallest member of the component they were last solved in
    uint8_t *solved;       // the event holds a partner and a dual from a perfect solve of its component
    uint8_t *regrown;      // the ball grew since the component was last looked at
    uint8_t *stale;        // indexed by pc: a member's new ball shows a violated pair inside the solved component
    uint8_t *comp_dirty, *used_flag;
    // dynamic region
    uint8_t *dyn;
    uint64_t dyn_cap;
    unsigned long long *prof;  // DecodeParams::phase_cycles (nullable)
    // lane 0 of every warp sub-allocates small blocks from a private chunk (one global atomic per chunk)
    uint64_t chunk_off;
    uint32_t chunk_left;

    CB_HD uint32_t *lidx_huge() const { return fillp; }  // free between the member fill and the path phase
    CB_HD uint32_t event_node(uint32_t i) const { return 2 * evdet[i >> 1] + (i & 1); }
    CB_HD bool node_is_event(uint32_t node) const {
        uint32_t d = node >> 1;
        return (words[d >> 5] >> (d & 31)) & 1;
    }
    CB_HD uint32_t node_event_index(uint32_t node) const {
        uint32_t d = node >> 1;
        uint32_t w = words[d >> 5] & ((1u << (d & 31)) - 1);
        return 2 * (word_rank[d >> 5] + popc32(w)) + (node & 1);
    }
    CB_HD void fail(uint32_t code) const { cb_atomic_max(&hdr->status, code); }
    // Thread-safe bump allocation from the dynamic region; returns nullptr (and flags overflow) when exhausted.
    CB_HD uint8_t *alloc(uint64_t bytes) const {
        bytes = align16(bytes);
        uint64_t off = cb_atomic_add64(&hdr->dyn_top, bytes);
        if (off + bytes > dyn_cap) {
            fail(SHOT_SCRATCH_OVERFLOW);
            return nullptr;
        }
        return dyn + off;
    }
    // Small allocation by lane 0 of a warp; amortises the bump-pointer atomic over a 4 KB chunk.
    CB_HD uint8_t *alloc_small(uint64_t bytes) {
        bytes = align16(bytes);
#ifdef CB_NO_CHUNK
        return alloc(bytes);
#endif
        if (bytes > 1024) return alloc(bytes);
        if (bytes > chunk_left) {
            uint8_t *c = alloc(4096);
            if (!c) return nullptr;
            chunk_off = (uint64_t)(c - dyn);
            chunk_left = 4096;
        }
        uint8_t *r = dyn + chunk_off;
        chunk_off += bytes;
        chunk_left -= (uint32_t)bytes;
        return r;
    }
};

template <class G>
CB_HD void bind_shot(Shot &s, const DecodeParams &p, const ScratchConfig &cfg, uint8_t *arena, uint8_t *shared, const G &g) {
    s.T = p.T;
    s.cfg = cfg;
    uint8_t *q = arena;
    s.hdr = (GroupHeader *)q;
    q += align16(sizeof(GroupHeader));
    s.scan_tmp = (uint32_t *)q;
    q += align16(SCAN_TMP * 4);
    uint64_t gb = ball_table_bytes(cfg.ht_log2, cfg.found_cap);
    bind_ball_tables(s.gl_ball, q + (uint64_t)g.warp() * gb, cfg.ht_log2, cfg.found_cap);
    q += (uint64_t)g.nwarps() * gb;
    uint64_t m = cfg.max_events;
    s.evdet = (uint32_t *)q; q += align16(m * 4);
    s.R = (int64_t *)q; q += align16(m * 8);
    s.h = (int64_t *)q; q += align16(m * 8);
    s.nb_off = (uint32_t *)q; q += align16(m * 4);
    s.nb_cnt = (uint32_t *)q; q += align16(m * 4);
    s.grow_ev = (uint32_t *)q; q += align16(m * 4);
    s.grow_to = (int64_t *)q; q += align16(m * 8);
    s.partner = (int32_t *)q; q += align16(m * 4);
    uint32_t **arrs[] = {&s.uf, &s.root, &s.cnt, &s.start, &s.fillp, &s.members, &s.lidx, &s.comp_root, &s.path_off, &s.pc, &s.blist};
    for (auto a : arrs) { *a = (uint32_t *)q; q += align16(m * 4); }
    s.path_len = s.fillp;  // fillp is dead once the matcher has finished; path_len is only used afterwards
    s.solved = q; q += align16(m);
    s.regrown = q; q += align16(m);
    s.stale = q; q += align16(m);
    s.comp_dirty = q; q += align16(m);
    s.used_flag = q; q += align16(m);
    s.dyn = q;
    s.dyn_cap = align16(cfg.dyn_bytes);
    s.prof = p.phase_cycles;
    s.chunk_off = 0;
    s.chunk_left = 0;
    s.M = 0;
    // shared memory
    s.nwords = (p.T.num_dets + 31) / 32;
    uint8_t *sh = shared;
    s.words = (uint32_t *)sh; sh += align16(uint64_t{s.nwords} * 4);
    s.word_rank = (uint32_t *)sh; sh += align16((uint64_t{s.nwords} + 1) * 4);
    uint64_t sb = ball_table_bytes(cfg.sh_ht_log2, cfg.sh_found);
    bind_ball_tables(s.sh_ball, sh + (uint64_t)g.warp() * sb, cfg.sh_ht_log2, cfg.sh_found);
    s.sh_tables = sh;
    s.sh_tables_bytes = (uint64_t)g.nwarps() * sb;
    if (align16(cfg.sh_pool) > s.sh_tables_bytes) s.sh_tables_bytes = align16(cfg.sh_pool);
    // hot per-event arrays in shared memory (every pass of a matcher round re-reads them right after other threads of
    // the group wrote them, which always misses L1)
    sh += s.sh_tables_bytes;
    if (cfg.sh_events & 1) {
        s.uf = (uint32_t *)sh; sh += align16(m * 4);
        s.root = (uint32_t *)sh; sh += align16(m * 4);
    }
    if (cfg.sh_events & 2) {
        s.pc = (uint32_t *)sh; sh += align16(m * 4);
        s.solved = sh; sh += align16(m);
        s.regrown = sh; sh += align16(m);
        s.stale = sh; sh += align16(m);
        s.comp_dirty = sh; sh += align16(m);
    }
    if (cfg.sh_events & 4) {
        s.h = (int64_t *)sh; sh += align16(m * 8);
    }
}

template <class G>
CB_HD void bind_shot(Shot &s, const DecodeParams &p, uint8_t *arena, uint8_t *shared, const G &g) {
    bind_shot(s, p, p.cfg, arena, shared, g);
}

// Called once per group before the first shot of a launch (shared memory does not survive launches).
template <class G>
CB_HD void init_group_shared(Shot &s, const G &g) {
    for (uint32_t k = g.lane(); k < s.sh_ball.H; k += g.lanes()) s.sh_ball.key[k] = HT_EMPTY;
    g.sync();
}
// Called once per arena layout.
template <class G>
CB_HD void init_group_arena(Shot &s, const G &g) {
    for (uint32_t k = g.lane(); k < s.gl_ball.H; k += g.lanes()) s.gl_ball.key[k] = HT_EMPTY;
    g.sync();
}

// Exclusive prefix sum of arr[0..n) in place, cooperatively; returns the total. Needs g.n() <= SCAN_TMP.
template <class G>
CB_HDN uint32_t group_exclusive_scan(Shot &s, const G &g, uint32_t *arr, uint32_t n) {
    // per-thread chunk sums -> inclusive scan inside the warp (shuffles) -> warp totals through scan_tmp
    uint32_t chunk = (n + g.n() - 1) / g.n();
    uint32_t lo = (uint32_t)g.tid() * chunk, hi = lo + chunk;
    if (lo > n) lo = n;
    if (hi > n) hi = n;
    uint32_t sum = 0;
    CB_LOOP
    for (uint32_t k = lo; k < hi; k++) sum += arr[k];
    uint32_t incl = sum;
    CB_LOOP
    for (int o = 1; o < g.lanes(); o <<= 1) {
        uint32_t t = g.shfl32(incl, g.lane() >= o ? g.lane() - o : g.lane());
        if (g.lane() >= o) incl += t;
    }
    if (g.lane() == g.lanes() - 1) s.scan_tmp[g.warp()] = incl;
    g.sync();
    uint32_t run = incl - sum, total = 0;
    CB_LOOP
    for (int w = 0; w < g.nwarps(); w++) {
        uint32_t v = s.scan_tmp[w];
        if (w < g.warp()) run += v;
        total += v;
    }
    CB_LOOP
    for (uint32_t k = lo; k < hi; k++) {
        uint32_t v = arr[k];
        arr[k] = run;
        run += v;
    }
    g.sync();
    return total;
}

// ---------------------------------------------------------------------------------------------------------------
// Stage 1: lift-in (decode/decoder.cc:231-247). The masked row doubles as the node -> event index map.
// Pad bits and ignored detectors are cleared by the ingest mask.
// ---------------------------------------------------------------------------------------------------------------
template <class G>
CB_HDN void ingest_shot(Shot &s, const uint8_t *row, const G &g) {
    const uint32_t nb = s.T.num_det_bytes;
    const uint32_t full = nb / 4;  // words that lie completely inside the row
    if ((((uintptr_t)row) & 3u) == 0 && full) {
        // 4-byte aligned row (every row of a contiguous batch whose detector byte count is a multiple of 4, e.g. 3076 at
        // d25 r100): words up to the first 16-byte boundary one by one, then 16-byte vector loads, four words each, then
        // the rest. The mask table is 16-byte aligned (tables.h) and is read the same way.
        const uint32_t *row32 = (const uint32_t *)row;
        const uint32_t *mask32 = (const uint32_t *)s.T.ingest_mask;
        uint32_t head = (uint32_t)(((16u - (uint32_t)((uintptr_t)row & 15u)) & 15u) >> 2);
        if (head > full) head = full;
        const uint32_t body = (full - head) / 4;
        CB_LOOP
        for (uint32_t w = g.tid(); w < head; w += g.n()) {
            const uint32_t v = row32[w] & mask32[w];
            s.words[w] = v;
            s.word_rank[w] = popc32(v);
        }
        const U4 *row128 = (const U4 *)(row32 + head);
        CB_LOOP
        for (uint32_t k = g.tid(); k < body; k += g.n()) {
            const U4 r = row128[k];
            const uint32_t w = head + 4 * k;
            const uint32_t v0 = r.x & mask32[w], v1 = r.y & mask32[w + 1], v2 = r.z & mask32[w + 2], v3 = r.w & mask32[w + 3];
            s.words[w] = v0;
            s.words[w + 1] = v1;
            s.words[w + 2] = v2;
            s.words[w + 3] = v3;
            s.word_rank[w] = popc32(v0);
            s.word_rank[w + 1] = popc32(v1);
            s.word_rank[w + 2] = popc32(v2);
            s.word_rank[w + 3] = popc32(v3);
        }
        CB_LOOP
        for (uint32_t w = head + 4 * body + g.tid(); w < s.nwords; w += g.n()) {
            uint32_t v = 0;
            if (w < full) {
                v = row32[w] & mask32[w];
            } else {
                CB_LOOP
                for (uint32_t b = 0; b < 4; b++) {
                    const uint32_t k = 4 * w + b;
                    if (k < nb) v |= (uint32_t)(row[k] & s.T.ingest_mask[k]) << (8 * b);
                }
            }
            s.words[w] = v;
            s.word_rank[w] = popc32(v);
        }
    } else {
        CB_LOOP
        for (uint32_t w = g.tid(); w < s.nwords; w += g.n()) {
            uint32_t v = 0;
            CB_LOOP
            for (uint32_t b = 0; b < 4; b++) {
                uint32_t k = 4 * w + b;
                if (k < nb) v |= (uint32_t)(row[k] & s.T.ingest_mask[k]) << (8 * b);
            }
            s.words[w] = v;
            s.word_rank[w] = popc32(v);
        }
    }
    if (g.tid() == 0) {
        s.hdr->status = 0;
        s.hdr->dyn_top = 0;
        s.hdr->weight = 0;
        s.hdr->obs = 0;
    }
    g.sync();
    uint32_t total = group_exclusive_scan(s, g, s.word_rank, s.nwords);
    if (g.tid() == 0) s.word_rank[s.nwords] = total;
    if (2 * total > s.cfg.max_events) {
        if (g.tid() == 0) s.fail(SHOT_SCRATCH_OVERFLOW);
        s.M = 0;
        g.sync();
        return;
    }
    CB_LOOP
    for (uint32_t w = g.tid(); w < s.nwords; w += g.n()) {
        uint32_t run = s.word_rank[w];
        uint32_t v = s.words[w];
        CB_LOOP
        while (v) {
            uint32_t b = ctz32(v);
            v &= v - 1;
            s.evdet[run++] = 32 * w + b;
        }
    }
    s.M = 2 * total;
    g.sync();
}

// ---------------------------------------------------------------------------------------------------------------
// Ball growth: truncated Dijkstra on the Moebius CSR, one warp per ball. The frontier is an unsorted list of hash
// slots; the minimum is found by a warp reduction, the popped node's adjacency row is relaxed one lane per edge.
// ---------------------------------------------------------------------------------------------------------------
CB_HD uint32_t ht_hash(uint32_t node, uint32_t H) { return ((node ^ (node >> 7)) * 2654435761u) & (H - 1); }

CB_HD int64_t ball_lookup(const BallTables &t, uint32_t node) {
    uint32_t k = ht_hash(node, t.H);
    while (true) {
        uint32_t cur = t.key[k];
        if (cur == node) return t.dist[k];
        if (cur == HT_EMPTY) return -1;
        k = (k + 1) & (t.H - 1);
    }
}
template <class G>
CB_HD void ball_clear(BallTables &t, const G &g) {
    g.wsync();
    uint32_t *const key = t.key;
    const uint32_t *const used = t.used;
    const uint32_t n = t.nused;
    for (uint32_t k = g.lane(); k < n; k += g.lanes()) key[used[k]] = HT_EMPTY;
    t.nused = t.nfront = 0;
    g.wsync();
}

enum BallResult : int { BALL_OK = 0, BALL_OVERFLOW = 1, BALL_EXHAUSTED = 2 };

// Adjacency row of a node: fixed-stride (ELL) rows when the graph's maximum degree allows, CSR otherwise.
struct RowRange {
    uint32_t e0, e1;
};
CB_HD RowRange row_of(const TablesView &T, uint32_t u) {
    RowRange r;
    if (T.ell_shift) {
        r.e0 = u << T.ell_shift;
        r.e1 = r.e0 + (1u << T.ell_shift);
    } else {
        r.e0 = T.row_ptr[u];
        r.e1 = T.row_ptr[u + 1];
    }
    return r;
}
CB_HD const AdjEntry *adj_of(const TablesView &T) { return T.ell_shift ? T.ell_adj : T.adj; }

// Grows a ball around `src` (all lanes of the warp call this together; every lane returns the same values).
// Nodes are settled in order of distance.
//   radius <  0 : stop as soon as more than `need_more_than + extra` other events have been settled; *limit_out =
//                 distance of the last one. Every event strictly closer than *limit_out has been reported. When the
//                 graph runs out first: R_UNBOUNDED if more than `need_more_than` were found, else BALL_EXHAUSTED.
//   radius >= 0, !inclusive : settle every node strictly closer than `radius`; *limit_out = radius.
//   radius >= 0,  inclusive : settle every node within `radius`, stopping early when `stop_node` is settled.
// Other events settled inside the ball are appended to t.found when collect_events is set.
// The tables are left populated (the caller clears them).
template <class G>
CB_HDN int ball_grow(const Shot &s, const G &g, BallTables &t, uint32_t src, int64_t radius, uint32_t need_more_than,
                     uint32_t stop_node, bool collect_events, bool inclusive, int64_t *limit_out, uint32_t *nfound_out,
                     uint32_t extra = 0) {
    // Everything the loop touches is copied into locals first: this routine is kept out of line, and fields reached
    // through the references would otherwise be re-read from local memory on every use.
    uint32_t *const key = t.key;
    int64_t *const dist = t.dist;
    uint32_t *const fpos = t.fpos;
    uint32_t *const used = t.used;
    uint64_t *const front = t.front;
    uint32_t *const fslot = t.fslot;
    NbEntry *const found = t.found;
    const uint32_t H = t.H, hmask = t.H - 1, cap = t.cap, fcap = t.fcap;
    const uint32_t lane = (uint32_t)g.lane(), lanes = (uint32_t)g.lanes();
    const uint32_t lane_lt = (1u << lane) - 1;
    const uint32_t ell_shift = s.T.ell_shift;
    const AdjEntry *const adj = ell_shift ? s.T.ell_adj : s.T.adj;
    const uint32_t *const row_ptr = s.T.row_ptr;
    const uint32_t *const words = s.words;
    const uint32_t *const word_rank = s.word_rank;
    uint32_t nused = 1, nfront = 1;
    if (lane == 0) {
        uint32_t k = ht_hash(src, H);
        key[k] = src;
        dist[k] = 0;
        fpos[k] = 0;
        used[0] = k;
        front[0] = src;
        fslot[0] = k;
    }
    g.wsync();
    uint32_t nfound = 0;
    int64_t limit = radius;
    int result = BALL_OK;
    // Pruning (changes no result). Nearest-event search (radius < 0, need_more_than == 0): `cand` is the smallest
    // tentative distance of an event seen so far; nothing beyond it can be settled before the search stops, so nodes
    // beyond it are not inserted and a settled node whose every edge leads beyond it is not relaxed at all.
    // Radius searches prune against the radius in the same way.
    const int64_t min_wt = s.T.min_wt;
    const bool nearest_only = radius < 0 && need_more_than == 0 && extra == 0 && collect_events;
    int64_t cand = INT64_MAX / 4;
    const unsigned long long prof_t0 = s.prof ? cb_clock() : 0;
    uint32_t prof_pops = 0;
    while (nfront) {
        prof_pops++;
        // pop the closest open node
        uint64_t best = ~uint64_t{0};
        uint32_t best_q = 0;
        for (uint32_t q = lane; q < nfront; q += lanes) {
            uint64_t cand = front[q];
            if (cand < best) {
                best = cand;
                best_q = q;
            }
        }
        const uint64_t mine = best;
        best = g.min64(best);
        const int64_t d = (int64_t)(best >> 24);
        const uint32_t u = (uint32_t)(best & 0xFFFFFF);
        if (limit >= 0 && (inclusive ? d > limit : d >= limit)) break;
        if (nearest_only && d + min_wt > cand) {
            // No relaxation can reach cand any more, so every tentative distance up to cand is final: the answer is the
            // open event at distance cand with the smallest node id -- exactly the node the settling order would reach.
            uint64_t hit = ~uint64_t{0};
            for (uint32_t q = lane; q < nfront; q += lanes) {
                const uint64_t cnd = front[q];
                const uint32_t node = (uint32_t)(cnd & 0xFFFFFF);
                const uint32_t det = node >> 1;
                if ((int64_t)(cnd >> 24) == cand && node != src && ((words[det >> 5] >> (det & 31)) & 1) && cnd < hit) hit = cnd;
            }
            hit = g.min64(hit);
            const uint32_t node = (uint32_t)(hit & 0xFFFFFF);
            const uint32_t det = node >> 1;
            if (lane == 0) {
                const uint32_t w = words[det >> 5];
                found[0].ev = 2 * (word_rank[det >> 5] + popc32(w & ((1u << (det & 31)) - 1))) + (node & 1);
                found[0].pos = NO_INDEX;
                found[0].d = cand;
            }
            nfound = 1;
            limit = cand;
            break;
        }
        // Start fetching the node's adjacency row now; the frontier bookkeeping below hides part of the latency.
        uint32_t e0, e1;
        if (ell_shift) {
            e0 = u << ell_shift;
            e1 = e0 + (1u << ell_shift);
        } else {
            e0 = row_ptr[u];
            e1 = row_ptr[u + 1];
        }
        AdjEntry first_seg;
        first_seg.nbr = NO_INDEX;
        first_seg.wt = 0;
        if (e0 + lane < e1) first_seg = adj[e0 + lane];
        // the lane that held the minimum knows its frontier position (keys are unique: one entry per node)
        const uint32_t owner = ctz32(g.ballot(mine == best));
        const uint32_t q = g.shfl32(best_q, (int)owner);
        const uint32_t slot = fslot[q];
        g.wsync();
        if (lane == 0) {
            const uint32_t last_slot = fslot[nfront - 1];
            front[q] = front[nfront - 1];
            fslot[q] = last_slot;
            fpos[last_slot] = q;
            fpos[slot] = FRONT_CLOSED;
        }
        nfront--;
        g.wsync();
        if (u == stop_node) break;
        if (collect_events && u != src) {
            const uint32_t det = u >> 1;
            const uint32_t w = words[det >> 5];
            if ((w >> (det & 31)) & 1) {
                if (nfound >= fcap) {
                    result = BALL_OVERFLOW;
                    break;
                }
                if (lane == 0) {
                    found[nfound].ev = 2 * (word_rank[det >> 5] + popc32(w & ((1u << (det & 31)) - 1))) + (u & 1);
                    found[nfound].pos = NO_INDEX;
                    found[nfound].d = d;
                }
                nfound++;
                if (limit < 0 && nfound > need_more_than + extra) {
                    limit = d;
                    break;
                }
            }
        }
        if (nearest_only ? d + min_wt > cand : (limit >= 0 && (inclusive ? d + min_wt > limit : d + min_wt >= limit))) continue;
        if (nused + (e1 - e0) > cap) {
            result = BALL_OVERFLOW;
            break;
        }
        int64_t step_cand = cand;
        for (uint32_t base = e0; base < e1; base += lanes) {
            const uint32_t e = base + lane;
            bool is_new = false, too_far = false;
            uint32_t my_slot = 0, my_node = 0;
            int64_t nd = 0;
            if (e < e1) {
                AdjEntry a = base == e0 ? first_seg : adj[e];
                nd = d + a.wt;
                my_node = a.nbr;
                if (a.nbr != NO_INDEX && !(limit >= 0 && (inclusive ? nd > limit : nd >= limit)) && nd <= cand) {
                    too_far = nd >= (int64_t{1} << 39);  // frontier keys hold 40 distance bits
                    if (nearest_only && a.nbr != src) {
                        const uint32_t det = a.nbr >> 1;
                        if (((words[det >> 5] >> (det & 31)) & 1) && nd < step_cand) step_cand = nd;
                    }
                    uint32_t k = ht_hash(a.nbr, H);
                    while (true) {
                        uint32_t old = cb_atomic_cas(&key[k], HT_EMPTY, a.nbr);
                        if (old == HT_EMPTY) {
                            is_new = true;
                            break;
                        }
                        if (old == a.nbr) break;
                        k = (k + 1) & hmask;
                    }
                    my_slot = k;
                    if (is_new) {
                        dist[k] = nd;
                    } else {
                        uint32_t fp = fpos[k];
                        if (fp != FRONT_CLOSED && nd < dist[k]) {
                            dist[k] = nd;
                            front[fp] = ((uint64_t)nd << 24) | a.nbr;
                        }
                    }
                }
            }
            const uint32_t fresh = g.ballot(is_new);
            if (is_new) {
                uint32_t idx = popc32(fresh & lane_lt);
                used[nused + idx] = my_slot;
                front[nfront + idx] = ((uint64_t)nd << 24) | my_node;
                fslot[nfront + idx] = my_slot;
                fpos[my_slot] = nfront + idx;
            }
            const uint32_t c = popc32(fresh);
            nused += c;
            nfront += c;
            g.wsync();
            if (g.ballot(too_far)) result = BALL_OVERFLOW;
        }
        if (nearest_only) cand = (int64_t)g.min64((uint64_t)step_cand);
        if (result != BALL_OK) break;
    }
    t.nused = nused;
    t.nfront = nfront;
#if !defined(CB_TAIL_PROF) && !defined(CB_SUBPROF)
    if (s.prof && lane == 0) {
        cb_atomic_add64((uint64_t *)&s.prof[19], prof_pops);
        cb_atomic_add64((uint64_t *)&s.prof[20], cb_clock() - prof_t0);
        cb_atomic_add64((uint64_t *)&s.prof[21], 1);
        cb_atomic_add64((uint64_t *)&s.prof[22], nused);
    }
#endif
    if (result == BALL_OK && limit < 0) {
        if (collect_events && nfound > need_more_than) limit = R_UNBOUNDED;
        else result = BALL_EXHAUSTED;
    }
    *limit_out = limit;
    *nfound_out = nfound;
    return result;
}

CB_HD const NbEntry *nb_list(const Shot &s, uint32_t i) { return (const NbEntry *)(s.dyn + ((uint64_t)s.nb_off[i] << 4)); }

// ---------------------------------------------------------------------------------------------------------------
// Near-table ball queries. The row of a node lists its closest nodes in the order a ball would settle them
// (tables.h), so a ball query is a scan of the row against the shot's event bits. Both routines return false when the
// row ends before the query is decided; the caller then runs the Dijkstra search, which gives the same answer.
// ---------------------------------------------------------------------------------------------------------------
CB_HD const NearEntry *near_at(const TablesView &T, uint32_t node, uint32_t p) {
    return T.near + ((uint64_t)(p >> 5) * T.num_nodes + node) * 32 + (p & 31);
}
CB_HD uint32_t nth_set_bit(uint32_t mask, uint32_t n) {  // position of the n-th (0-based) set bit
    while (n--) mask &= mask - 1;
    return ctz32(mask);
}

// Same contract as ball_grow(..., collect_events = true, inclusive = false) followed by the copy into the neighbour
// list of event i: on success nb_off / nb_cnt / R / regrown of event i are written.
template <class G>
CB_HDN bool explore_event_near(Shot &s, const G &g, uint32_t i, int64_t radius, uint32_t need_more_than, uint32_t extra) {
    const uint32_t stride = s.T.near_stride;
    if (!stride) return false;
    const uint32_t src = s.event_node(i);
    const uint32_t lane = (uint32_t)g.lane(), lanes = (uint32_t)g.lanes();
    const uint32_t lane_lt = (1u << lane) - 1;
    const uint32_t *const words = s.words;
    const uint32_t *const word_rank = s.word_rank;
    // pass 1: how far the row has to be read, and how many events that reports
    uint32_t nfound = 0, end_pos = 0;  // entries [0, end_pos) take part
    int64_t limit = radius;
    bool decided = false;
    CB_LOOP
    for (uint32_t pos = 0; pos < stride && !decided; pos += lanes) {
        const NearEntry e = *near_at(s.T, src, pos + lane);
        const bool pad = e.dist_flag == NEAR_PAD;
        const int64_t d = (int64_t)(e.dist_flag & 0x7FFFFFFFu);
        const uint32_t det = (e.node_pred & 0xFFFFFFu) >> 1;
        const bool in_range = !pad && (radius < 0 || d < radius);
        const bool is_ev = in_range && ((words[det >> 5] >> (det & 31)) & 1);
        const uint32_t evs = g.ballot(is_ev);
        const uint32_t outs = g.ballot(!in_range);
        if (radius < 0) {
            const uint32_t usable = outs ? evs & ((1u << ctz32(outs)) - 1) : evs;  // events before the first pad entry
            if (nfound + popc32(usable) > need_more_than + extra) {
                const uint32_t l = nth_set_bit(usable, need_more_than + extra - nfound);
                limit = (int64_t)(g.shfl32(e.dist_flag, (int)l) & 0x7FFFFFFFu);
                nfound = need_more_than + extra + 1;
                end_pos = pos + l + 1;
                decided = true;
            } else if (outs) {
                // the row ends here: the graph has no further node. Enough events for the request itself -> everything
                // reachable has been seen; otherwise the search reports the exhaustion
                if (nfound + popc32(usable) <= need_more_than) return false;
                nfound += popc32(usable);
                limit = R_UNBOUNDED;
                end_pos = pos + ctz32(outs);
                decided = true;
            } else {
                nfound += popc32(evs);
            }
        } else {
            if (outs) {
                const uint32_t l = ctz32(outs);  // sorted by distance: everything from here on is outside the radius
                nfound += popc32(evs & ((1u << l) - 1));
                end_pos = pos + l;
                decided = true;
            } else {
                nfound += popc32(evs);
            }
        }
    }
    if (!decided) return false;
    uint64_t off = 0;
    if (lane == 0) {
        uint8_t *dst = s.alloc_small(uint64_t{nfound} * sizeof(NbEntry) + 16);
        off = dst ? (uint64_t)(dst - s.dyn) + 1 : 0;
    }
    off = g.shfl64(off, 0);
    if (!off) return true;  // alloc flagged the overflow
    NbEntry *dst = (NbEntry *)(s.dyn + (off - 1));
    uint32_t run = 0;
    CB_LOOP
    for (uint32_t pos = 0; pos < end_pos; pos += lanes) {
        const uint32_t p = pos + lane;
        bool is_ev = false;
        NearEntry e;
        e.node_pred = 0;
        e.dist_flag = 0;
        if (p < end_pos) {
            e = *near_at(s.T, src, p);
            const uint32_t det = (e.node_pred & 0xFFFFFFu) >> 1;
            is_ev = (words[det >> 5] >> (det & 31)) & 1;
        }
        const uint32_t evs = g.ballot(is_ev);
        if (is_ev) {
            const uint32_t node = e.node_pred & 0xFFFFFFu, det = node >> 1;
            NbEntry out;
            out.ev = 2 * (word_rank[det >> 5] + popc32(words[det >> 5] & ((1u << (det & 31)) - 1))) + (node & 1);
            out.pos = p;
            out.d = (int64_t)(e.dist_flag & 0x7FFFFFFFu);
            dst[run + popc32(evs & lane_lt)] = out;
        }
        run += popc32(evs);
    }
    if (lane == 0) {
        s.regrown[i] = 1;
        s.nb_off[i] = (uint32_t)((off - 1) >> 4);
        s.nb_cnt[i] = nfound;
        s.R[i] = limit;
#if !defined(CB_TAIL_PROF) && !defined(CB_SUBPROF)
        if (s.prof) cb_atomic_add64((uint64_t *)&s.prof[23], 1);
#endif
    }
    g.wsync();
    return true;
}

// Grows the ball of event i (one warp). Writes nb_off/nb_cnt/R of event i and flags it as regrown.
template <class G>
CB_HDN void explore_event(Shot &s, const G &g, uint32_t i, int64_t radius, uint32_t need_more_than, uint32_t extra = 0) {
    if (explore_event_near(s, g, i, radius, need_more_than, extra)) return;
    const uint32_t src = s.event_node(i);
    int64_t limit = 0;
    uint32_t nfound = 0;
    BallTables *t = nullptr;
    int r = BALL_OVERFLOW;
    for (int attempt = s.sh_ball.H ? 0 : 1; attempt < 2 && r == BALL_OVERFLOW; attempt++) {
        if (t) ball_clear(*t, g);
        t = attempt ? &s.gl_ball : &s.sh_ball;
        r = ball_grow(s, g, *t, src, radius, need_more_than, NO_INDEX, true, false, &limit, &nfound, extra);
    }
    ball_clear(*t, g);
    if (r == BALL_OVERFLOW) {
        if (g.lane() == 0) s.fail(SHOT_SCRATCH_OVERFLOW);
        return;
    }
    if (r == BALL_EXHAUSTED) {
        // Nothing more is reachable: this event cannot be paired (odd component of the Moebius graph).
        if (g.lane() == 0) s.fail(SHOT_NO_PERFECT_MATCHING);
        return;
    }
    uint64_t off = 0;
    if (g.lane() == 0) {
        uint8_t *dst = s.alloc_small(uint64_t{nfound} * sizeof(NbEntry) + 16);
        off = dst ? (uint64_t)(dst - s.dyn) + 1 : 0;
    }
    off = g.shfl64(off, 0);
    if (!off) return;
    NbEntry *dst = (NbEntry *)(s.dyn + (off - 1));
    for (uint32_t k = g.lane(); k < nfound; k += g.lanes()) {
        dst[k] = t->found[k];
    }
    if (g.lane() == 0) {
        s.regrown[i] = 1;
        s.nb_off[i] = (uint32_t)((off - 1) >> 4);
        s.nb_cnt[i] = nfound;
        s.R[i] = limit;
    }
    g.wsync();
}

// First round: the nearest event of every event, one THREAD per event. Each thread walks its near-table row two
// entries (16 bytes) at a time until it meets an event; the one-entry neighbour lists live in one block of the dynamic
// region (entry i belongs to event i). Events whose row holds no event are left with nb_cnt == 0 for explore_event.
template <class G>
CB_HDN void explore_first_near(Shot &s, const G &g) {
    const uint32_t M = s.M, stride = s.T.near_stride;
    if (g.tid() == 0) {
        uint8_t *mem = stride ? s.alloc(uint64_t{M} * sizeof(NbEntry)) : nullptr;
        s.hdr->bcast = mem ? (uint32_t)((mem - s.dyn) >> 4) + 1 : 0;
    }
    g.sync();
    const uint32_t base = s.hdr->bcast;
    g.sync();
    if (!base) return;
    NbEntry *const lists = (NbEntry *)(s.dyn + ((uint64_t)(base - 1) << 4));
    const uint32_t *const words = s.words;
    for (uint32_t i = (uint32_t)g.tid(); i < M; i += (uint32_t)g.n()) {
        const uint32_t src = s.event_node(i);
        for (uint32_t p = 0; p < stride; p += 4) {
            // rows are 256-byte aligned and p is a multiple of four: two independent 16-byte loads per step
            const U4 pair0 = *(const U4 *)near_at(s.T, src, p);
            const U4 pair1 = *(const U4 *)near_at(s.T, src, p + 2);
            const NearEntry e[4] = {{pair0.x, pair0.y}, {pair0.z, pair0.w}, {pair1.x, pair1.y}, {pair1.z, pair1.w}};
            bool stop = false;
            for (int k = 0; k < 4 && !stop; k++) {
                if (e[k].dist_flag == NEAR_PAD) {
                    stop = true;  // the row ends without an event: the search decides (it reports an exhausted graph)
                    break;
                }
                const uint32_t node = e[k].node_pred & 0xFFFFFFu, det = node >> 1;
                const uint32_t w = words[det >> 5];
                if ((w >> (det & 31)) & 1) {
                    NbEntry out;
                    out.ev = 2 * (s.word_rank[det >> 5] + popc32(w & ((1u << (det & 31)) - 1))) + (node & 1);
                    out.pos = p + (uint32_t)k;
                    out.d = (int64_t)(e[k].dist_flag & 0x7FFFFFFFu);
                    lists[i] = out;
                    s.regrown[i] = 1;
                    s.nb_off[i] = base - 1 + i;
                    s.nb_cnt[i] = 1;
                    s.R[i] = out.d;
                    stop = true;
                }
            }
            if (stop) break;
        }
    }
    g.sync();
}

// ---------------------------------------------------------------------------------------------------------------
// Dense primal-dual blossom (maximum weight among maximum-cardinality matchings) over one component.
// Vertices 1..n, blossoms n+1..NX-1. Edge (u, v) exists iff W[u][v] > 0. Doubled duals `lab` as in the classic
// O(n^3) formulation: lab[u] + lab[v] >= 2 W(u, v), tight on matched edges.
// All storage is caller provided; recursion is replaced by explicit stacks.
//
// Execution: the primal-dual bookkeeping (trees, blossoms, augmentations) is sequential and runs on lane 0 of the warp
// ("master"). Its O(n) / O(n^2) inner loops -- slack recomputation, the dual step, blossom row merges, initialisation --
// are data parallel and are executed by all 32 lanes: lanes 1..31 sit in worker_loop() and execute the helper the
// master announces through a small command block, synchronising with __syncwarp (which does not require the lanes to
// be at the same program counter). Every helper reproduces the sequential loop's result exactly (ties go to the
// lowest index), so the matching and the duals do not depend on the number of lanes. Small components run on the
// master alone (`par == false`), and so does the host build (one lane).
// ---------------------------------------------------------------------------------------------------------------
#ifndef CB_BLOSSOM_PAR_MIN
#define CB_BLOSSOM_PAR_MIN 33
#endif
constexpr int BLOSSOM_PAR_MIN = CB_BLOSSOM_PAR_MIN;  // components with at least this many events use the warp helpers
#ifndef CB_BLOSSOM_PAR_MIN_FIXED
#define CB_BLOSSOM_PAR_MIN_FIXED 33  // > 32: the fixed 32-vertex layout never uses them
#endif
constexpr int BLOSSOM_PAR_MIN_FIXED = CB_BLOSSOM_PAR_MIN_FIXED;
#ifndef CB_WARP_MIN_N
#define CB_WARP_MIN_N 17
#endif
constexpr uint32_t WARP_SOLVER_MIN_N = CB_WARP_MIN_N;  // fixed-layout components from this size on: warp-synchronous solver

template <class G, int CAP = 0>
struct Blossom {
    enum Op : uint32_t { OP_EXIT = 0, OP_SET_SLACK, OP_DUAL_MIN, OP_LAB_UPDATE, OP_NEXT_TIGHT, OP_BLOSSOM_CLEAR, OP_MERGE_ROW, OP_PHASE_INIT };
    int n, n_x, NX, FL, AW;
    // Storage. CAP == 0: sized for n at run time, addressed through the pointers below (any memory). CAP > 0: sized for CAP
    // vertices at compile time, every array at a constant offset from the object itself, which then always lives in shared
    // memory: the accessors below turn into shared-memory loads with immediate offsets.
    int64_t *W_p;        // (n+1) x (n+1)
    uint8_t *ge_p;       // NX x NX x 2 : original endpoints (eu, ev) of the best edge between two nodes; eu == 0: none
    int64_t *lab_p;      // NX
    uint16_t *match_p, *slack_p, *st_p, *pa_p;
    int8_t *S_p;
    uint32_t *vis_p;
    uint16_t *ff_p;      // flower_from: NX x (n+1)
    uint16_t *fl_p;      // flower lists: (NX - n) x FL, row b - n
    uint16_t *fl_len_p;  // NX
    uint16_t *q_p;       // circular queue, capacity QC
    uint16_t *stk_p;     // explicit stack, capacity 2*NX
    uint32_t *adjm_p;    // (n+1) x AW adjacency bit rows: bit v of row u <=> W[u][v] > 0
    uint32_t *cmd_p;     // [8] command block (master -> workers)
    int qh, qt, QC;
    uint32_t vis_t;
    bool broken;
    bool par;
    bool wide;         // endpoints stored as u16

    static constexpr uint32_t c16(uint64_t x) { return (uint32_t)((x + 15) & ~uint64_t{15}); }
    static constexpr int NXC = CAP + CAP / 2 + 2, WSC = CAP + 1, FLC = CAP + 2, QCC = 2 * CAP + 4, AWC = (CAP + 32) / 32;
    static constexpr uint32_t HDR = 256;  // room for the object itself (checked below the class)
    static constexpr uint32_t OFF_W = HDR;
    static constexpr uint32_t OFF_GE = OFF_W + c16(uint64_t(WSC) * WSC * 8);
    static constexpr uint32_t OFF_LAB = OFF_GE + c16(uint64_t(NXC) * NXC * 2);
    static constexpr uint32_t OFF_MATCH = OFF_LAB + c16(uint64_t(NXC) * 8);
    static constexpr uint32_t OFF_SLACK = OFF_MATCH + c16(uint64_t(NXC) * 2);
    static constexpr uint32_t OFF_ST = OFF_SLACK + c16(uint64_t(NXC) * 2);
    static constexpr uint32_t OFF_PA = OFF_ST + c16(uint64_t(NXC) * 2);
    static constexpr uint32_t OFF_S = OFF_PA + c16(uint64_t(NXC) * 2);
    static constexpr uint32_t OFF_VIS = OFF_S + c16(uint64_t(NXC));
    static constexpr uint32_t OFF_FF = OFF_VIS + c16(uint64_t(NXC) * 4);
    static constexpr uint32_t OFF_FL = OFF_FF + c16(uint64_t(NXC) * WSC * 2);
    static constexpr uint32_t OFF_FLLEN = OFF_FL + c16(uint64_t(CAP / 2 + 2) * FLC * 2);
    static constexpr uint32_t OFF_Q = OFF_FLLEN + c16(uint64_t(NXC) * 2);
    static constexpr uint32_t OFF_STK = OFF_Q + c16(uint64_t(QCC) * 2);
    static constexpr uint32_t OFF_ADJM = OFF_STK + c16(uint64_t(2 * NXC + 4) * 2);
    static constexpr uint32_t OFF_CMD = OFF_ADJM + c16(uint64_t(WSC) * AWC * 4);
    static constexpr uint32_t FIXED_BYTES = OFF_CMD + 32;
    CB_HD uint8_t *self() const { return (uint8_t *)this; }
    CB_HD int64_t *W() const { return CAP ? (int64_t *)(self() + OFF_W) : W_p; }
    CB_HD uint8_t *ge() const { return CAP ? self() + OFF_GE : ge_p; }
    CB_HD int64_t *lab() const { return CAP ? (int64_t *)(self() + OFF_LAB) : lab_p; }
    CB_HD uint16_t *match() const { return CAP ? (uint16_t *)(self() + OFF_MATCH) : match_p; }
    CB_HD uint16_t *slack() const { return CAP ? (uint16_t *)(self() + OFF_SLACK) : slack_p; }
    CB_HD uint16_t *st() const { return CAP ? (uint16_t *)(self() + OFF_ST) : st_p; }
    CB_HD uint16_t *pa() const { return CAP ? (uint16_t *)(self() + OFF_PA) : pa_p; }
    CB_HD int8_t *S() const { return CAP ? (int8_t *)(self() + OFF_S) : S_p; }
    CB_HD uint32_t *vis() const { return CAP ? (uint32_t *)(self() + OFF_VIS) : vis_p; }
    CB_HD uint16_t *ff() const { return CAP ? (uint16_t *)(self() + OFF_FF) : ff_p; }
    CB_HD uint16_t *fl() const { return CAP ? (uint16_t *)(self() + OFF_FL) : fl_p; }
    CB_HD uint16_t *fl_len() const { return CAP ? (uint16_t *)(self() + OFF_FLLEN) : fl_len_p; }
    CB_HD uint16_t *q() const { return CAP ? (uint16_t *)(self() + OFF_Q) : q_p; }
    CB_HD uint16_t *stk() const { return CAP ? (uint16_t *)(self() + OFF_STK) : stk_p; }
    CB_HD uint32_t *adjm() const { return CAP ? (uint32_t *)(self() + OFF_ADJM) : adjm_p; }
    CB_HD uint32_t *cmd() const { return CAP ? (uint32_t *)(self() + OFF_CMD) : cmd_p; }
    // row strides / capacities
    CB_HD int ws() const { return CAP ? WSC : n + 1; }
    CB_HD int nxs() const { return CAP ? NXC : NX; }
    CB_HD int fls() const { return CAP ? FLC : FL; }
    CB_HD int aws() const { return CAP ? AWC : AW; }
    CB_HD int qcs() const { return CAP ? QCC : QC; }
    // Tells the compiler the address space of a fixed-layout object (no-op otherwise).
    CB_HD void in_shared() const {
#ifdef __CUDA_ARCH__
        if (CAP) __builtin_assume(__isShared(this));
#endif
    }

    static CB_HD uint64_t bytes_needed(int n) {
        if (CAP) return FIXED_BYTES;
        int NX = n + n / 2 + 2;
        int FL = n + 2;
        int AW = (n + 32) / 32;
        uint64_t b = 0;
        b += align16(uint64_t(n + 1) * (n + 1) * 8);
        b += align16(uint64_t(NX) * NX * 2 * (n > (int)NARROW_COMPONENT ? 2 : 1));
        b += align16(uint64_t(NX) * 8);
        b += align16(uint64_t(NX) * 2) * 4;
        b += align16(uint64_t(NX));
        b += align16(uint64_t(NX) * 4);
        b += align16(uint64_t(NX) * (n + 1) * 2);
        b += align16(uint64_t(NX - n) * FL * 2);
        b += align16(uint64_t(NX) * 2);
        b += align16(uint64_t(2 * n + 4) * 2);
        b += align16(uint64_t(2 * NX + 4) * 2);
        b += align16(uint64_t(n + 1) * AW * 4);
        b += 32;
        b += align16(sizeof(Blossom));  // the solver state itself lives at the front of its scratch memory
        return b;
    }
    // The object is placed in the scratch memory it manages (shared memory for all but giant components): its fields
    // are read on every step of the sequential bookkeeping, and a copy in a lane's local memory would cost one cache
    // line per 4 bytes.
    static CB_HD Blossom *place(uint8_t *mem) { return (Blossom *)mem; }
    CB_HD void bind(uint8_t *mem, int n_, int lanes) {
        n = n_;
        NX = n + n / 2 + 2;
        FL = n + 2;
        QC = 2 * n + 4;
        AW = (n + 32) / 32;
        wide = !CAP && n > (int)NARROW_COMPONENT;
        broken = false;
        par = PAR_CLASS && lanes > 1 && n >= (CAP ? BLOSSOM_PAR_MIN_FIXED : BLOSSOM_PAR_MIN);
        n_x = n;
        if (CAP) return;
        uint8_t *p = mem + align16(sizeof(Blossom));
        W_p = (int64_t *)p; p += align16(uint64_t(n + 1) * (n + 1) * 8);
        ge_p = p; p += align16(uint64_t(NX) * NX * 2 * (wide ? 2 : 1));
        lab_p = (int64_t *)p; p += align16(uint64_t(NX) * 8);
        match_p = (uint16_t *)p; p += align16(uint64_t(NX) * 2);
        slack_p = (uint16_t *)p; p += align16(uint64_t(NX) * 2);
        st_p = (uint16_t *)p; p += align16(uint64_t(NX) * 2);
        pa_p = (uint16_t *)p; p += align16(uint64_t(NX) * 2);
        S_p = (int8_t *)p; p += align16(uint64_t(NX));
        vis_p = (uint32_t *)p; p += align16(uint64_t(NX) * 4);
        ff_p = (uint16_t *)p; p += align16(uint64_t(NX) * (n + 1) * 2);
        fl_p = (uint16_t *)p; p += align16(uint64_t(NX - n) * FL * 2);
        fl_len_p = (uint16_t *)p; p += align16(uint64_t(NX) * 2);
        q_p = (uint16_t *)p; p += align16(uint64_t(QC) * 2);
        stk_p = (uint16_t *)p; p += align16(uint64_t(2 * NX + 4) * 2);
        adjm_p = (uint32_t *)p; p += align16(uint64_t(n + 1) * AW * 4);
        cmd_p = (uint32_t *)p;
    }
    CB_HD int64_t &w(int u, int v) { return W()[u * ws() + v]; }
    CB_HD bool is_wide() const { return !CAP && wide; }
    // The run-time layout always has the warp helpers; of the fixed layouts only the 32-vertex one (and only from
    // BLOSSOM_PAR_MIN_FIXED events on): its O(n) loops are what a round waits for.
    static constexpr bool PAR_CLASS = CAP == 0 || (CAP == 32 && BLOSSOM_PAR_MIN_FIXED <= 32);
    // -DCB_WARP_SOLVER: fixed-layout components of at least CB_WARP_MIN_N events are solved by the warp-synchronous
    // solver (solve_w) on 32-lane warps. Off by default -- measured at d25 r100 (profiles/r02_kernel_experiments.txt):
    // the 17..32-event class gets 25-33 % faster, the <= 8-event class 8 % slower, and the busier warps slow every
    // other phase of the blocks sharing the SM by 4-6 %: 174.9 (all sizes) / 176.4 (>= 17) / 187.4 (>= 9) ms vs 171.1.
    // The run-time layout (components of more than 32 events, 0.2 per shot but ~1 M cycles each on one lane) always uses
    // it unless -DCB_NO_WARP_SOLVER_BIG: there the neighbour scan is what the time goes into, and nothing else of the
    // block runs meanwhile.
#ifdef CB_WARP_SOLVER
    static constexpr bool WARP_SOLVER_FIXED = true;
#else
    static constexpr bool WARP_SOLVER_FIXED = false;
#endif
#ifdef CB_NO_WARP_SOLVER_BIG
    static constexpr bool WARP_SOLVER_BIG = false;
#else
    static constexpr bool WARP_SOLVER_BIG = true;
#endif
    static constexpr bool WARP_SOLVER = CAP > 0 ? WARP_SOLVER_FIXED : WARP_SOLVER_BIG;
    CB_HD bool is_par() const { return PAR_CLASS && par; }
    CB_HD int geu(int a, int b) const { return is_wide() ? ((const uint16_t *)ge())[(a * nxs() + b) * 2] : ge()[(a * nxs() + b) * 2]; }
    CB_HD int gev(int a, int b) const { return is_wide() ? ((const uint16_t *)ge())[(a * nxs() + b) * 2 + 1] : ge()[(a * nxs() + b) * 2 + 1]; }
    CB_HD bool has(int a, int b) const { return geu(a, b) != 0; }
    CB_HD void set_ge(int a, int b, int eu, int ev) {
        if (is_wide()) {
            ((uint16_t *)ge())[(a * nxs() + b) * 2] = (uint16_t)eu;
            ((uint16_t *)ge())[(a * nxs() + b) * 2 + 1] = (uint16_t)ev;
        } else {
            ((uint16_t *)ge())[a * nxs() + b] = (uint16_t)(eu | (ev << 8));  // both bytes with one store
        }
    }
    CB_HD int64_t edist(int a, int b) {  // slack of the best edge between nodes a and b
        int eu, ev;
        if (is_wide()) {
            eu = geu(a, b);
            ev = gev(a, b);
        } else {
            const uint32_t pair = ((const uint16_t *)ge())[a * nxs() + b];
            eu = (int)(pair & 0xFF);
            ev = (int)(pair >> 8);
        }
        return lab()[eu] + lab()[ev] - 2 * W()[eu * ws() + ev];
    }
    CB_HD uint16_t *flower(int b) { return fl() + (b - n) * fls(); }
    CB_HD uint16_t &ffrom(int b, int x) { return ff()[b * ws() + x]; }

    // ---- data-parallel kernels: executed by every participating lane (`lane`, `lanes` = 0, 1 when running alone) ----
    CB_HD void k_init(int lane, int lanes) {
        CB_LOOP
        for (int u = lane; u < NX; u += lanes) {
            match()[u] = 0;
            st()[u] = (uint16_t)(u <= n ? u : 0);
            fl_len()[u] = 0;
            vis()[u] = 0;
            slack()[u] = 0;
            pa()[u] = 0;
            S()[u] = -1;
            lab()[u] = 0;
        }
        CB_LOOP
        for (int u = 1 + lane; u <= n; u += lanes) {
            CB_LOOP
            for (int k = 0; k < aws(); k++) {
                uint32_t acc = 0;  // adjacency bits of this 32-vertex chunk, stored once
                const int v0 = k ? 32 * k : 1, v1 = 32 * k + 31 < n ? 32 * k + 31 : n;
                CB_LOOP
                for (int v = v0; v <= v1; ++v) {
                    ffrom(u, v) = (uint16_t)(u == v ? u : 0);
                    int64_t x = W()[u * ws() + v];
                    if (x > 0) {
                        set_ge(u, v, u, v);
                        acc |= 1u << (v & 31);
                    } else {
                        set_ge(u, v, 0, 0);
                    }
                }
                adjm()[u * aws() + k] = acc;
            }
        }
    }
    // Starting labels: lab()[u] = max_v W()[u][v], feasible because lab()[u] + lab()[v] >= 2 W()[u][v] for every edge.
    CB_HD void k_row_max_lab(int lane, int lanes) {
        CB_LOOP
        for (int u = 1 + lane; u <= n; u += lanes) {
            int64_t m = 0;
            CB_LOOP
            for (int v = 1; v <= n; ++v) {
                int64_t x = W()[u * ws() + v];
                if (x > m) m = x;
            }
            lab()[u] = m;
        }
    }
    // (distance << 16 | u) of the best eligible vertex for slack()[x]; ~0 when none. Lowest u wins ties.
    CB_HD uint64_t k_set_slack(int lane, int lanes, int x) {
        uint64_t best = ~uint64_t{0};
        CB_LOOP
        for (int u = 1 + lane; u <= n; u += lanes)
            if (has(u, x) && st()[u] != x && S()[st()[u]] == 0) {
                uint64_t key = ((uint64_t)edist(u, x) << 16) | (uint64_t)u;
                if (key < best) best = key;
            }
        return best;
    }
    CB_HD uint64_t k_dual_min(int lane, int lanes, int nx) {
        uint64_t d = (uint64_t)(INT64_MAX / 4);
        CB_LOOP
        for (int b = n + 1 + lane; b <= nx; b += lanes)
            if (st()[b] == b && S()[b] == 1) {
                uint64_t v = (uint64_t)(lab()[b] / 2);
                if (v < d) d = v;
            }
        CB_LOOP
        for (int x = 1 + lane; x <= nx; x += lanes)
            if (st()[x] == x && slack()[x]) {
                int64_t sd = edist(slack()[x], x);
                if (S()[x] == -1) {
                    if ((uint64_t)sd < d) d = (uint64_t)sd;
                } else if (S()[x] == 0) {
                    if ((uint64_t)(sd / 2) < d) d = (uint64_t)(sd / 2);
                }
            }
        return d;
    }
    CB_HD void k_lab_update(int lane, int lanes, int nx, int64_t d) {
        CB_LOOP
        for (int u = 1 + lane; u <= n; u += lanes) {
            if (S()[st()[u]] == 0) lab()[u] -= d;
            else if (S()[st()[u]] == 1) lab()[u] += d;
        }
        CB_LOOP
        for (int b = n + 1 + lane; b <= nx; b += lanes)
            if (st()[b] == b) {
                if (S()[b] == 0) lab()[b] += d * 2;
                else if (S()[b] == 1) lab()[b] -= d * 2;
            }
    }
    CB_HD bool tight_candidate(int x) {
        return st()[x] == x && slack()[x] && st()[slack()[x]] != x && edist(slack()[x], x) == 0;
    }
    CB_HD void k_blossom_clear(int lane, int lanes, int b, int nx) {
        CB_LOOP
        for (int x = 1 + lane; x <= nx; x += lanes) {
            set_ge(b, x, 0, 0);
            set_ge(x, b, 0, 0);
        }
        CB_LOOP
        for (int x = 1 + lane; x <= n; x += lanes) ffrom(b, x) = 0;
    }
    CB_HD void k_merge_row(int lane, int lanes, int b, int xs, int nx) {
        CB_LOOP
        for (int x = 1 + lane; x <= nx; x += lanes)
            if (has(xs, x) && (!has(b, x) || edist(xs, x) < edist(b, x))) {
                set_ge(b, x, geu(xs, x), gev(xs, x));
                set_ge(x, b, geu(x, xs), gev(x, xs));
            }
        CB_LOOP
        for (int x = 1 + lane; x <= n; x += lanes)
            if (ffrom(xs, x)) ffrom(b, x) = (uint16_t)xs;
    }
    CB_HD void k_phase_init(int lane, int lanes, int nx) {
        CB_LOOP
        for (int x = 1 + lane; x <= nx; x += lanes) {
            S()[x] = -1;
            slack()[x] = 0;
        }
    }

    // ---- warp dispatch: every lane runs the same helper; results are reduced with warp primitives ----
    CB_HDN uint64_t dispatch(const G &g, uint32_t op, uint32_t a0, uint32_t a1, uint32_t a2) {
        in_shared();
        const int lane = g.lane(), lanes = g.lanes();
        uint64_t r = 0;
        switch (op) {
            case OP_SET_SLACK:
                r = g.min64(k_set_slack(lane, lanes, (int)a0));
                break;
            case OP_DUAL_MIN:
                r = g.min64(k_dual_min(lane, lanes, (int)a0));
                break;
            case OP_LAB_UPDATE:
                k_lab_update(lane, lanes, (int)a0, (int64_t)(((uint64_t)a2 << 32) | a1));
                break;
            case OP_NEXT_TIGHT: {
                // smallest x >= a0 that is a tight candidate, or a1 + 1
                int nx = (int)a1;
                r = (uint64_t)(nx + 1);
                CB_LOOP
                for (int base = (int)a0; base <= nx; base += lanes) {
                    int x = base + lane;
                    uint32_t hits = g.ballot(x <= nx && tight_candidate(x));
                    if (hits) {
                        r = (uint64_t)(base + (int)ctz32(hits));
                        break;
                    }
                }
                break;
            }
            case OP_BLOSSOM_CLEAR:
                k_blossom_clear(lane, lanes, (int)a0, (int)a1);
                break;
            case OP_MERGE_ROW:
                k_merge_row(lane, lanes, (int)a0, (int)a1, (int)a2);
                break;
            case OP_PHASE_INIT:
                k_phase_init(lane, lanes, (int)a0);
                break;
            default:
                break;
        }
        return r;
    }
    // master side: announce, run together, wait for everyone's writes
    CB_HD uint64_t call(const G &g, uint32_t op, uint32_t a0 = 0, uint32_t a1 = 0, uint32_t a2 = 0) {
        cmd()[0] = op;
        cmd()[1] = a0;
        cmd()[2] = a1;
        cmd()[3] = a2;
        g.wsync();
        uint64_t r = dispatch(g, op, a0, a1, a2);
        g.wsync();
        return r;
    }
    // lanes 1..31
    CB_HDN void worker_loop(const G &g) {
        in_shared();
        CB_LOOP
        while (true) {
            g.wsync();
            uint32_t op = cmd()[0];
            if (op == OP_EXIT) break;
            dispatch(g, op, cmd()[1], cmd()[2], cmd()[3]);
            g.wsync();
        }
    }
    CB_HD void release_workers(const G &g) {
        if (is_par()) {
            cmd()[0] = OP_EXIT;
            g.wsync();
        }
    }

    // ---- sequential bookkeeping (master) ----
    CB_HD void update_slack(int u, int x) {
        if (!slack()[x] || edist(u, x) < edist(slack()[x], x)) slack()[x] = (uint16_t)u;
    }
    CB_HDN void set_slack(const G &g, int x) {
        in_shared();
        uint64_t best = is_par() ? call(g, OP_SET_SLACK, (uint32_t)x) : k_set_slack(0, 1, x);
        slack()[x] = best == ~uint64_t{0} ? 0 : (uint16_t)(best & 0xFFFF);
    }
    CB_HD void q_put(int v) {
        q()[qt] = (uint16_t)v;
        qt = qt + 1 == qcs() ? 0 : qt + 1;
        if (qt == qh) broken = true;
    }
    // depth-first, children in order
    CB_HDN void q_push(int x) {
        in_shared();
        int sp = 0;
        stk()[sp++] = (uint16_t)x;
        CB_LOOP
        while (sp) {
            int y = stk()[--sp];
            if (y <= n) {
                q_put(y);
                continue;
            }
            uint16_t *f = flower(y);
            CB_LOOP
            for (int i = fl_len()[y] - 1; i >= 0; --i) stk()[sp++] = f[i];
        }
    }
    CB_HDN void set_st(int x, int b) {
        in_shared();
        int sp = 0;
        stk()[sp++] = (uint16_t)x;
        CB_LOOP
        while (sp) {
            int y = stk()[--sp];
            st()[y] = (uint16_t)b;
            if (y <= n) continue;
            uint16_t *f = flower(y);
            CB_LOOP
            for (int i = fl_len()[y] - 1; i >= 0; --i) stk()[sp++] = f[i];
        }
    }
    CB_HD void reverse_tail(uint16_t *f, int len) {  // reverse f[1..len)
        CB_LOOP
        for (int a = 1, b = len - 1; a < b; ++a, --b) {
            uint16_t t = f[a];
            f[a] = f[b];
            f[b] = t;
        }
    }
    CB_HDN int get_pr(int b, int xr) {
        in_shared();
        uint16_t *f = flower(b);
        int len = fl_len()[b];
        int pr = 0;
        CB_LOOP
        while (pr < len && f[pr] != xr) pr++;
        if (pr % 2 == 1) {
            reverse_tail(f, len);
            return len - pr;
        }
        return pr;
    }
    CB_HD void rotate_left(uint16_t *f, int len, int k) {  // three reversals
        if (k <= 0 || k >= len) return;
        CB_LOOP
        for (int a = 0, b = k - 1; a < b; ++a, --b) { uint16_t t = f[a]; f[a] = f[b]; f[b] = t; }
        CB_LOOP
        for (int a = k, b = len - 1; a < b; ++a, --b) { uint16_t t = f[a]; f[a] = f[b]; f[b] = t; }
        CB_LOOP
        for (int a = 0, b = len - 1; a < b; ++a, --b) { uint16_t t = f[a]; f[a] = f[b]; f[b] = t; }
    }
    // set_match(u, v) with an explicit stack of (u, v) pairs.
    CB_HDN void set_match(int u0, int v0) {
        in_shared();
        int sp = 0;
        stk()[sp++] = (uint16_t)u0;
        stk()[sp++] = (uint16_t)v0;
        CB_LOOP
        while (sp) {
            int v = stk()[--sp];
            int u = stk()[--sp];
            match()[u] = (uint16_t)gev(u, v);
            if (u <= n) continue;
            int eu = geu(u, v);
            int xr = ffrom(u, eu);
            int pr = get_pr(u, xr);
            uint16_t *f = flower(u);
            CB_LOOP
            for (int i = 0; i < pr; ++i) {
                stk()[sp++] = f[i];
                stk()[sp++] = f[i ^ 1];
            }
            stk()[sp++] = (uint16_t)xr;
            stk()[sp++] = (uint16_t)v;
            rotate_left(f, fl_len()[u], pr);
        }
    }
    CB_HDN void augment(int u, int v) {
        in_shared();
        CB_LOOP
        for (;;) {
            int xnv = st()[match()[u]];
            set_match(u, v);
            if (!xnv) return;
            set_match(xnv, st()[pa()[xnv]]);
            u = st()[pa()[xnv]];
            v = xnv;
        }
    }
    CB_HDN int get_lca(int u, int v) {
        in_shared();
        CB_LOOP
        for (++vis_t; u || v;) {
            if (u != 0) {
                if (vis()[u] == vis_t) return u;
                vis()[u] = vis_t;
                u = st()[match()[u]];
                if (u) u = st()[pa()[u]];
            }
            int t = u; u = v; v = t;
        }
        return 0;
    }
    CB_HDN void add_blossom(const G &g, int u, int lca, int v) {
        in_shared();
        int b = n + 1;
        CB_LOOP
        while (b <= n_x && st()[b]) ++b;
        if (b > n_x) ++n_x;
        if (b >= NX) { broken = true; return; }
        lab()[b] = 0;
        S()[b] = 0;
        match()[b] = match()[lca];
        uint16_t *f = flower(b);
        int len = 0;
        f[len++] = (uint16_t)lca;
        CB_LOOP
        for (int x = u, y; x != lca; x = st()[pa()[y]]) {
            if (len + 2 > fls()) { broken = true; return; }
            f[len++] = (uint16_t)x;
            f[len++] = (uint16_t)(y = st()[match()[x]]);
            q_push(y);
        }
        reverse_tail(f, len);
        CB_LOOP
        for (int x = v, y; x != lca; x = st()[pa()[y]]) {
            if (len + 2 > fls()) { broken = true; return; }
            f[len++] = (uint16_t)x;
            f[len++] = (uint16_t)(y = st()[match()[x]]);
            q_push(y);
        }
        fl_len()[b] = (uint16_t)len;
        set_st(b, b);
        if (is_par()) call(g, OP_BLOSSOM_CLEAR, (uint32_t)b, (uint32_t)n_x);
        else k_blossom_clear(0, 1, b, n_x);
        CB_LOOP
        for (int i = 0; i < len; ++i) {
            if (is_par()) call(g, OP_MERGE_ROW, (uint32_t)b, (uint32_t)f[i], (uint32_t)n_x);
            else k_merge_row(0, 1, b, f[i], n_x);
        }
        set_slack(g, b);
    }
    CB_HDN void expand_blossom(const G &g, int b) {
        in_shared();
        uint16_t *f = flower(b);
        int len = fl_len()[b];
        CB_LOOP
        for (int i = 0; i < len; ++i) set_st(f[i], f[i]);
        int xr = ffrom(b, geu(b, pa()[b]));
        int pr = get_pr(b, xr);
        CB_LOOP
        for (int i = 0; i < pr; i += 2) {
            int xs = f[i], xns = f[i + 1];
            pa()[xs] = (uint16_t)geu(xns, xs);
            S()[xs] = 1;
            S()[xns] = 0;
            slack()[xs] = 0;
            set_slack(g, xns);
            q_push(xns);
        }
        S()[xr] = 1;
        pa()[xr] = pa()[b];
        CB_LOOP
        for (int i = pr + 1; i < len; ++i) {
            int xs = f[i];
            S()[xs] = -1;
            set_slack(g, xs);
        }
        st()[b] = 0;
    }
    // e = best edge between nodes a (row) and b: endpoints (eu in a, ev in b)
    CB_HDN bool on_found_edge(const G &g, int eu, int ev) {
        in_shared();
        int u = st()[eu], v = st()[ev];
        if (S()[v] == -1) {
            pa()[v] = (uint16_t)eu;
            S()[v] = 1;
            int nu = st()[match()[v]];
            slack()[v] = slack()[nu] = 0;
            S()[nu] = 0;
            q_push(nu);
        } else if (S()[v] == 0) {
            int lca = get_lca(u, v);
            if (!lca) {
                augment(u, v);
                augment(v, u);
                return true;
            }
            add_blossom(g, u, lca, v);
        }
        return false;
    }
    CB_HDN bool matching(const G &g) {
        in_shared();
        const int64_t INF = INT64_MAX / 4;
        if (is_par()) call(g, OP_PHASE_INIT, (uint32_t)n_x);
        else k_phase_init(0, 1, n_x);
        qh = qt = 0;
        CB_LOOP
        for (int x = 1; x <= n_x; ++x)
            if (st()[x] == x && !match()[x]) {
                pa()[x] = 0;
                S()[x] = 0;
                q_push(x);
            }
        if (qh == qt) return false;
        CB_LOOP
        for (;;) {
            CB_LOOP
            while (qh != qt) {
                if (broken) return false;
                int u = q()[qh];
                qh = qh + 1 == qcs() ? 0 : qh + 1;
                if (S()[st()[u]] == 1) continue;
                // neighbours of u in ascending order (adjacency bit rows; W never changes during a solve)
                CB_LOOP
                for (int k = 0; k < aws(); ++k) {
                    uint32_t bits = adjm()[u * aws() + k];
                    CB_LOOP
                    while (bits) {
                        int v = 32 * k + (int)ctz32(bits);
                        bits &= bits - 1;
                        if (st()[u] != st()[v]) {
                            if (lab()[u] + lab()[v] - 2 * W()[u * ws() + v] == 0) {
                                if (on_found_edge(g, u, v)) return true;
                            } else {
                                update_slack(u, st()[v]);
                            }
                        }
                    }
                }
            }
            if (broken) return false;
            int64_t d = (int64_t)(is_par() ? call(g, OP_DUAL_MIN, (uint32_t)n_x) : k_dual_min(0, 1, n_x));
            if (d >= INF) return false;
            if (is_par()) call(g, OP_LAB_UPDATE, (uint32_t)n_x, (uint32_t)((uint64_t)d & 0xFFFFFFFFu), (uint32_t)((uint64_t)d >> 32));
            else k_lab_update(0, 1, n_x, d);
            qh = qt = 0;
            if (is_par()) {
                CB_LOOP
                for (int x = (int)call(g, OP_NEXT_TIGHT, 1, (uint32_t)n_x); x <= n_x; x = (int)call(g, OP_NEXT_TIGHT, (uint32_t)(x + 1), (uint32_t)n_x))
                    if (tight_candidate(x) && on_found_edge(g, geu(slack()[x], x), gev(slack()[x], x))) return true;
            } else {
                CB_LOOP
                for (int x = 1; x <= n_x; ++x)
                    if (tight_candidate(x))
                        if (on_found_edge(g, geu(slack()[x], x), gev(slack()[x], x))) return true;
            }
            CB_LOOP
            for (int b = n + 1; b <= n_x; ++b)
                if (st()[b] == b && S()[b] == 1 && lab()[b] == 0) expand_blossom(g, b);
        }
    }
    // ---- warp-synchronous solver (fixed layouts) ------------------------------------------------------------------
    // Every lane of the warp runs these functions with the same control flow. The O(n) loops of the method (slack
    // search, dual step, label update, neighbour scan, row merges) are spread over the lanes -- vertex / node
    // x = 1 + lane (+ 32) -- and reduced with warp primitives; the sequential bookkeeping (queue, flowers, augmenting)
    // is done by lane 0 alone between two wsync()s. All solver state lives in shared memory, so after a wsync() every
    // lane reads what lane 0 wrote; results lane 0 has to hand over travel through the command block. The steps and
    // their order are those of the sequential solver above (lowest index first wherever it looks for "the first"), so
    // matching and labels come out identical (tests/sim/warp_blossom_check.cc runs both side by side).
    CB_HD void set_slack_w(const G &g, int x) {
        const uint64_t best = g.min64(k_set_slack(g.lane(), g.lanes(), x));
        if (g.lane() == 0) slack()[x] = best == ~uint64_t{0} ? 0 : (uint16_t)(best & 0xFFFF);
        g.wsync();
    }
    CB_HDN void add_blossom_w(const G &g, int u, int lca, int v) {
        in_shared();
        if (g.lane() == 0) {
            int b = n + 1;
            CB_LOOP
            while (b <= n_x && st()[b]) ++b;
            if (b > n_x) ++n_x;
            if (b >= NX) {
                broken = true;
            } else {
                lab()[b] = 0;
                S()[b] = 0;
                match()[b] = match()[lca];
                uint16_t *f = flower(b);
                int len = 0;
                f[len++] = (uint16_t)lca;
                CB_LOOP
                for (int x = u, y; x != lca && !broken; x = st()[pa()[y]]) {
                    if (len + 2 > fls()) { broken = true; break; }
                    f[len++] = (uint16_t)x;
                    f[len++] = (uint16_t)(y = st()[match()[x]]);
                    q_push(y);
                }
                reverse_tail(f, len);
                CB_LOOP
                for (int x = v, y; x != lca && !broken; x = st()[pa()[y]]) {
                    if (len + 2 > fls()) { broken = true; break; }
                    f[len++] = (uint16_t)x;
                    f[len++] = (uint16_t)(y = st()[match()[x]]);
                    q_push(y);
                }
                if (!broken) {
                    fl_len()[b] = (uint16_t)len;
                    set_st(b, b);
                }
            }
            cmd()[4] = broken ? 0u : (uint32_t)b;
        }
        g.wsync();
        const int b = (int)cmd()[4];
        if (!b) return;
        const int nx = n_x, len = fl_len()[b];
        const uint16_t *f = flower(b);
        // a lane owns the same x in the clearing pass and in every row merge: no barrier in between
        k_blossom_clear(g.lane(), g.lanes(), b, nx);
        CB_LOOP
        for (int i = 0; i < len; ++i) k_merge_row(g.lane(), g.lanes(), b, f[i], nx);
        g.wsync();
        set_slack_w(g, b);
    }
    CB_HDN void expand_blossom_w(const G &g, int b) {
        in_shared();
        if (g.lane() == 0) {
            uint16_t *f0 = flower(b);
            const int len0 = fl_len()[b];
            CB_LOOP
            for (int i = 0; i < len0; ++i) set_st(f0[i], f0[i]);
            const int xr0 = ffrom(b, geu(b, pa()[b]));
            cmd()[4] = (uint32_t)xr0;
            cmd()[5] = (uint32_t)get_pr(b, xr0);
        }
        g.wsync();
        const int xr = (int)cmd()[4], pr = (int)cmd()[5];
        const uint16_t *f = flower(b);
        const int len = fl_len()[b];
        CB_LOOP
        for (int i = 0; i < pr; i += 2) {
            const int xs = f[i], xns = f[i + 1];
            if (g.lane() == 0) {
                pa()[xs] = (uint16_t)geu(xns, xs);
                S()[xs] = 1;
                S()[xns] = 0;
                slack()[xs] = 0;
            }
            g.wsync();
            set_slack_w(g, xns);
            if (g.lane() == 0) q_push(xns);
        }
        if (g.lane() == 0) {
            S()[xr] = 1;
            pa()[xr] = pa()[b];
        }
        CB_LOOP
        for (int i = pr + 1; i < len; ++i) {
            const int xs = f[i];
            if (g.lane() == 0) S()[xs] = -1;
            g.wsync();
            set_slack_w(g, xs);
        }
        if (g.lane() == 0) st()[b] = 0;
        g.wsync();
    }
    CB_HDN bool on_found_edge_w(const G &g, int eu, int ev) {
        in_shared();
        g.wsync();  // every lane is done with the state (and the command block) lane 0 is about to change
        if (g.lane() == 0) {
            const int u = st()[eu], v = st()[ev];
            uint32_t code = 0;
            if (S()[v] == -1) {
                pa()[v] = (uint16_t)eu;
                S()[v] = 1;
                const int nu = st()[match()[v]];
                slack()[v] = slack()[nu] = 0;
                S()[nu] = 0;
                q_push(nu);
            } else if (S()[v] == 0) {
                const int lca = get_lca(u, v);
                if (!lca) {
                    augment(u, v);
                    augment(v, u);
                    code = 1;
                } else {
                    code = 2;
                    cmd()[1] = (uint32_t)u;
                    cmd()[2] = (uint32_t)lca;
                    cmd()[3] = (uint32_t)v;
                }
            }
            cmd()[0] = code;
        }
        g.wsync();
        const uint32_t code = cmd()[0];
        if (code == 1) return true;
        if (code == 2) add_blossom_w(g, (int)cmd()[1], (int)cmd()[2], (int)cmd()[3]);
        return false;
    }
    CB_HDN bool matching_w(const G &g) {
        in_shared();
        const int lane = g.lane(), lanes = g.lanes();
        const int64_t INF = INT64_MAX / 4;
        k_phase_init(lane, lanes, n_x);
        if (lane == 0) qh = qt = 0;
        g.wsync();
        // free top-level nodes become roots, ascending
        CB_LOOP
        for (int base = 1; base <= n_x; base += lanes) {
            const int x = base + lane;
            uint32_t hits = g.ballot(x <= n_x && st()[x] == x && !match()[x]);
            if (lane == 0) {
                CB_LOOP
                while (hits) {
                    const int r = base + (int)ctz32(hits);
                    hits &= hits - 1;
                    pa()[r] = 0;
                    S()[r] = 0;
                    q_push(r);
                }
            }
        }
        g.wsync();
        if (qt == 0) return false;
        int head = 0;  // queue head, the same in every lane (the member copy only serves q_put's overflow check)
        CB_LOOP
        for (;;) {
            CB_LOOP
            while (true) {
                // (a wsync precedes every pass through here)
                if (broken) return false;
                if (head == qt) break;
                const int u = q()[head];
                head = head + 1 == qcs() ? 0 : head + 1;
                if (S()[st()[u]] == 1) continue;
                g.wsync();  // everyone has read the queue state before lane 0 changes it
                if (lane == 0) qh = head;
                // neighbours of u in ascending order: the lanes test 32 of them at a time; everything below the first
                // tight edge only updates slacks (order-free), the tight edge itself is handled by the warp, and the
                // rest is tested again afterwards -- handling an edge changes S / st / slack.
                CB_LOOP
                for (int k = 0; k < aws(); ++k) {
                    uint32_t pending = adjm()[u * aws() + k];
                    CB_LOOP
                    while (pending) {
                        const int v = 32 * k + lane;
                        bool cand = false;
                        int sv = 0;
                        int64_t dist = 1;
                        if ((pending >> lane) & 1u) {
                            sv = st()[v];
                            cand = st()[u] != sv;
                            if (cand) dist = lab()[u] + lab()[v] - 2 * W()[u * ws() + v];
                        }
                        const uint32_t tight = g.ballot(cand && dist == 0);
                        const uint32_t below = tight ? ((1u << ctz32(tight)) - 1u) : 0xFFFFFFFFu;
                        if (cand && dist != 0 && ((below >> lane) & 1u)) update_slack(u, sv);
                        g.wsync();
                        if (!tight) break;
                        const int fbit = (int)ctz32(tight);
                        if (on_found_edge_w(g, u, 32 * k + fbit)) return true;
                        pending &= ~(below | (1u << fbit));
                    }
                }
            }
            if (broken) return false;
            const int64_t d = (int64_t)g.min64(k_dual_min(lane, lanes, n_x));
            if (d >= INF) return false;
            g.wsync();  // every lane has read the labels the minimum came from
            k_lab_update(lane, lanes, n_x, d);
            head = 0;
            if (lane == 0) qh = qt = 0;
            g.wsync();
            CB_LOOP
            for (int x0 = 1; x0 <= n_x;) {
                int r = n_x + 1;
                CB_LOOP
                for (int base = x0; base <= n_x; base += lanes) {
                    const int x = base + lane;
                    const uint32_t hits = g.ballot(x <= n_x && tight_candidate(x));
                    if (hits) {
                        r = base + (int)ctz32(hits);
                        break;
                    }
                }
                if (r > n_x) break;
                const int su = slack()[r];
                if (on_found_edge_w(g, geu(su, r), gev(su, r))) return true;
                x0 = r + 1;
            }
            CB_LOOP
            for (int b = n + 1; b <= n_x; ++b)
                if (st()[b] == b && S()[b] == 1 && lab()[b] == 0) expand_blossom_w(g, b);
        }
    }
    // All lanes, after prepare(). Returns the number of matched pairs (the same value in every lane).
    CB_HDN int solve_w(const G &g) {
        in_shared();
        const int lane = g.lane();
        if (lane == 0) {
            n_x = n;
            vis_t = 0;
        }
        g.wsync();
        int cnt = 0;
        // greedy start, first pass (see solve()): the first free neighbour joined by a mutual-best edge
        CB_LOOP
        for (int u = 1; u <= n; ++u) {
            if (match()[u] || lab()[u] <= 0) continue;
            CB_LOOP
            for (int k = 0; k < aws(); ++k) {
                const int v = 32 * k + lane;
                const bool ok = ((adjm()[u * aws() + k] >> lane) & 1u) && !match()[v] && W()[u * ws() + v] == lab()[u] && lab()[v] == lab()[u];
                const uint32_t hits = g.ballot(ok);
                if (hits) {
                    if (lane == 0) {
                        const int bv = 32 * k + (int)ctz32(hits);
                        match()[u] = (uint16_t)bv;
                        match()[bv] = (uint16_t)u;
                    }
                    ++cnt;
                    break;
                }
            }
            g.wsync();
        }
        // second pass: lowest feasible label, first maximum wins
        CB_LOOP
        for (int u = 1; u <= n; ++u) {
            if (match()[u] || lab()[u] <= 0) continue;
            bool any = false;
            int64_t best = 0;
            int bv = 0;
            CB_LOOP
            for (int k = 0; k < aws(); ++k) {
                const int v = 32 * k + lane;
                const bool ok = ((adjm()[u * aws() + k] >> lane) & 1u) && v != u;
                const int64_t t = ok ? 2 * W()[u * ws() + v] - lab()[v] : INT64_MIN;
                // maximum over the lanes (order-preserving map to unsigned), then the lowest lane that holds it
                const uint64_t key = (uint64_t)t ^ (uint64_t{1} << 63);
                const uint64_t kmax = ~g.min64(~key);
                const uint32_t hits = g.ballot(ok && key == kmax);
                if (hits) {
                    const int64_t tm = (int64_t)(kmax ^ (uint64_t{1} << 63));
                    if (!any || tm > best) {
                        any = true;
                        best = tm;
                        bv = 32 * k + (int)ctz32(hits);
                    }
                }
            }
            if (!any) continue;
            const bool free_end = !match()[bv];
            g.wsync();  // every lane has read labels / partners before lane 0 changes them
            if (lane == 0) {
                lab()[u] = best;
                if (free_end) {
                    match()[u] = (uint16_t)bv;
                    match()[bv] = (uint16_t)u;
                }
            }
            if (free_end) ++cnt;
            g.wsync();
        }
        CB_LOOP
        while (matching_w(g)) ++cnt;
        return cnt;
    }

    // W must be filled (symmetric, 0 = absent). Master only. Returns the number of matched pairs; match()[1..n] = partners.
    // Every lane of the warp: per-vertex state, edge endpoints, adjacency rows and the starting labels (all O(n^2), data
    // parallel). W must be filled. Ends with the lanes synchronised.
    CB_HD void prepare(const G &g) {
        in_shared();
        k_init(g.lane(), g.lanes());
        g.wsync();  // k_init clears lab[] with a different lane-to-vertex mapping
        k_row_max_lab(g.lane(), g.lanes());
        g.wsync();
    }
    // Master only, after prepare().
    CB_HDN int solve(const G &g) {
        in_shared();
        n_x = n;
        vis_t = 0;
        // Greedy start on the edges that are tight under the row-maximum labels (mutual best partners), lowest indices
        // first: a valid primal-dual state that leaves only the remaining vertices to the augmentation phases.
        int cnt = 0;
        CB_LOOP
        for (int u = 1; u <= n; ++u) {
            if (match()[u] || lab()[u] <= 0) continue;
            bool done = false;
            CB_LOOP
            for (int k = 0; k < aws() && !done; ++k) {
                uint32_t bits = adjm()[u * aws() + k];
                CB_LOOP
                while (bits) {
                    int v = 32 * k + (int)ctz32(bits);
                    bits &= bits - 1;
                    if (!match()[v] && W()[u * ws() + v] == lab()[u] && lab()[v] == lab()[u]) {
                        match()[u] = (uint16_t)v;
                        match()[v] = (uint16_t)u;
                        ++cnt;
                        done = true;
                        break;
                    }
                }
            }
        }
        // Second pass: every vertex that is still free takes the lowest label its edges allow,
        // lab[u] = max_v (2 W(u, v) - lab[v]) -- what the first dual steps of the phases would do one O(n)-scan step at a
        // time (measured at d25 r100: 3.2 -> 1.3 dual steps and 1.2 -> 0.6 augmentation phases per component). The edge
        // that defines it (lowest v on ties) is tight afterwards; if its other end is free as well the pair is matched.
        CB_LOOP
        for (int u = 1; u <= n; ++u) {
            if (match()[u] || lab()[u] <= 0) continue;
            bool any = false;
            int64_t best = 0;
            int bv = 0;
            CB_LOOP
            for (int k = 0; k < aws(); ++k) {
                uint32_t bits = adjm()[u * aws() + k];
                CB_LOOP
                while (bits) {
                    int v = 32 * k + (int)ctz32(bits);
                    bits &= bits - 1;
                    if (v == u) continue;
                    int64_t t = 2 * W()[u * ws() + v] - lab()[v];
                    if (!any || t > best) {
                        any = true;
                        best = t;
                        bv = v;
                    }
                }
            }
            if (!any) continue;
            lab()[u] = best;
            if (!match()[bv]) {
                match()[u] = (uint16_t)bv;
                match()[bv] = (uint16_t)u;
                ++cnt;
            }
        }
        CB_LOOP
        while (matching(g)) ++cnt;
        return cnt;
    }
};
static_assert(sizeof(Blossom<HostGroup, 8>) <= Blossom<HostGroup, 8>::HDR, "Blossom header room");

// ---------------------------------------------------------------------------------------------------------------
// Stage 2 driver
// ---------------------------------------------------------------------------------------------------------------
CB_HD uint32_t uf_find(uint32_t *uf, uint32_t x) {
    while (true) {
        uint32_t p = uf[x];
        if (p == x) return x;
        x = p;
    }
}
CB_HD void uf_union(uint32_t *uf, uint32_t a, uint32_t b) {
    while (true) {
        a = uf_find(uf, a);
        b = uf_find(uf, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }
        if (cb_atomic_cas(&uf[a], a, b) == a) return;  // hook the larger root under the smaller
    }
}

// Exact distance between two events of the same component, taken from whichever ball saw the other.
CB_HD int64_t seen_distance(const Shot &s, uint32_t i, uint32_t j) {
    // both lists side by side (a pair seen from both sides carries the same distance in both)
    const NbEntry *li = nb_list(s, i), *lj = nb_list(s, j);
    const uint32_t ni = s.nb_cnt[i], nj = s.nb_cnt[j];
    const uint32_t nmax = ni > nj ? ni : nj;
    CB_LOOP
    for (uint32_t k = 0; k < nmax; k++) {
        uint32_t vi = NO_INDEX, vj = NO_INDEX;
        int64_t xi = 0, xj = 0;
        if (k < ni) { vi = li[k].ev; xi = li[k].d; }
        if (k < nj) { vj = lj[k].ev; xj = lj[k].d; }
        if (vi == j) return xi;
        if (vj == i) return xj;
    }
    return -1;
}

CB_HD void request_growth(Shot &s, uint32_t ev, int64_t to) {
    uint32_t k = cb_atomic_add(&s.hdr->ngrow, 1);
    s.grow_ev[k] = ev;
    s.grow_to[k] = to;
}

// Four events: the perfect-matching LP of K4 needs no blossom inequalities (an odd set of three is the complement of a
// single vertex), so vertex duals alone certify the best of the three pairings. With the pairing {ab, cd} optimal, the
// doubled duals h_a = t, h_b = 2 d_ab - t, h_c = s, h_d = 2 d_cd - s are feasible iff
//     2 (d_ab + d_cd - d_bd) <= t + s <= 2 d_ac      and      2 (d_ab - d_bc) <= t - s <= 2 (d_ad - d_cd),
// a non-empty box by optimality. The balanced point (d_ab, d_cd) is clamped into it. Returns false (generic solver)
// when the seen edges admit no perfect matching.
// The six distances live in named scalars and the three pairings are spelled out: indexing a 4 x 4 local array with
// run-time indices made the compiler build select trees over every element (27 KB of code for this routine, run by
// divergent single threads).
CB_HDN bool solve_component_four(Shot &s, const uint32_t *mem) {
    const uint32_t e0 = mem[0], e1 = mem[1], e2 = mem[2], e3 = mem[3];
    int64_t d01 = -1, d02 = -1, d03 = -1, d12 = -1, d13 = -1, d23 = -1;
    {
        // the four lists side by side: one entry of each per step, so that their (independent) global loads overlap
        const NbEntry *l0 = nb_list(s, e0), *l1 = nb_list(s, e1), *l2 = nb_list(s, e2), *l3 = nb_list(s, e3);
        const uint32_t n0 = s.nb_cnt[e0], n1 = s.nb_cnt[e1], n2 = s.nb_cnt[e2], n3 = s.nb_cnt[e3];
        uint32_t nmax = n0 > n1 ? n0 : n1;
        if (n2 > nmax) nmax = n2;
        if (n3 > nmax) nmax = n3;
        CB_LOOP
        for (uint32_t k = 0; k < nmax; k++) {
            uint32_t v0 = NO_INDEX, v1 = NO_INDEX, v2 = NO_INDEX, v3 = NO_INDEX;
            int64_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;
            if (k < n0) { v0 = l0[k].ev; x0 = l0[k].d; }
            if (k < n1) { v1 = l1[k].ev; x1 = l1[k].d; }
            if (k < n2) { v2 = l2[k].ev; x2 = l2[k].d; }
            if (k < n3) { v3 = l3[k].ev; x3 = l3[k].d; }
            if (v0 == e1) d01 = x0;
            else if (v0 == e2) d02 = x0;
            else if (v0 == e3) d03 = x0;
            if (v1 == e0) d01 = x1;
            else if (v1 == e2) d12 = x1;
            else if (v1 == e3) d13 = x1;
            if (v2 == e0) d02 = x2;
            else if (v2 == e1) d12 = x2;
            else if (v2 == e3) d23 = x2;
            if (v3 == e0) d03 = x3;
            else if (v3 == e1) d13 = x3;
            else if (v3 == e2) d23 = x3;
        }
    }
    // pairings in fixed preference order: {01,23}, {02,13}, {03,12}; the first cheapest one wins
    int best = -1;
    int64_t best_cost = 0;
    if (d01 >= 0 && d23 >= 0) {
        best = 0;
        best_cost = d01 + d23;
    }
    if (d02 >= 0 && d13 >= 0 && (best < 0 || d02 + d13 < best_cost)) {
        best = 1;
        best_cost = d02 + d13;
    }
    if (d03 >= 0 && d12 >= 0 && (best < 0 || d03 + d12 < best_cost)) {
        best = 2;
        best_cost = d03 + d12;
    }
    if (best < 0) {
        // No pairing has both edges: the seen edges form a star (the only connected graph on four vertices without a
        // perfect matching). The blossom's greedy start pairs the centre with its closest leaf (lowest index on ties) and
        // no augmenting path exists afterwards, so the other two leaves are left over and look further.
        int c = -1;
        int64_t x0 = 0, x1 = 0, x2 = 0;  // distances from the centre to the other three vertices, in index order
        if (d01 >= 0 && d02 >= 0 && d03 >= 0) { c = 0; x0 = d01; x1 = d02; x2 = d03; }
        else if (d01 >= 0 && d12 >= 0 && d13 >= 0) { c = 1; x0 = d01; x1 = d12; x2 = d13; }
        else if (d02 >= 0 && d12 >= 0 && d23 >= 0) { c = 2; x0 = d02; x1 = d12; x2 = d23; }
        else if (d03 >= 0 && d13 >= 0 && d23 >= 0) { c = 3; x0 = d03; x1 = d13; x2 = d23; }
        if (c < 0) return false;
        int leaf_pos = 0;  // position among the three non-centre vertices (ascending index); first minimum
        int64_t leaf_d = x0;
        if (x1 < leaf_d) { leaf_pos = 1; leaf_d = x1; }
        if (x2 < leaf_d) { leaf_pos = 2; leaf_d = x2; }
        const int leaf = leaf_pos + (leaf_pos >= c ? 1 : 0);
        s.solved[e0] = 0;
        if (c != 0 && leaf != 0) request_growth(s, e0, -1);
        s.solved[e1] = 0;
        if (c != 1 && leaf != 1) request_growth(s, e1, -1);
        s.solved[e2] = 0;
        if (c != 2 && leaf != 2) request_growth(s, e2, -1);
        s.solved[e3] = 0;
        if (c != 3 && leaf != 3) request_growth(s, e3, -1);
        return true;
    }
    // (A, B, C, D) = (0,1,2,3), (0,2,1,3) or (0,3,1,2)
    uint32_t eB, eC, eD;
    int64_t dab, dcd, dac, dbd, dad, dbc;
    if (best == 0) {
        eB = e1; eC = e2; eD = e3;
        dab = d01; dcd = d23; dac = d02; dbd = d13; dad = d03; dbc = d12;
    } else if (best == 1) {
        eB = e2; eC = e1; eD = e3;
        dab = d02; dcd = d13; dac = d01; dbd = d23; dad = d03; dbc = d12;
    } else {
        eB = e3; eC = e1; eD = e2;
        dab = d03; dcd = d12; dac = d01; dbd = d23; dad = d02; dbc = d13;
    }
    int64_t P = dab + dcd, Q = dab - dcd;
    if (dac >= 0 && P > 2 * dac) P = 2 * dac;
    if (dbd >= 0 && P < 2 * (dab + dcd - dbd)) P = 2 * (dab + dcd - dbd);
    if (dad >= 0 && Q > 2 * (dad - dcd)) Q = 2 * (dad - dcd);
    if (dbc >= 0 && Q < 2 * (dab - dbc)) Q = 2 * (dab - dbc);
    const int64_t t = (P + Q) / 2, u = (P - Q) / 2;
    const int64_t hA = t, hB = 2 * dab - t, hC = u, hD = 2 * dcd - u;
    s.h[e0] = hA;
    s.solved[e0] = 1;
    s.pc[e0] = e0;
    if (hA > s.R[e0]) request_growth(s, e0, hA);
    s.h[eB] = hB;
    s.solved[eB] = 1;
    s.pc[eB] = e0;
    if (hB > s.R[eB]) request_growth(s, eB, hB);
    s.h[eC] = hC;
    s.solved[eC] = 1;
    s.pc[eC] = e0;
    if (hC > s.R[eC]) request_growth(s, eC, hC);
    s.h[eD] = hD;
    s.solved[eD] = 1;
    s.pc[eD] = e0;
    if (hD > s.R[eD]) request_growth(s, eD, hD);
    s.partner[e0] = (int32_t)eB;
    s.partner[eB] = (int32_t)e0;
    s.partner[eC] = (int32_t)eD;
    s.partner[eD] = (int32_t)eC;
    return true;
}

// First pass over component c, one thread per component: sorts the members, handles odd components (everyone grows)
// and single-edge components (closed form). Returns false when the component needs the blossom solver.
CB_HDN bool solve_component_small(Shot &s, uint32_t c) {
    uint32_t r = s.comp_root[c];
    uint32_t n = s.cnt[r];
    uint32_t *mem = s.members + s.start[r];
    // deterministic member order: insertion sort (components are small)
    CB_LOOP
    for (uint32_t a = 1; a < n; a++) {
        uint32_t v = mem[a];
        uint32_t b = a;
        CB_LOOP
        while (b > 0 && mem[b - 1] > v) {
            mem[b] = mem[b - 1];
            b--;
        }
        mem[b] = v;
    }
    if (n & 1) {
        // nobody can be left over: every member looks further
        CB_LOOP
        for (uint32_t a = 0; a < n; a++) {
            s.solved[mem[a]] = 0;
            request_growth(s, mem[a], -1);
        }
        return true;
    }
    if (n == 2) {
        // Closed form of the blossom on one edge: y = d/2 each, so both balls need radius >= d.
        uint32_t i = mem[0], j = mem[1];
        int64_t d = seen_distance(s, i, j);
        s.h[i] = s.h[j] = d;
        s.solved[i] = s.solved[j] = 1;
        s.pc[i] = s.pc[j] = i;
        s.partner[i] = (int32_t)j;
        s.partner[j] = (int32_t)i;
        if (d > s.R[i]) request_growth(s, i, d);
        if (d > s.R[j]) request_growth(s, j, d);
        return true;
    }
    if (n == 4) return solve_component_four(s, mem);
    return false;
}

// Scratch bytes of the solver class that handles n events: fixed layouts of 16 and 32 vertices (shared memory only), the
// run-time sized layout beyond. A third fixed class for <= 8 vertices (-DCB_CAP8) saves pool space but costs a fourth copy
// of the 41 KB solver body; without it every class runs 5-12 % faster (instruction cache) and, with a 32 KB pool, the
// step gets 2.5 % shorter (measured at d25 r100, profiles/r02_kernel_experiments.txt).
template <class G>
CB_HD uint64_t blossom_bytes(uint32_t n) {
#ifdef CB_ONE_BODY
    return Blossom<G, 0>::bytes_needed((int)n);  // experiment: the run-time layout for every size (one code copy)
#endif
#ifdef CB_CAP8
    if (n <= 8) return Blossom<G, 8>::bytes_needed((int)n);
#endif
    if (n <= 16) return Blossom<G, 16>::bytes_needed((int)n);
    if (n <= 32) return Blossom<G, 32>::bytes_needed((int)n);
    return Blossom<G, 0>::bytes_needed((int)n);
}

// Set-up shared by every solver class, all lanes: clears the weight matrix W (row stride ws) and fills it from the
// members' neighbour lists (global memory). Only pairs inside the component take part; pairs seen across components are
// guarded by the dual check. Weights are BLOSSOM_BIG - distance: the solver only ever looks at label sums minus doubled
// weights, so its run does not depend on the constant (the oracle uses max distance + 2 and takes the same steps) --
// a fixed one saves the extra pass over the lists that a data-dependent one needs. Distances are < 2^49 (node ids have
// 24 bits, edge weights 25). The lists are walked by lanes / n lanes per member (n < lanes), each entry by one lane: the
// walk is a chain of three dependent global loads per entry, so its length is what the set-up costs.
constexpr int64_t BLOSSOM_BIG = int64_t{1} << 52;
template <class G>
CB_HDN int64_t blossom_fill_weights(Shot &s, const G &g, uint32_t r, uint32_t n, const uint32_t *mem, int64_t *W, uint32_t ws) {
    const uint32_t lane = (uint32_t)g.lane(), lanes = (uint32_t)g.lanes();
    CB_LOOP
    for (uint32_t k = lane; k < (n + 1) * ws; k += lanes) W[k] = 0;
    CB_LOOP
    for (uint32_t a = lane; a < n; a += lanes) s.lidx[mem[a]] = a + 1;
    g.wsync();
    const uint32_t per = n < lanes ? lanes / n : 1u;  // lanes per member
    const uint32_t sweep = lanes / per;               // members per sweep
    const uint32_t sub = lane % per;
    CB_LOOP
    for (uint32_t a = lane / per; a < n; a += sweep) {
        const NbEntry *l = nb_list(s, mem[a]);
        const uint32_t cnt = s.nb_cnt[mem[a]];
        CB_LOOP
        for (uint32_t k = sub; k < cnt; k += per) {
            const uint32_t ev = l[k].ev;
            if (s.root[ev] != r) continue;
            const uint32_t b = s.lidx[ev];
            const int64_t w = BLOSSOM_BIG - l[k].d;
            W[(a + 1) * ws + b] = w;  // a pair seen from both sides writes the same value twice
            W[b * ws + (a + 1)] = w;
        }
    }
    g.wsync();
    return BLOSSOM_BIG;
}

// Write-back shared by every solver class, all lanes, one member each.
template <class G>
CB_HDN void blossom_write_back(Shot &s, const G &g, uint32_t n, const uint32_t *mem, uint32_t pairs, int64_t big,
                               const uint16_t *match, const int64_t *lab) {
    const uint32_t lane = (uint32_t)g.lane(), lanes = (uint32_t)g.lanes();
    if (pairs * 2 != n) {
        CB_LOOP
        for (uint32_t a = lane; a < n; a += lanes) {
            s.solved[mem[a]] = 0;
            if (!match[a + 1]) request_growth(s, mem[a], -1);
        }
        g.wsync();
        return;
    }
    CB_LOOP
    for (uint32_t a = lane; a < n; a += lanes) {
        uint32_t e = mem[a];
        int64_t need = big - lab[a + 1];  // = 2 * y_a
        s.h[e] = need;
        s.solved[e] = 1;
        s.pc[e] = mem[0];
        s.partner[e] = (int32_t)mem[match[a + 1] - 1];
        if (need > s.R[e]) request_growth(s, e, need);
    }
    g.wsync();
}

template <class G, int CAP>
CB_HDN void solve_component_blossom_t(Shot &s, const G &g, uint32_t c, uint8_t *scratch) {
    typedef Blossom<G, CAP> BL;
    uint32_t r = s.comp_root[c];
    uint32_t n = s.cnt[r];
    uint32_t *mem = s.members + s.start[r];
    if (n > MAX_COMPONENT) {
        if (g.lane() == 0) s.fail(SHOT_COMPONENT_TOO_LARGE);
        return;
    }
    uint64_t off = 0;
    if (g.lane() == 0) {
        uint8_t *memory = scratch ? scratch : s.alloc(BL::bytes_needed((int)n));
        off = memory ? (uint64_t)(uintptr_t)memory : 0;
    }
    off = g.shfl64(off, 0);
    if (!off) return;
    BL &bl = *BL::place((uint8_t *)(uintptr_t)off);
    bl.in_shared();
    if (g.lane() == 0) bl.bind((uint8_t *)(uintptr_t)off, (int)n, g.lanes());
    g.wsync();
    const unsigned long long prof_t0 = s.prof ? cb_clock() : 0;
    const int64_t big = blossom_fill_weights(s, g, r, n, mem, bl.W(), (uint32_t)bl.ws());
#ifdef CB_SUBPROF  // diagnostic build: set-up / prepare / solve / write-back cycles of one size class in slots 16..19
    const bool subprof = s.prof && g.lane() == 0 && (n <= 8 ? 0 : (n <= 16 ? 1 : (n <= 32 ? 2 : 3))) == CB_SUBPROF;
    unsigned long long sub_t = cb_clock();
    if (subprof) {
        cb_atomic_add64((uint64_t *)&s.prof[16], sub_t - prof_t0);
        cb_atomic_add64((uint64_t *)&s.prof[22], n);
        cb_atomic_add64((uint64_t *)&s.prof[23], 1);
    }
#endif
    bl.prepare(g);
#ifdef CB_SUBPROF
    if (subprof) {
        unsigned long long now_ = cb_clock();
        cb_atomic_add64((uint64_t *)&s.prof[17], now_ - sub_t);
        sub_t = now_;
    }
#endif
    uint32_t pairs = 0;
    if (BL::WARP_SOLVER && g.lanes() == 32 && (CAP == 0 || n >= WARP_SOLVER_MIN_N)) {
        // fixed layouts on a 32-lane warp: every lane runs the warp-synchronous solver
        pairs = (uint32_t)bl.solve_w(g);
        if (g.lane() == 0 && s.prof) {
            int bucket = n <= 8 ? 0 : (n <= 16 ? 1 : (n <= 32 ? 2 : 3));
            cb_atomic_add64((uint64_t *)&s.prof[8 + bucket], cb_clock() - prof_t0);
#ifndef CB_XPROF2
            cb_atomic_add64((uint64_t *)&s.prof[12 + bucket], 1);
#endif
#ifdef CB_SUBPROF
            if (subprof) {
                unsigned long long now_ = cb_clock();
                cb_atomic_add64((uint64_t *)&s.prof[18], now_ - sub_t);
                sub_t = now_;
            }
#endif
        }
        if (bl.broken) pairs = NO_INDEX;
    } else if (g.lane() == 0) {
        pairs = (uint32_t)bl.solve(g);
        bl.release_workers(g);
#ifdef CB_SUBPROF
        if (subprof) {
            unsigned long long now_ = cb_clock();
            cb_atomic_add64((uint64_t *)&s.prof[18], now_ - sub_t);
            sub_t = now_;
        }
#endif
        if (s.prof) {
            int bucket = n <= 8 ? 0 : (n <= 16 ? 1 : (n <= 32 ? 2 : 3));
            cb_atomic_add64((uint64_t *)&s.prof[8 + bucket], cb_clock() - prof_t0);
#ifndef CB_XPROF2
            cb_atomic_add64((uint64_t *)&s.prof[12 + bucket], 1);
#endif
        }
        if (bl.broken) pairs = NO_INDEX;
    } else if (bl.is_par()) {
        bl.worker_loop(g);
    }
    g.wsync();
    pairs = g.shfl32(pairs, 0);
    if (pairs == NO_INDEX) {
        if (g.lane() == 0) s.fail(SHOT_SCRATCH_OVERFLOW);
        return;
    }
    blossom_write_back(s, g, n, mem, pairs, big, bl.match(), bl.lab());
#ifdef CB_SUBPROF
    if (subprof) cb_atomic_add64((uint64_t *)&s.prof[19], cb_clock() - sub_t);
#endif
}

// Exact solve of component c by the dense blossom; every lane of the warp calls this together. Lane 0 drives the solver,
// the other lanes fill the weight matrix, serve the data-parallel helpers (large components only) and write back.
// `scratch`: a slice of the shared-memory pool holding blossom_bytes(n) bytes, or nullptr (dynamic region, run-time
// sized layout).
template <class G>
CB_HD void solve_component_blossom(Shot &s, const G &g, uint32_t c, uint8_t *scratch) {
    const uint32_t n = s.cnt[s.comp_root[c]];
#ifdef CB_ONE_BODY
    solve_component_blossom_t<G, 0>(s, g, c, scratch);
    return;
#endif
#ifdef CB_CAP8
    if (scratch && n <= 8) solve_component_blossom_t<G, 8>(s, g, c, scratch);
    else
#endif
    if (scratch && n <= 16) solve_component_blossom_t<G, 16>(s, g, c, scratch);
    else if (scratch && n <= 32) solve_component_blossom_t<G, 32>(s, g, c, scratch);
    else solve_component_blossom_t<G, 0>(s, g, c, scratch);
}

// Serves growth requests [task_next, limit) of this round, one warp per request, until none is left. Requests are
// claimed with a compare-and-swap so that a warp never takes a ticket beyond `limit` (later requests are still being
// produced by running blossoms and are served after them).
template <class G>
CB_HDN void grow_some(Shot &s, const G &g, uint32_t limit) {
    while (true) {
        uint32_t k = NO_INDEX;
        if (g.lane() == 0) {
            uint32_t cur = cb_load_volatile(&s.hdr->task_next);
            while (cur < limit) {
                const uint32_t old = cb_atomic_cas(&s.hdr->task_next, cur, cur + 1);
                if (old == cur) {
                    k = cur;
                    break;
                }
                cur = old;
            }
        }
        k = g.shfl32(k, 0);
        if (k == NO_INDEX) return;
        const uint32_t ev = s.grow_ev[k];
        explore_event(s, g, ev, s.grow_to[k], s.nb_cnt[ev], s.grow_to[k] < 0 ? GROW_EXTRA_EVENTS : 0u);
    }
}

template <class G, class PT>
CB_HDN uint32_t match_shot(Shot &s, const G &g, PT &pt) {
    const uint32_t M = s.M;
    s.chunk_left = 0;  // the dynamic region was reset for this shot
    CB_LOOP
    for (uint32_t i = g.tid(); i < M; i += g.n()) {
        s.partner[i] = -1;
        s.solved[i] = 0;
        s.regrown[i] = 0;
        s.stale[i] = 0;
        s.pc[i] = NO_INDEX;
        s.nb_cnt[i] = 0;
        s.used_flag[i] = 0;
        s.uf[i] = i;
        s.comp_dirty[i] = 0;
        s.cnt[i] = 0;
    }
    if (g.tid() == 0) s.hdr->ncomp = 0;
#ifdef CB_TAIL_PROF
    if (g.tid() == 0) s.hdr->pad_hdr[0] = 0;  // largest component of the shot (diagnostic build)
#endif
    g.sync();
#ifdef CB_XPROF
    unsigned long long xt = cb_clock();
#endif
    explore_first_near(s, g);
    CB_XLAP(s, g, xt, 12)
    // the few events whose row held no event: one warp each (the lanes look at 32 events at a time)
    CB_LOOP
    for (uint32_t base = (uint32_t)g.warp() * (uint32_t)g.lanes(); base < M; base += (uint32_t)(g.nwarps() * g.lanes())) {
        const uint32_t i = base + (uint32_t)g.lane();
        uint32_t todo = g.ballot(i < M && s.nb_cnt[i] == 0);
        CB_LOOP
        while (todo) {
            explore_event(s, g, base + ctz32(todo), -1, 0);
            todo &= todo - 1;
        }
    }
    CB_XLAP(s, g, xt, 13)
    uint32_t round = 0;
    uint32_t *const act = s.lidx;  // events of this round's dirty components (lidx is only used inside the blossom solves)
    CB_LOOP
    while (true) {
        g.sync();
        pt.lap(round == 0 ? 1 : 4);
        if (s.hdr->status) return round;
        if (++round > MAX_ROUNDS) {
            if (g.tid() == 0) s.fail(SHOT_ROUND_LIMIT);
            g.sync();
            return round;
        }
        // The union-find persists across rounds: every union ever made stays valid (a solved component stays whole, an
        // unsolved one re-joins everything it has seen), so a round only adds unions. What the previous round's dirty
        // components used is reset here; everything else below touches all events in three passes only.
#ifdef CB_XPROF2
        unsigned long long xt2 = cb_clock();
#define CB_XLAP2(slot) if (s.prof && g.tid() == 0) { unsigned long long now_ = cb_clock(); cb_atomic_add64((uint64_t *)&s.prof[slot], now_ - xt2); xt2 = now_; }
#else
#define CB_XLAP2(slot)
#endif
        const uint32_t prev_ncomp = s.hdr->ncomp;
        CB_LOOP
        for (uint32_t c = g.tid(); c < prev_ncomp; c += g.n()) {
            const uint32_t r = s.comp_root[c];
            s.cnt[r] = 0;
            s.comp_dirty[r] = 0;
        }
        g.sync();
        if (g.tid() == 0) {
            s.hdr->ncomp = 0;
            s.hdr->mem_top = 0;
            s.hdr->ngrow = 0;
            s.hdr->nbig = 0;
            s.hdr->nhuge = 0;  // number of active events
            s.hdr->task_next = 0;
        }
        g.sync();
        // Lazy merging: two events are joined when either has no dual yet or their duals overlap (h_i + h_j > 2 d).
        // A violated pair inside one solved component is either an old blossom-internal edge or a pair first seen after
        // the solve -- only a ball that grew since can show the latter, which marks it stale.
        CB_LOOP
        for (uint32_t i = g.tid(); i < M; i += g.n()) {
            const bool si = s.solved[i] != 0;
            const NbEntry *l = nb_list(s, i);
            CB_LOOP
            for (uint32_t k = 0; k < s.nb_cnt[i]; k++) {
                uint32_t j = l[k].ev;
                if (!si || !s.solved[j] || s.h[i] + s.h[j] > 2 * l[k].d) {
                    if (si && s.solved[j] && s.pc[i] == s.pc[j] && s.regrown[i]) s.stale[s.pc[i]] = 1;
                    uf_union(s.uf, i, j);
                }
            }
        }
        g.sync();
        CB_XLAP2(12)
        CB_LOOP
        for (uint32_t i = g.tid(); i < M; i += g.n()) {
            uint32_t r = uf_find(s.uf, i);
            s.root[i] = r;
            s.uf[i] = r;  // path compression (concurrent finds still see an ancestor)
            s.regrown[i] = 0;
            // a component is (re)solved when a member has no dual, when it joins events of different solved components,
            // or when a solved component went stale
            if (!s.solved[i] || s.stale[s.pc[i]] || !s.solved[r] || s.pc[r] != s.pc[i]) s.comp_dirty[r] = 1;
        }
        g.sync();
        CB_XLAP2(13)
        CB_LOOP
        for (uint32_t i = g.tid(); i < M; i += g.n())
            if (s.comp_dirty[s.root[i]]) {
                cb_atomic_add(&s.cnt[s.root[i]], 1);
                act[cb_atomic_add(&s.hdr->nhuge, 1)] = i;
            }
        g.sync();
        CB_XLAP2(14)
        const uint32_t nact = s.hdr->nhuge;
        CB_LOOP
        for (uint32_t k = g.tid(); k < nact; k += g.n()) {
            const uint32_t i = act[k];
            if (s.root[i] == i) {
                uint32_t slot = cb_atomic_add(&s.hdr->ncomp, 1);
                s.comp_root[slot] = i;
                s.start[i] = cb_atomic_add(&s.hdr->mem_top, s.cnt[i]);
                s.fillp[i] = 0;
#ifdef CB_TAIL_PROF
                cb_atomic_max(&s.hdr->pad_hdr[0], s.cnt[i]);
#endif
            }
        }
        g.sync();
        CB_LOOP
        for (uint32_t k = g.tid(); k < nact; k += g.n()) {
            const uint32_t i = act[k];
            const uint32_t r = s.root[i];
            uint32_t p = cb_atomic_add(&s.fillp[r], 1);
            s.members[s.start[r] + p] = i;
            if (s.solved[i]) s.stale[s.pc[i]] = 0;
        }
        g.sync();
        CB_XLAP2(15)
        const uint32_t ncomp = s.hdr->ncomp;
        pt.lap(2);
        // small components: one thread each; the rest is queued for the blossom
        CB_LOOP
        for (uint32_t c = g.tid(); c < ncomp; c += g.n())
            if (!solve_component_small(s, c)) s.blist[cb_atomic_add(&s.hdr->nbig, 1)] = c;
        g.sync();
        pt.lap(8);
        const uint32_t nbig = s.hdr->nbig;
#ifdef CB_WAVE_PROF
        const unsigned long long wave_prof_before = pt.t[9];
#endif
        // Blossom components share the group's shared-memory table space as one pool: largest first, each gets its own
        // slice, and what does not fit next to the others waits for the next wave (components beyond the whole pool use
        // the dynamic region). s.blist holds the components, s.fillp their pool offset / 16 (0xFFFF: dynamic region),
        // s.scan_tmp the first blossom of every wave.
        if (nbig) {
            // order: (size descending, root ascending), by counting predecessors -- one thread per component.
            // s.path_off (unused until the path phase) receives the ordered list, s.fillp the sizes.
            CB_LOOP
            for (uint32_t k = (uint32_t)g.tid(); k < nbig; k += (uint32_t)g.n()) {
                const uint32_t v = s.blist[k], vr = s.comp_root[v], vn = s.cnt[vr];
                uint32_t rank = 0;
                CB_LOOP
                for (uint32_t j = 0; j < nbig; j++) {
                    const uint32_t ur = s.comp_root[s.blist[j]], un = s.cnt[ur];
                    rank += (un > vn || (un == vn && ur < vr)) ? 1u : 0u;
                }
                s.path_off[rank] = v;
                s.fillp[rank] = vn;
            }
            g.sync();
            if (g.tid() == 0) {
                const uint64_t pool = s.sh_tables_bytes;
                uint32_t wave = 0;
                uint64_t top = 0;
                s.scan_tmp[0] = 0;  // scan_tmp[w] = first blossom of wave w
                CB_LOOP
                for (uint32_t k = 0; k < nbig; k++) {
                    const uint64_t need = align16(blossom_bytes<G>(s.fillp[k]));
                    if (need > pool) {
                        s.fillp[k] = 0xFFFFu;
                        continue;
                    }
                    if (top + need > pool && wave + 2 < SCAN_TMP) {
                        wave++;
                        top = 0;
                        s.scan_tmp[wave] = k;
                    }
                    s.fillp[k] = top + need > pool ? 0xFFFFu : (uint32_t)(top >> 4);
                    top += need;
                }
                s.scan_tmp[wave + 1] = nbig;
                s.hdr->bcast = wave + 1;
            }
            g.sync();
            const uint32_t nwaves = s.hdr->bcast;
            const uint32_t ngrow_cf = s.hdr->ngrow;  // complete: the closed-form pass ended at a barrier, no blossom has run yet
            g.sync();
            CB_LOOP
            for (uint32_t wave = 0; wave < nwaves; wave++) {
                const uint32_t k0 = s.scan_tmp[wave], k1 = s.scan_tmp[wave + 1];
                // largest first, dealt out in serpentine order (warp 0 gets the 1st and the 16th, warp 7 the 8th and 9th)
                CB_LOOP
                for (uint32_t base = k0, pass = 0; base < k1; base += g.nwarps(), pass++) {
                    const uint32_t k = base + ((pass & 1) ? (uint32_t)(g.nwarps() - 1 - g.warp()) : (uint32_t)g.warp());
                    if (k >= k1) continue;
                    const uint32_t off16 = s.fillp[k];
                    solve_component_blossom(s, g, s.path_off[k], off16 == 0xFFFFu ? nullptr : s.sh_tables + ((uint64_t)off16 << 4));
                    g.wsync();
                }
                // A warp without (further) blossoms in this wave serves the growth requests the closed forms made (not
                // when the pool doubles as the warps' shared ball tables, i.e. without a near table).
                if (!s.sh_ball.H) grow_some(s, g, ngrow_cf);
                g.sync();
            }
            pt.lap(9);
#ifdef CB_WAVE_PROF
            if (s.prof && g.tid() == 0) cb_atomic_add64((uint64_t *)&s.prof[12 + (round < 4 ? round - 1 : 3)], pt.t[9] - wave_prof_before);
#endif
            if (s.sh_ball.H)
                CB_LOOP
                for (uint32_t k = g.lane(); k < s.sh_ball.H; k += g.lanes()) s.sh_ball.key[k] = HT_EMPTY;
        }
        g.sync();
        pt.lap(10);
        const uint32_t st = s.hdr->status;
        const uint32_t ngrow = s.hdr->ngrow;
        g.sync();  // everyone has read the verdict before the next phase may change it
        if (st) return round;
        if (ngrow == 0) return round;
        grow_some(s, g, ngrow);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Matched pair -> graph edges along the canonical shortest path: a direct edge of the pair's distance if there is one,
// otherwise, among equal-length paths, walking back from the larger event index towards the smaller one always steps
// to the smallest-id tight predecessor. Stored as the node
// sequence src .. dst in the dynamic region. One warp per pair.
// ---------------------------------------------------------------------------------------------------------------
// Writes the path src .. dst of event i's matched pair, dst being entry `at` of src's near-table row: a direct edge if
// the entry is flagged, otherwise the chain of stored canonical predecessors. One thread. Returns false when the chain
// runs through a predecessor whose position is not stored.
CB_HDN bool near_path_emit(Shot &s, uint32_t i, uint32_t src, uint32_t dst, uint32_t at, bool chunked) {
    const NearEntry last = *near_at(s.T, src, at);
    uint32_t len = 2;
    if (!(last.dist_flag & 0x80000000u))
        for (uint32_t q = last.node_pred >> 24; q != NEAR_PRED_SRC; q = near_at(s.T, src, q)->node_pred >> 24) {
            if (q == NEAR_PRED_FAR) return false;  // predecessor position not stored: the caller searches
            len++;
        }
    uint32_t *out = (uint32_t *)(chunked ? s.alloc_small(uint64_t{len} * 4) : s.alloc(uint64_t{len} * 4));
    if (!out) return true;  // alloc flagged the overflow
    out[0] = src;
    out[len - 1] = dst;
    uint32_t k = len - 2;
    if (len > 2)
        for (uint32_t q = last.node_pred >> 24; q != NEAR_PRED_SRC; k--) {
            const uint32_t np = near_at(s.T, src, q)->node_pred;
            out[k] = np & 0xFFFFFFu;
            q = np >> 24;
        }
    s.path_off[i] = (uint32_t)(((uint8_t *)out - s.dyn) >> 4);
    s.path_len[i] = len;
    return true;
}

// Path of a matched pair from the near table of `src`: find `dst` in the row, then follow the stored canonical
// predecessors. Returns false when dst is not in the row (the caller searches instead).
template <class G>
CB_HDN bool extract_path_near(Shot &s, const G &g, uint32_t i, uint32_t src, uint32_t dst, int64_t d_total) {
    const uint32_t stride = s.T.near_stride;
    if (!stride) return false;
    const uint32_t lane = (uint32_t)g.lane(), lanes = (uint32_t)g.lanes();
    uint32_t at = NO_INDEX;
    for (uint32_t pos = 0; pos < stride; pos += lanes) {
        const NearEntry e = *near_at(s.T, src, pos + lane);
        const bool pad = e.dist_flag == NEAR_PAD;
        const int64_t d = (int64_t)(e.dist_flag & 0x7FFFFFFFu);
        const uint32_t hits = g.ballot(!pad && (e.node_pred & 0xFFFFFFu) == dst && d == d_total);
        if (hits) {
            at = pos + ctz32(hits);
            break;
        }
        if (g.ballot(pad || d > d_total)) return false;
    }
    if (at == NO_INDEX) return false;
    uint32_t ok = 0;
    if (lane == 0) ok = near_path_emit(s, i, src, dst, at, true) ? 1 : 0;
    g.wsync();
    return g.shfl32(ok, 0) != 0;
}

template <class G>
CB_HDN void extract_path(Shot &s, const G &g, uint32_t i, uint32_t j, int64_t d_total) {
    const uint32_t src = s.event_node(i), dst = s.event_node(j);
    if (extract_path_near(s, g, i, src, dst, d_total)) return;
    // Most matched pairs are graph neighbours: a direct edge of exactly the pair's distance is the path.
    {
        const RowRange row = row_of(s.T, src);
        const AdjEntry *adj = adj_of(s.T);
        bool direct = false;
        for (uint32_t base = row.e0; base < row.e1 && !direct; base += g.lanes()) {
            uint32_t e = base + g.lane();
            bool hit = false;
            if (e < row.e1) {
                AdjEntry a = adj[e];
                hit = a.nbr == dst && a.wt == d_total;
            }
            direct = g.ballot(hit) != 0;
        }
        if (direct) {
            uint64_t off = 0;
            if (g.lane() == 0) {
                uint8_t *mem = s.alloc_small(8);
                off = mem ? (uint64_t)(mem - s.dyn) + 1 : 0;
                if (mem) {
                    ((uint32_t *)mem)[0] = src;
                    ((uint32_t *)mem)[1] = dst;
                    s.path_off[i] = (uint32_t)((off - 1) >> 4);
                    s.path_len[i] = 2;
                }
            }
            g.wsync();
            return;
        }
    }
    int64_t limit = 0;
    uint32_t nfound = 0;
    BallTables *t = nullptr;
    int r = BALL_OVERFLOW;
    for (int attempt = s.sh_ball.H ? 0 : 1; attempt < 2 && r == BALL_OVERFLOW; attempt++) {
        if (t) ball_clear(*t, g);
        t = attempt ? &s.gl_ball : &s.sh_ball;
        r = ball_grow(s, g, *t, src, d_total, 0, dst, false, true, &limit, &nfound);
    }
    bool ok = r == BALL_OK;
    uint32_t len = 1;
    uint32_t *out = nullptr;
    // two walks: count, then fill back to front
    for (int pass = 0; pass < 2 && ok; pass++) {
        uint32_t cur = dst;
        int64_t dc = d_total;
        uint32_t pos = len;
        if (pass == 1 && g.lane() == 0) out[pos - 1] = cur;
        pos--;
        uint32_t steps = 0;
        while (cur != src) {
            const RowRange row = row_of(s.T, cur);
            const AdjEntry *adj = adj_of(s.T);
            uint32_t best = NO_INDEX;
            int64_t bd = 0;
            for (uint32_t base = row.e0; base < row.e1 && best == NO_INDEX; base += g.lanes()) {
                uint32_t e = base + g.lane();
                bool hit = false;
                int64_t dv = -1;
                uint32_t nbr = 0;
                if (e < row.e1) {
                    AdjEntry a = adj[e];
                    nbr = a.nbr;
                    if (nbr != NO_INDEX) {
                        dv = ball_lookup(*t, nbr);
                        hit = dv >= 0 && dv + a.wt == dc;
                    }
                }
                uint32_t hits = g.ballot(hit);
                if (hits) {
                    int first = (int)ctz32(hits);  // rows are sorted by neighbour id: the first hit is the smallest id
                    best = g.shfl32(nbr, first);
                    bd = (int64_t)g.shfl64((uint64_t)dv, first);
                }
            }
            if (best == NO_INDEX || ++steps > s.T.num_nodes) {
                ok = false;
                break;
            }
            cur = best;
            dc = bd;
            if (pass == 1 && g.lane() == 0) out[pos - 1] = cur;
            pos--;
            if (pass == 0) len++;
        }
        if (pass == 0 && ok) {
            uint64_t off = 0;
            if (g.lane() == 0) {
                uint8_t *mem = s.alloc_small(uint64_t{len} * 4);
                off = mem ? (uint64_t)(mem - s.dyn) + 1 : 0;
            }
            off = g.shfl64(off, 0);
            if (!off) {
                ball_clear(*t, g);
                return;  // alloc already flagged the overflow
            }
            out = (uint32_t *)(s.dyn + (off - 1));
        }
    }
    ball_clear(*t, g);
    if (!ok) {
        if (g.lane() == 0) s.fail(SHOT_SCRATCH_OVERFLOW);
        return;
    }
    if (g.lane() == 0) {
        s.path_off[i] = (uint32_t)(((uint8_t *)out - s.dyn) >> 4);
        s.path_len[i] = len;
    }
    g.wsync();
}

template <class G>
CB_HDN void paths_shot(Shot &s, const G &g) {
    for (uint32_t i = g.tid(); i < s.M; i += g.n()) s.path_len[i] = 0;
    g.sync();
#ifdef CB_XPROF
    unsigned long long xt = cb_clock();
#endif
    // Pairs whose partner sits in the near-table row of the smaller event: one thread each.
    uint64_t wsum = 0;
    for (uint32_t i = (uint32_t)g.tid(); i < s.M; i += (uint32_t)g.n()) {
        const int32_t j = s.partner[i];
        if (j <= (int32_t)i) continue;
        const NbEntry *l = nb_list(s, i);
        uint32_t pos = NO_INDEX;
        int64_t d = -1;
        for (uint32_t k = 0; k < s.nb_cnt[i]; k++)
            if (l[k].ev == (uint32_t)j) {
                pos = l[k].pos;
                d = l[k].d;
                break;
            }
        const uint32_t src = s.event_node(i), dst = s.event_node((uint32_t)j);
        if (d < 0 && s.T.near_stride) {
            // seen from the partner's side only: take the distance from there and look the partner up in this event's
            // row (the canonical path is defined from the smaller event), four entries per step
            const NbEntry *lj = nb_list(s, (uint32_t)j);
            for (uint32_t k = 0; k < s.nb_cnt[j]; k++)
                if (lj[k].ev == i) {
                    d = lj[k].d;
                    break;
                }
            for (uint32_t p = 0; d >= 0 && p < s.T.near_stride && pos == NO_INDEX; p += 4) {
                const U4 pair0 = *(const U4 *)near_at(s.T, src, p);
                const U4 pair1 = *(const U4 *)near_at(s.T, src, p + 2);
                const NearEntry e[4] = {{pair0.x, pair0.y}, {pair0.z, pair0.w}, {pair1.x, pair1.y}, {pair1.z, pair1.w}};
                bool past = false;
                for (int k = 0; k < 4; k++) {
                    if (e[k].dist_flag == NEAR_PAD || (int64_t)(e[k].dist_flag & 0x7FFFFFFFu) > d) {
                        past = true;
                        break;
                    }
                    if ((e[k].node_pred & 0xFFFFFFu) == dst && (int64_t)(e[k].dist_flag & 0x7FFFFFFFu) == d) {
                        pos = p + (uint32_t)k;
                        break;
                    }
                }
                if (past) break;
            }
        }
        if (pos != NO_INDEX && near_path_emit(s, i, src, dst, pos, false)) wsum += (uint64_t)d;
    }
    if (wsum) cb_atomic_add64(&s.hdr->weight, wsum);
    g.sync();
    CB_XLAP(s, g, xt, 14)
    // The rest (seen only from the partner's side, or found by a search): one warp each.
    wsum = 0;
    for (uint32_t base = (uint32_t)g.warp() * (uint32_t)g.lanes(); base < s.M; base += (uint32_t)(g.nwarps() * g.lanes())) {
        const uint32_t mine = base + (uint32_t)g.lane();  // the lanes look at 32 events at a time
        uint32_t todo = g.ballot(mine < s.M && s.partner[mine] > (int32_t)mine && s.path_len[mine] == 0);
        while (todo) {
            const uint32_t i = base + ctz32(todo);
            todo &= todo - 1;
            const uint32_t j = (uint32_t)s.partner[i];
            int64_t d = seen_distance(s, i, j);
            wsum += (uint64_t)d;
            extract_path(s, g, i, j, d);
        }
    }
    if (wsum && g.lane() == 0) cb_atomic_add64(&s.hdr->weight, wsum);
    g.sync();
    CB_XLAP(s, g, xt, 15)
}

// ---------------------------------------------------------------------------------------------------------------
// Stages 3 + 4: Euler tours (graph/euler_tours.h:89-112, euler_tours.cc:21-97) over matched edges + one doubling
// edge per detection event, each closed tour discharged through the drag table (decode/decoder.cc:80-203) and
// XOR-ed into the prediction (decoder.cc:268-276).
//
// The reference builds per-node neighbour lists in edge-list order and burns components in the order their first node
// appears in the matcher's edge list. Components are independent, so here every connected component is toured by its
// own thread; its neighbour lists hold the half-edges in the same (edge-list) order and its tour starts at the same
// node, which makes each cycle identical to the reference's and the XOR of the cycle results order independent.
// Half-edge q = 2 * edge + side; edges are numbered in emission order, matched edges first, then doubling edges.
// ---------------------------------------------------------------------------------------------------------------

CB_HD uint32_t find_half_edge(const TablesView &T, uint32_t from, uint32_t to) {
    for (uint32_t e = T.row_ptr[from]; e < T.row_ptr[from + 1]; e++)
        if (T.adj[e].nbr == to) return e;
    return NO_INDEX;
}

// decoder.cc:80-138. `used` bookkeeping is a per-detection-event flag instead of a linear search.
// cyc_di[k] is the DragEntry index of the step that arrives at cycle[k] (the drag-table lookup of the reference's
// `mmm.find`, precomputed per edge by lift_shot). Nothing the loop loads depends on the running charge state except the
// observable mask itself, so the next step's node, drag index and valid bits are fetched one step ahead.
CB_HDN bool discharge_single_start(Shot &s, const uint32_t *cycle, const uint32_t *cyc_di, uint32_t len, uint32_t start, uint64_t *out) {
    // The four optional charge states live in named scalars (`has` bit c set <=> state c holds a value): an array
    // indexed by the run-time colour would sit in local memory, and this loop was a quarter of the kernel's local traffic.
    uint32_t has = 1u << start;
    uint64_t v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    uint32_t cur_node = cycle[len - 1];
    uint32_t cur_loc = cur_node >> 1;
    uint32_t pf_node = cycle[0], pf_di = cyc_di[0];
    for (uint32_t k = 0; k < len; k++) {
        const uint32_t next_node = pf_node, di = pf_di;
        const uint32_t next_loc = next_node >> 1;
        const uint32_t valid = di != NO_INDEX ? s.T.drag[di].valid : 0;
        if (k + 1 < len) {
            pf_node = cycle[k + 1];
            pf_di = cyc_di[k + 1];
        }
        bool has_event = s.node_is_event(cur_node);
        uint32_t evi = has_event ? (s.node_event_index(cur_node) >> 1) : 0;
        uint32_t ahas = 0;
        uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        auto get = [&](uint32_t i) -> uint64_t { return i == 0 ? v0 : (i == 1 ? v1 : (i == 2 ? v2 : v3)); };
        auto set = [&](uint32_t i, uint64_t x) {
            a0 = i == 0 ? x : a0;
            a1 = i == 1 ? x : a1;
            a2 = i == 2 ? x : a2;
            a3 = i == 3 ? x : a3;
            ahas |= 1u << i;
        };
        if (next_loc == cur_loc && has_event && !s.used_flag[evi]) {
            s.used_flag[evi] = 1;
            uint8_t info = s.T.det_info[cur_loc];
            uint32_t dc = info & DET_COLOR_MASK;
            // after[dc] = cur[NEUTRAL]; after[NEUTRAL] = cur[dc] (dc == NEUTRAL is unreachable for valid tables: events
            // never sit on NEUTRAL detectors; the optional semantics then keep cur[NEUTRAL])
            if (has & 1u) set(dc, v0);
            if (has & (1u << dc)) set(0, get(dc));
            if (info & DET_REP_WEIGHT3) {
                uint32_t c1 = dc % 3 + 1, c2 = c1 % 3 + 1;
                uint64_t f = s.T.rep_obs[cur_loc];
                if (has & (1u << c1)) set(c2, get(c1) ^ f);
                if (has & (1u << c2)) set(c1, get(c2) ^ f);
            }
        } else if (di != NO_INDEX) {
            const bool forward = cur_loc <= next_loc;  // walking the stored (min, max) orientation
            const DragEntry &de = s.T.drag[di];
            CB_UNROLL
            for (uint32_t cc = 0; cc < 4; cc++) {
                if (!(has & (1u << cc))) continue;
                const uint64_t vc = cc == 0 ? v0 : (cc == 1 ? v1 : (cc == 2 ? v2 : v3));
                CB_UNROLL
                for (uint32_t nc = 0; nc < 4; nc++) {
                    uint32_t bit = forward ? 4 * cc + nc : 4 * nc + cc;
                    if (valid & (1u << bit)) {
                        const uint64_t x = vc ^ de.obs[bit];
                        if (nc == 0) a0 = x;
                        else if (nc == 1) a1 = x;
                        else if (nc == 2) a2 = x;
                        else a3 = x;
                        ahas |= 1u << nc;
                    }
                }
            }
        }
        has = ahas;
        v0 = a0; v1 = a1; v2 = a2; v3 = a3;
        cur_node = next_node;
        cur_loc = next_loc;
    }
    // reset the used flags touched by this attempt
    for (uint32_t k = 0; k < len; k++) {
        uint32_t nd = cycle[k];
        if (s.node_is_event(nd)) s.used_flag[s.node_event_index(nd) >> 1] = 0;
    }
    if (has & (1u << start)) {
        *out = start == 0 ? v0 : (start == 1 ? v1 : (start == 2 ? v2 : v3));
        return true;
    }
    return false;
}

struct LiftArrays {
    uint32_t *own;      // [HE] node at this end of half-edge q
    uint32_t *slot;     // [HE] id of own[q]'s adjacency list (= smallest half-edge at that node)
    uint32_t *lstart;   // [HE + 1] list start per slot id (exclusive scan of degrees)
    uint32_t *fillc;    // [HE] insertion counters, later the tour cursor (look_next_neighbor state)
    uint32_t *lst;      // [HE] half-edges grouped by list, ascending q inside a list
    uint32_t *pos_of;   // [HE] position of half-edge q in lst
    uint32_t *uf;       // [HE] union-find over slot ids
    uint32_t *cmin;     // [HE] per component root: smallest matched half-edge (tour start)
    uint32_t *cedges;   // [HE] per component root: number of edges
    uint32_t *comps;    // [HE] list of component roots
    uint32_t *cyc;      // cycle buffers
    uint32_t *cyc_e;    // parallel to cyc: the edge that arrives at cyc[k]; replaced by its drag index before the discharge
    uint32_t *mate;     // [HE] the other half-edge at own[q] when that node has exactly two, else NO_INDEX
    uint32_t *edrag;    // [E] DragEntry index of edge e (self-drag entry for an edge inside one detector), NO_INDEX: none
    uint32_t *node_min; // [M + HK] smallest half-edge per compact node id: event nodes by event index, the other (path
                        // interior) nodes by their slot in the hash below + M
    uint32_t *hkey;     // [HK] open-addressing hash of the interior nodes (HK a power of two)
    uint32_t HK, M;
    uint8_t *dead;      // [HE] voided list positions
    uint8_t *simple;    // [HE] per component root: every node has exactly two half-edges (the component is one plain cycle)
    uint32_t HE, EM;    // half-edges, matched edges
};

// Compact id of a node that takes part in the shot's Euler graph (see LiftArrays::node_min). `insert` adds unknown
// interior nodes; lookups of nodes that were never inserted return NO_INDEX.
CB_HD uint32_t lift_node_id(Shot &s, const LiftArrays &L, uint32_t node, bool insert) {
    if (s.node_is_event(node)) return s.node_event_index(node);
    uint32_t k = ht_hash(node, L.HK);
    while (true) {
        uint32_t cur = insert ? cb_atomic_cas(&L.hkey[k], HT_EMPTY, node) : L.hkey[k];
        if (cur == node || (insert && cur == HT_EMPTY)) return L.M + k;
        if (cur == HT_EMPTY) return NO_INDEX;
        k = (k + 1) & (L.HK - 1);
    }
}
CB_HD uint32_t lift_node_slot(Shot &s, const LiftArrays &L, uint32_t node) { return L.node_min[lift_node_id(s, L, node, false)]; }

// Tours and discharges one component. Runs on one thread.
CB_HDN void lift_component(Shot &s, const LiftArrays &L, uint32_t root, uint32_t *cyc, uint32_t *cyc_e) {
    uint32_t clen = 0;
    if (L.simple[root]) {
        // Every node has two half-edges: the component is one plain cycle and the reference's depth-first extension
        // (euler_tours.cc:64-76) just walks it, starting at the first node along its first half-edge -- the smallest
        // matched half-edge of the component. One dependent load per step instead of the list bookkeeping below.
        const uint32_t q0 = L.cmin[root];
        const uint32_t n0 = L.own[q0];
        cyc[clen++] = n0;
        uint32_t q = q0;
        while (true) {
            const uint32_t far = q ^ 1;
            const uint32_t node = L.own[far];
            const uint32_t nq = L.mate[far];
            if (node == n0) {
                cyc_e[0] = q >> 1;  // the closing edge arrives at the first node
                break;
            }
            cyc_e[clen] = q >> 1;
            cyc[clen++] = node;
            q = nq;
            if (clen > L.cedges[root] + 1) {
                s.fail(SHOT_EULER_FAILED);
                return;
            }
        }
    } else {
        auto look_next = [&](uint32_t sl) -> uint32_t {  // euler_tours.cc:34-45
            while (true) {
                uint32_t p = L.lstart[sl] + L.fillc[sl];
                if (p >= L.lstart[sl + 1]) return NO_INDEX;
                if (L.dead[p]) {
                    L.fillc[sl]++;
                    continue;
                }
                return p;
            }
        };
        uint32_t n0 = L.own[L.cmin[root]];
        cyc[clen++] = n0;
        cyc_e[0] = NO_INDEX;
        while (true) {
            // extend_cycle_depth_first (euler_tours.cc:64-76)
            while (true) {
                uint32_t sl = lift_node_slot(s, L, cyc[clen - 1]);
                uint32_t p = look_next(sl);
                if (p == NO_INDEX) break;
                L.fillc[sl]++;
                uint32_t q = L.lst[p];
                cyc_e[clen] = q >> 1;
                cyc[clen++] = L.own[q ^ 1];
                L.dead[L.pos_of[q ^ 1]] = 1;
            }
            // rotate_cycle_to_end_with_unfinished_node (euler_tours.cc:78-97)
            if (cyc[clen - 1] != cyc[0]) {
                s.fail(SHOT_EULER_FAILED);
                return;
            }
            clen--;
            cyc_e[0] = cyc_e[clen];  // the dropped duplicate of the first node carried its arrival edge
            uint32_t ck = 1;
            while (ck < clen && look_next(lift_node_slot(s, L, cyc[ck])) == NO_INDEX) ck++;
            if (ck >= clen) break;
            // cyc = cyc[ck:] + cyc[:ck+1]; the arrival edges move with their nodes
            for (uint32_t *arr : {cyc, cyc_e}) {
                for (uint32_t a = 0, b = ck - 1; a < b; ++a, --b) { uint32_t t = arr[a]; arr[a] = arr[b]; arr[b] = t; }
                for (uint32_t a = ck, b = clen - 1; a < b; ++a, --b) { uint32_t t = arr[a]; arr[a] = arr[b]; arr[b] = t; }
                for (uint32_t a = 0, b = clen - 1; a < b; ++a, --b) { uint32_t t = arr[a]; arr[a] = arr[b]; arr[b] = t; }
            }
            cyc[clen] = cyc[0];
            cyc_e[clen] = cyc_e[0];
            clen++;
        }
    }
    // arrival edges -> drag indices (independent loads, pipelined by the hardware)
    for (uint32_t k = 0; k < clen; k++) cyc_e[k] = L.edrag[cyc_e[k]];
    // discharge_cycle (decoder.cc:140-203): start charges N, R, G, B; the first that closes wins
    for (uint32_t c = 0; c < 4; c++) {
        uint64_t v;
        if (discharge_single_start(s, cyc, cyc_e, clen, c, &v)) {
            if (v) cb_atomic_xor64(&s.hdr->obs, v);
            return;
        }
    }
    s.fail(SHOT_DISCHARGE_FAILED);
}

// ---------------------------------------------------------------------------------------------------------------
// Plain-cycle lift. When no Moebius node takes part in more than one matched path -- no path runs through another
// pair's event node and no two paths share an interior node -- every node of the shot's Euler graph has exactly two
// half-edges (event nodes: their path and their doubling edge), every component is one plain cycle, and the
// reference's tour (euler_tours.cc:64-97) is forced: it starts at the first node of the component's first matcher
// edge, which is the node of the component's smallest event index (pairs are emitted in ascending order of their
// smaller event, each path from the smaller event to the larger), leaves along that event's path, and alternates
// paths and doubling edges until it is back. That cycle is produced here straight from the pair list and the stored
// paths, one thread per detection event: the thread walks from its even event and gives up as soon as it meets a
// smaller event index, so exactly the component's minimum finishes the walk, stores the cycle and discharges it.
// None of the half-edge arrays of the general lift (below) are built: the scratch of this stage drops from ~65 to
// ~10 bytes per half-edge. At d25 r100 nearly every matched pair is a direct edge (784 path edges for 771 pairs) and
// shots with a junction are rare; those take the general lift, which gives identical cycles for plain components.
// ---------------------------------------------------------------------------------------------------------------
// Drag-table index of the step a -> b (decoder.cc:104-127 looks the detector pair up in mmm).
CB_HD uint32_t step_drag_index(const TablesView &T, uint32_t a, uint32_t b) {
    if ((a >> 1) == (b >> 1)) return T.self_drag[a >> 1];
    const uint32_t he = find_half_edge(T, a, b);
    return he != NO_INDEX ? T.adj_drag[he] : NO_INDEX;
}

// Length (number of nodes) of the cycle that starts at event i0 (an even event index), or 0 when a smaller event index
// turns up (i0 is not the component's minimum) or the walk does not close within `limit` nodes. Follows the partner
// links only -- one dependent load per event, no path is read.
CB_HDN uint32_t plain_cycle_length(Shot &s, uint32_t i0, uint32_t limit) {
    uint32_t len = 0;
    uint32_t c = i0;
    while (true) {
        const int32_t pj = s.partner[c];
        if (pj < 0) return 0;
        const uint32_t j = (uint32_t)pj;
        if (j < i0 || (j ^ 1u) < i0) return 0;
        const uint32_t plen = s.path_len[c < j ? c : j];
        if (plen < 2) return 0;
        len += plen;  // plen - 1 path nodes after the current one, plus the twin reached through the doubling edge
        if (len > limit) return 0;
        c = j ^ 1u;
        if (c == i0) return len;
    }
}

// Stores the cycle that starts at event i0 (the minimum of its component, `len` nodes as counted above): the nodes and,
// for each, the drag index of the step that arrives at it (cyc_di[0] belongs to the closing step).
CB_HDN void plain_cycle_fill(Shot &s, uint32_t i0, uint32_t *cyc, uint32_t *cyc_di, uint32_t len) {
    uint32_t at = 1;
    uint32_t cur_node = s.event_node(i0);
    cyc[0] = cur_node;
    uint32_t c = i0;
    while (at <= len) {
        const uint32_t j = (uint32_t)s.partner[c];
        const uint32_t owner = c < j ? c : j;
        const uint32_t plen = s.path_len[owner];
        const uint32_t *p = (const uint32_t *)(s.dyn + ((uint64_t)s.path_off[owner] << 4));
        for (uint32_t k = 1; k < plen; k++) {
            const uint32_t node = c < j ? p[k] : p[plen - 1 - k];
            cyc[at] = node;
            cyc_di[at] = step_drag_index(s.T, cur_node, node);
            cur_node = node;
            at++;
        }
        // doubling edge of the event just reached
        const uint32_t di = s.T.self_drag[cur_node >> 1];
        c = j ^ 1u;
        if (c == i0) {
            cyc_di[0] = di;
            return;
        }
        cur_node ^= 1u;
        cyc[at] = cur_node;
        cyc_di[at] = di;
        at++;
    }
}

// Returns false when the shot has a junction (the caller then runs the general lift); true when the shot was lifted here
// (or failed for good).
template <class G>
CB_HDN bool lift_shot_plain(Shot &s, const G &g) {
    const uint32_t M = s.M;
    // a. junction test over the interior nodes of the paths (most paths are a single edge and have none)
    if (g.tid() == 0) {
        s.hdr->scan_total = 0;   // number of interior nodes
        s.hdr->lift_ncomp = 0;   // junction flag
        s.hdr->lift_cyc_top = 0; // cycle buffer bump pointer
        s.hdr->bcast = 0;        // total path edges
    }
    g.sync();
    {
        uint32_t interior = 0, edges = 0;
        CB_LOOP
        for (uint32_t i = g.tid(); i < M; i += g.n()) {
            const uint32_t len = s.path_len[i];
            if (len >= 2) edges += len - 1;
            if (len > 2) interior += len - 2;
        }
        if (interior) cb_atomic_add(&s.hdr->scan_total, interior);
        if (edges) cb_atomic_add(&s.hdr->bcast, edges);
    }
    g.sync();
    const uint32_t ninterior = s.hdr->scan_total;
    const uint32_t E = s.hdr->bcast + M / 2;
    uint32_t HK = 0;
    if (ninterior) {
        HK = 16;
        CB_LOOP
        while (HK < 2 * ninterior) HK <<= 1;
    }
    g.sync();
    if (g.tid() == 0) {
        uint8_t *mem = s.alloc((uint64_t{2} * E + 8 + HK) * 4);
        s.hdr->scan_total = mem ? (uint32_t)((mem - s.dyn) >> 4) + 1 : 0;
    }
    g.sync();
    const uint32_t moff = s.hdr->scan_total;
    g.sync();
    if (!moff) return true;  // alloc flagged the overflow
    uint32_t *cyc = (uint32_t *)(s.dyn + ((uint64_t)(moff - 1) << 4));
    uint32_t *cyc_di = cyc + E + 4;
    uint32_t *hkey = cyc_di + E + 4;
    if (HK) {
        CB_LOOP
        for (uint32_t k = g.tid(); k < HK; k += g.n()) hkey[k] = HT_EMPTY;
        g.sync();
        CB_LOOP
        for (uint32_t i = g.tid(); i < M; i += g.n()) {
            const uint32_t len = s.path_len[i];
            if (len <= 2) continue;
            const uint32_t *p = (const uint32_t *)(s.dyn + ((uint64_t)s.path_off[i] << 4));
            CB_LOOP
            for (uint32_t k = 1; k + 1 < len; k++) {
                const uint32_t node = p[k];
                bool clash = s.node_is_event(node);
                uint32_t h = ht_hash(node, HK);
                CB_LOOP
                while (!clash) {
                    const uint32_t old = cb_atomic_cas(&hkey[h], HT_EMPTY, node);
                    if (old == HT_EMPTY) break;
                    if (old == node) clash = true;
                    h = (h + 1) & (HK - 1);
                }
                if (clash) s.hdr->lift_ncomp = 1;
            }
        }
        g.sync();
        const bool junction = s.hdr->lift_ncomp != 0;
        g.sync();
        if (junction) return false;
    }
    // b. one thread per detection event: the component's smallest event walks, stores and discharges its cycle
    CB_LOOP
    for (uint32_t i = 2 * (uint32_t)g.tid(); i < M; i += 2 * (uint32_t)g.n()) {
        const uint32_t len = plain_cycle_length(s, i, E);
        if (!len) {
            // not the minimum of its component -- or a broken structure, which the minimum's own walk reports
            if (s.partner[i] < 0) s.fail(SHOT_EULER_FAILED);
            continue;
        }
        const uint32_t off = cb_atomic_add(&s.hdr->lift_cyc_top, len);
        if (off + len > E) {
            s.fail(SHOT_EULER_FAILED);
            continue;
        }
        plain_cycle_fill(s, i, cyc + off, cyc_di + off, len);
        bool done = false;
        CB_LOOP
        for (uint32_t c = 0; c < 4 && !done; c++) {
            uint64_t v;
            if (discharge_single_start(s, cyc + off, cyc_di + off, len, c, &v)) {
                if (v) cb_atomic_xor64(&s.hdr->obs, v);
                done = true;
            }
        }
        if (!done) s.fail(SHOT_DISCHARGE_FAILED);
    }
    g.sync();
    // every half-edge belongs to exactly one cycle: the cycles together hold E nodes, or some walk did not close
    if (g.tid() == 0 && !s.hdr->status && s.hdr->lift_cyc_top != E) s.fail(SHOT_EULER_FAILED);
    g.sync();
    return true;
}

template <class G>
CB_HDN void lift_shot(Shot &s, const G &g, int64_t *edge_dump, uint64_t edge_dump_cap, uint64_t *edge_dump_len) {
#ifdef CB_LIFT_PROF
    unsigned long long lp_t = cb_clock();
#define CB_LIFT_LAP(k) if (s.prof && g.tid() == 0) { unsigned long long now = cb_clock(); cb_atomic_add64((uint64_t *)&s.prof[k], now - lp_t); lp_t = now; }
#else
#define CB_LIFT_LAP(k)
#endif
    const uint32_t M = s.M;
    // edge numbering: exclusive scan of (path_len - 1) over events -> s.start
    for (uint32_t i = g.tid(); i < M; i += g.n()) s.start[i] = s.path_len[i] ? s.path_len[i] - 1 : 0;
    g.sync();
    LiftArrays L;
    L.EM = group_exclusive_scan(s, g, s.start, M);
    const uint32_t E = L.EM + M / 2;
    L.HE = 2 * E;
    const uint32_t HE = L.HE;
    // every pair has one path: at most EM - M/2 interior nodes
    L.M = M;
    L.HK = 16;
    while (L.HK < 2 * (L.EM > M / 2 ? L.EM - M / 2 : 0)) L.HK <<= 1;
    if (g.tid() == 0) {
        uint8_t *mem = s.alloc((uint64_t{HE} * 11 + 1) * 4 + 16 * 20 + (uint64_t{3} * E + 8) * 8 + uint64_t{E} * 4 + uint64_t{HE} * 2 +
                               (uint64_t{M} + 2 * uint64_t{L.HK}) * 4);
        s.hdr->scan_total = mem ? (uint32_t)((mem - s.dyn) >> 4) + 1 : 0;
        s.hdr->lift_ncomp = 0;
        s.hdr->lift_cyc_top = 0;
        if (edge_dump_len) *edge_dump_len = 2 * (uint64_t)L.EM;
    }
    g.sync();
    const uint32_t moff = s.hdr->scan_total;
    g.sync();
    if (!moff) return;
    {
        uint8_t *p = s.dyn + ((uint64_t)(moff - 1) << 4);
        uint32_t **arrs[] = {&L.own, &L.slot, &L.fillc, &L.lst, &L.pos_of, &L.uf, &L.cmin, &L.cedges, &L.comps, &L.mate};
        for (auto a : arrs) { *a = (uint32_t *)p; p += align16(uint64_t{HE} * 4); }
        L.lstart = (uint32_t *)p; p += align16((uint64_t{HE} + 1) * 4);
        L.cyc = (uint32_t *)p; p += align16((uint64_t{3} * E + 8) * 4);
        L.cyc_e = (uint32_t *)p; p += align16((uint64_t{3} * E + 8) * 4);
        L.edrag = (uint32_t *)p; p += align16(uint64_t{E} * 4);
        L.dead = p; p += align16(uint64_t{HE});
        L.simple = p; p += align16(uint64_t{HE});
        L.node_min = (uint32_t *)p; p += align16((uint64_t{M} + L.HK) * 4);
        L.hkey = (uint32_t *)p;
    }
    for (uint32_t k = g.tid(); k < M + L.HK; k += g.n()) L.node_min[k] = NO_INDEX;
    for (uint32_t k = g.tid(); k < L.HK; k += g.n()) L.hkey[k] = HT_EMPTY;
    // a. endpoints of every half-edge
    for (uint32_t i = g.tid(); i < M; i += g.n()) {
        uint32_t len = s.path_len[i];
        if (len) {
            const uint32_t *p = (const uint32_t *)(s.dyn + ((uint64_t)s.path_off[i] << 4));
            uint32_t e = s.start[i];
            for (uint32_t k = 0; k + 1 < len; k++, e++) {
                L.own[2 * e] = p[k];
                L.own[2 * e + 1] = p[k + 1];
                if (edge_dump && 2 * (uint64_t)e + 2 <= edge_dump_cap) {
                    edge_dump[2 * e] = p[k];
                    edge_dump[2 * e + 1] = p[k + 1];
                }
            }
        }
        if ((i & 1) == 0) {
            uint32_t e = L.EM + i / 2;
            L.own[2 * e] = s.event_node(i);
            L.own[2 * e + 1] = s.event_node(i + 1);
        }
    }
    g.sync();
    // b. list id of a node = its smallest half-edge
    for (uint32_t q = g.tid(); q < HE; q += g.n()) {
        const uint32_t id = lift_node_id(s, L, L.own[q], true);
        L.slot[q] = id;  // compact node id for now, the list id after the next pass
        cb_atomic_min(&L.node_min[id], q);
        L.lstart[q] = 0;
        L.fillc[q] = 0;
        L.uf[q] = q;
        L.cmin[q] = NO_INDEX;
        L.cedges[q] = 0;
        L.dead[q] = 0;
        L.mate[q] = NO_INDEX;
        L.simple[q] = 1;
    }
    g.sync();
    // c. degrees and connectivity
    for (uint32_t q = g.tid(); q < HE; q += g.n()) {
        uint32_t sl = L.node_min[L.slot[q]];
        L.slot[q] = sl;
        cb_atomic_add(&L.lstart[sl], 1);
    }
    g.sync();
    for (uint32_t e = g.tid(); e < E; e += g.n()) {
        uf_union(L.uf, L.slot[2 * e], L.slot[2 * e + 1]);
        // the drag-table entry the discharge will need for this edge (decoder.cc:104-127: mmm.find of the detector pair)
        const uint32_t a = L.own[2 * e], b = L.own[2 * e + 1];
        uint32_t di;
        if ((a >> 1) == (b >> 1)) {
            di = s.T.self_drag[a >> 1];
        } else {
            const uint32_t he = find_half_edge(s.T, a, b);
            di = he != NO_INDEX ? s.T.adj_drag[he] : NO_INDEX;
        }
        L.edrag[e] = di;
    }
    g.sync();
    CB_LIFT_LAP(12)
    // d. list offsets
    uint32_t total = group_exclusive_scan(s, g, L.lstart, HE);
    if (g.tid() == 0) L.lstart[HE] = total;
    g.sync();
    // e. fill
    for (uint32_t q = g.tid(); q < HE; q += g.n()) {
        uint32_t sl = L.slot[q];
        uint32_t p = L.lstart[sl] + cb_atomic_add(&L.fillc[sl], 1);
        L.lst[p] = q;
    }
    g.sync();
    // f. restore edge-list order inside every list; component bookkeeping
    for (uint32_t q = g.tid(); q < HE; q += g.n()) {
        if (L.slot[q] == q) {
            uint32_t a0 = L.lstart[q], a1 = L.lstart[q + 1];
            for (uint32_t a = a0 + 1; a < a1; a++) {
                uint32_t v = L.lst[a];
                uint32_t b = a;
                while (b > a0 && L.lst[b - 1] > v) {
                    L.lst[b] = L.lst[b - 1];
                    b--;
                }
                L.lst[b] = v;
            }
            for (uint32_t a = a0; a < a1; a++) L.pos_of[L.lst[a]] = a;
            L.fillc[q] = 0;  // becomes next_neighbor
            if (a1 - a0 == 2) {
                L.mate[L.lst[a0]] = L.lst[a0 + 1];
                L.mate[L.lst[a0 + 1]] = L.lst[a0];
            } else {
                L.simple[uf_find(L.uf, q)] = 0;
            }
        }
        uint32_t r = uf_find(L.uf, L.slot[q]);
        if (q < 2 * L.EM) cb_atomic_min(&L.cmin[r], q);
        if ((q & 1) == 0) cb_atomic_add(&L.cedges[r], 1);
    }
    g.sync();
    // g. component list and cycle buffers
    for (uint32_t q = g.tid(); q < HE; q += g.n())
        if (L.slot[q] == q && uf_find(L.uf, q) == q) {
            uint32_t k = cb_atomic_add(&s.hdr->lift_ncomp, 1);
            L.comps[k] = q;
        }
    g.sync();
    const uint32_t ncomp = s.hdr->lift_ncomp;
    CB_LIFT_LAP(13)
    // h. one thread per component
    for (uint32_t c = g.tid(); c < ncomp; c += g.n()) {
        uint32_t root = L.comps[c];
        uint32_t off = cb_atomic_add(&s.hdr->lift_cyc_top, L.cedges[root] + 2);
        if (L.cmin[root] == NO_INDEX) {
            s.fail(SHOT_EULER_FAILED);
            continue;
        }
        lift_component(s, L, root, L.cyc + off, L.cyc_e + off);
    }
    g.sync();
    CB_LIFT_LAP(14)
    CB_LIFT_LAP(15)
}

// One shot, start to finish. `shot` indexes the caller's arrays.
template <class G, class PT = PhaseTimer>
CB_HD void decode_shot(Shot &s, const DecodeParams &p, uint32_t shot, const G &g) {
    PT pt;
    pt.start();
    ingest_shot(s, p.dets + (uint64_t)shot * p.stride, g);
    pt.lap(0);
    uint32_t rounds = 0;
    if (!s.hdr->status && s.M) {
        rounds = match_shot(s, g, pt);
        if (!s.hdr->status) paths_shot(s, g);
    }
    g.sync();
    pt.lap(5);
    const bool liftable = !s.hdr->status && s.M;
    g.sync();
    if (liftable) {
        // the plain-cycle lift covers shots without junctions; the edge dump (a debugging hook) needs the general one
        const bool plain = !p.edge_dump && !p.edge_dump_len && !(s.cfg.pad & 1u) && lift_shot_plain(s, g);
        if (!plain) lift_shot(s, g, p.edge_dump, p.edge_dump_cap, p.edge_dump_len);
    } else if (g.tid() == 0 && p.edge_dump_len) {
        *p.edge_dump_len = 0;
    }
    g.sync();
    if (g.tid() == 0) {
        pt.lap(6);
        if (p.phase_cycles) {
            pt.t[3] = pt.t[8] + pt.t[9] + pt.t[10];
            for (int k = 0; k < 7; k++) cb_atomic_add64((uint64_t *)&p.phase_cycles[k], pt.t[k]);
            cb_atomic_add64((uint64_t *)&p.phase_cycles[7], 1);
#ifndef CB_SUBPROF
            for (int k = 8; k < 11; k++) cb_atomic_add64((uint64_t *)&p.phase_cycles[8 + k], pt.t[k]);
#endif
#if defined(CB_TAIL_PROF) && defined(__CUDA_ARCH__)
            // diagnostic build: slot 19 = longest shot (cycles), slot 23 = shots longer than 8 M cycles (~4 ms)
            unsigned long long tot_ = 0;
            for (int k = 0; k < 7; k++) tot_ += pt.t[k];
            atomicMax((unsigned long long *)&p.phase_cycles[19], tot_);
            if (tot_ > 8000000ull) cb_atomic_add64((uint64_t *)&p.phase_cycles[23], 1);
#endif
        }
        uint32_t st = s.hdr->status;
        p.status_out[shot] = st;
        p.obs_out[shot] = st ? 0 : s.hdr->obs;
        if (p.weight_out) p.weight_out[shot] = st ? 0 : (int64_t)s.hdr->weight;
#if defined(CB_TAIL_PROF) && defined(__CUDA_ARCH__)
        if (p.weight_out && p.phase_cycles) {  // diagnostic build: (cycles >> 10) | largest component << 32 | rounds << 56
            unsigned long long tot_ = 0;
            for (int k = 0; k < 7; k++) tot_ += pt.t[k];
            p.weight_out[shot] = (int64_t)(((tot_ >> 10) & 0xFFFFFFFFull) | ((unsigned long long)(s.hdr->pad_hdr[0] & 0xFFFFFFu) << 32) |
                                           ((unsigned long long)(rounds & 0x7Fu) << 56));
        }
#endif
        if (p.rounds_out) p.rounds_out[shot] = rounds;
    }
    g.sync();
}

}  // namespace cb200

#endif
