// Flat, read-only decoder tables shared by the host compile step and the CUDA kernels.
//
// These replace the reference's pointer-rich containers with arrays that can be uploaded once and stay
// L2-resident:
//   node_colors / rgb_reps        -> det_info[D] + rep_obs[D]     (decode/decoder.h:66-74)
//   drag_graph.mmm (std::map)     -> DragEntry[] indexed per CSR half-edge / per detector (graph/drag_graph.h:78)
//   PyMatching matching graph     -> CSR of the Moebius graph with even integer weights (decode/pymatcher.cc:22-25)
#ifndef CHROMOBIUS_B200_TABLES_H
#define CHROMOBIUS_B200_TABLES_H

#include <cstdint>

namespace cb200 {

constexpr uint32_t NO_INDEX = 0xFFFFFFFFu;

// det_info bit layout
constexpr uint8_t DET_COLOR_MASK = 0x03;   // Charge: 0 NEUTRAL, 1 R, 2 G, 3 B
constexpr uint8_t DET_REP_WEIGHT3 = 0x04;  // rgb_reps[d].weight() == 3
constexpr uint8_t DET_BASIS_SHIFT = 4;     // bits 4-5: 0 unknown, 1 X, 2 Z
constexpr uint8_t DET_IGNORED = 0x80;

// One entry per unordered detector pair {n1 <= n2} that appears in the drag graph (n1 == n2 allowed).
// valid bit (4*c1 + c2) set  <=>  mmm[{n1, n2, c1, c2}] exists, with value obs[4*c1 + c2].
// The reference always inserts {n1,n2,c1,c2} and {n2,n1,c2,c1} together with the same value
// (graph/drag_graph.cc:117-120), so walking the pair in the n2 -> n1 direction uses bit (4*c2 + c1).
struct DragEntry {
    uint64_t obs[16];
    uint32_t valid;
    uint32_t pad[3];
};
static_assert(sizeof(DragEntry) == 144, "DragEntry layout");

// Adjacency entry of the Moebius matching graph: neighbour node and even integer weight.
struct AdjEntry {
    uint32_t nbr;
    int32_t wt;
};

// Neighbourhood table ("near table"): for every Moebius node the NEAR_K closest other nodes in (distance, node id)
// order -- exactly the order in which a Dijkstra ball around the node settles them -- with the exact distance and the
// canonical shortest-path predecessor. A ball query that stays inside the table is one or two coalesced row reads and a
// ballot against the shot's event bits instead of a Dijkstra search; queries that leave it fall back to the search.
//   node_pred : bits 0-23 node id, bits 24-31 position (within the same row) of the canonical predecessor on the
//               shortest path from the row's node, NEAR_PRED_SRC when the predecessor is the row's node itself,
//               NEAR_PRED_FAR when its position does not fit 8 bits (paths through it are searched instead).
//               Canonical = the smallest-id neighbour w of the entry with dist(w) + wt(w, entry) == dist(entry).
//   dist_flag : bits 0-30 exact distance, bit 31 set when the row's node and the entry are joined by a direct edge of
//               exactly that weight. Rows shorter than the stride are padded with dist_flag == NEAR_PAD.
// Layout: chunk-major, 32 entries per chunk: entry p of node u sits at ((p >> 5) * num_nodes + u) * 32 + (p & 31), so
// the first chunks of all rows (the hot part) are contiguous.
struct NearEntry {
    uint32_t node_pred;
    uint32_t dist_flag;
};
constexpr uint32_t NEAR_PAD = 0xFFFFFFFFu;
constexpr uint32_t NEAR_PRED_SRC = 0xFFu;
constexpr uint32_t NEAR_PRED_FAR = 0xFEu;
constexpr uint32_t NEAR_MAX_STRIDE = 512;

// Views used by the per-shot code (host build and device build alike).
struct TablesView {
    uint32_t num_dets;
    uint32_t num_nodes;       // 2 * num_dets
    uint32_t num_obs;
    uint32_t num_det_bytes;   // ceil(num_dets / 8)
    const uint32_t *row_ptr;  // [num_nodes + 1]
    const AdjEntry *adj;      // [row_ptr[num_nodes]] sorted by nbr within a row
    const AdjEntry *ell_adj;  // [num_nodes << ell_shift] the same rows at a fixed stride, padded with nbr == NO_INDEX
    uint32_t ell_shift;       // 0: no fixed-stride copy (maximum degree above 32)
    int32_t min_wt;           // smallest edge weight of the graph (> 0)
    const uint32_t *adj_drag; // [row_ptr[num_nodes]] DragEntry index of the detector pair of this half-edge
    const uint8_t *det_info;  // [num_dets]
    const uint64_t *rep_obs;  // [num_dets] rgb_reps[d].obs_flip
    const uint32_t *self_drag;// [num_dets] DragEntry index of {d, d} or NO_INDEX
    const uint8_t *ingest_mask; // [num_det_bytes] 1-bits for detectors that exist and are not ignored
    const DragEntry *drag;    // [num_drag]
    const NearEntry *near;    // [near_stride * num_nodes] chunk-major (nullable)
    uint32_t near_stride;     // entries per node, a multiple of 32 (0: no near table)
};

}  // namespace cb200

#endif
