// Host side of the decoder: DEM text -> flat tables (runs once per DEM, never on the GPU).
// Mirrors the reference's Decoder::from_dem pipeline (src/chromobius/decode/decoder.cc:24-69).
#ifndef CHROMOBIUS_B200_HOST_COMPILE_H
#define CHROMOBIUS_B200_HOST_COMPILE_H

#include <cstdint>
#include <string>
#include <vector>

#include "tables.h"

namespace cb200 {

// A DEM error after `repeat` unrolling and `shift_detectors` application.
struct FlatDemError {
    double p;
    uint32_t first_target;  // into FlatDem::targets
    uint32_t num_targets;
    uint32_t line;          // source line, for messages
};
// Target encoding: detector d -> d ; observable k -> OBS_FLAG | k ; separator -> SEPARATOR
constexpr uint64_t TGT_OBS_FLAG = 1ull << 62;
constexpr uint64_t TGT_SEPARATOR = 1ull << 63;

struct FlatDetector {
    uint64_t id;
    uint32_t num_coords;
    double coords[8];       // shifted; only the first 8 are kept
    uint32_t line;
};

struct FlatDem {
    uint64_t num_detectors = 0;
    uint64_t num_observables = 0;
    std::vector<FlatDemError> errors;
    std::vector<uint64_t> targets;
    std::vector<FlatDetector> detectors;
    std::vector<std::string> lines;  // trimmed source lines for error messages

    // Throws std::invalid_argument on malformed text.
    static FlatDem parse(const std::string &text);
};

struct MobiusDemError {
    double p;
    uint32_t first_edge;  // into HostTables::mobius_edges (pairs)
    uint32_t num_edges;
};

struct RgbRep {
    uint32_t node[3];  // red, green, blue (NO_INDEX if absent)
    uint64_t obs_flip;
    uint8_t charge_flip;
};

struct HostTables {
    uint32_t num_dets = 0;
    uint32_t num_obs = 0;
    double normalising_constant = 1.0;

    // Reference-shaped intermediate results (kept for describe_decoder and for parity tests).
    std::vector<uint8_t> color, basis, ignored;            // node_colors
    std::vector<uint32_t> atom_dets;                        // 3 per atomic error (NO_INDEX padded), map order
    std::vector<uint64_t> atom_obs;
    std::vector<MobiusDemError> mobius_errors;
    std::vector<uint32_t> mobius_edges;                     // interleaved pairs
    std::vector<RgbRep> rgb_reps;
    std::vector<std::vector<std::pair<uint32_t, uint64_t>>> charge_graph;  // per node, sorted by neighbour
    uint64_t drag_ambiguities = 0;

    // Device-shaped tables.
    std::vector<uint32_t> row_ptr;
    std::vector<AdjEntry> adj;
    std::vector<AdjEntry> ell_adj;   // fixed-stride copy of adj (1 << ell_shift entries per node), see TablesView
    uint32_t ell_shift = 0;
    int32_t min_wt = 0;              // smallest edge weight (0 when the graph has no edges)
    std::vector<uint32_t> adj_drag;
    std::vector<uint8_t> det_info;
    std::vector<uint64_t> rep_obs;
    std::vector<uint32_t> self_drag;
    std::vector<uint8_t> ingest_mask;
    std::vector<DragEntry> drag;
    std::vector<uint32_t> drag_pair;  // (n1, n2) per DragEntry, n1 <= n2
    std::vector<NearEntry> near;      // neighbourhood table, see tables.h
    uint32_t near_stride = 0;

    TablesView view() const;

    // Same text layout as the reference's `chromobius describe_decoder` for the charge graph, rgb reps and
    // drag graph sections (decode/decoder.cc:279-291); the Moebius DEM is printed as plain `error(p) Da Db ^ ...` lines.
    std::string describe() const;
    std::string mobius_dem_text() const;
};

// Throws std::invalid_argument with the reference's messages where tests pin them.
// near_stride: entries per node of the neighbourhood table (multiple of 32, <= NEAR_MAX_STRIDE); 0 = no table;
// NEAR_STRIDE_AUTO sizes it from a byte budget (environment CB200_NEAR_STRIDE overrides).
constexpr uint32_t NEAR_STRIDE_AUTO = 0xFFFFFFFFu;
HostTables compile_tables(const FlatDem &dem, uint32_t near_stride = NEAR_STRIDE_AUTO);

// Test hook: the order in which the reference's charge-graph container (libstdc++ std::unordered_map<uint32_t, uint64_t>,
// src/chromobius/graph/charge_graph.h:52) iterates distinct keys inserted in the given sequence, as restated by the
// compile step (the drag-table search depends on it, src/chromobius/graph/drag_graph.cc:87).
std::vector<uint32_t> reference_container_visit_order(const std::vector<uint32_t> &keys_in_insertion_order);
// (Re)builds t.near for the given stride from t.row_ptr / t.adj (multi-threaded).
void build_near_table(HostTables &t, uint32_t near_stride);

// DEM text -> tables, through an on-disk cache when the environment variable CHROMOBIUS_B200_CACHE names a directory
// (the compile step is paid once per DEM *per worker process*; sinter starts many). The cache key is a
// hash of the DEM text, the table format version and the near-table setting; a file that fails its checks is ignored
// and rewritten. `from_cache` (nullable) reports whether the tables were loaded.
HostTables compile_tables_cached(const std::string &dem_text, bool *from_cache = nullptr);
void save_tables(const HostTables &t, const std::string &path, uint64_t key, const std::string &dem_text);
bool load_tables(HostTables &t, const std::string &path, uint64_t key, const std::string &dem_text);

}  // namespace cb200

#endif
