// Host "simulation" build of the device decode path (chromobius_b200/csrc/decode_core.h) with a group of one thread.
// TEST INFRASTRUCTURE ONLY: lets the CPU-only test tier exercise the exact per-shot logic the CUDA kernels run.
// The product library never links or loads this file.
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../chromobius_b200/csrc/decode_core.h"
#include "../../chromobius_b200/csrc/host_compile.h"

using namespace cb200;

struct Sim {
    HostTables tables;
    std::vector<uint8_t> arena;
    ScratchConfig cfg{};
};

static thread_local std::string g_err;

extern "C" {

const char *cbsim_last_error() { return g_err.c_str(); }

void *cbsim_compile(const char *dem_text) {
    try {
        auto *s = new Sim();
        s->tables = compile_tables_cached(dem_text);
        return s;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void cbsim_free(void *h) { delete (Sim *)h; }

// Visit order of n distinct keys inserted in sequence: the compile step's restatement (which = 0) or the real
// std::unordered_map<uint32_t, uint64_t> of this toolchain's libstdc++ (which = 1), the container the reference uses.
void cbsim_container_visit_order(const uint32_t *keys, uint64_t n, int which, uint32_t *out) {
    std::vector<uint32_t> seq(keys, keys + n);
    if (which == 0) {
        auto order = reference_container_visit_order(seq);
        memcpy(out, order.data(), n * 4);
    } else {
        std::unordered_map<uint32_t, uint64_t> m;
        for (uint32_t k : seq) m[k] = k;
        uint64_t i = 0;
        for (const auto &kv : m) out[i++] = kv.first;
    }
}
uint64_t cbsim_num_detectors(void *h) { return ((Sim *)h)->tables.num_dets; }
uint64_t cbsim_num_observables(void *h) { return ((Sim *)h)->tables.num_obs; }
double cbsim_normalising_constant(void *h) { return ((Sim *)h)->tables.normalising_constant; }
static char *dup_string(const std::string &s) {
    char *r = (char *)malloc(s.size() + 1);
    memcpy(r, s.c_str(), s.size() + 1);
    return r;
}
char *cbsim_describe(void *h) { return dup_string(((Sim *)h)->tables.describe()); }
char *cbsim_mobius_dem(void *h) { return dup_string(((Sim *)h)->tables.mobius_dem_text()); }
void cbsim_free_string(char *s) { free(s); }
uint64_t cbsim_graph_half_edges(void *h) { return ((Sim *)h)->tables.adj.size(); }
void cbsim_graph_export(void *h, uint32_t *row_ptr, uint32_t *nbr, int32_t *wt) {
    auto &t = ((Sim *)h)->tables;
    memcpy(row_ptr, t.row_ptr.data(), t.row_ptr.size() * 4);
    for (size_t k = 0; k < t.adj.size(); k++) {
        nbr[k] = t.adj[k].nbr;
        wt[k] = t.adj[k].wt;
    }
}

// Decodes shots one by one. Returns 0. status_out[k] != 0 marks failed shots.
// edges_out (nullable): matched edges of the LAST shot (interleaved), *num_edge_ints receives the count.
int cbsim_decode(void *h, const uint8_t *dets, uint64_t num_shots, uint64_t stride, uint64_t *obs_out, int64_t *weight_out,
                 uint32_t *status_out, uint32_t *rounds_out, uint32_t ht_log2, uint32_t sh_ht_log2, uint64_t dyn_bytes, int64_t *edges_out,
                 uint64_t edges_cap, uint64_t *num_edge_ints) {
    auto *s = (Sim *)h;
    DecodeParams p{};
    p.T = s->tables.view();
    p.dets = dets;
    p.stride = stride;
    p.num_shots = (uint32_t)num_shots;
    p.obs_out = obs_out;
    p.weight_out = weight_out;
    p.status_out = status_out;
    p.rounds_out = rounds_out;
    p.edge_dump = edges_out;
    p.edge_dump_cap = edges_cap;
    p.edge_dump_len = num_edge_ints;
    p.cfg.sh_ht_log2 = sh_ht_log2;
    p.cfg.sh_found = 8;
    p.cfg.sh_pool = 20u << 10;  // room for one fixed-layout blossom of any class; larger components use the run-time layout
    if (const char *env = getenv("CBSIM_POOL_KB")) p.cfg.sh_pool = (uint32_t)strtoul(env, nullptr, 10) << 10;
    p.cfg.ht_log2 = ht_log2;
    p.cfg.found_cap = 256;
    p.cfg.max_events = 2 * s->tables.num_dets + 2;
    p.cfg.dyn_bytes = dyn_bytes;
    if (getenv("CB200_NO_PLAIN_LIFT")) p.cfg.pad |= 1u;  // tests: always the general lift
    uint64_t bytes = arena_bytes_per_group(p.cfg, p.T.num_nodes, 1);
    s->arena.assign(bytes, 0);
    std::vector<uint8_t> shared(shared_bytes_per_group(p.cfg, p.T.num_dets, 1) + 16, 0);
    p.arena = s->arena.data();
    p.arena_stride = bytes;
    HostGroup g;
    Shot shot;
    bind_shot(shot, p, p.arena, shared.data(), g);
    init_group_arena(shot, g);
    init_group_shared(shot, g);
    for (uint32_t k = 0; k < num_shots; k++) decode_shot(shot, p, k, g);
    return 0;
}

}  // extern "C"
