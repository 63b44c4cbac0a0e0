// chromobius_oracle.cc -- CPU restatement of the reference decoder.  TEST INFRASTRUCTURE ONLY.
//
// This file is the parity oracle for the B200 decode path. Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / `--impl reference` legs may load it; the product library never does.
//
// It restates, function by function, the reference's per-shot path and the compile passes that
// define its tables (citations are relative to /root/reference/src/chromobius):
//
//   parse / flatten DEM text         stim DetectorErrorModel (not in tree; grammar per SURVEY App. C.1)
//   collect_nodes                    graph/collect_nodes.cc:21-131, datatypes/stim_integration.cc:21-67
//   collect_atomic_errors            graph/collect_atomic_errors.cc:19-115
//   collect_composite_errors         graph/collect_composite_errors.cc:21-399
//   mobius edges of an atom          datatypes/atomic_error.h:104-140
//   choose_rgb_reps                  graph/choose_rgb_reps.cc:19-92
//   charge graph                     graph/charge_graph.cc:80-160
//   drag graph                       graph/drag_graph.cc:24-203
//   matcher graph + integer weights  decode/pymatcher.cc:22-25 -> PyMatching v2.2.2 (NOT in tree; restated
//                                    from its published behaviour)
//   lift-in                          decode/decoder.cc:231-247
//   match                            decode/pymatcher.cc:27-42 -> exact MWPM (own dense blossom, see below)
//   euler tours                      graph/euler_tours.h:34-112, graph/euler_tours.cc:21-97
//   discharge                        decode/decoder.cc:80-203
//   decode_detection_events          decode/decoder.cc:249-277
//
// PARITY STATUS. Tables, Euler tours, discharge and end-to-end predictions are pinned against the
// reference's own golden tests (tests/test_oracle_golden.py). The matcher is NOT PyMatching: it is an
// exact minimum-weight perfect matching (all-pairs truncated Dijkstra + O(n^3) primal-dual blossom with a
// dual-feasibility certificate for the truncated pairs), pinned against networkx.max_weight_matching and
// brute force. The integer edge weights follow PyMatching's discretisation as recalled (App. B); that
// discretisation is "parity unpinned" because PyMatching is not available in this image.
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <queue>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace orc {

typedef uint64_t obsmask;
typedef uint32_t node_t;
constexpr node_t BOUNDARY = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// DEM text model (restates the subset of stim::DetectorErrorModel the reference uses).
// ---------------------------------------------------------------------------------------------
enum class InsType { ERROR, DETECTOR, SHIFT, REPEAT, OBSERVABLE };
struct Target {
    // kind: 0 detector, 1 observable, 2 separator, 3 raw integer (shift amount)
    int kind;
    uint64_t val;
};
struct Instruction {
    InsType type;
    std::vector<double> args;
    std::vector<Target> targets;
    uint64_t reps = 0;
    std::vector<Instruction> body;
    std::string text;
};

static std::string trim(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

static std::vector<Instruction> parse_block(const std::vector<std::string> &lines, size_t &pos, bool nested) {
    std::vector<Instruction> out;
    while (pos < lines.size()) {
        std::string line = lines[pos++];
        size_t hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        line = trim(line);
        if (line.empty()) continue;
        if (line == "}") {
            if (!nested) throw std::invalid_argument("Unmatched '}' in detector error model.");
            return out;
        }
        Instruction ins;
        ins.text = line;
        size_t k = 0;
        while (k < line.size() && (isalpha((unsigned char)line[k]) || line[k] == '_')) k++;
        std::string name = line.substr(0, k);
        for (auto &c : name) c = (char)tolower(c);
        // optional tag [..]
        if (k < line.size() && line[k] == '[') {
            size_t e = line.find(']', k);
            if (e == std::string::npos) throw std::invalid_argument("Bad tag in: " + line);
            k = e + 1;
        }
        if (k < line.size() && line[k] == '(') {
            size_t e = line.find(')', k);
            if (e == std::string::npos) throw std::invalid_argument("Missing ')' in: " + line);
            std::string inside = line.substr(k + 1, e - k - 1);
            std::stringstream ss(inside);
            std::string item;
            while (std::getline(ss, item, ',')) {
                item = trim(item);
                if (item.empty()) continue;
                char *endp = nullptr;
                double v = strtod(item.c_str(), &endp);
                if (endp == item.c_str() || *endp != 0) throw std::invalid_argument("Bad number in: " + line);
                ins.args.push_back(v);
            }
            k = e + 1;
        }
        std::string rest = trim(line.substr(k));
        if (name == "repeat") {
            if (rest.empty() || rest.back() != '{') throw std::invalid_argument("Expected '{' in: " + line);
            rest = trim(rest.substr(0, rest.size() - 1));
            ins.type = InsType::REPEAT;
            ins.reps = strtoull(rest.c_str(), nullptr, 10);
            ins.body = parse_block(lines, pos, true);
            out.push_back(std::move(ins));
            continue;
        }
        if (name == "error") ins.type = InsType::ERROR;
        else if (name == "detector") ins.type = InsType::DETECTOR;
        else if (name == "shift_detectors") ins.type = InsType::SHIFT;
        else if (name == "logical_observable") ins.type = InsType::OBSERVABLE;
        else throw std::invalid_argument("Unrecognized instruction type: " + line);
        std::stringstream ts(rest);
        std::string tok;
        while (ts >> tok) {
            Target t;
            if (tok == "^") {
                t = {2, 0};
            } else if (tok[0] == 'D' || tok[0] == 'd') {
                t = {0, strtoull(tok.c_str() + 1, nullptr, 10)};
            } else if (tok[0] == 'L' || tok[0] == 'l') {
                t = {1, strtoull(tok.c_str() + 1, nullptr, 10)};
            } else if (isdigit((unsigned char)tok[0])) {
                t = {3, strtoull(tok.c_str(), nullptr, 10)};
            } else {
                throw std::invalid_argument("Unrecognized target '" + tok + "' in: " + line);
            }
            ins.targets.push_back(t);
        }
        if (ins.type == InsType::ERROR && ins.args.size() != 1)
            throw std::invalid_argument("error instruction needs one probability: " + line);
        out.push_back(std::move(ins));
    }
    if (nested) throw std::invalid_argument("Missing '}' in detector error model.");
    return out;
}

struct Dem {
    std::vector<Instruction> instructions;
    static Dem parse(const std::string &text) {
        std::vector<std::string> lines;
        std::stringstream ss(text);
        std::string l;
        while (std::getline(ss, l)) lines.push_back(l);
        size_t pos = 0;
        Dem d;
        d.instructions = parse_block(lines, pos, false);
        return d;
    }
};

struct FlatError {
    double p;
    std::vector<Target> targets;  // detector ids absolute
    std::string text;
};

static void flatten_block(
    const std::vector<Instruction> &block, uint64_t &det_offset, std::vector<FlatError> *errors, uint64_t *max_det,
    uint64_t *max_obs, bool *any_det, bool *any_obs) {
    for (const auto &ins : block) {
        switch (ins.type) {
            case InsType::ERROR: {
                FlatError fe;
                fe.p = ins.args[0];
                fe.text = ins.text;
                for (auto t : ins.targets) {
                    if (t.kind == 0) {
                        t.val += det_offset;
                        *max_det = std::max(*max_det, t.val);
                        *any_det = true;
                    } else if (t.kind == 1) {
                        *max_obs = std::max(*max_obs, t.val);
                        *any_obs = true;
                    }
                    fe.targets.push_back(t);
                }
                if (errors) errors->push_back(std::move(fe));
                break;
            }
            case InsType::DETECTOR:
                for (auto t : ins.targets) {
                    if (t.kind == 0) {
                        *max_det = std::max(*max_det, t.val + det_offset);
                        *any_det = true;
                    }
                }
                break;
            case InsType::OBSERVABLE:
                for (auto t : ins.targets) {
                    if (t.kind == 1) {
                        *max_obs = std::max(*max_obs, t.val);
                        *any_obs = true;
                    }
                }
                break;
            case InsType::SHIFT:
                if (!ins.targets.empty()) det_offset += ins.targets[0].val;
                break;
            case InsType::REPEAT:
                for (uint64_t r = 0; r < ins.reps; r++)
                    flatten_block(ins.body, det_offset, errors, max_det, max_obs, any_det, any_obs);
                break;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Datatypes (datatypes/color_basis.h:28-72, atomic_error.h:28-140, rgb_edge.h:32-55)
// ---------------------------------------------------------------------------------------------
enum Charge : uint8_t { NEUTRAL = 0, R = 1, G = 2, B = 3 };
enum Basis : uint8_t { UNKNOWN_BASIS = 0, X = 1, Z = 2 };
struct ColorBasis {
    uint8_t color = 0;
    uint8_t basis = 0;
    bool ignored = false;
};
static inline uint8_t next_non_neutral(uint8_t c) { return (uint8_t)(c % 3 + 1); }

struct AtomKey {
    std::array<node_t, 3> d;
    AtomKey() : d{BOUNDARY, BOUNDARY, BOUNDARY} {}
    AtomKey(node_t a, node_t b, node_t c) : d{a, b, c} { std::sort(d.begin(), d.end()); }
    explicit AtomKey(const std::vector<node_t> &v) : d{BOUNDARY, BOUNDARY, BOUNDARY} {
        for (size_t k = 0; k < v.size() && k < 3; k++) d[k] = v[k];
        std::sort(d.begin(), d.end());
    }
    bool operator<(const AtomKey &o) const { return d < o.d; }
    bool operator==(const AtomKey &o) const { return d == o.d; }
    int weight() const { return (d[0] != BOUNDARY) + (d[1] != BOUNDARY) + (d[2] != BOUNDARY); }
    uint8_t net_charge(const std::vector<ColorBasis> &nc) const {
        uint8_t c = 0;
        for (auto x : d)
            if (x != BOUNDARY) c ^= nc[x].color;
        return c;
    }
};
typedef std::map<AtomKey, obsmask> AtomMap;
// charge_graph.h:37-53: adjacency as the reference's own container type (its iteration order matters, see short_path_flip).
typedef std::vector<std::unordered_map<node_t, obsmask>> ChargeGraph;

struct RgbRep {
    node_t rgb[3] = {BOUNDARY, BOUNDARY, BOUNDARY};
    obsmask obs_flip = 0;
    uint8_t charge_flip = 0;
    int weight() const { return (rgb[0] != BOUNDARY) + (rgb[1] != BOUNDARY) + (rgb[2] != BOUNDARY); }
    node_t color_node(uint8_t c) const { return c == 0 ? BOUNDARY : rgb[c - 1]; }
};

// atomic_error.h:104-140 -- the Moebius doubling rule.
static void atom_mobius_edges(const AtomKey &a, const std::vector<ColorBasis> &nc, std::vector<std::pair<node_t, node_t>> *out) {
    node_t n1 = a.d[0], n2 = a.d[1], n3 = a.d[2];
    if (n1 == BOUNDARY) return;
    if (n2 == BOUNDARY) {
        out->push_back({n1 * 2, n1 * 2 + 1});
    } else if (n3 == BOUNDARY) {
        uint32_t flip = ((nc[n1].color ^ nc[n2].color) == G) ? 1u : 0u;
        out->push_back({n1 * 2 + 0, (n2 * 2 + 0) ^ flip});
        out->push_back({n1 * 2 + 1, (n2 * 2 + 1) ^ flip});
    } else {
        node_t rgb[3] = {BOUNDARY, BOUNDARY, BOUNDARY};
        rgb[nc[n1].color - 1] = n1;
        rgb[nc[n2].color - 1] = n2;
        rgb[nc[n3].color - 1] = n3;
        node_t r = rgb[0], g = rgb[1], b = rgb[2];
        // subgraph offsets (color_basis.h:55-60): Red:{NotGreen:0, NotBlue:1} Green:{NotRed:0, NotBlue:1} Blue:{NotRed:0, NotGreen:1}
        std::pair<node_t, node_t> e0{r * 2 + 1, g * 2 + 1};  // NotBlue subgraph
        std::pair<node_t, node_t> e1{g * 2 + 0, b * 2 + 0};  // NotRed subgraph
        std::pair<node_t, node_t> e2{r * 2 + 0, b * 2 + 1};  // NotGreen subgraph
        for (auto *e : {&e0, &e1, &e2}) {
            if (e->first > e->second) std::swap(e->first, e->second);
            out->push_back(*e);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Compile passes
// ---------------------------------------------------------------------------------------------
static const char *COLOR_MSG =
    "Expected all detectors to have at least 4 coordinates, with the 4th identifying the basis and color "
    "(RedX=0, GreenX=1, BlueX=2, RedZ=3, GreenZ=4, BlueZ=5), but got ";

// stim_integration.cc:21-67
static ColorBasis detector_to_color_basis(const Instruction &ins, const std::vector<double> &coord_offsets) {
    bool failed = false;
    double c = -2;
    if (ins.args.size() > 3) {
        c = ins.args[3];
        if (coord_offsets.size() > 3) c += coord_offsets[3];
    } else {
        failed = true;
    }
    int r = 0;
    if (c < -1 || c > 5) {
        failed = true;
    } else {
        r = (int)c;
        if (r != c) failed = true;
    }
    if (failed) throw std::invalid_argument(std::string(COLOR_MSG) + ins.text);
    if (r == -1) return ColorBasis{NEUTRAL, UNKNOWN_BASIS, true};
    return ColorBasis{(uint8_t)(r % 3 + 1), (uint8_t)(r / 3 + 1), false};
}

// collect_nodes.cc:73-131
static void collect_nodes_helper(
    const std::vector<Instruction> &block, uint64_t *det_offset, std::vector<double> *coord_offsets,
    std::vector<ColorBasis> *out) {
    for (const auto &ins : block) {
        switch (ins.type) {
            case InsType::DETECTOR: {
                ColorBasis cb = detector_to_color_basis(ins, *coord_offsets);
                for (const auto &t : ins.targets) (*out)[t.val + *det_offset] = cb;
                break;
            }
            case InsType::SHIFT:
                if (!ins.targets.empty()) *det_offset += ins.targets[0].val;
                for (size_t k = 0; k < ins.args.size(); k++) {
                    if (coord_offsets->size() < ins.args.size()) coord_offsets->push_back(0);
                    (*coord_offsets)[k] += ins.args[k];
                }
                break;
            case InsType::REPEAT:
                for (uint64_t k = 0; k < ins.reps; k++) collect_nodes_helper(ins.body, det_offset, coord_offsets, out);
                break;
            default:
                break;
        }
    }
}

// collect_atomic_errors.cc:57-100
static void extract_obs_and_dets(
    const FlatError &e, const std::vector<ColorBasis> &nc, std::vector<node_t> *dets, obsmask *obs) {
    dets->clear();
    *obs = 0;
    for (const auto &t : e.targets) {
        if (t.kind == 0) {
            if (t.val > 0xFFFFFFFFull) throw std::invalid_argument("The detector error model is too large.");
            if (!nc[t.val].ignored) {
                if (nc[t.val].color == NEUTRAL) throw std::invalid_argument(std::string(COLOR_MSG) + e.text);
                dets->push_back((node_t)t.val);
            }
        } else if (t.kind == 1) {
            if (t.val >= 64) {
                throw std::invalid_argument("Max logical observable is L63 but a larger one appeared in '" + e.text + "'");
            }
            *obs ^= obsmask{1} << t.val;
        }
    }
    // SparseXorVec semantics: sorted, pairs cancel.
    std::sort(dets->begin(), dets->end());
    size_t n = 0;
    for (size_t k = 0; k < dets->size(); k++) {
        if (n > 0 && (*dets)[n - 1] == (*dets)[k]) n--;
        else (*dets)[n++] = (*dets)[k];
    }
    dets->resize(n);
}

// collect_atomic_errors.cc:19-55
static AtomKey extract_atomic(const std::vector<node_t> &dets, obsmask obs, const std::vector<ColorBasis> &nc, AtomMap *out) {
    if (dets.size() == 1) {
        AtomKey key(dets[0], BOUNDARY, BOUNDARY);
        (*out)[key] = obs;
        return key;
    }
    if (dets.size() == 2) {
        if (nc[dets[0]].basis == nc[dets[1]].basis) {
            AtomKey key(dets[0], dets[1], BOUNDARY);
            (*out)[key] = obs;
            return key;
        }
        return AtomKey();
    }
    if (dets.size() == 3) {
        auto c0 = nc[dets[0]], c1 = nc[dets[1]], c2 = nc[dets[2]];
        if ((c0.color ^ c1.color ^ c2.color) == NEUTRAL && c0.basis == c1.basis && c1.basis == c2.basis) {
            AtomKey key(dets[0], dets[1], dets[2]);
            (*out)[key] = obs;
            return key;
        }
        return AtomKey();
    }
    return AtomKey();
}

// collect_composite_errors.cc:21-42
static void try_grow(
    const AtomKey &e1, const AtomKey &e2, const std::vector<ColorBasis> &nc, const AtomMap &atoms, AtomKey *out, int &best) {
    bool c1 = atoms.count(e1) > 0;
    bool c2 = atoms.count(e2) > 0;
    int score = (int)c1 + 2 * (int)c2;
    if (score <= best) return;
    if (score == 1 && e2.weight() == 3 && e2.net_charge(nc) != NEUTRAL) return;
    if (score == 2 && e1.weight() == 3 && e1.net_charge(nc) != NEUTRAL) return;
    *out = c2 ? e2 : e1;
    best = score;
}

// The complement picker used by collect_composite_errors.cc:70-190: the j-th detector not among the chosen ones.
static node_t nth_unchosen(const std::vector<node_t> &dets, std::initializer_list<size_t> chosen, size_t j) {
    size_t seen = 0;
    for (size_t i = 0; i < dets.size(); i++) {
        bool is_chosen = false;
        for (size_t c : chosen) is_chosen |= (c == i);
        if (is_chosen) continue;
        if (seen == j) return dets[i];
        seen++;
    }
    return BOUNDARY;
}

// collect_composite_errors.cc:44-228
static AtomKey decompose_single_basis(const std::vector<node_t> &dets, const std::vector<ColorBasis> &nc, const AtomMap &atoms) {
    size_t n = dets.size();
    if (n <= 3) {
        AtomKey solo(dets);
        if (atoms.count(solo)) return solo;
    }
    int best = 0;
    AtomKey result;
    switch (n) {
        case 2: {
            // Quirk (collect_composite_errors.cc:49-50): only the FIRST detector is ever tried as a singlet.
            AtomKey a1(dets[0], BOUNDARY, BOUNDARY);
            if (atoms.count(a1)) return a1;
            return AtomKey();
        }
        case 3:
            for (size_t k1 = 0; k1 < n; k1++)
                try_grow(AtomKey(dets[k1], BOUNDARY, BOUNDARY),
                         AtomKey(nth_unchosen(dets, {k1}, 0), nth_unchosen(dets, {k1}, 1), BOUNDARY), nc, atoms, &result, best);
            return result;
        case 4:
            for (size_t k1 = 0; k1 < n && best < 2; k1++)
                for (size_t k2 = k1 + 1; k2 < n; k2++)
                    try_grow(AtomKey(dets[k1], dets[k2], BOUNDARY),
                             AtomKey(nth_unchosen(dets, {k1, k2}, 0), nth_unchosen(dets, {k1, k2}, 1), BOUNDARY), nc, atoms, &result, best);
            for (size_t k1 = 0; k1 < n; k1++)
                try_grow(AtomKey(dets[k1], BOUNDARY, BOUNDARY),
                         AtomKey(nth_unchosen(dets, {k1}, 0), nth_unchosen(dets, {k1}, 1), nth_unchosen(dets, {k1}, 2)), nc, atoms, &result, best);
            return result;
        case 5:
            for (size_t k1 = 0; k1 < n && best < 2; k1++)
                for (size_t k2 = k1 + 1; k2 < n; k2++)
                    try_grow(AtomKey(dets[k1], dets[k2], BOUNDARY),
                             AtomKey(nth_unchosen(dets, {k1, k2}, 0), nth_unchosen(dets, {k1, k2}, 1), nth_unchosen(dets, {k1, k2}, 2)),
                             nc, atoms, &result, best);
            return result;
        case 6:
            for (size_t k1 = 0; k1 < n && best < 2; k1++)
                for (size_t k2 = k1 + 1; k2 < n; k2++)
                    for (size_t k3 = k2 + 1; k3 < n; k3++)
                        try_grow(AtomKey(dets[k1], dets[k2], dets[k3]),
                                 AtomKey(nth_unchosen(dets, {k1, k2, k3}, 0), nth_unchosen(dets, {k1, k2, k3}, 1),
                                         nth_unchosen(dets, {k1, k2, k3}, 2)),
                                 nc, atoms, &result, best);
            return result;
        default:
            return AtomKey();
    }
}

struct MobiusError {
    double p;
    std::vector<std::pair<node_t, node_t>> edges;  // '^'-separated components
};

// collect_composite_errors.cc:230-335 (default options: drop remnants, do not ignore failures)
static void decompose_dets_into_atoms(
    const std::vector<node_t> &dets, obsmask obs, const std::vector<ColorBasis> &nc, const AtomMap &atoms,
    const FlatError &for_msg, std::vector<AtomKey> *out_atoms, AtomMap *remnants) {
    std::vector<node_t> xz[2];
    for (auto t : dets) xz[nc[t].basis == X ? 0 : 1].push_back(t);
    out_atoms->clear();
    for (int rep = 0; rep < 3; rep++) {
        for (int b = 0; b < 2; b++) {
            auto &bd = xz[b];
            AtomKey removed = (rep == 2) ? extract_atomic(bd, obs, nc, remnants) : decompose_single_basis(bd, nc, atoms);
            int w = removed.weight();
            if (!w) continue;
            for (int k = 0; k < w; k++) {
                for (size_t i = 0; i < bd.size(); i++) {
                    if (bd[i] == removed.d[k]) {
                        bd[i] = bd.back();
                        bd.pop_back();
                        break;
                    }
                }
            }
            auto f = atoms.find(removed);
            obs ^= (f != atoms.end()) ? f->second : remnants->at(removed);
            out_atoms->push_back(removed);
        }
    }
    if (!xz[0].empty() || !xz[1].empty()) {
        std::stringstream ss;
        ss << "Failed to decompose a complex error instruction into basic errors.\n";
        ss << "    The instruction (after shifting): " << for_msg.text << "\n";
        throw std::invalid_argument(ss.str());
    }
}

struct Tables {
    uint64_t num_dets = 0, num_obs = 0;
    std::vector<ColorBasis> node_colors;
    AtomMap atomic_errors;
    std::vector<MobiusError> mobius;
    std::vector<RgbRep> rgb_reps;
    ChargeGraph charge_graph;  // neighbor -> obs flip (BOUNDARY key allowed)
    std::map<std::array<uint32_t, 4>, obsmask> drag;      // {n1, n2, c1, c2}
    std::unordered_map<uint64_t, obsmask> drag_lookup;    // same content, hashed (filled after the map is complete)
    static uint64_t drag_key(uint32_t n1, uint32_t n2, uint32_t c1, uint32_t c2) {
        return ((uint64_t)n1 << 34) | ((uint64_t)n2 << 4) | (c1 << 2) | c2;
    }
    uint64_t drag_ambiguities = 0;
    // matcher graph (CSR over 2*num_dets nodes)
    std::vector<uint32_t> row_ptr, nbr;
    std::vector<int32_t> wt;
    double normalising_constant = 1;
};

// choose_rgb_reps.cc:19-92
static std::vector<RgbRep> choose_rgb_reps(const AtomMap &atoms, const std::vector<ColorBasis> &nc) {
    std::vector<RgbRep> result(nc.size());
    for (const auto &[err, obs] : atoms) {
        RgbRep rep;
        rep.obs_flip = obs;
        int weight = 0;
        for (auto n : err.d) {
            if (n != BOUNDARY) {
                uint8_t c = nc[n].color;
                rep.rgb[c - 1] = n;
                rep.charge_flip ^= c;
                weight++;
            }
        }
        if (rep.weight() != weight) continue;  // a colour appeared more than once
        for (auto n : err.d)
            if (n != BOUNDARY && weight > result[n].weight()) result[n] = rep;
    }
    for (const auto &[e, obs] : atoms) {
        if (e.weight() != 2) continue;
        uint8_t c1 = nc[e.d[0]].color, c2 = nc[e.d[1]].color;
        if (c1 != c2) continue;
        int w1 = result[e.d[0]].weight(), w2 = result[e.d[1]].weight();
        if (w1 == 0 && w2 > 0) {
            result[e.d[0]] = result[e.d[1]];
            result[e.d[0]].rgb[c1 - 1] = e.d[0];
            result[e.d[0]].obs_flip ^= obs;
        }
        if (w2 == 0 && w1 > 0) {
            result[e.d[1]] = result[e.d[0]];
            result[e.d[1]].rgb[c2 - 1] = e.d[1];
            result[e.d[1]].obs_flip ^= obs;
        }
    }
    return result;
}

// charge_graph.cc:80-160
static ChargeGraph build_charge_graph(const AtomMap &atoms, size_t num_nodes) {
    ChargeGraph g(num_nodes);
    auto add_edge = [&](node_t a, node_t b, obsmask f) {
        if (a != BOUNDARY) g[a][b] = f;
        if (b != BOUNDARY) g[b][a] = f;
    };
    for (size_t k = 0; k < num_nodes; k++) g[k][(node_t)k] = 0;
    for (const auto &[err, obs] : atoms)
        if (err.d[2] == BOUNDARY) add_edge(err.d[0], err.d[1], obs);
    std::map<node_t, std::vector<AtomKey>> node2errs;
    for (const auto &[err, obs] : atoms)
        for (auto n : err.d)
            if (n != BOUNDARY) node2errs[n].push_back(err);
    std::vector<node_t> buf;
    for (const auto &[_, errs] : node2errs) {
        for (size_t k1 = 0; k1 < errs.size(); k1++) {
            for (size_t k2 = k1 + 1; k2 < errs.size(); k2++) {
                const auto &e1 = errs[k1], &e2 = errs[k2];
                if (e1.weight() < 3 && e2.weight() < 3) continue;
                buf.assign({e1.d[0], e1.d[1], e1.d[2], e2.d[0], e2.d[1], e2.d[2]});
                std::sort(buf.begin(), buf.end());
                std::vector<node_t> x;
                for (auto v : buf) {
                    if (!x.empty() && x.back() == v) x.pop_back();
                    else x.push_back(v);
                }
                node_t a, b;
                if (x.size() == 1) {
                    a = x[0];
                    b = BOUNDARY;
                } else if (x.size() == 2 || (x.size() == 3 && x.back() == BOUNDARY)) {
                    a = x[0];
                    b = x[1];
                } else {
                    continue;
                }
                add_edge(a, b, atoms.at(e1) ^ atoms.at(e2));
            }
        }
    }
    return g;
}

// drag_graph.cc:24-105 (BfsSearcher::find_shortest_path_obs_flip, max_cost = 2). The reference walks
// `graph.nodes[n].neighbors`, a std::unordered_map<uint32_t, uint64_t> (charge_graph.h:52), and returns the FIRST path
// of at most two edges it meets. Several two-edge paths with different observable masks exist on some circuits
// (toric d12: 4000 searches, midout488 d9: 380, superdense d5: 180), so the container's iteration order decides 1-17 %
// of the predictions there. ChargeGraph here IS that container (same key/value types, same insertion sequence, the
// libstdc++ the reference's CI builds with), and the search below visits in the reference's order: level-1 nodes are
// pushed in src's iteration order (src itself included through its self loop -- src is never tagged as seen) and
// popped from the back; each popped node's neighbours are scanned in its own iteration order.
// `ambiguities` additionally counts the searches whose answer depends on that order (diagnostic only).
static std::optional<obsmask> short_path_flip(const ChargeGraph &g, node_t src, node_t dst, uint64_t *ambiguities) {
    if (src == dst) return obsmask{0};
    auto it = g[src].find(dst);
    if (it != g[src].end()) return it->second;
    std::vector<std::pair<node_t, obsmask>> level1;
    std::set<node_t> seen;
    for (const auto &[nb, f] : g[src]) {
        if (nb == dst) return f;
        if (nb == BOUNDARY) continue;
        if (!seen.insert(nb).second) continue;
        level1.push_back({nb, f});
    }
    std::optional<obsmask> found;
    for (size_t i = level1.size(); i-- > 0;) {
        auto [n, pf] = level1[i];
        for (const auto &[nb, f] : g[n]) {
            if (nb == dst) {
                obsmask v = pf ^ f;
                if (!found.has_value()) found = v;
                else if (*found != v) {
                    (*ambiguities)++;
                    return found;
                }
                break;
            }
        }
    }
    return found;
}

// drag_graph.cc:107-203
static void build_drag_graph(Tables &t) {
    const auto &nc = t.node_colors;
    const auto &reps = t.rgb_reps;
    std::set<std::pair<node_t, node_t>> decomposed;
    auto add_edge = [&](node_t n1, node_t n2, uint8_t c1, uint8_t c2, obsmask f) {
        t.drag[{n1, n2, c1, c2}] = f;
        t.drag[{n2, n1, c2, c1}] = f;
    };
    auto sorted_pair = [](node_t a, node_t b) { return a < b ? std::make_pair(a, b) : std::make_pair(b, a); };
    auto add_boundary_dumping_edge = [&](node_t a, node_t b, obsmask ab_flip) {
        if (reps[a].weight() != 3) return;
        uint8_t ca = nc[a].color, cb = nc[b].color, c = ca ^ cb;
        if (c == NEUTRAL) return;
        auto r1 = short_path_flip(t.charge_graph, reps[a].color_node(ca), a, &t.drag_ambiguities);
        auto r2 = short_path_flip(t.charge_graph, reps[a].color_node(cb), b, &t.drag_ambiguities);
        if (r1.has_value() && r2.has_value()) add_edge(a, b, c, NEUTRAL, *r1 ^ *r2 ^ reps[a].obs_flip ^ ab_flip);
    };
    for (const auto &[err, obs] : t.atomic_errors) {
        int w = err.weight();
        if (w == 3) {
            decomposed.insert(sorted_pair(err.d[0], err.d[1]));
            decomposed.insert(sorted_pair(err.d[0], err.d[2]));
            decomposed.insert(sorted_pair(err.d[1], err.d[2]));
        } else if (w == 2) {
            node_t a = err.d[0], b = err.d[1];
            uint8_t ca = nc[a].color, cb = nc[b].color;
            obsmask p = t.charge_graph[a].at(b);
            add_edge(a, b, ca, cb, p);
            add_edge(a, b, NEUTRAL, NEUTRAL, 0);
            add_boundary_dumping_edge(a, b, obs);
            add_boundary_dumping_edge(b, a, obs);
            decomposed.insert(sorted_pair(a, b));
        } else if (w == 1) {
            node_t n = err.d[0];
            uint8_t c = nc[n].color;
            add_edge(n, n, c, NEUTRAL, obs);
            add_edge(n, n, NEUTRAL, NEUTRAL, 0);
            if (reps[n].weight() == 3) {
                uint8_t c1 = next_non_neutral(c), c2 = next_non_neutral(c1);
                add_edge(n, n, c1, c2, reps[n].obs_flip ^ obs);
            }
        }
    }
    for (const auto &[n1, n2] : decomposed) {
        for (uint8_t c = 1; c < 4; c++) {
            node_t r1 = reps[n1].color_node(c), r2 = reps[n2].color_node(c);
            if (r1 != BOUNDARY && r2 != BOUNDARY) {
                auto res = short_path_flip(t.charge_graph, r1, r2, &t.drag_ambiguities);
                if (res.has_value()) add_edge(n1, n2, c, c, *res);
            }
        }
        add_edge(n1, n2, NEUTRAL, NEUTRAL, 0);
    }
}

// PyMatching v2.2.2 user-graph semantics as recalled.  decode/pymatcher.cc:22-25.
static double merge_weights(double a, double b) {
    double sgn = std::copysign(1.0, a) * std::copysign(1.0, b);
    double signed_min = sgn * std::min(std::fabs(a), std::fabs(b));
    return signed_min + std::log(1 + std::exp(-std::fabs(a + b))) - std::log(1 + std::exp(-std::fabs(a - b)));
}

static void build_matcher_graph(Tables &t) {
    size_t n = t.num_dets * 2;
    std::map<std::pair<node_t, node_t>, double> w;
    std::vector<std::pair<node_t, node_t>> order;
    for (const auto &me : t.mobius) {
        if (!(me.p > 0)) continue;
        if (me.p > 0.5) throw std::invalid_argument("Mobius error with probability > 0.5 is not supported.");
        double lw = std::log((1 - me.p) / me.p);
        for (auto e : me.edges) {
            if (e.first == e.second) continue;
            if (e.first > e.second) std::swap(e.first, e.second);
            auto it = w.find(e);
            if (it == w.end()) {
                w[e] = lw;
                order.push_back(e);
            } else {
                it->second = merge_weights(it->second, lw);
            }
        }
    }
    double max_abs = 0;
    bool all_integral = true;
    for (auto &kv : w) {
        max_abs = std::max(max_abs, std::fabs(kv.second));
        if (std::round(kv.second) != kv.second) all_integral = false;
    }
    double c = all_integral ? 1.0 : (double)((1 << 24) - 1) / max_abs;
    if (w.empty()) c = 1.0;
    t.normalising_constant = c * 2;
    std::vector<std::vector<std::pair<uint32_t, int32_t>>> adj(n);
    for (auto &kv : w) {
        int64_t iw = (int64_t)std::llround(kv.second * c) * 2;
        adj[kv.first.first].push_back({kv.first.second, (int32_t)iw});
        adj[kv.first.second].push_back({kv.first.first, (int32_t)iw});
    }
    t.row_ptr.assign(n + 1, 0);
    for (size_t i = 0; i < n; i++) {
        std::sort(adj[i].begin(), adj[i].end());
        t.row_ptr[i + 1] = t.row_ptr[i] + (uint32_t)adj[i].size();
        for (auto &e : adj[i]) {
            t.nbr.push_back(e.first);
            t.wt.push_back(e.second);
        }
    }
}

// decoder.cc:24-69
static std::unique_ptr<Tables> compile(const std::string &dem_text) {
    Dem dem = Dem::parse(dem_text);
    auto t = std::make_unique<Tables>();
    std::vector<FlatError> errors;
    {
        uint64_t off = 0, max_det = 0, max_obs = 0;
        bool any_det = false, any_obs = false;
        flatten_block(dem.instructions, off, &errors, &max_det, &max_obs, &any_det, &any_obs);
        t->num_dets = any_det ? max_det + 1 : 0;
        t->num_obs = any_obs ? max_obs + 1 : 0;
    }
    t->node_colors.resize(t->num_dets);
    {
        uint64_t off = 0;
        std::vector<double> coord_offsets;
        collect_nodes_helper(dem.instructions, &off, &coord_offsets, &t->node_colors);
    }
    const auto &nc = t->node_colors;
    std::vector<node_t> dets;
    obsmask obs;
    for (const auto &e : errors) {
        extract_obs_and_dets(e, nc, &dets, &obs);
        extract_atomic(dets, obs, nc, &t->atomic_errors);
    }
    AtomMap remnants;
    std::vector<AtomKey> atoms;
    for (const auto &e : errors) {
        extract_obs_and_dets(e, nc, &dets, &obs);
        decompose_dets_into_atoms(dets, obs, nc, t->atomic_errors, e, &atoms, &remnants);
        if (!remnants.empty()) {  // drop_mobius_errors_involving_remnant_errors = true
            atoms.clear();
            remnants.clear();
        }
        MobiusError me;
        bool has_corner = false;
        for (const auto &a : atoms) {
            has_corner |= a.d[1] == BOUNDARY;
            atom_mobius_edges(a, nc, &me.edges);
        }
        if (!me.edges.empty()) {
            me.p = e.p;
            if (has_corner) me.p *= me.p;
            t->mobius.push_back(std::move(me));
        }
    }
    t->rgb_reps = choose_rgb_reps(t->atomic_errors, nc);
    t->charge_graph = build_charge_graph(t->atomic_errors, nc.size());
    build_drag_graph(*t);
    t->drag_lookup.reserve(t->drag.size() * 2);
    for (const auto &kv : t->drag) t->drag_lookup[Tables::drag_key(kv.first[0], kv.first[1], kv.first[2], kv.first[3])] = kv.second;
    build_matcher_graph(*t);
    return t;
}

// ---------------------------------------------------------------------------------------------
// Exact min-weight perfect matching on a small dense graph (primal-dual blossom, O(n^3)).
// Maximises sum of w over maximum-cardinality matchings. Vertices 1..n. w > 0 marks an existing edge.
// Duals: lab[] (doubled units): lab[u] + lab[v] >= 2 w(u,v) for every edge, with equality on matched edges
// (plus blossom terms for edges inside blossoms).
// ---------------------------------------------------------------------------------------------
struct DenseBlossom {
    struct E {
        int u, v;
        int64_t w;
    };
    int n = 0, n_x = 0;
    std::vector<std::vector<E>> g;
    std::vector<int64_t> lab;
    std::vector<int> match, slack, st, pa, S, vis;
    std::vector<std::vector<int>> flower_from, flower;
    std::deque<int> q;
    int vis_t = 0;
    static constexpr int64_t INF = INT64_MAX / 4;

    void init(int n_) {
        n = n_;
        int m = n + n / 2 + 3;  // at most (n - 1) / 2 blossoms are alive at once
        if ((int)g.size() < m) g.resize(m);
        for (int u = 0; u < m; u++) {
            if ((int)g[u].size() < m) g[u].resize(m);
            for (int v = 0; v < m; v++) g[u][v] = E{u, v, 0};
        }
        lab.assign(m, 0);
        match.assign(m, 0);
        slack.assign(m, 0);
        st.assign(m, 0);
        pa.assign(m, 0);
        S.assign(m, 0);
        vis.assign(m, 0);
        if ((int)flower_from.size() < m) flower_from.resize(m);
        for (int u = 0; u < m; u++) flower_from[u].assign(n + 1, 0);
        if ((int)flower.size() < m) flower.resize(m);
        for (int u = 0; u < m; u++) flower[u].clear();
    }
    void add_edge(int u, int v, int64_t w) {
        g[u][v].w = g[v][u].w = w;
    }
    int64_t dist(const E &e) const { return lab[e.u] + lab[e.v] - g[e.u][e.v].w * 2; }
    void update_slack(int u, int x) {
        if (!slack[x] || dist(g[u][x]) < dist(g[slack[x]][x])) slack[x] = u;
    }
    void set_slack(int x) {
        slack[x] = 0;
        for (int u = 1; u <= n; ++u)
            if (g[u][x].w > 0 && st[u] != x && S[st[u]] == 0) update_slack(u, x);
    }
    void q_push(int x) {
        if (x <= n) {
            q.push_back(x);
            return;
        }
        for (int y : flower[x]) q_push(y);
    }
    void set_st(int x, int b) {
        st[x] = b;
        if (x <= n) return;
        for (int y : flower[x]) set_st(y, b);
    }
    int get_pr(int b, int xr) {
        int pr = (int)(std::find(flower[b].begin(), flower[b].end(), xr) - flower[b].begin());
        if (pr % 2 == 1) {
            std::reverse(flower[b].begin() + 1, flower[b].end());
            return (int)flower[b].size() - pr;
        }
        return pr;
    }
    void set_match(int u, int v) {
        match[u] = g[u][v].v;
        if (u <= n) return;
        E e = g[u][v];
        int xr = flower_from[u][e.u], pr = get_pr(u, xr);
        for (int i = 0; i < pr; ++i) set_match(flower[u][i], flower[u][i ^ 1]);
        set_match(xr, v);
        std::rotate(flower[u].begin(), flower[u].begin() + pr, flower[u].end());
    }
    void augment(int u, int v) {
        for (;;) {
            int xnv = st[match[u]];
            set_match(u, v);
            if (!xnv) return;
            set_match(xnv, st[pa[xnv]]);
            u = st[pa[xnv]];
            v = xnv;
        }
    }
    int get_lca(int u, int v) {
        for (++vis_t; u || v; std::swap(u, v)) {
            if (u == 0) continue;
            if (vis[u] == vis_t) return u;
            vis[u] = vis_t;
            u = st[match[u]];
            if (u) u = st[pa[u]];
        }
        return 0;
    }
    void add_blossom(int u, int lca, int v) {
        int b = n + 1;
        while (b <= n_x && st[b]) ++b;
        if (b > n_x) ++n_x;
        lab[b] = 0;
        S[b] = 0;
        match[b] = match[lca];
        flower[b].clear();
        flower[b].push_back(lca);
        for (int x = u, y; x != lca; x = st[pa[y]]) {
            flower[b].push_back(x);
            flower[b].push_back(y = st[match[x]]);
            q_push(y);
        }
        std::reverse(flower[b].begin() + 1, flower[b].end());
        for (int x = v, y; x != lca; x = st[pa[y]]) {
            flower[b].push_back(x);
            flower[b].push_back(y = st[match[x]]);
            q_push(y);
        }
        set_st(b, b);
        for (int x = 1; x <= n_x; ++x) g[b][x].w = g[x][b].w = 0;
        for (int x = 1; x <= n; ++x) flower_from[b][x] = 0;
        for (size_t i = 0; i < flower[b].size(); ++i) {
            int xs = flower[b][i];
            for (int x = 1; x <= n_x; ++x)
                if (g[xs][x].w > 0 && (g[b][x].w == 0 || dist(g[xs][x]) < dist(g[b][x]))) {  // sparse-safe: never copy a non-edge
                    g[b][x] = g[xs][x];
                    g[x][b] = g[x][xs];
                }
            for (int x = 1; x <= n; ++x)
                if (flower_from[xs][x]) flower_from[b][x] = xs;
        }
        set_slack(b);
    }
    void expand_blossom(int b) {
        for (size_t i = 0; i < flower[b].size(); ++i) set_st(flower[b][i], flower[b][i]);
        int xr = flower_from[b][g[b][pa[b]].u], pr = get_pr(b, xr);
        for (int i = 0; i < pr; i += 2) {
            int xs = flower[b][i], xns = flower[b][i + 1];
            pa[xs] = g[xns][xs].u;
            S[xs] = 1;
            S[xns] = 0;
            slack[xs] = 0;
            set_slack(xns);
            q_push(xns);
        }
        S[xr] = 1;
        pa[xr] = pa[b];
        for (size_t i = pr + 1; i < flower[b].size(); ++i) {
            int xs = flower[b][i];
            S[xs] = -1;
            set_slack(xs);
        }
        st[b] = 0;
    }
    bool on_found_edge(const E &e) {
        int u = st[e.u], v = st[e.v];
        if (S[v] == -1) {
            pa[v] = e.u;
            S[v] = 1;
            int nu = st[match[v]];
            slack[v] = slack[nu] = 0;
            S[nu] = 0;
            q_push(nu);
        } else if (S[v] == 0) {
            int lca = get_lca(u, v);
            if (!lca) {
                augment(u, v);
                augment(v, u);
                return true;
            }
            add_blossom(u, lca, v);
        }
        return false;
    }
    bool matching() {
        std::fill(S.begin() + 1, S.begin() + 1 + n_x, -1);
        std::fill(slack.begin() + 1, slack.begin() + 1 + n_x, 0);
        q.clear();
        for (int x = 1; x <= n_x; ++x)
            if (st[x] == x && !match[x]) {
                pa[x] = 0;
                S[x] = 0;
                q_push(x);
            }
        if (q.empty()) return false;
        for (;;) {
            while (!q.empty()) {
                int u = q.front();
                q.pop_front();
                if (S[st[u]] == 1) continue;
                for (int v = 1; v <= n; ++v)
                    if (g[u][v].w > 0 && st[u] != st[v]) {
                        if (dist(g[u][v]) == 0) {
                            if (on_found_edge(g[u][v])) return true;
                        } else {
                            update_slack(u, st[v]);
                        }
                    }
            }
            int64_t d = INF;
            for (int b = n + 1; b <= n_x; ++b)
                if (st[b] == b && S[b] == 1) d = std::min(d, lab[b] / 2);
            for (int x = 1; x <= n_x; ++x)
                if (st[x] == x && slack[x]) {
                    if (S[x] == -1) d = std::min(d, dist(g[slack[x]][x]));
                    else if (S[x] == 0) d = std::min(d, dist(g[slack[x]][x]) / 2);
                }
            if (d >= INF) return false;  // maximum-cardinality mode: vertex duals may go negative, never stop on them
            for (int u = 1; u <= n; ++u) {
                if (S[st[u]] == 0) lab[u] -= d;
                else if (S[st[u]] == 1) lab[u] += d;
            }
            for (int b = n + 1; b <= n_x; ++b)
                if (st[b] == b) {
                    if (S[st[b]] == 0) lab[b] += d * 2;
                    else if (S[st[b]] == 1) lab[b] -= d * 2;
                }
            q.clear();
            for (int x = 1; x <= n_x; ++x)
                if (st[x] == x && slack[x] && st[slack[x]] != x && dist(g[slack[x]][x]) == 0)
                    if (on_found_edge(g[slack[x]][x])) return true;
            for (int b = n + 1; b <= n_x; ++b)
                if (st[b] == b && S[b] == 1 && lab[b] == 0) expand_blossom(b);
        }
    }
    // returns number of matched pairs
    int solve() {
        n_x = n;
        std::fill(match.begin(), match.end(), 0);
        for (int u = 0; u <= n; ++u) {
            st[u] = u;
            flower[u].clear();
        }
        int64_t w_max = 0;
        for (int u = 1; u <= n; ++u)
            for (int v = 1; v <= n; ++v) {
                flower_from[u][v] = (u == v ? u : 0);
                w_max = std::max(w_max, g[u][v].w);
            }
        (void)w_max;
        // Start from per-vertex labels lab[u] = max_v w(u, v) (feasible: lab[u] + lab[v] >= 2 w(u, v)) and a greedy
        // matching on the edges that are tight under them (mutual best partners), lowest indices first. Any feasible
        // start with tight matched edges is a valid state of the primal-dual method; this one leaves only the
        // vertices without a mutual best partner to the second pass below and to the augmentation phases.
        // (decode_core.h Blossom::solve does the same, step by step.)
        for (int u = 1; u <= n; ++u) {
            int64_t m = 0;
            for (int v = 1; v <= n; ++v) m = std::max(m, g[u][v].w);
            lab[u] = m;
        }
        int cnt = 0;
        for (int u = 1; u <= n; ++u) {
            if (match[u] || lab[u] <= 0) continue;
            for (int v = 1; v <= n; ++v)
                if (v != u && !match[v] && g[u][v].w == lab[u] && lab[v] == lab[u]) {
                    match[u] = v;
                    match[v] = u;
                    ++cnt;
                    break;
                }
        }
        // Second pass: every vertex that is still free takes the lowest label its edges allow,
        // lab[u] = max_v (2 w(u, v) - lab[v]) -- what the first dual steps of the phases would do one step at a time. The edge
        // that defines it (lowest v on ties) is tight afterwards, and if its other end is free as well the pair is matched.
        // Vertices are visited in ascending order and see the labels their predecessors were given.
        for (int u = 1; u <= n; ++u) {
            if (match[u] || lab[u] <= 0) continue;
            bool any = false;
            int64_t best = 0;
            int bv = 0;
            for (int v = 1; v <= n; ++v) {
                if (v == u || g[u][v].w <= 0) continue;
                int64_t t = 2 * g[u][v].w - lab[v];
                if (!any || t > best) {
                    any = true;
                    best = t;
                    bv = v;
                }
            }
            if (!any) continue;
            lab[u] = best;
            if (!match[bv]) {
                match[u] = bv;
                match[bv] = u;
                ++cnt;
            }
        }
        while (matching()) ++cnt;
        return cnt;
    }
};

// ---------------------------------------------------------------------------------------------
// Per-shot path
// ---------------------------------------------------------------------------------------------
struct MatchResult {
    std::vector<int64_t> edges;  // interleaved graph edges on the matched paths
    int64_t weight = 0;
};

// std::priority_queue semantics (min-heap) over storage that survives between calls.
struct ReusableHeap {
    typedef std::pair<int64_t, node_t> QE;
    std::vector<QE> v;
    void clear() { v.clear(); }
    bool empty() const { return v.empty(); }
    const QE &top() const { return v.front(); }
    void push(QE e) {
        v.push_back(e);
        std::push_heap(v.begin(), v.end(), std::greater<QE>());
    }
    void pop() {
        std::pop_heap(v.begin(), v.end(), std::greater<QE>());
        v.pop_back();
    }
};

struct Decoder {
    std::unique_ptr<Tables> t;
    ReusableHeap heap_buf;
    // scratch for dijkstra
    std::vector<int64_t> dist;
    std::vector<uint32_t> touched;
    std::vector<int32_t> event_index;  // node -> index in current event list or -1
    // euler scratch (euler_tours.h:34-49)
    struct Nb {
        node_t node;
        uint32_t back;
    };
    std::vector<std::vector<Nb>> en;
    std::vector<size_t> enext;
    std::string last_error;
    // statistics
    uint64_t stat_radius_retries = 0;
    // 0 = dense all-pairs matcher up to 160 events, certified-local with lazy merging beyond; 1 = certified-local, eager
    // merging; 2 = always dense; 3 = always certified-local with lazy merging (the kernels' algorithm)
    int matcher_mode = 0;
    int odd_policy = 0;  // 0: every member of an odd component grows; 1: only the vertices a maximum matching leaves exposed
    int growth_extra = 1;  // events an open-ended growth request asks for beyond the required one (kernels: GROW_EXTRA_EVENTS = 1;
                           // 0 is the cheaper setting on a CPU, where matcher rounds cost little: bench.py's CPU legs use it)
    DenseBlossom local_blossom;
    std::vector<int> loc_buf;

    explicit Decoder(std::unique_ptr<Tables> tt) : t(std::move(tt)) {
        size_t n = t->num_dets * 2;
        dist.assign(n, -1);
        event_index.assign(n, -1);
        en.resize(n);
        enext.assign(n, 0);
    }

    // Truncated Dijkstra from src; reports (event index, distance) of every other event within `radius`.
    void dijkstra_events(node_t src, int64_t radius, std::vector<std::pair<int, int64_t>> *found) {
        typedef std::pair<int64_t, node_t> QE;
        std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
        touched.clear();
        dist[src] = 0;
        touched.push_back(src);
        pq.push({0, src});
        while (!pq.empty()) {
            auto [d, u] = pq.top();
            pq.pop();
            if (d != dist[u]) continue;
            if (u != src && event_index[u] >= 0) found->push_back({event_index[u], d});
            for (uint32_t k = t->row_ptr[u]; k < t->row_ptr[u + 1]; k++) {
                node_t v = t->nbr[k];
                int64_t nd = d + t->wt[k];
                if (radius >= 0 && nd > radius) continue;
                if (dist[v] < 0 || nd < dist[v]) {
                    if (dist[v] < 0) touched.push_back(v);
                    dist[v] = nd;
                    pq.push({nd, v});
                }
            }
        }
        for (auto v : touched) dist[v] = -1;
    }

    // Canonical shortest path src -> dst. A direct graph edge of length d(src,dst) is taken as it is; otherwise: full
    // distances from src (bounded by d(src,dst)), then walk back from dst always taking the smallest-id tight
    // predecessor. Emits edges in walking order from src to dst.
    void canonical_path(node_t src, node_t dst, int64_t d_total, std::vector<int64_t> *out_edges) {
        for (uint32_t k = t->row_ptr[src]; k < t->row_ptr[src + 1]; k++)
            if (t->nbr[k] == dst && t->wt[k] == d_total) {
                out_edges->push_back(src);
                out_edges->push_back(dst);
                return;
            }
        ReusableHeap &pq = heap_buf;
        pq.clear();
        touched.clear();
        dist[src] = 0;
        touched.push_back(src);
        pq.push({0, src});
        while (!pq.empty()) {
            auto [d, u] = pq.top();
            pq.pop();
            if (d != dist[u]) continue;
            for (uint32_t k = t->row_ptr[u]; k < t->row_ptr[u + 1]; k++) {
                node_t v = t->nbr[k];
                int64_t nd = d + t->wt[k];
                if (nd > d_total) continue;
                if (dist[v] < 0 || nd < dist[v]) {
                    if (dist[v] < 0) touched.push_back(v);
                    dist[v] = nd;
                    pq.push({nd, v});
                }
            }
        }
        std::vector<node_t> rev;
        node_t cur = dst;
        rev.push_back(cur);
        while (cur != src) {
            node_t best = BOUNDARY;
            for (uint32_t k = t->row_ptr[cur]; k < t->row_ptr[cur + 1]; k++) {
                node_t v = t->nbr[k];
                if (dist[v] >= 0 && dist[v] + t->wt[k] == dist[cur]) {
                    best = v;  // neighbours are sorted ascending: first hit is the smallest id
                    break;
                }
            }
            assert(best != BOUNDARY);
            cur = best;
            rev.push_back(cur);
        }
        for (size_t i = rev.size(); i-- > 1;) {
            out_edges->push_back(rev[i]);
            out_edges->push_back(rev[i - 1]);
        }
        for (auto v : touched) dist[v] = -1;
    }

    // Exact MWPM of the events (decode/pymatcher.cc:27-42 contract: graph edges of matched paths + total weight).
    bool match_events(const std::vector<uint64_t> &events, MatchResult *out, std::vector<std::pair<int, int>> *out_pairs = nullptr) {
        out->edges.clear();
        out->weight = 0;
        int m = (int)events.size();
        if (m == 0) return true;
        if (matcher_mode == 1 || matcher_mode == 3 || (matcher_mode == 0 && m > 160)) {
            const bool lazy = matcher_mode != 1;
            std::vector<int> partner;
            std::vector<int64_t> pdist;
            if (lazy ? !match_events_lazy(events, &partner, &pdist) : !match_events_local(events, &partner, &pdist)) return false;
            for (int i = 0; i < m; i++) {
                int j = partner[i];
                if (j > i) {
                    out->weight += pdist[i];
                    canonical_path((node_t)events[i], (node_t)events[j], pdist[i], &out->edges);
                    if (out_pairs) out_pairs->push_back({i, j});
                }
            }
            return true;
        }
        for (int i = 0; i < m; i++) event_index[events[i]] = i;
        int64_t max_w = 0;
        for (auto w : t->wt) max_w = std::max<int64_t>(max_w, w);
        int64_t radius = 3 * max_w;
        bool ok = false;
        std::vector<int> partner(m, -1);
        std::vector<std::vector<std::pair<int, int64_t>>> near(m);
        for (int attempt = 0; attempt < 40 && !ok; attempt++) {
            bool unbounded = attempt >= 30;
            for (int i = 0; i < m; i++) {
                near[i].clear();
                dijkstra_events((node_t)events[i], unbounded ? -1 : radius, &near[i]);
            }
            DenseBlossom bl;
            bl.init(m);
            int64_t big = (unbounded ? 0 : radius) + 2;
            if (unbounded)
                for (int i = 0; i < m; i++)
                    for (auto &pr : near[i]) big = std::max(big, pr.second + 2);
            for (int i = 0; i < m; i++)
                for (auto &pr : near[i]) bl.add_edge(i + 1, pr.first + 1, big - pr.second);
            int cnt = bl.solve();
            if (cnt * 2 != m) {
                if (unbounded) break;
                radius *= 2;
                stat_radius_retries++;
                continue;
            }
            // certificate for the pairs that were truncated away: rad(a) = (big - lab[a]) / 2 <= radius / 2
            bool cert = true;
            if (!unbounded)
                for (int i = 1; i <= m; i++)
                    if (big - bl.lab[i] > radius) cert = false;
            if (!cert) {
                radius *= 2;
                stat_radius_retries++;
                continue;
            }
            for (int i = 0; i < m; i++) partner[i] = bl.match[i + 1] - 1;
            ok = true;
        }
        for (int i = 0; i < m; i++) event_index[events[i]] = -1;
        if (!ok) {
            last_error = "No perfect matching exists for the detection events.";
            return false;
        }
        for (int i = 0; i < m; i++) {
            int j = partner[i];
            if (j > i) {
                int64_t d = -1;
                for (auto &pr : near[i])
                    if (pr.first == j) d = pr.second;
                assert(d >= 0);
                out->weight += d;
                canonical_path((node_t)events[i], (node_t)events[j], d, &out->edges);
                if (out_pairs) out_pairs->push_back({i, j});
            }
        }
        return true;
    }

    // Open ball around one event (truncated Dijkstra). radius >= 0: every node strictly closer than radius. radius < 0:
    // until more than `need_more_than` other events have been settled; the radius is the distance of the last one, and
    // every event strictly closer than it has been reported. Returns the radius (-1 when nothing more is reachable).
    // `extra` (open-ended searches only): keep going until `extra` events beyond the required one have been settled, if
    // the graph has them; a search that exhausts the graph after the required event returns R_UNBOUNDED (everything
    // reachable has been seen).
    static constexpr int64_t R_UNBOUNDED = INT64_MAX / 4;
    int64_t ball(node_t src, int64_t radius, size_t need_more_than, std::vector<std::pair<int, int64_t>> *found, size_t extra = 0) {
        ReusableHeap &pq = heap_buf;
        pq.clear();
        touched.clear();
        found->clear();
        dist[src] = 0;
        touched.push_back(src);
        pq.push({0, src});
        int64_t limit = radius;
        while (!pq.empty()) {
            auto [d, u] = pq.top();
            pq.pop();
            if (d != dist[u]) continue;
            if (limit >= 0 && d >= limit) break;  // open ball: every event strictly closer than the radius is known
            if (u != src && event_index[u] >= 0) {
                found->push_back({event_index[u], d});
                if (limit < 0 && found->size() > need_more_than + extra) {
                    limit = d;
                    break;
                }
            }
            for (uint32_t k = t->row_ptr[u]; k < t->row_ptr[u + 1]; k++) {
                node_t v = t->nbr[k];
                int64_t nd = d + t->wt[k];
                if (limit >= 0 && nd >= limit) continue;
                if (dist[v] < 0 || nd < dist[v]) {
                    if (dist[v] < 0) touched.push_back(v);
                    dist[v] = nd;
                    pq.push({nd, v});
                }
            }
        }
        for (auto v : touched) dist[v] = -1;
        if (limit < 0 && found->size() > need_more_than) limit = R_UNBOUNDED;
        return limit;
    }

    // Four events: vertex duals alone certify the best of the three pairings (K4 has no non-trivial odd-set cut).
    // With {ab, cd} optimal the doubled duals (t, 2 d_ab - t, u, 2 d_cd - u) are feasible inside the box
    // 2 (d_ab + d_cd - d_bd) <= t + u <= 2 d_ac,  2 (d_ab - d_bc) <= t - u <= 2 (d_ad - d_cd); the balanced point
    // (d_ab, d_cd) is clamped into it. Returns false when the seen edges admit no perfect matching.
    bool solve_four(const std::vector<int> &mem, const std::vector<std::vector<std::pair<int, int64_t>>> &nb, int partner[4],
                    int64_t h[4], int64_t pdist[4]) {
        int64_t d[4][4];
        for (auto &row : d)
            for (auto &x : row) x = -1;
        for (int a = 0; a < 4; a++)
            for (auto &p : nb[mem[a]])
                for (int b = 0; b < 4; b++)
                    if (p.first == mem[b]) d[a][b] = d[b][a] = p.second;
        const int pa[3][4] = {{0, 1, 2, 3}, {0, 2, 1, 3}, {0, 3, 1, 2}};
        int best = -1;
        int64_t best_cost = 0;
        for (int k = 0; k < 3; k++) {
            int64_t x = d[pa[k][0]][pa[k][1]], y = d[pa[k][2]][pa[k][3]];
            if (x < 0 || y < 0) continue;
            if (best < 0 || x + y < best_cost) {
                best = k;
                best_cost = x + y;
            }
        }
        if (best < 0) return false;
        const int A = pa[best][0], B = pa[best][1], C = pa[best][2], D = pa[best][3];
        const int64_t dab = d[A][B], dcd = d[C][D];
        int64_t P = dab + dcd, Q = dab - dcd;
        if (d[A][C] >= 0) P = std::min(P, 2 * d[A][C]);
        if (d[B][D] >= 0) P = std::max(P, 2 * (dab + dcd - d[B][D]));
        if (d[A][D] >= 0) Q = std::min(Q, 2 * (d[A][D] - dcd));
        if (d[B][C] >= 0) Q = std::max(Q, 2 * (dab - d[B][C]));
        const int64_t t = (P + Q) / 2, u = (P - Q) / 2;
        h[A] = t; h[B] = 2 * dab - t; h[C] = u; h[D] = 2 * dcd - u;
        partner[A] = B; partner[B] = A; partner[C] = D; partner[D] = C;
        pdist[A] = pdist[B] = dab; pdist[C] = pdist[D] = dcd;
        return true;
    }

    // Certified local matching: the same optimum as match_events_dense without the all-pairs step, usable at
    // d25 r100 sizes. Balls grow until the blossom duals of every component fit inside them (2 y_a <= R_a), which keeps
    // the duals feasible for every pair that was never looked at; see the argument in DESIGN.md ("Matcher").
    bool match_events_local(const std::vector<uint64_t> &events, std::vector<int> *partner_out, std::vector<int64_t> *pdist_out) {
        int m = (int)events.size();
        partner_out->assign(m, -1);
        pdist_out->assign(m, 0);
        if (m == 0) return true;
        for (int i = 0; i < m; i++) event_index[events[i]] = i;
        std::vector<int64_t> R(m);
        std::vector<std::vector<std::pair<int, int64_t>>> nb(m);
        std::vector<char> dirty(m, 1);
        bool ok = true;
        for (int i = 0; i < m && ok; i++) {
            R[i] = ball((node_t)events[i], -1, 0, &nb[i]);
            if (R[i] < 0) ok = false;
        }
        std::vector<int> uf(m), comp_of(m);
        std::vector<std::pair<int, int64_t>> grow;
        for (int round = 0; ok; round++) {
            if (round > 64) {
                ok = false;
                break;
            }
            for (int i = 0; i < m; i++) uf[i] = i;
            auto find = [&](int x) {
                while (uf[x] != x) {
                    uf[x] = uf[uf[x]];
                    x = uf[x];
                }
                return x;
            };
            for (int i = 0; i < m; i++)
                for (auto &p : nb[i]) {
                    int a = find(i), b = find(p.first);
                    if (a != b) uf[std::max(a, b)] = std::min(a, b);
                }
            std::vector<std::pair<int, int>> by_root(m);
            for (int i = 0; i < m; i++) by_root[i] = {find(i), i};
            std::sort(by_root.begin(), by_root.end());
            grow.clear();
            std::vector<int> mem;
            for (int a = 0; a < m;) {
                int b = a;
                mem.clear();
                while (b < m && by_root[b].first == by_root[a].first) mem.push_back(by_root[b++].second);
                a = b;
                bool is_dirty = false;
                for (int i : mem) is_dirty |= dirty[i] != 0;
                if (!is_dirty) continue;
                int n = (int)mem.size();
                if (n % 2 && (odd_policy == 0 || n == 1)) {
                    for (int i : mem) grow.push_back({i, -1});
                    continue;
                }
                if (n == 2) {
                    // one edge: the blossom sets y = d/2 on both ends, so both balls must reach d
                    int i = mem[0], j = mem[1];
                    int64_t d = -1;
                    for (auto &p : nb[i])
                        if (p.first == j) d = p.second;
                    for (auto &p : nb[j])
                        if (p.first == i) d = p.second;
                    bool certified = true;
                    for (int e : mem)
                        if (d > R[e]) {
                            grow.push_back({e, d});
                            certified = false;
                        }
                    if (certified) {
                        (*partner_out)[i] = j;
                        (*partner_out)[j] = i;
                        (*pdist_out)[i] = (*pdist_out)[j] = d;
                    }
                    continue;
                }
                if (n == 4 && odd_policy != 99) {
                    int p4[4];
                    int64_t h4[4], d4[4];
                    if (solve_four(mem, nb, p4, h4, d4)) {
                        bool certified = true;
                        for (int k = 0; k < 4; k++)
                            if (h4[k] > R[mem[k]]) certified = false;
                        // growth requests in the order the kernel issues them: A, B, C, D of the chosen pairing
                        int first = 0;
                        (void)first;
                        std::vector<int> order;
                        for (int k = 0; k < 4; k++)
                            if (p4[k] > k) order.push_back(k), order.push_back(p4[k]);
                        for (int k : order)
                            if (h4[k] > R[mem[k]]) grow.push_back({mem[k], h4[k]});
                        if (certified)
                            for (int k = 0; k < 4; k++) {
                                (*partner_out)[mem[k]] = mem[p4[k]];
                                (*pdist_out)[mem[k]] = d4[k];
                            }
                        continue;
                    }
                }
                stat_blossoms[n <= 8 ? 0 : (n <= 16 ? 1 : (n <= 32 ? 2 : 3))]++;
                stat_max_comp = std::max<uint64_t>(stat_max_comp, (uint64_t)n);
                std::vector<int> &loc = loc_buf;
                loc.resize(m);
                for (int k = 0; k < n; k++) loc[mem[k]] = k + 1;
                int64_t big = 2;
                for (int i : mem)
                    for (auto &p : nb[i]) big = std::max(big, p.second + 2);
                DenseBlossom &bl = local_blossom;
                bl.init(n);
                for (int i : mem)
                    for (auto &p : nb[i]) bl.add_edge(loc[i], loc[p.first], big - p.second);
                int cnt = bl.solve();
                if (cnt * 2 != n) {
                    for (int i : mem)
                        if (!bl.match[loc[i]]) grow.push_back({i, -1});
                    continue;
                }
                bool certified = true;
                for (int i : mem) {
                    int64_t need = big - bl.lab[loc[i]];
                    if (need > R[i]) {
                        grow.push_back({i, need});
                        certified = false;
                    }
                }
                if (certified)
                    for (int i : mem) {
                        int j = mem[bl.match[loc[i]] - 1];
                        (*partner_out)[i] = j;
                        for (auto &p : nb[i])
                            if (p.first == j) (*pdist_out)[i] = (*pdist_out)[j] = p.second;
                    }
            }
            std::fill(dirty.begin(), dirty.end(), 0);
            if (grow.empty()) break;
            for (auto &[i, to] : grow) {
                R[i] = ball((node_t)events[i], to, nb[i].size(), &nb[i]);
                if (R[i] < 0) {
                    ok = false;
                    break;
                }
                dirty[i] = 1;
                for (auto &p : nb[i]) dirty[p.first] = 1;
            }
        }
        for (int i = 0; i < m; i++) event_index[events[i]] = -1;
        if (!ok) last_error = "No perfect matching exists for the detection events.";
        return ok;
    }

    // Certified local matching with lazy merging. Same certificate as match_events_local, but solved components keep
    // their doubled duals h (= 2 y): balls that grow only force a re-solve when a newly seen pair violates
    // h_a + h_b <= 2 d(a, b); otherwise the old component solutions stay optimal together and nothing is re-solved.
    // Final state: every event solved, h_a <= R_a for all a (covers unseen pairs) and h_a + h_b <= 2 d(a, b) for every
    // seen pair in different components, so the union of the component matchings is a global optimum.
    uint64_t stat_blossoms[4] = {0, 0, 0, 0};  // component sizes <= 8, <= 16, <= 32, > 32
    uint64_t stat_rounds = 0, stat_pairs_closed_form = 0, stat_max_comp = 0;
    bool match_events_lazy(const std::vector<uint64_t> &events, std::vector<int> *partner_out, std::vector<int64_t> *pdist_out) {
        int m = (int)events.size();
        partner_out->assign(m, -1);
        pdist_out->assign(m, 0);
        if (m == 0) return true;
        for (int i = 0; i < m; i++) event_index[events[i]] = i;
        std::vector<int64_t> R(m), h(m, 0);
        std::vector<std::vector<std::pair<int, int64_t>>> nb(m);
        std::vector<char> solved(m, 0), regrown(m, 0), stale(m, 0);
        std::vector<int> pc(m, -1);
        bool ok = true;
        for (int i = 0; i < m && ok; i++) {
            R[i] = ball((node_t)events[i], -1, 0, &nb[i]);
            if (R[i] < 0) ok = false;
        }
        std::vector<int> uf(m);
        std::vector<std::pair<int, int64_t>> grow;
        std::vector<std::pair<int, int>> by_root(m);
        std::vector<int> mem;
        for (int round = 0; ok; round++) {
            if (round > 64) {
                ok = false;
                break;
            }
            stat_rounds++;
            for (int i = 0; i < m; i++) uf[i] = i;
            auto find = [&](int x) {
                while (uf[x] != x) {
                    uf[x] = uf[uf[x]];
                    x = uf[x];
                }
                return x;
            };
            auto unite = [&](int a, int b) {
                a = find(a);
                b = find(b);
                if (a != b) uf[std::max(a, b)] = std::min(a, b);
            };
            for (int i = 0; i < m; i++)
                if (solved[i]) unite(i, pc[i]);
            for (int i = 0; i < m; i++)
                for (auto &p : nb[i]) {
                    int j = p.first;
                    if (!solved[i] || !solved[j] || h[i] + h[j] > 2 * p.second) {
                        // a violated pair inside one solved component is either an old blossom-internal edge or a pair
                        // first seen after the solve; only a ball that grew since then can have produced the latter
                        if (solved[i] && solved[j] && pc[i] == pc[j] && regrown[i]) stale[pc[i]] = 1;
                        unite(i, j);
                    }
                }
            for (int i = 0; i < m; i++) by_root[i] = {find(i), i};
            std::sort(by_root.begin(), by_root.end());
            grow.clear();
            for (int a = 0; a < m;) {
                int b = a;
                mem.clear();
                while (b < m && by_root[b].first == by_root[a].first) mem.push_back(by_root[b++].second);
                a = b;
                bool needs_solve = false;
                for (int i : mem) needs_solve |= !solved[i] || pc[i] != pc[mem[0]] || stale[pc[i]];
                for (int i : mem) regrown[i] = 0;
                if (!needs_solve) continue;
                for (int i : mem)
                    if (pc[i] >= 0) stale[pc[i]] = 0;
                int n = (int)mem.size();
                if (n % 2) {
                    for (int i : mem) {
                        grow.push_back({i, -1});
                        solved[i] = 0;
                    }
                    continue;
                }
                if (n == 2) {
                    int i = mem[0], j = mem[1];
                    int64_t d = -1;
                    for (auto &p : nb[i])
                        if (p.first == j) d = p.second;
                    for (auto &p : nb[j])
                        if (p.first == i) d = p.second;
                    stat_pairs_closed_form++;
                    for (int e : mem) {
                        h[e] = d;
                        solved[e] = 1;
                        pc[e] = mem[0];
                        if (d > R[e]) grow.push_back({e, d});
                    }
                    (*partner_out)[i] = j;
                    (*partner_out)[j] = i;
                    (*pdist_out)[i] = (*pdist_out)[j] = d;
                    continue;
                }
                if (n == 4) {
                    int p4[4];
                    int64_t h4[4], d4[4];
                    if (solve_four(mem, nb, p4, h4, d4)) {
                        std::vector<int> order;
                        for (int k = 0; k < 4; k++)
                            if (p4[k] > k) order.push_back(k), order.push_back(p4[k]);
                        for (int k : order) {
                            int e = mem[k];
                            h[e] = h4[k];
                            solved[e] = 1;
                            pc[e] = mem[0];
                            (*partner_out)[e] = mem[p4[k]];
                            (*pdist_out)[e] = d4[k];
                            if (h4[k] > R[e]) grow.push_back({e, h4[k]});
                        }
                        continue;
                    }
                }
                stat_blossoms[n <= 8 ? 0 : (n <= 16 ? 1 : (n <= 32 ? 2 : 3))]++;
                stat_max_comp = std::max<uint64_t>(stat_max_comp, (uint64_t)n);
                std::vector<int> &loc = loc_buf;
                loc.resize(m);
                for (int k = 0; k < n; k++) loc[mem[k]] = k + 1;
                int64_t big = 2;
                for (int i : mem)
                    for (auto &p : nb[i])
                        if (find(p.first) == by_root[a - 1].first) big = std::max(big, p.second + 2);
                DenseBlossom &bl = local_blossom;
                bl.init(n);
                for (int i : mem)
                    for (auto &p : nb[i])
                        if (find(p.first) == by_root[a - 1].first) bl.add_edge(loc[i], loc[p.first], big - p.second);
                int cnt = bl.solve();
                if (cnt * 2 != n) {
                    for (int i : mem) {
                        solved[i] = 0;
                        if (!bl.match[loc[i]]) grow.push_back({i, -1});
                    }
                    continue;
                }
                for (int i : mem) {
                    int j = mem[bl.match[loc[i]] - 1];
                    (*partner_out)[i] = j;
                    for (auto &p : nb[i])
                        if (p.first == j) (*pdist_out)[i] = (*pdist_out)[j] = p.second;
                    h[i] = big - bl.lab[loc[i]];
                    solved[i] = 1;
                    pc[i] = mem[0];
                    if (h[i] > R[i]) grow.push_back({i, h[i]});
                }
            }
            if (grow.empty()) break;
            // An open-ended request (odd component, exposed vertex) asks for ONE event more than it needs: 20 % fewer
            // matcher rounds at d25 r100 (4.49 -> 3.60 per shot) for slightly larger components (GROW_EXTRA_EVENTS in
            // the kernels' grow_some).
            for (auto &[i, to] : grow) {
                R[i] = ball((node_t)events[i], to, nb[i].size(), &nb[i], to < 0 ? (size_t)growth_extra : 0);
                if (R[i] < 0) {
                    ok = false;
                    break;
                }
                regrown[i] = 1;
            }
        }
        for (int i = 0; i < m; i++) event_index[events[i]] = -1;
        if (ok)
            for (int i = 0; i < m; i++) ok &= solved[i] != 0;
        if (!ok) last_error = "No perfect matching exists for the detection events.";
        return ok;
    }

    // euler_tours.cc:21-32
    void euler_add_edge(node_t a, node_t b) {
        uint32_t na = (uint32_t)en[a].size(), nb = (uint32_t)en[b].size();
        en[a].push_back({b, nb});
        en[b].push_back({a, na});
    }
    // euler_tours.cc:34-45
    size_t look_next(node_t n) {
        while (true) {
            if (enext[n] >= en[n].size()) return SIZE_MAX;
            if (en[n][enext[n]].node == BOUNDARY) {
                enext[n]++;
                continue;
            }
            return enext[n];
        }
    }
    void euler_reset(const std::vector<int64_t> &edges, const std::vector<uint64_t> &events) {
        for (auto n : edges) {
            en[n].clear();
            enext[n] = 0;
        }
        for (auto n : events) {
            en[n].clear();
            enext[n] = 0;
        }
    }
    // euler_tours.h:89-112 + euler_tours.cc:64-97. Calls cb(cycle) per connected component.
    bool euler_tours(
        const std::vector<int64_t> &edges, const std::vector<uint64_t> &events,
        const std::function<bool(const std::vector<node_t> &)> &cb) {
        for (size_t k = 0; k < edges.size(); k += 2) euler_add_edge((node_t)edges[k], (node_t)edges[k + 1]);
        for (size_t k = 0; k < events.size(); k += 2) euler_add_edge((node_t)events[k], (node_t)events[k + 1]);
        std::vector<node_t> cyc, cyc2;
        bool ok = true;
        for (auto n64 : edges) {
            node_t n = (node_t)n64;
            if (look_next(n) == SIZE_MAX) continue;
            cyc.clear();
            cyc.push_back(n);
            while (true) {
                // extend_cycle_depth_first
                while (true) {
                    node_t cur = cyc.back();
                    size_t k = look_next(cur);
                    if (k == SIZE_MAX) break;
                    enext[cur]++;
                    Nb nb = en[cur][k];
                    cyc.push_back(nb.node);
                    en[nb.node][nb.back].node = BOUNDARY;
                }
                // rotate_cycle_to_end_with_unfinished_node
                if (cyc.back() != cyc.front()) {
                    last_error = "Graph didn't decompose into Euler tours.";
                    euler_reset(edges, events);
                    return false;
                }
                cyc.pop_back();
                size_t ck = 1;
                for (; ck < cyc.size() && look_next(cyc[ck]) == SIZE_MAX; ck++) {
                }
                if (ck < cyc.size()) {
                    cyc2.clear();
                    cyc2.insert(cyc2.end(), cyc.begin() + ck, cyc.end());
                    cyc2.insert(cyc2.end(), cyc.begin(), cyc.begin() + ck + 1);
                    cyc.swap(cyc2);
                } else {
                    break;
                }
            }
            if (!cb(cyc)) ok = false;
        }
        euler_reset(edges, events);
        return ok;
    }

    // decoder.cc:80-138
    std::optional<obsmask> discharge_single_start(const uint8_t *dets, const std::vector<node_t> &cycle, uint8_t start, std::vector<node_t> *used) {
        used->clear();
        std::array<std::optional<obsmask>, 4> cur;
        cur[start] = obsmask{0};
        node_t cur_loc = cycle.back() >> 1;
        for (size_t k = 0; k < cycle.size(); k++) {
            node_t next_loc = cycle[k] >> 1;
            bool has_event = (dets[cur_loc >> 3] >> (cur_loc & 7)) & 1;
            if (next_loc == cur_loc && has_event && std::find(used->begin(), used->end(), cur_loc) == used->end()) {
                used->push_back(cur_loc);
                uint8_t dc = t->node_colors[cur_loc].color;
                std::array<std::optional<obsmask>, 4> after;
                after[dc] = cur[NEUTRAL];
                after[NEUTRAL] = cur[dc];
                const RgbRep &r = t->rgb_reps[cur_loc];
                if (r.weight() == 3) {
                    uint8_t c1 = next_non_neutral(dc), c2 = next_non_neutral(c1);
                    if (cur[c1].has_value()) after[c2] = *cur[c1] ^ r.obs_flip;
                    if (cur[c2].has_value()) after[c1] = *cur[c2] ^ r.obs_flip;
                }
                cur = after;
            } else {
                std::array<std::optional<obsmask>, 4> after;
                for (uint32_t cc = 0; cc < 4; cc++) {
                    if (!cur[cc].has_value()) continue;
                    for (uint32_t ncg = 0; ncg < 4; ncg++) {
                        auto f = t->drag_lookup.find(Tables::drag_key(cur_loc, next_loc, cc, ncg));
                        if (f != t->drag_lookup.end()) after[ncg] = *cur[cc] ^ f->second;
                    }
                }
                cur = after;
            }
            cur_loc = next_loc;
        }
        return cur[start];
    }
    // decoder.cc:140-203
    std::optional<obsmask> discharge_cycle(const uint8_t *dets, const std::vector<node_t> &cycle, std::vector<node_t> *used) {
        for (uint8_t c = 0; c < 4; c++) {
            auto v = discharge_single_start(dets, cycle, c, used);
            if (v.has_value()) return v;
        }
        return {};
    }

    // decoder.cc:231-277. Returns false on failure (last_error set).
    bool decode(const uint8_t *dets, size_t num_bytes, obsmask *out_obs, int64_t *out_weight, std::vector<int64_t> *out_edges) {
        std::vector<uint64_t> events;
        for (size_t k = 0; k < num_bytes; k++) {
            for (unsigned b = dets[k], k2 = 0; b; b >>= 1, k2++) {
                if (b & 1) {
                    size_t d = k * 8 + k2;
                    if (d >= t->num_dets) {
                        last_error = "Detection event bit set past the last detector.";
                        return false;
                    }
                    if (!t->node_colors[d].ignored) {
                        events.push_back(d * 2);
                        events.push_back(d * 2 + 1);
                    }
                }
            }
        }
        MatchResult mr;
        if (!match_events(events, &mr)) return false;
        if (out_weight) *out_weight = mr.weight;
        if (out_edges) *out_edges = mr.edges;
        obsmask solution = 0;
        std::vector<node_t> used;
        bool ok = euler_tours(mr.edges, events, [&](const std::vector<node_t> &cycle) {
            auto v = discharge_cycle(dets, cycle, &used);
            if (!v.has_value()) {
                last_error =
                    "Failed to lift a flattened edge cycle from the matcher into an explanation of the detection events in the cycle.";
                return false;
            }
            solution ^= *v;
            return true;
        });
        if (!ok) return false;
        *out_obs = solution;
        return true;
    }
};

static const char *charge_name(uint32_t c) {
    static const char *names[4] = {"NEUTRAL", "R", "G", "B"};
    return names[c & 3];
}

// decoder.cc:279-291 + the operator<< of each table: the `describe_decoder` dump.
static std::string describe(const Tables &t) {
    std::stringstream out;
    out << "chromobius::Decoder{\n\n.charge_graph=ChargeGraph{.nodes={\n";
    for (size_t k = 0; k < t.charge_graph.size(); k++) {
        out << "    ChargeGraphNode{.neighbors={";
        bool first = true;
        std::map<node_t, obsmask> sorted_nbrs(t.charge_graph[k].begin(), t.charge_graph[k].end());  // charge_graph.cc:43 sorts
        for (const auto &[nb, f] : sorted_nbrs) {
            if (!first) out << ", ";
            first = false;
            out << "{";
            if (nb == BOUNDARY) out << "BOUNDARY_NODE";
            else out << nb;
            out << "," << f << "}";
        }
        out << "}}, // node " << k << "\n";
    }
    out << "}}\n\n.rgb_reps={";
    for (size_t k = 0; k < t.rgb_reps.size(); k++) {
        const auto &r = t.rgb_reps[k];
        out << "\n    RgbEdge{";
        const char *nm[3] = {".red_node=", ", .green_node=", ", .blue_node="};
        for (int c = 0; c < 3; c++) {
            out << nm[c];
            if (r.rgb[c] == BOUNDARY) out << "BOUNDARY_NODE";
            else out << r.rgb[c];
        }
        out << ", .obs_flip=" << r.obs_flip << ", .charge_flip=" << charge_name(r.charge_flip) << "} // rep " << k;
    }
    out << "\n}\n\n.drag_graph=DragGraph{.mmm={\n";
    for (const auto &[k, v] : t.drag)
        out << "    " << charge_name(k[2]) << "@" << k[0] << ":" << charge_name(k[3]) << "@" << k[1] << " = " << v << "\n";
    out << "}}\n\n.mobius_dem=\n";
    char buf[64];
    for (const auto &me : t.mobius) {
        snprintf(buf, sizeof(buf), "%.17g", me.p);
        out << "error(" << buf << ")";
        for (size_t i = 0; i < me.edges.size(); i++) {
            if (i) out << " ^";
            out << " D" << me.edges[i].first << " D" << me.edges[i].second;
        }
        out << "\n";
    }
    // decoder.cc:48-52: padded with a detector declaration when the errors do not mention the last Moebius node.
    uint64_t mentioned = 0;
    for (const auto &me : t.mobius)
        for (const auto &e : me.edges) mentioned = std::max<uint64_t>(mentioned, (uint64_t)std::max(e.first, e.second) + 1);
    if (mentioned < t.num_dets * 2) out << "detector D" << (t.num_dets * 2 - 1) << "\n";
    out << "\n}\n";
    return out.str();
}

}  // namespace orc

// ---------------------------------------------------------------------------------------------
// C entry points for ctypes (tests / bench only)
// ---------------------------------------------------------------------------------------------
using orc::Decoder;

static thread_local std::string g_err;

extern "C" {

const char *orc_last_error() { return g_err.c_str(); }

void *orc_compile(const char *dem_text) {
    try {
        auto t = orc::compile(dem_text);
        return new Decoder(std::move(t));
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

void *orc_clone(void *h) {
    auto *d = (Decoder *)h;
    return new Decoder(std::make_unique<orc::Tables>(*d->t));
}

void orc_free(void *h) { delete (Decoder *)h; }

uint64_t orc_num_detectors(void *h) { return ((Decoder *)h)->t->num_dets; }
uint64_t orc_num_observables(void *h) { return ((Decoder *)h)->t->num_obs; }
double orc_normalising_constant(void *h) { return ((Decoder *)h)->t->normalising_constant; }
uint64_t orc_drag_ambiguities(void *h) { return ((Decoder *)h)->t->drag_ambiguities; }
uint64_t orc_radius_retries(void *h) { return ((Decoder *)h)->stat_radius_retries; }
void orc_set_matcher_mode(void *h, int mode) { ((Decoder *)h)->matcher_mode = mode; }
void orc_set_odd_policy(void *h, int p) { ((Decoder *)h)->odd_policy = p; }
void orc_set_growth_extra(void *h, int k) { ((Decoder *)h)->growth_extra = k < 0 ? 0 : k; }
void orc_matcher_stats(void *h, uint64_t *out6 /* 7 values */, int reset) {
    auto *d = (Decoder *)h;
    for (int k = 0; k < 4; k++) out6[k] = d->stat_blossoms[k];
    out6[4] = d->stat_rounds;
    out6[5] = d->stat_pairs_closed_form;
    out6[6] = d->stat_max_comp;
    if (reset) {
        for (int k = 0; k < 4; k++) d->stat_blossoms[k] = 0;
        d->stat_rounds = d->stat_pairs_closed_form = d->stat_max_comp = 0;
    }
}

// Caller frees with orc_free_str.
char *orc_describe(void *h) {
    std::string s = orc::describe(*((Decoder *)h)->t);
    char *r = (char *)malloc(s.size() + 1);
    memcpy(r, s.c_str(), s.size() + 1);
    return r;
}
void orc_free_str(char *s) { free(s); }

// graph export: sizes then arrays
uint64_t orc_graph_num_half_edges(void *h) { return ((Decoder *)h)->t->nbr.size(); }
void orc_graph_export(void *h, uint32_t *row_ptr, uint32_t *nbr, int32_t *wt) {
    auto &t = *((Decoder *)h)->t;
    memcpy(row_ptr, t.row_ptr.data(), t.row_ptr.size() * 4);
    memcpy(nbr, t.nbr.data(), t.nbr.size() * 4);
    memcpy(wt, t.wt.data(), t.wt.size() * 4);
}
// node table: per detector [color, basis, ignored]
void orc_node_colors(void *h, uint8_t *out) {
    auto &t = *((Decoder *)h)->t;
    for (size_t k = 0; k < t.node_colors.size(); k++) {
        out[3 * k] = t.node_colors[k].color;
        out[3 * k + 1] = t.node_colors[k].basis;
        out[3 * k + 2] = t.node_colors[k].ignored;
    }
}
// rgb reps: per detector 3 x u32 nodes, u64 obs (as two u32), u32 charge_flip -> 6 u32
void orc_rgb_reps(void *h, uint32_t *out) {
    auto &t = *((Decoder *)h)->t;
    for (size_t k = 0; k < t.rgb_reps.size(); k++) {
        const auto &r = t.rgb_reps[k];
        out[6 * k + 0] = r.rgb[0];
        out[6 * k + 1] = r.rgb[1];
        out[6 * k + 2] = r.rgb[2];
        out[6 * k + 3] = (uint32_t)r.obs_flip;
        out[6 * k + 4] = (uint32_t)(r.obs_flip >> 32);
        out[6 * k + 5] = r.charge_flip;
    }
}
uint64_t orc_drag_size(void *h) { return ((Decoder *)h)->t->drag.size(); }
// drag entries sorted by (n1,n2,c1,c2): 4 x u32 key + u64 value
void orc_drag_export(void *h, uint32_t *keys, uint64_t *vals) {
    auto &t = *((Decoder *)h)->t;
    size_t i = 0;
    for (const auto &[k, v] : t.drag) {
        memcpy(keys + 4 * i, k.data(), 16);
        vals[i] = v;
        i++;
    }
}
uint64_t orc_num_atomic_errors(void *h) { return ((Decoder *)h)->t->atomic_errors.size(); }
uint64_t orc_num_mobius_errors(void *h) { return ((Decoder *)h)->t->mobius.size(); }

// Decode a batch. obs_out: u64 per shot. weight_out (nullable): int64 per shot. Returns 0 or -(1+failing shot).
int64_t orc_decode_batch(
    void *h, const uint8_t *dets, uint64_t num_shots, uint64_t stride, uint64_t *obs_out, int64_t *weight_out) {
    auto *d = (Decoder *)h;
    size_t nb = (d->t->num_dets + 7) / 8;
    for (uint64_t s = 0; s < num_shots; s++) {
        orc::obsmask o = 0;
        int64_t w = 0;
        if (!d->decode(dets + s * stride, nb, &o, &w, nullptr)) {
            g_err = d->last_error;
            return -(int64_t)(1 + s);
        }
        obs_out[s] = o;
        if (weight_out) weight_out[s] = w;
    }
    return 0;
}

// Decode one shot and return the matched edge list (interleaved). Returns number of int64 written, or -1.
int64_t orc_decode_edges(void *h, const uint8_t *dets, int64_t *edges_out, uint64_t cap, uint64_t *obs_out, int64_t *weight_out) {
    auto *d = (Decoder *)h;
    size_t nb = (d->t->num_dets + 7) / 8;
    std::vector<int64_t> edges;
    orc::obsmask o = 0;
    int64_t w = 0;
    if (!d->decode(dets, nb, &o, &w, &edges)) {
        g_err = d->last_error;
        return -1;
    }
    *obs_out = o;
    *weight_out = w;
    size_t n = std::min<size_t>(cap, edges.size());
    memcpy(edges_out, edges.data(), n * 8);
    return (int64_t)edges.size();
}

// Stage-isolation hooks -----------------------------------------------------------------------
// Euler tours of an arbitrary interleaved edge list (euler_tours.test.cc cases): writes cycles as
// [len, nodes..., len, nodes..., ] ; returns count of ints written or -1 on failure.
int64_t orc_euler_tours(uint64_t num_nodes, const int64_t *edges, uint64_t num_edge_ints, int64_t *out, uint64_t cap) {
    auto t = std::make_unique<orc::Tables>();
    t->num_dets = (num_nodes + 1) / 2;
    Decoder d(std::move(t));
    std::vector<int64_t> e(edges, edges + num_edge_ints);
    std::vector<uint64_t> none;
    size_t n = 0;
    bool ok = d.euler_tours(e, none, [&](const std::vector<orc::node_t> &c) {
        if (n + 1 + c.size() <= cap) {
            out[n++] = (int64_t)c.size();
            for (auto v : c) out[n++] = v;
        }
        return true;
    });
    if (!ok) {
        g_err = d.last_error;
        return -1;
    }
    return (int64_t)n;
}

// Lift the given matched edges (instead of running the matcher): isolates stages 3-4.
int64_t orc_lift_edges(void *h, const uint8_t *dets, const int64_t *edges, uint64_t num_edge_ints, uint64_t *obs_out) {
    auto *d = (Decoder *)h;
    size_t nb = (d->t->num_dets + 7) / 8;
    std::vector<uint64_t> events;
    for (size_t k = 0; k < nb; k++)
        for (unsigned b = dets[k], k2 = 0; b; b >>= 1, k2++)
            if ((b & 1) && !d->t->node_colors[k * 8 + k2].ignored) {
                events.push_back((k * 8 + k2) * 2);
                events.push_back((k * 8 + k2) * 2 + 1);
            }
    std::vector<int64_t> e(edges, edges + num_edge_ints);
    orc::obsmask sol = 0;
    std::vector<orc::node_t> used;
    bool ok = d->euler_tours(e, events, [&](const std::vector<orc::node_t> &cycle) {
        auto v = d->discharge_cycle(dets, cycle, &used);
        if (!v.has_value()) return false;
        sol ^= *v;
        return true;
    });
    if (!ok) return -1;
    *obs_out = sol;
    return 0;
}

// Match only: event list in, pairs + weight out (for cross-checking against networkx).
int64_t orc_match_events(void *h, const uint64_t *events, uint64_t m, int32_t *partner_out, int64_t *weight_out) {
    auto *d = (Decoder *)h;
    std::vector<uint64_t> ev(events, events + m);
    orc::MatchResult mr;
    std::vector<std::pair<int, int>> pairs;
    if (!d->match_events(ev, &mr, &pairs)) {
        g_err = d->last_error;
        return -1;
    }
    for (auto &p : pairs) {
        partner_out[p.first] = p.second;
        partner_out[p.second] = p.first;
    }
    *weight_out = mr.weight;
    return 0;
}

}  // extern "C"
