// Host compile step: DEM text -> flat decoder tables.  See host_compile.h.
//
// Restates (does not copy) the reference passes, with flat sorted arrays + hashing in place of the
// reference's std::map / std::unordered_map containers. Citations relative to /root/reference/src/chromobius.
#include "host_compile.h"

#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <unordered_map>

namespace cb200 {

// =================================================================================================
// DEM text -> FlatDem.  Grammar: SURVEY.md App. C.1 (stim's DetectorErrorModel text format).
// =================================================================================================
namespace {

enum Op : uint8_t { OP_ERROR, OP_DETECTOR, OP_SHIFT, OP_OBSERVABLE, OP_REPEAT, OP_END };

struct ParsedLine {
    Op op;
    uint32_t line;
    uint32_t first_arg, num_args;
    uint32_t first_tgt, num_tgts;
    uint64_t count;     // repeat count
    uint32_t partner;   // OP_REPEAT -> index of matching OP_END, and back
};

struct Lexer {
    const std::string &s;
    size_t i = 0;
    explicit Lexer(const std::string &str) : s(str) {}
    void skip_ws() {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\r')) i++;
    }
    bool done() {
        skip_ws();
        return i >= s.size();
    }
};

[[noreturn]] void fail(const std::string &msg) { throw std::invalid_argument(msg); }

}  // namespace

FlatDem FlatDem::parse(const std::string &text) {
    FlatDem out;
    std::vector<ParsedLine> prog;
    std::vector<double> args;
    std::vector<uint64_t> tgts;
    std::vector<uint32_t> open_blocks;

    size_t pos = 0;
    while (pos <= text.size()) {
        size_t eol = text.find('\n', pos);
        if (eol == std::string::npos) eol = text.size();
        std::string line = text.substr(pos, eol - pos);
        pos = eol + 1;
        size_t hash = line.find('#');
        if (hash != std::string::npos) line.resize(hash);
        Lexer lx(line);
        if (lx.done()) {
            if (pos > text.size()) break;
            continue;
        }
        uint32_t line_index = (uint32_t)out.lines.size();
        {
            size_t a = lx.i, b = line.size();
            while (b > a && isspace((unsigned char)line[b - 1])) b--;
            out.lines.push_back(line.substr(a, b - a));
        }
        const std::string &shown = out.lines.back();
        ParsedLine pl{};
        pl.line = line_index;
        if (line[lx.i] == '}') {
            lx.i++;
            if (!lx.done()) fail("Unexpected text after '}' in: " + shown);
            if (open_blocks.empty()) fail("Unmatched '}' in detector error model.");
            pl.op = OP_END;
            pl.partner = open_blocks.back();
            prog[open_blocks.back()].partner = (uint32_t)prog.size();
            open_blocks.pop_back();
            prog.push_back(pl);
            continue;
        }
        size_t name_start = lx.i;
        while (lx.i < line.size() && (isalpha((unsigned char)line[lx.i]) || line[lx.i] == '_')) lx.i++;
        std::string name = line.substr(name_start, lx.i - name_start);
        for (auto &c : name) c = (char)tolower((unsigned char)c);
        if (lx.i < line.size() && line[lx.i] == '[') {  // instruction tag: ignored
            size_t close = line.find(']', lx.i);
            if (close == std::string::npos) fail("Unterminated tag in: " + shown);
            lx.i = close + 1;
        }
        pl.first_arg = (uint32_t)args.size();
        if (lx.i < line.size() && line[lx.i] == '(') {
            lx.i++;
            while (true) {
                lx.skip_ws();
                if (lx.i < line.size() && line[lx.i] == ')') {
                    lx.i++;
                    break;
                }
                const char *start = line.c_str() + lx.i;
                char *end = nullptr;
                double v = strtod(start, &end);
                if (end == start) fail("Bad numeric argument in: " + shown);
                args.push_back(v);
                lx.i += (size_t)(end - start);
                lx.skip_ws();
                if (lx.i < line.size() && line[lx.i] == ',') lx.i++;
                else if (lx.i >= line.size()) fail("Missing ')' in: " + shown);
            }
        }
        pl.num_args = (uint32_t)args.size() - pl.first_arg;

        if (name == "error") pl.op = OP_ERROR;
        else if (name == "detector") pl.op = OP_DETECTOR;
        else if (name == "shift_detectors") pl.op = OP_SHIFT;
        else if (name == "logical_observable") pl.op = OP_OBSERVABLE;
        else if (name == "repeat") pl.op = OP_REPEAT;
        else fail("Unrecognized instruction type: " + shown);

        pl.first_tgt = (uint32_t)tgts.size();
        bool saw_brace = false;
        while (!lx.done()) {
            char c = line[lx.i];
            if (c == '^') {
                tgts.push_back(TGT_SEPARATOR);
                lx.i++;
            } else if (c == '{') {
                saw_brace = true;
                lx.i++;
            } else if (c == 'D' || c == 'd' || c == 'L' || c == 'l' || isdigit((unsigned char)c)) {
                bool is_d = (c == 'D' || c == 'd'), is_l = (c == 'L' || c == 'l');
                if (is_d || is_l) lx.i++;
                if (lx.i >= line.size() || !isdigit((unsigned char)line[lx.i])) fail("Bad target in: " + shown);
                uint64_t v = 0;
                while (lx.i < line.size() && isdigit((unsigned char)line[lx.i])) v = v * 10 + (uint64_t)(line[lx.i++] - '0');
                if (is_l) v |= TGT_OBS_FLAG;
                tgts.push_back(v);
            } else {
                fail("Unrecognized target in: " + shown);
            }
        }
        pl.num_tgts = (uint32_t)tgts.size() - pl.first_tgt;
        if (pl.op == OP_REPEAT) {
            if (!saw_brace || pl.num_tgts != 1) fail("Expected 'repeat N {' but got: " + shown);
            pl.count = tgts[pl.first_tgt];
            open_blocks.push_back((uint32_t)prog.size());
        } else if (saw_brace) {
            fail("Unexpected '{' in: " + shown);
        }
        if (pl.op == OP_ERROR && pl.num_args != 1) fail("An error instruction needs exactly one probability: " + shown);
        prog.push_back(pl);
    }
    if (!open_blocks.empty()) fail("Missing '}' in detector error model.");

    // Execute the program: unroll repeats, apply shifts.
    struct Loop {
        uint32_t body_start;
        uint64_t remaining;
    };
    std::vector<Loop> loops;
    uint64_t det_shift = 0;
    std::vector<double> coord_shift;
    bool any_det = false, any_obs = false;
    uint64_t max_det = 0, max_obs = 0;
    for (uint32_t pc = 0; pc < prog.size();) {
        const ParsedLine &pl = prog[pc];
        switch (pl.op) {
            case OP_REPEAT:
                if (pl.count == 0) {
                    pc = pl.partner + 1;
                } else {
                    loops.push_back({pc + 1, pl.count});
                    pc++;
                }
                continue;
            case OP_END:
                if (--loops.back().remaining > 0) {
                    pc = loops.back().body_start;
                } else {
                    loops.pop_back();
                    pc++;
                }
                continue;
            case OP_ERROR: {
                FlatDemError e;
                e.p = args[pl.first_arg];
                e.first_target = (uint32_t)out.targets.size();
                e.line = pl.line;
                for (uint32_t k = 0; k < pl.num_tgts; k++) {
                    uint64_t t = tgts[pl.first_tgt + k];
                    if (t == TGT_SEPARATOR) {
                    } else if (t & TGT_OBS_FLAG) {
                        any_obs = true;
                        max_obs = std::max(max_obs, t & ~TGT_OBS_FLAG);
                    } else {
                        t += det_shift;
                        any_det = true;
                        max_det = std::max(max_det, t);
                    }
                    out.targets.push_back(t);
                }
                e.num_targets = pl.num_tgts;
                out.errors.push_back(e);
                break;
            }
            case OP_DETECTOR:
                for (uint32_t k = 0; k < pl.num_tgts; k++) {
                    uint64_t t = tgts[pl.first_tgt + k];
                    if (t == TGT_SEPARATOR || (t & TGT_OBS_FLAG)) fail("Bad detector target in: " + out.lines[pl.line]);
                    FlatDetector fd{};
                    fd.id = t + det_shift;
                    fd.line = pl.line;
                    fd.num_coords = pl.num_args;
                    for (uint32_t a = 0; a < pl.num_args && a < 8; a++) {
                        fd.coords[a] = args[pl.first_arg + a] + (a < coord_shift.size() ? coord_shift[a] : 0.0);
                    }
                    any_det = true;
                    max_det = std::max(max_det, fd.id);
                    out.detectors.push_back(fd);
                }
                break;
            case OP_SHIFT:
                if (pl.num_tgts > 0) det_shift += tgts[pl.first_tgt];
                if (coord_shift.size() < pl.num_args) coord_shift.resize(pl.num_args, 0.0);
                for (uint32_t a = 0; a < pl.num_args; a++) coord_shift[a] += args[pl.first_arg + a];
                break;
            case OP_OBSERVABLE:
                for (uint32_t k = 0; k < pl.num_tgts; k++) {
                    uint64_t t = tgts[pl.first_tgt + k];
                    if (t & TGT_OBS_FLAG) {
                        any_obs = true;
                        max_obs = std::max(max_obs, t & ~TGT_OBS_FLAG);
                    }
                }
                break;
        }
        pc++;
    }
    out.num_detectors = any_det ? max_det + 1 : 0;
    out.num_observables = any_obs ? max_obs + 1 : 0;
    return out;
}

// =================================================================================================
// Compile passes
// =================================================================================================
namespace {

constexpr uint32_t NONE = NO_INDEX;
enum : uint8_t { NEUTRAL = 0, RED = 1, GREEN = 2, BLUE = 3 };
enum : uint8_t { BASIS_UNKNOWN = 0, BASIS_X = 1, BASIS_Z = 2 };

const char *COLOR_MSG =
    "Expected all detectors to have at least 4 coordinates, with the 4th identifying the basis and color "
    "(RedX=0, GreenX=1, BlueX=2, RedZ=3, GreenZ=4, BlueZ=5), but got ";

inline uint8_t next_color(uint8_t c) { return (uint8_t)(c % 3 + 1); }

// Sorted triple of detectors, NONE-padded at the end (datatypes/atomic_error.h:46-66).
struct Atom {
    uint32_t d[3];
    bool operator==(const Atom &o) const { return d[0] == o.d[0] && d[1] == o.d[1] && d[2] == o.d[2]; }
    bool operator<(const Atom &o) const {
        if (d[0] != o.d[0]) return d[0] < o.d[0];
        if (d[1] != o.d[1]) return d[1] < o.d[1];
        return d[2] < o.d[2];
    }
    int weight() const { return (d[0] != NONE) + (d[1] != NONE) + (d[2] != NONE); }
    bool empty() const { return d[0] == NONE; }
};
inline Atom make_atom(uint32_t a, uint32_t b = NONE, uint32_t c = NONE) {
    if (a > b) std::swap(a, b);
    if (b > c) std::swap(b, c);
    if (a > b) std::swap(a, b);
    return Atom{{a, b, c}};
}
struct AtomHash {
    size_t operator()(const Atom &a) const {
        uint64_t h = a.d[0] * 0x9E3779B97F4A7C15ull;
        h ^= (a.d[1] + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
        h ^= (a.d[2] + 0x165667B1ull) * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        return (size_t)h;
    }
};
typedef std::unordered_map<Atom, uint64_t, AtomHash> AtomTable;

// The order in which libstdc++'s std::unordered_map<uint32_t, T> (identity hash, maximum load factor 1, prime bucket
// counts) iterates its keys after a sequence of first-time insertions. The reference's charge graph is such a container
// (graph/charge_graph.h:52) and its two-step path search returns the first path met while iterating it
// (graph/drag_graph.cc:87), so this order decides 1-17 % of the predictions on some circuits. It is restated here, not
// taken from the platform's <unordered_map>, so the tables do not depend on the standard library the product is built
// with; tests/test_decode_path_cpu.py checks it against the real container (which the oracle uses).
//   - an element whose bucket is empty becomes the head of the whole list; otherwise it goes to the front of its bucket;
//   - growth happens before an insertion that would exceed the bucket count: 1 -> 13 -> 29 -> 59 -> 127 -> ... (the
//     smallest listed prime >= twice the old count), and re-buckets the elements in list order by the same rule.
class StdHashOrder {
    static constexpr int32_t NIL = -1, HEAD = -2;
    std::vector<uint32_t> keys_;
    std::vector<int32_t> next_;
    std::vector<int32_t> before_{NIL};  // per bucket: the element its run follows (HEAD: the run starts the list), NIL: empty
    int32_t head_ = NIL;
    size_t next_resize_ = 0;

    int32_t &link(int32_t prev) { return prev == HEAD ? head_ : next_[prev]; }
    static size_t next_bucket_count(size_t at_least) {
        static const size_t growth[] = {13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753, 42043, 85229, 172933,
                                        351061, 712697, 1447153};
        for (size_t c : growth)
            if (c >= at_least) return c;
        fail("charge graph node with more than a million neighbours");
        return 0;
    }
    void place(int32_t i, std::vector<int32_t> &before, size_t &first_bucket, bool rehashing) {
        size_t b = keys_[i] % before.size();
        if (before[b] == NIL) {
            next_[i] = head_;
            head_ = i;
            before[b] = HEAD;
            if (next_[i] != NIL) before[rehashing ? first_bucket : keys_[next_[i]] % before.size()] = i;
            first_bucket = b;
        } else {
            next_[i] = link(before[b]);
            link(before[b]) = i;
        }
    }

  public:
    // `key` must not have been inserted before.
    void insert_new(uint32_t key) {
        size_t n = keys_.size();
        if (n + 1 > next_resize_) {
            size_t want = std::max<size_t>(n + 1, next_resize_ ? 0 : 11);
            if (want >= before_.size()) {
                size_t count = next_bucket_count(std::max(want + 1, before_.size() * 2));
                std::vector<int32_t> before(count, NIL);
                int32_t p = head_;
                head_ = NIL;
                size_t first_bucket = 0;
                while (p != NIL) {
                    int32_t nx = next_[p];
                    place(p, before, first_bucket, true);
                    p = nx;
                }
                before_.swap(before);
                next_resize_ = count;
            } else {
                next_resize_ = before_.size();
            }
        }
        keys_.push_back(key);
        next_.push_back(NIL);
        size_t unused = 0;
        place((int32_t)n, before_, unused, false);
    }
    template <class F>
    void for_each(F f) const {
        for (int32_t p = head_; p != NIL; p = next_[p]) f(keys_[p]);
    }
};

struct Ctx {
    const FlatDem &dem;
    HostTables &t;
    AtomTable atoms;
    std::vector<StdHashOrder> charge_visit_order;  // per charge-graph node: the reference container's iteration order
    Ctx(const FlatDem &d, HostTables &tt) : dem(d), t(tt) {}

    uint8_t color_of(uint32_t d) const { return t.color[d]; }
    uint8_t net_charge(const Atom &a) const {
        uint8_t c = 0;
        for (int k = 0; k < 3; k++)
            if (a.d[k] != NONE) c ^= t.color[a.d[k]];
        return c;
    }
};

// graph/collect_nodes.cc:21-131 + datatypes/stim_integration.cc:21-67
void pass_collect_nodes(Ctx &cx) {
    size_t n = cx.dem.num_detectors;
    cx.t.color.assign(n, NEUTRAL);
    cx.t.basis.assign(n, BASIS_UNKNOWN);
    cx.t.ignored.assign(n, 0);
    for (const auto &fd : cx.dem.detectors) {
        bool ok = fd.num_coords > 3;
        double c = ok ? fd.coords[3] : -2;
        int r = 0;
        if (c < -1 || c > 5) {
            ok = false;
        } else {
            r = (int)c;
            if ((double)r != c) ok = false;
        }
        if (!ok) fail(std::string(COLOR_MSG) + cx.dem.lines[fd.line]);
        if (r == -1) {
            cx.t.color[fd.id] = NEUTRAL;
            cx.t.basis[fd.id] = BASIS_UNKNOWN;
            cx.t.ignored[fd.id] = 1;
        } else {
            cx.t.color[fd.id] = (uint8_t)(r % 3 + 1);
            cx.t.basis[fd.id] = (uint8_t)(r / 3 + 1);
            cx.t.ignored[fd.id] = 0;
        }
    }
}

// graph/collect_atomic_errors.cc:57-100: symptoms of one DEM error (xor-cancelled, sorted), observable mask.
void symptoms_of(Ctx &cx, const FlatDemError &e, std::vector<uint32_t> &dets, uint64_t &obs) {
    dets.clear();
    obs = 0;
    for (uint32_t k = 0; k < e.num_targets; k++) {
        uint64_t tg = cx.dem.targets[e.first_target + k];
        if (tg == TGT_SEPARATOR) continue;
        if (tg & TGT_OBS_FLAG) {
            uint64_t id = tg & ~TGT_OBS_FLAG;
            if (id >= 64) {
                fail("Max logical observable is L63 but a larger one appeared in '" + cx.dem.lines[e.line] + "'");
            }
            obs ^= 1ull << id;
            continue;
        }
        if (tg > 0xFFFFFFFEull) {
            std::stringstream ss;
            ss << "The detector error model is too large. It has a detector with index " << tg
               << " but the max supported by chromobius is " << 0xFFFFFFFFull;
            fail(ss.str());
        }
        if (cx.t.ignored[tg]) continue;
        if (cx.t.color[tg] == NEUTRAL) fail(std::string(COLOR_MSG) + cx.dem.lines[e.line]);
        dets.push_back((uint32_t)tg);
    }
    std::sort(dets.begin(), dets.end());
    size_t w = 0;
    for (size_t k = 0; k < dets.size(); k++) {
        if (w > 0 && dets[w - 1] == dets[k]) w--;
        else dets[w++] = dets[k];
    }
    dets.resize(w);
}

// graph/collect_atomic_errors.cc:19-55. Registers (dets -> obs) in `table` when the symptom set is atomic.
Atom register_if_atomic(Ctx &cx, const std::vector<uint32_t> &dets, uint64_t obs, AtomTable &table) {
    const Atom none{{NONE, NONE, NONE}};
    size_t n = dets.size();
    if (n == 0 || n > 3) return none;
    for (size_t k = 1; k < n; k++)
        if (cx.t.basis[dets[k]] != cx.t.basis[dets[0]]) return none;
    if (n == 3 && (cx.color_of(dets[0]) ^ cx.color_of(dets[1]) ^ cx.color_of(dets[2])) != NEUTRAL) return none;
    Atom a = make_atom(dets[0], n > 1 ? dets[1] : NONE, n > 2 ? dets[2] : NONE);
    table[a] = obs;
    return a;
}

// graph/collect_composite_errors.cc:21-42
struct Grow {
    Atom best{{NONE, NONE, NONE}};
    int score = 0;
};
void consider(Ctx &cx, const Atom &e1, const Atom &e2, Grow &g) {
    bool h1 = cx.atoms.count(e1) != 0, h2 = cx.atoms.count(e2) != 0;
    int score = (int)h1 + 2 * (int)h2;
    if (score <= g.score) return;
    if (score == 1 && e2.weight() == 3 && cx.net_charge(e2) != NEUTRAL) return;
    if (score == 2 && e1.weight() == 3 && cx.net_charge(e1) != NEUTRAL) return;
    g.best = h2 ? e2 : e1;
    g.score = score;
}

// graph/collect_composite_errors.cc:44-228: split `dets` (one basis, sorted) into a chosen subset + the rest and
// score the split. `pick` enumerates subsets of the given size in lexicographic index order.
Atom split_single_basis(Ctx &cx, const std::vector<uint32_t> &dets) {
    const Atom none{{NONE, NONE, NONE}};
    size_t n = dets.size();
    if (n >= 1 && n <= 3) {
        Atom solo = make_atom(dets[0], n > 1 ? dets[1] : NONE, n > 2 ? dets[2] : NONE);
        if (cx.atoms.count(solo)) return solo;
    }
    if (n == 2) {
        // Reference quirk (collect_composite_errors.cc:49-50): only dets[0] is tried as a singlet.
        Atom a = make_atom(dets[0]);
        return cx.atoms.count(a) ? a : none;
    }
    if (n < 3 || n > 6) return none;

    Grow g;
    auto run = [&](size_t pick, bool early_stop) {
        // iterate index subsets of size `pick` in lexicographic order
        size_t idx[3] = {0, 1, 2};
        while (true) {
            uint32_t chosen[3] = {NONE, NONE, NONE}, rest[3] = {NONE, NONE, NONE};
            size_t rc = 0;
            for (size_t i = 0, c = 0; i < n; i++) {
                if (c < pick && idx[c] == i) {
                    chosen[c++] = dets[i];
                } else if (rc < 3) {
                    rest[rc++] = dets[i];
                }
            }
            consider(cx, make_atom(chosen[0], chosen[1], chosen[2]), make_atom(rest[0], rest[1], rest[2]), g);
            // advance
            size_t old0 = idx[0];
            int k = (int)pick - 1;
            while (k >= 0 && idx[k] == n - pick + (size_t)k) k--;
            if (k < 0) break;
            idx[k]++;
            for (size_t j = (size_t)k + 1; j < pick; j++) idx[j] = idx[j - 1] + 1;
            // The reference tests `best_score < 2` only in the OUTER (k1) loop condition, i.e. it finishes the inner
            // loops of the current k1 first; so stop only when idx[0] advances.
            if (early_stop && idx[0] != old0 && g.score >= 2) break;
        }
    };
    switch (n) {
        case 3:
            run(1, false);  // 1:2
            break;
        case 4:
            run(2, true);   // 2:2 (outer loop stops once best_score >= 2)
            run(1, false);  // 1:3
            break;
        case 5:
            run(2, true);  // 2:3
            break;
        case 6:
            run(3, true);  // 3:3
            break;
    }
    return g.best;
}

// graph/collect_composite_errors.cc:230-399 with the default options of the Python API
// (drop_mobius_errors_involving_remnant_errors=true, ignore_decomposition_failures=false).
void pass_mobius_dem(Ctx &cx) {
    std::vector<uint32_t> dets, part[2];
    std::vector<Atom> pieces;
    AtomTable remnants;
    uint64_t obs;
    for (const auto &e : cx.dem.errors) {
        symptoms_of(cx, e, dets, obs);
        part[0].clear();
        part[1].clear();
        for (uint32_t d : dets) part[cx.t.basis[d] == BASIS_X ? 0 : 1].push_back(d);
        pieces.clear();
        remnants.clear();
        for (int round = 0; round < 3; round++) {
            for (int b = 0; b < 2; b++) {
                std::vector<uint32_t> &pd = part[b];
                Atom got = (round == 2) ? register_if_atomic(cx, pd, obs, remnants) : split_single_basis(cx, pd);
                if (got.empty()) continue;
                for (int k = 0; k < got.weight(); k++) {
                    for (size_t i = 0; i < pd.size(); i++) {
                        if (pd[i] == got.d[k]) {
                            pd[i] = pd.back();
                            pd.pop_back();
                            break;
                        }
                    }
                }
                auto it = cx.atoms.find(got);
                obs ^= (it != cx.atoms.end()) ? it->second : remnants.at(got);
                pieces.push_back(got);
            }
        }
        if (!part[0].empty() || !part[1].empty()) {
            std::stringstream ss;
            ss << "Failed to decompose a complex error instruction into basic errors.\n";
            ss << "    The instruction (after shifting): error(" << e.p << ")";
            for (uint32_t k = 0; k < e.num_targets; k++) {
                uint64_t tg = cx.dem.targets[e.first_target + k];
                if (tg == TGT_SEPARATOR) ss << " ^";
                else if (tg & TGT_OBS_FLAG) ss << " L" << (tg & ~TGT_OBS_FLAG);
                else ss << " D" << tg;
            }
            ss << "\n    The undecomposed X detectors: [";
            for (size_t i = 0; i < part[0].size(); i++) ss << (i ? ", " : "") << part[0][i];
            ss << "]\n    The undecomposed Z detectors: [";
            for (size_t i = 0; i < part[1].size(); i++) ss << (i ? ", " : "") << part[1][i];
            ss << "]\n";
            ss << "This problem can unfortunately be quite difficult to debug. Likely causes are:\n";
            ss << "    (1) The source circuit has detectors with invalid color/basis annotations.\n";
            ss << "    (2) The source circuit is producing errors too complex to decompose (e.g. more than 6 symptoms in "
                  "one basis).\n";
            ss << "    (3) chromobius is missing logic for a corner case present in the source circuit; a corner case that "
                  "didn't appear in the test circuits used during development.\n";
            fail(ss.str());
        }
        if (!remnants.empty()) pieces.clear();  // errors that needed a remnant are dropped entirely

        MobiusDemError me;
        me.first_edge = (uint32_t)(cx.t.mobius_edges.size() / 2);
        bool corner = false;
        for (const Atom &a : pieces) {
            // datatypes/atomic_error.h:104-140 -- the Moebius doubling rule.
            int w = a.weight();
            auto push = [&](uint32_t x, uint32_t y) {
                cx.t.mobius_edges.push_back(x);
                cx.t.mobius_edges.push_back(y);
            };
            if (w == 1) {
                corner = true;
                push(2 * a.d[0], 2 * a.d[0] + 1);
            } else if (w == 2) {
                uint32_t swap_sub = ((cx.color_of(a.d[0]) ^ cx.color_of(a.d[1])) == GREEN) ? 1u : 0u;
                push(2 * a.d[0], (2 * a.d[1]) ^ swap_sub);
                push(2 * a.d[0] + 1, (2 * a.d[1] + 1) ^ swap_sub);
            } else {
                uint32_t by_color[4] = {NONE, NONE, NONE, NONE};
                for (int k = 0; k < 3; k++) by_color[cx.color_of(a.d[k])] = a.d[k];
                uint32_t r = by_color[RED], g = by_color[GREEN], b = by_color[BLUE];
                // node = 2*det + subgraph offset; offsets from datatypes/color_basis.h:55-60
                uint32_t e[3][2] = {{2 * r + 1, 2 * g + 1}, {2 * g, 2 * b}, {2 * r, 2 * b + 1}};
                for (auto &p : e) push(std::min(p[0], p[1]), std::max(p[0], p[1]));
            }
        }
        me.num_edges = (uint32_t)(cx.t.mobius_edges.size() / 2) - me.first_edge;
        if (me.num_edges) {
            me.p = corner ? e.p * e.p : e.p;
            cx.t.mobius_errors.push_back(me);
        }
    }
}

// graph/choose_rgb_reps.cc:19-92 (iteration in atom key order)
void pass_rgb_reps(Ctx &cx, const std::vector<Atom> &order, const std::vector<uint64_t> &obs) {
    RgbRep blank{{NONE, NONE, NONE}, 0, NEUTRAL};
    cx.t.rgb_reps.assign(cx.t.num_dets, blank);
    auto rep_weight = [](const RgbRep &r) { return (r.node[0] != NONE) + (r.node[1] != NONE) + (r.node[2] != NONE); };
    for (size_t i = 0; i < order.size(); i++) {
        RgbRep rep = blank;
        rep.obs_flip = obs[i];
        int w = order[i].weight();
        for (int k = 0; k < w; k++) {
            uint8_t c = cx.color_of(order[i].d[k]);
            rep.node[c - 1] = order[i].d[k];
            rep.charge_flip ^= c;
        }
        if (rep_weight(rep) != w) continue;
        for (int k = 0; k < w; k++) {
            RgbRep &cur = cx.t.rgb_reps[order[i].d[k]];
            if (w > rep_weight(cur)) cur = rep;
        }
    }
    for (size_t i = 0; i < order.size(); i++) {
        if (order[i].weight() != 2) continue;
        uint32_t a = order[i].d[0], b = order[i].d[1];
        uint8_t ca = cx.color_of(a), cb = cx.color_of(b);
        if (ca != cb) continue;
        int wa = rep_weight(cx.t.rgb_reps[a]), wb = rep_weight(cx.t.rgb_reps[b]);
        if (wa == 0 && wb > 0) {
            cx.t.rgb_reps[a] = cx.t.rgb_reps[b];
            cx.t.rgb_reps[a].node[ca - 1] = a;
            cx.t.rgb_reps[a].obs_flip ^= obs[i];
        }
        if (wb == 0 && wa > 0) {
            cx.t.rgb_reps[b] = cx.t.rgb_reps[a];
            cx.t.rgb_reps[b].node[cb - 1] = b;
            cx.t.rgb_reps[b].obs_flip ^= obs[i];
        }
    }
}

// graph/charge_graph.cc:80-160. Per-node neighbour lists kept sorted; writes overwrite (last writer wins). The order in
// which the reference's container would iterate each node's neighbours is tracked beside it (Ctx::charge_visit_order).
void pass_charge_graph(Ctx &cx, const std::vector<Atom> &order, const std::vector<uint64_t> &obs) {
    auto &g = cx.t.charge_graph;
    g.assign(cx.t.num_dets, {});
    cx.charge_visit_order.assign(cx.t.num_dets, {});
    auto put = [&](uint32_t at, uint32_t nb, uint64_t f) {
        auto &row = g[at];
        auto it = std::lower_bound(row.begin(), row.end(), std::make_pair(nb, (uint64_t)0),
                                   [](const auto &x, const auto &y) { return x.first < y.first; });
        if (it != row.end() && it->first == nb) {
            it->second = f;
        } else {
            row.insert(it, {nb, f});
            cx.charge_visit_order[at].insert_new(nb);
        }
    };
    auto link = [&](uint32_t a, uint32_t b, uint64_t f) {
        if (a != NONE) put(a, b, f);
        if (b != NONE) put(b, a, f);
    };
    for (uint32_t k = 0; k < cx.t.num_dets; k++) put(k, k, 0);
    for (size_t i = 0; i < order.size(); i++)
        if (order[i].d[2] == NONE) link(order[i].d[0], order[i].d[1], obs[i]);
    // errors incident to each detector, in atom order
    std::vector<uint32_t> deg(cx.t.num_dets + 1, 0);
    for (const Atom &a : order)
        for (int k = 0; k < a.weight(); k++) deg[a.d[k] + 1]++;
    for (size_t i = 1; i < deg.size(); i++) deg[i] += deg[i - 1];
    std::vector<uint32_t> incident(deg.back()), fill(deg.begin(), deg.end() - 1);
    for (uint32_t i = 0; i < order.size(); i++)
        for (int k = 0; k < order[i].weight(); k++) incident[fill[order[i].d[k]]++] = i;
    for (uint32_t n = 0; n < cx.t.num_dets; n++) {
        for (uint32_t x = deg[n]; x < deg[n + 1]; x++) {
            for (uint32_t y = x + 1; y < deg[n + 1]; y++) {
                const Atom &e1 = order[incident[x]], &e2 = order[incident[y]];
                if (e1.weight() < 3 && e2.weight() < 3) continue;
                uint32_t all[6] = {e1.d[0], e1.d[1], e1.d[2], e2.d[0], e2.d[1], e2.d[2]};
                std::sort(all, all + 6);
                uint32_t odd[6];
                int m = 0;
                for (int k = 0; k < 6; k++) {
                    if (m > 0 && odd[m - 1] == all[k]) m--;
                    else odd[m++] = all[k];
                }
                uint32_t a, b;
                if (m == 1) {
                    a = odd[0];
                    b = NONE;
                } else if (m == 2 || (m == 3 && odd[2] == NONE)) {
                    a = odd[0];
                    b = odd[1];
                } else {
                    continue;
                }
                link(a, b, obs[incident[x]] ^ obs[incident[y]]);
            }
        }
    }
}

// graph/drag_graph.cc:24-105 (BfsSearcher::find_shortest_path_obs_flip with max_cost = 2): obs flip of a path of at
// most 2 charge-graph edges. The reference pushes src's neighbours in its container's iteration order (src itself
// included through its self loop; the boundary is skipped), pops them from the back and returns the first one that has
// dst as a neighbour. Searches whose answer depends on that order are counted in drag_ambiguities.
bool short_path(Ctx &cx, uint32_t src, uint32_t dst, uint64_t &flip) {
    const auto &g = cx.t.charge_graph;
    if (src == dst) {
        flip = 0;
        return true;
    }
    auto find_in = [&](uint32_t at, uint32_t nb, uint64_t &f) {
        const auto &row = g[at];
        auto it = std::lower_bound(row.begin(), row.end(), std::make_pair(nb, (uint64_t)0),
                                   [](const auto &x, const auto &y) { return x.first < y.first; });
        if (it == row.end() || it->first != nb) return false;
        f = it->second;
        return true;
    };
    if (find_in(src, dst, flip)) return true;
    std::vector<uint32_t> level1;
    cx.charge_visit_order[src].for_each([&](uint32_t nb) {
        if (nb != NONE) level1.push_back(nb);
    });
    bool found = false;
    for (size_t i = level1.size(); i-- > 0;) {
        uint64_t f1 = 0, f2;
        if (!find_in(level1[i], dst, f2)) continue;
        find_in(src, level1[i], f1);
        uint64_t v = f1 ^ f2;
        if (!found) {
            found = true;
            flip = v;
        } else if (v != flip) {
            cx.t.drag_ambiguities++;
            break;
        }
    }
    return found;
}

struct PairHash {
    size_t operator()(uint64_t k) const { return (size_t)(k * 0x9E3779B97F4A7C15ull >> 7); }
};

// graph/drag_graph.cc:107-203, emitting DragEntry records instead of a std::map.
void pass_drag(Ctx &cx, const std::vector<Atom> &order, const std::vector<uint64_t> &obs) {
    auto &t = cx.t;
    std::unordered_map<uint64_t, uint32_t, PairHash> index;
    auto entry_for = [&](uint32_t n1, uint32_t n2) -> DragEntry & {
        uint64_t key = ((uint64_t)std::min(n1, n2) << 32) | std::max(n1, n2);
        auto it = index.find(key);
        if (it == index.end()) {
            it = index.emplace(key, (uint32_t)t.drag.size()).first;
            DragEntry blank{};
            t.drag.push_back(blank);
            t.drag_pair.push_back(std::min(n1, n2));
            t.drag_pair.push_back(std::max(n1, n2));
        }
        return t.drag[it->second];
    };
    // mmm[{n1,n2,c1,c2}] = mmm[{n2,n1,c2,c1}] = f
    auto set = [&](uint32_t n1, uint32_t n2, uint8_t c1, uint8_t c2, uint64_t f) {
        DragEntry &e = entry_for(n1, n2);
        if (n1 > n2) std::swap(c1, c2);  // store in the (min, max) orientation
        e.obs[4 * c1 + c2] = f;
        e.valid |= 1u << (4 * c1 + c2);
        if (n1 == n2) {
            e.obs[4 * c2 + c1] = f;
            e.valid |= 1u << (4 * c2 + c1);
        }
    };
    auto rep_weight = [&](uint32_t d) {
        const RgbRep &r = t.rgb_reps[d];
        return (r.node[0] != NONE) + (r.node[1] != NONE) + (r.node[2] != NONE);
    };
    auto charge_edge = [&](uint32_t a, uint32_t b) -> uint64_t {
        const auto &row = t.charge_graph[a];
        for (const auto &p : row)
            if (p.first == b) return p.second;
        fail("internal: missing charge graph edge");
    };
    auto dump_to_boundary = [&](uint32_t a, uint32_t b, uint64_t ab_flip) {
        if (rep_weight(a) != 3) return;
        uint8_t ca = cx.color_of(a), cb = cx.color_of(b), c = ca ^ cb;
        if (c == NEUTRAL) return;
        uint64_t f1, f2;
        bool ok1 = short_path(cx, t.rgb_reps[a].node[ca - 1], a, f1);
        bool ok2 = short_path(cx, t.rgb_reps[a].node[cb - 1], b, f2);
        if (ok1 && ok2) set(a, b, c, NEUTRAL, f1 ^ f2 ^ t.rgb_reps[a].obs_flip ^ ab_flip);
    };
    std::vector<uint64_t> decomposed;  // (min << 32 | max) pairs, sorted+unique later (std::set order)
    auto note_pair = [&](uint32_t a, uint32_t b) { decomposed.push_back(((uint64_t)std::min(a, b) << 32) | std::max(a, b)); };
    for (size_t i = 0; i < order.size(); i++) {
        const Atom &a = order[i];
        int w = a.weight();
        if (w == 3) {
            note_pair(a.d[0], a.d[1]);
            note_pair(a.d[0], a.d[2]);
            note_pair(a.d[1], a.d[2]);
        } else if (w == 2) {
            uint32_t x = a.d[0], y = a.d[1];
            set(x, y, cx.color_of(x), cx.color_of(y), charge_edge(x, y));
            set(x, y, NEUTRAL, NEUTRAL, 0);
            dump_to_boundary(x, y, obs[i]);
            dump_to_boundary(y, x, obs[i]);
            note_pair(x, y);
        } else if (w == 1) {
            uint32_t n = a.d[0];
            uint8_t c = cx.color_of(n);
            set(n, n, c, NEUTRAL, obs[i]);
            set(n, n, NEUTRAL, NEUTRAL, 0);
            if (rep_weight(n) == 3) {
                uint8_t c1 = next_color(c), c2 = next_color(c1);
                set(n, n, c1, c2, t.rgb_reps[n].obs_flip ^ obs[i]);
            }
        }
    }
    std::sort(decomposed.begin(), decomposed.end());
    decomposed.erase(std::unique(decomposed.begin(), decomposed.end()), decomposed.end());
    for (uint64_t key : decomposed) {
        uint32_t n1 = (uint32_t)(key >> 32), n2 = (uint32_t)key;
        for (uint8_t c = 1; c < 4; c++) {
            uint32_t r1 = t.rgb_reps[n1].node[c - 1], r2 = t.rgb_reps[n2].node[c - 1];
            uint64_t f;
            if (r1 != NONE && r2 != NONE && short_path(cx, r1, r2, f)) set(n1, n2, c, c, f);
        }
        set(n1, n2, NEUTRAL, NEUTRAL, 0);
    }
    // per-detector and per-half-edge indices
    t.self_drag.assign(t.num_dets, NO_INDEX);
    for (uint32_t d = 0; d < t.num_dets; d++) {
        auto it = index.find(((uint64_t)d << 32) | d);
        if (it != index.end()) t.self_drag[d] = it->second;
    }
    t.adj_drag.assign(t.adj.size(), NO_INDEX);
    for (uint32_t u = 0; u + 1 < t.row_ptr.size(); u++) {
        for (uint32_t k = t.row_ptr[u]; k < t.row_ptr[u + 1]; k++) {
            uint32_t a = u >> 1, b = t.adj[k].nbr >> 1;
            auto it = index.find(((uint64_t)std::min(a, b) << 32) | std::max(a, b));
            if (it != index.end()) t.adj_drag[k] = it->second;
        }
    }
}

// decode/pymatcher.cc:22-25 -> PyMatching v2.2.2 `detector_error_model_to_mwpm(dem, 1 << 24, true)`, restated from its
// published behaviour: w = ln((1-p)/p); parallel edges merged as independent mechanisms in DEM order;
// integer weight = 2 * round(w * (2^24 - 1) / max|w|); normalising constant = 2 * (2^24 - 1) / max|w|.
double merge_independent(double a, double b) {
    double sign = std::copysign(1.0, a) * std::copysign(1.0, b);
    double m = sign * std::min(std::fabs(a), std::fabs(b));
    return m + std::log1p(std::exp(-std::fabs(a + b))) - std::log1p(std::exp(-std::fabs(a - b)));
}

void pass_matching_graph(Ctx &cx) {
    auto &t = cx.t;
    uint32_t n = t.num_dets * 2;
    std::unordered_map<uint64_t, uint32_t, PairHash> slot;
    std::vector<uint64_t> keys;
    std::vector<double> weights;
    for (const auto &me : t.mobius_errors) {
        if (!(me.p > 0)) continue;
        if (me.p > 0.5) fail("A mobius dem error has probability above 0.5; negative edge weights are not supported.");
        double w = std::log((1 - me.p) / me.p);
        for (uint32_t k = 0; k < me.num_edges; k++) {
            uint32_t a = t.mobius_edges[2 * (me.first_edge + k)], b = t.mobius_edges[2 * (me.first_edge + k) + 1];
            if (a == b) continue;
            uint64_t key = ((uint64_t)std::min(a, b) << 32) | std::max(a, b);
            auto it = slot.find(key);
            if (it == slot.end()) {
                slot.emplace(key, (uint32_t)keys.size());
                keys.push_back(key);
                weights.push_back(w);
            } else {
                weights[it->second] = merge_independent(weights[it->second], w);
            }
        }
    }
    double max_abs = 0;
    bool integral = true;
    for (double w : weights) {
        max_abs = std::max(max_abs, std::fabs(w));
        if (std::round(w) != w) integral = false;
    }
    double scale = (weights.empty() || integral) ? 1.0 : (double)((1 << 24) - 1) / max_abs;
    t.normalising_constant = 2 * scale;
    std::vector<uint32_t> deg(n + 1, 0);
    for (uint64_t key : keys) {
        deg[(key >> 32) + 1]++;
        deg[(uint32_t)key + 1]++;
    }
    for (uint32_t i = 0; i < n; i++) deg[i + 1] += deg[i];
    t.row_ptr = deg;
    t.adj.assign(deg[n], AdjEntry{0, 0});
    std::vector<uint32_t> fill(deg.begin(), deg.end() - 1);
    for (size_t i = 0; i < keys.size(); i++) {
        uint32_t a = (uint32_t)(keys[i] >> 32), b = (uint32_t)keys[i];
        int32_t iw = (int32_t)(2 * (int64_t)std::llround(weights[i] * scale));
        t.adj[fill[a]++] = AdjEntry{b, iw};
        t.adj[fill[b]++] = AdjEntry{a, iw};
    }
    for (uint32_t u = 0; u < n; u++)
        std::sort(t.adj.begin() + t.row_ptr[u], t.adj.begin() + t.row_ptr[u + 1],
                  [](const AdjEntry &x, const AdjEntry &y) { return x.nbr < y.nbr; });
}

const char *charge_name(unsigned c) {
    static const char *names[] = {"NEUTRAL", "R", "G", "B"};
    return names[c & 3];
}

}  // namespace

// Neighbourhood table (tables.h): one truncated Dijkstra per node, popped in (distance, node id) order -- with positive
// weights every node at distance d is open before the first of them is popped, so the pop order is the sorted order
// and the first `stride` pops are exactly the `stride` closest nodes. Independent per node: spread over host threads.
void build_near_table(HostTables &t, uint32_t stride) {
    t.near.clear();
    t.near_stride = 0;
    const uint32_t n = t.num_dets * 2;
    if (stride == 0 || n == 0 || t.adj.empty()) return;
    stride = std::min<uint32_t>((stride + 31) / 32 * 32, NEAR_MAX_STRIDE);
    t.near.assign((size_t)stride * n, NearEntry{NEAR_PAD, NEAR_PAD});
    t.near_stride = stride;
    unsigned nthreads = std::thread::hardware_concurrency();
    if (nthreads == 0) nthreads = 1;
    nthreads = std::min<unsigned>(nthreads, 32);
    if (n < 4096) nthreads = 1;
    const uint32_t *row_ptr = t.row_ptr.data();
    const AdjEntry *adj = t.adj.data();
    NearEntry *near = t.near.data();
    auto worker = [=](unsigned tid) {
        constexpr uint16_t OPEN = 0xFFFF, IS_SRC = 0xFFFE;
        std::vector<int64_t> dist(n, -1);
        std::vector<uint16_t> pos(n, OPEN);
        std::vector<uint32_t> touched;
        std::vector<std::pair<int64_t, uint32_t>> heap;
        auto cmp = [](const std::pair<int64_t, uint32_t> &a, const std::pair<int64_t, uint32_t> &b) { return a > b; };
        for (uint32_t u = tid; u < n; u += nthreads) {
            touched.clear();
            heap.clear();
            dist[u] = 0;
            pos[u] = IS_SRC;
            touched.push_back(u);
            auto relax = [&](uint32_t v, int64_t dv) {
                for (uint32_t e = row_ptr[v]; e < row_ptr[v + 1]; e++) {
                    const uint32_t w = adj[e].nbr;
                    const int64_t nd = dv + adj[e].wt;
                    if (dist[w] < 0) touched.push_back(w);
                    else if (nd >= dist[w]) continue;
                    dist[w] = nd;
                    heap.emplace_back(nd, w);
                    std::push_heap(heap.begin(), heap.end(), cmp);
                }
            };
            relax(u, 0);
            uint32_t count = 0;
            while (!heap.empty() && count < stride) {
                std::pop_heap(heap.begin(), heap.end(), cmp);
                const int64_t d = heap.back().first;
                const uint32_t v = heap.back().second;
                heap.pop_back();
                if (pos[v] != OPEN || d != dist[v]) continue;
                if (d >= 0x7FFFFFFF) break;
                uint32_t pred = NO_INDEX;
                bool direct = false;
                for (uint32_t e = row_ptr[v]; e < row_ptr[v + 1]; e++) {
                    const uint32_t w = adj[e].nbr;
                    if (pos[w] == OPEN || dist[w] + adj[e].wt != d) continue;
                    if (w == u) direct = true;
                    if (pred == NO_INDEX) pred = pos[w] == IS_SRC ? NEAR_PRED_SRC : std::min<uint32_t>(pos[w], NEAR_PRED_FAR);
                }
                NearEntry &out = near[((size_t)(count >> 5) * n + u) * 32 + (count & 31)];
                out.node_pred = v | (pred << 24);
                out.dist_flag = (uint32_t)d | (direct ? 0x80000000u : 0u);
                pos[v] = (uint16_t)count;
                count++;
                if (count < stride) relax(v, d);
            }
            for (uint32_t w : touched) {
                dist[w] = -1;
                pos[w] = OPEN;
            }
        }
    };
    if (nthreads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> pool;
        for (unsigned k = 0; k < nthreads; k++) pool.emplace_back(worker, k);
        for (auto &th : pool) th.join();
    }
}

std::vector<uint32_t> reference_container_visit_order(const std::vector<uint32_t> &keys_in_insertion_order) {
    StdHashOrder order;
    for (uint32_t k : keys_in_insertion_order) order.insert_new(k);
    std::vector<uint32_t> out;
    order.for_each([&](uint32_t k) { out.push_back(k); });
    return out;
}

HostTables compile_tables(const FlatDem &dem, uint32_t near_stride) {
    HostTables t;
    if (dem.num_detectors > 0x7FFFFFFEull) fail("The detector error model is too large.");
    if (dem.num_detectors * 2 > (1ull << 24))
        fail("The detector error model is too large: the B200 matcher supports at most 2^23 detectors.");
    t.num_dets = (uint32_t)dem.num_detectors;
    t.num_obs = (uint32_t)dem.num_observables;
    Ctx cx(dem, t);
    pass_collect_nodes(cx);

    // graph/collect_atomic_errors.cc:102-115
    {
        std::vector<uint32_t> dets;
        uint64_t obs;
        for (const auto &e : dem.errors) {
            symptoms_of(cx, e, dets, obs);
            register_if_atomic(cx, dets, obs, cx.atoms);
        }
    }
    pass_mobius_dem(cx);

    // Freeze the atomic errors in std::map key order (the order every later pass iterates in).
    std::vector<Atom> order;
    order.reserve(cx.atoms.size());
    for (const auto &kv : cx.atoms) order.push_back(kv.first);
    std::sort(order.begin(), order.end());
    std::vector<uint64_t> obs(order.size());
    for (size_t i = 0; i < order.size(); i++) {
        obs[i] = cx.atoms.at(order[i]);
        t.atom_dets.insert(t.atom_dets.end(), order[i].d, order[i].d + 3);
    }
    t.atom_obs = obs;

    pass_rgb_reps(cx, order, obs);
    pass_charge_graph(cx, order, obs);
    pass_matching_graph(cx);
    pass_drag(cx, order, obs);

    // Fixed-stride adjacency rows: one aligned row per node, so a warp fetches a row with a single request and the
    // row_ptr indirection disappears from the ball-growth loop.
    {
        uint32_t max_deg = 0;
        for (uint32_t u = 0; u + 1 < t.row_ptr.size(); u++) max_deg = std::max(max_deg, t.row_ptr[u + 1] - t.row_ptr[u]);
        t.ell_shift = 0;
        if (max_deg >= 1 && max_deg <= 32) {
            uint32_t sh = 1;
            while ((1u << sh) < max_deg) sh++;
            t.ell_shift = sh;
            t.ell_adj.assign((size_t)(t.num_dets * 2) << sh, AdjEntry{NO_INDEX, 0});
            for (uint32_t u = 0; u + 1 < t.row_ptr.size(); u++)
                for (uint32_t k = t.row_ptr[u]; k < t.row_ptr[u + 1]; k++)
                    t.ell_adj[((size_t)u << sh) + (k - t.row_ptr[u])] = t.adj[k];
        }
    }
    {
        int64_t mw = 0;
        for (const auto &a : t.adj) {
            if (a.wt <= 0) fail("The matcher graph has an edge of non-positive weight; such detector error models are not supported.");
            if (mw == 0 || a.wt < mw) mw = a.wt;
        }
        t.min_wt = (int32_t)mw;
    }
    t.det_info.assign(t.num_dets, 0);
    t.rep_obs.assign(t.num_dets, 0);
    t.ingest_mask.assign((t.num_dets + 7) / 8, 0);
    for (uint32_t d = 0; d < t.num_dets; d++) {
        const RgbRep &r = t.rgb_reps[d];
        bool w3 = r.node[0] != NO_INDEX && r.node[1] != NO_INDEX && r.node[2] != NO_INDEX;
        t.det_info[d] = (uint8_t)(t.color[d] | (w3 ? DET_REP_WEIGHT3 : 0) | (t.basis[d] << DET_BASIS_SHIFT) |
                                  (t.ignored[d] ? DET_IGNORED : 0));
        t.rep_obs[d] = r.obs_flip;
        if (!t.ignored[d]) t.ingest_mask[d >> 3] |= (uint8_t)(1u << (d & 7));
    }
    if (near_stride == NEAR_STRIDE_AUTO) {
        // as many entries per node as a 256 MB table holds (HBM is plentiful; only the first chunks of the rows are hot)
        const uint64_t budget = 256ull << 20;
        const uint64_t nodes = std::max<uint64_t>(uint64_t{t.num_dets} * 2, 1);
        near_stride = (uint32_t)std::min<uint64_t>(budget / (nodes * sizeof(NearEntry)) / 32 * 32, NEAR_MAX_STRIDE);
        if (const char *env = std::getenv("CB200_NEAR_STRIDE")) near_stride = (uint32_t)std::strtoul(env, nullptr, 10);
    }
    build_near_table(t, near_stride);
    return t;
}

// =================================================================================================
// On-disk cache of compiled tables.
// File: magic, version, key, then every field in declaration order; vectors as (u64 count, raw elements).
// =================================================================================================
namespace {
constexpr uint64_t CACHE_MAGIC = 0x3030324253454c42ull;  // "BLESB200"
constexpr uint64_t CACHE_VERSION = 5;  // 4: drag table follows the reference container's visit order; 5: model text stored

struct Writer {
    FILE *f;
    bool ok = true;
    void raw(const void *p, size_t n) { ok = ok && (n == 0 || fwrite(p, 1, n, f) == n); }
    template <class T> void pod(const T &v) { raw(&v, sizeof(T)); }
    template <class T> void vec(const std::vector<T> &v) {
        pod<uint64_t>(v.size());
        raw(v.data(), v.size() * sizeof(T));
    }
};
struct Reader {
    FILE *f;
    bool ok = true;
    void raw(void *p, size_t n) { ok = ok && (n == 0 || fread(p, 1, n, f) == n); }
    template <class T> void pod(T &v) { raw(&v, sizeof(T)); }
    template <class T> void vec(std::vector<T> &v) {
        uint64_t n = 0;
        pod(n);
        if (!ok || n > (uint64_t{1} << 34) / sizeof(T)) {
            ok = false;
            return;
        }
        v.resize((size_t)n);
        raw(v.data(), (size_t)n * sizeof(T));
    }
};

template <class IO, class T>
void tables_io(IO &io, T &t) {
    io.pod(t.num_dets);
    io.pod(t.num_obs);
    io.pod(t.normalising_constant);
    io.vec(t.color);
    io.vec(t.basis);
    io.vec(t.ignored);
    io.vec(t.atom_dets);
    io.vec(t.atom_obs);
    io.vec(t.mobius_errors);
    io.vec(t.mobius_edges);
    io.vec(t.rgb_reps);
    io.pod(t.drag_ambiguities);
    io.vec(t.row_ptr);
    io.vec(t.adj);
    io.vec(t.ell_adj);
    io.pod(t.ell_shift);
    io.pod(t.min_wt);
    io.vec(t.adj_drag);
    io.vec(t.det_info);
    io.vec(t.rep_obs);
    io.vec(t.self_drag);
    io.vec(t.ingest_mask);
    io.vec(t.drag);
    io.vec(t.drag_pair);
    io.vec(t.near);
    io.pod(t.near_stride);
}

uint64_t fnv1a(const void *p, size_t n, uint64_t h = 0xcbf29ce484222325ull) {
    const uint8_t *b = (const uint8_t *)p;
    for (size_t i = 0; i < n; i++) {
        h ^= b[i];
        h *= 0x100000001b3ull;
    }
    return h;
}
}  // namespace

void save_tables(const HostTables &t, const std::string &path, uint64_t key, const std::string &dem_text) {
    // written next to the final name and renamed, so a concurrent reader never sees half a file; the temporary name is
    // unique per process and call
    static std::atomic<unsigned long long> counter{0};
    std::string tmp = path + ".tmp." + std::to_string((unsigned long long)getpid()) + "." + std::to_string(counter.fetch_add(1));
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    Writer w{f};
    w.pod(CACHE_MAGIC);
    w.pod(CACHE_VERSION);
    w.pod(key);
    // the model itself: a file is only ever used for the exact text it was compiled from (the 64-bit key names the file,
    // it does not identify the model)
    std::vector<char> text(dem_text.begin(), dem_text.end());
    w.vec(text);
    tables_io(w, t);
    // the charge graph is nested: flatten it
    w.pod<uint64_t>(t.charge_graph.size());
    for (const auto &row : t.charge_graph) w.vec(row);
    w.pod(CACHE_MAGIC);
    bool ok = w.ok;
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
}

bool load_tables(HostTables &t, const std::string &path, uint64_t key, const std::string &dem_text) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    Reader r{f};
    uint64_t magic = 0, version = 0, k = 0, tail = 0, rows = 0;
    r.pod(magic);
    r.pod(version);
    r.pod(k);
    bool ok = r.ok && magic == CACHE_MAGIC && version == CACHE_VERSION && k == key;
    if (ok) {
        std::vector<char> text;
        r.vec(text);
        ok = r.ok && text.size() == dem_text.size() && memcmp(text.data(), dem_text.data(), text.size()) == 0;
    }
    if (ok) {
        tables_io(r, t);
        r.pod(rows);
        ok = r.ok && rows <= (uint64_t{1} << 32);
        if (ok) {
            t.charge_graph.resize((size_t)rows);
            for (auto &row : t.charge_graph) r.vec(row);
        }
        r.pod(tail);
        ok = ok && r.ok && tail == CACHE_MAGIC && t.row_ptr.size() == (size_t)t.num_dets * 2 + 1 &&
             t.near.size() == (size_t)t.near_stride * t.num_dets * 2;
    }
    fclose(f);
    if (!ok) t = HostTables{};
    return ok;
}

HostTables compile_tables_cached(const std::string &dem_text, bool *from_cache) {
    if (from_cache) *from_cache = false;
    const char *dir = std::getenv("CHROMOBIUS_B200_CACHE");
    if (!dir || !*dir) return compile_tables(FlatDem::parse(dem_text));
    const char *stride_env = std::getenv("CB200_NEAR_STRIDE");
    uint64_t key = fnv1a(dem_text.data(), dem_text.size());
    key = fnv1a(&CACHE_VERSION, sizeof(CACHE_VERSION), key);
    if (stride_env) key = fnv1a(stride_env, strlen(stride_env), key);
    char name[64];
    snprintf(name, sizeof(name), "/cb200_%016llx.tables", (unsigned long long)key);
    const std::string path = std::string(dir) + name;
    HostTables t;
    if (load_tables(t, path, key, dem_text)) {
        if (from_cache) *from_cache = true;
        return t;
    }
    t = compile_tables(FlatDem::parse(dem_text));
    save_tables(t, path, key, dem_text);
    return t;
}

TablesView HostTables::view() const {
    TablesView v{};
    v.num_dets = num_dets;
    v.num_nodes = num_dets * 2;
    v.num_obs = num_obs;
    v.num_det_bytes = (num_dets + 7) / 8;
    v.row_ptr = row_ptr.data();
    v.adj = adj.data();
    v.ell_adj = ell_adj.data();
    v.ell_shift = ell_shift;
    v.min_wt = min_wt;
    v.adj_drag = adj_drag.data();
    v.det_info = det_info.data();
    v.rep_obs = rep_obs.data();
    v.self_drag = self_drag.data();
    v.ingest_mask = ingest_mask.data();
    v.drag = drag.data();
    v.near = near_stride ? near.data() : nullptr;
    v.near_stride = near_stride;
    return v;
}

std::string HostTables::mobius_dem_text() const {
    std::string out;
    char buf[64];
    for (const auto &me : mobius_errors) {
        snprintf(buf, sizeof(buf), "error(%.17g)", me.p);
        out += buf;
        for (uint32_t k = 0; k < me.num_edges; k++) {
            snprintf(buf, sizeof(buf), "%s D%u D%u", k ? " ^" : "", mobius_edges[2 * (me.first_edge + k)],
                     mobius_edges[2 * (me.first_edge + k) + 1]);
            out += buf;
        }
        out += "\n";
    }
    // decoder.cc:48-52: padded with a detector declaration when the errors do not mention the last Moebius node.
    uint32_t mentioned = 0;
    for (uint32_t v : mobius_edges) mentioned = std::max(mentioned, v + 1);
    if (mentioned < num_dets * 2) {
        snprintf(buf, sizeof(buf), "detector D%u\n", num_dets * 2 - 1);
        out += buf;
    }
    return out;
}

std::string HostTables::describe() const {
    std::stringstream out;
    auto node = [&](uint32_t n) {
        if (n == NO_INDEX) out << "BOUNDARY_NODE";
        else out << n;
    };
    out << "chromobius::Decoder{\n\n.charge_graph=ChargeGraph{.nodes={\n";
    for (size_t k = 0; k < charge_graph.size(); k++) {
        out << "    ChargeGraphNode{.neighbors={";
        for (size_t i = 0; i < charge_graph[k].size(); i++) {
            out << (i ? ", " : "") << "{";
            node(charge_graph[k][i].first);
            out << "," << charge_graph[k][i].second << "}";
        }
        out << "}}, // node " << k << "\n";
    }
    out << "}}\n\n.rgb_reps={";
    for (size_t k = 0; k < rgb_reps.size(); k++) {
        out << "\n    RgbEdge{.red_node=";
        node(rgb_reps[k].node[0]);
        out << ", .green_node=";
        node(rgb_reps[k].node[1]);
        out << ", .blue_node=";
        node(rgb_reps[k].node[2]);
        out << ", .obs_flip=" << rgb_reps[k].obs_flip << ", .charge_flip=" << charge_name(rgb_reps[k].charge_flip)
            << "} // rep " << k;
    }
    out << "\n}\n\n.drag_graph=DragGraph{.mmm={\n";
    // Re-expand the pair entries into the reference's (n1, n2, c1, c2) key order.
    struct Row {
        uint32_t n1, n2, c1, c2;
        uint64_t v;
    };
    std::vector<Row> rows;
    for (size_t i = 0; i < drag.size(); i++) {
        uint32_t a = drag_pair[2 * i], b = drag_pair[2 * i + 1];
        for (uint32_t c1 = 0; c1 < 4; c1++)
            for (uint32_t c2 = 0; c2 < 4; c2++)
                if (drag[i].valid >> (4 * c1 + c2) & 1) {
                    rows.push_back({a, b, c1, c2, drag[i].obs[4 * c1 + c2]});
                    if (a != b) rows.push_back({b, a, c2, c1, drag[i].obs[4 * c1 + c2]});
                }
    }
    std::sort(rows.begin(), rows.end(), [](const Row &x, const Row &y) {
        if (x.n1 != y.n1) return x.n1 < y.n1;
        if (x.n2 != y.n2) return x.n2 < y.n2;
        if (x.c1 != y.c1) return x.c1 < y.c1;
        return x.c2 < y.c2;
    });
    for (const auto &r : rows)
        out << "    " << charge_name(r.c1) << "@" << r.n1 << ":" << charge_name(r.c2) << "@" << r.n2 << " = " << r.v << "\n";
    out << "}}\n\n.mobius_dem=\n" << mobius_dem_text() << "\n}\n";
    return out.str();
}

}  // namespace cb200
