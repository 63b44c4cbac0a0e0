// Side-by-side check of the two dense blossom solvers of chromobius_b200/csrc/decode_core.h on random components:
// the sequential one (Blossom::solve, a group of one lane) against the warp-synchronous one (Blossom::solve_w) run by 32
// host threads that stand in for the lanes of a warp (wsync / ballot / min64 are barriers). Matching, labels and pair
// count must be identical. TEST INFRASTRUCTURE ONLY; built and run by tests/test_warp_blossom_cpu.py.
#include <pthread.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../chromobius_b200/csrc/decode_core.h"

using namespace cb200;

struct WarpCtx {
    pthread_barrier_t bar;
    uint64_t slot[32];
};
struct WarpEmu {
    int l;
    WarpCtx *c;
    int lane() const { return l; }
    int lanes() const { return 32; }
    void wsync() const { pthread_barrier_wait(&c->bar); }
    uint32_t ballot(bool p) const {
        c->slot[l] = p ? 1 : 0;
        pthread_barrier_wait(&c->bar);
        uint32_t r = 0;
        for (int k = 0; k < 32; k++) r |= (uint32_t)c->slot[k] << k;
        pthread_barrier_wait(&c->bar);
        return r;
    }
    uint64_t min64(uint64_t v) const {
        c->slot[l] = v;
        pthread_barrier_wait(&c->bar);
        uint64_t r = ~uint64_t{0};
        for (int k = 0; k < 32; k++) r = c->slot[k] < r ? c->slot[k] : r;
        pthread_barrier_wait(&c->bar);
        return r;
    }
};

template <int CAP>
struct Job {
    int n;
    std::vector<int64_t> w;  // (n+1) x (n+1)
    uint8_t *mem;
    WarpCtx *ctx;
    int lane;
    int pairs;
};

template <int CAP>
static void fill(Blossom<WarpEmu, CAP> &bl, const Job<CAP> &j) {
    for (int u = 0; u <= j.n; u++)
        for (int v = 0; v <= j.n; v++) bl.W()[u * bl.ws() + v] = j.w[u * (j.n + 1) + v];
}

template <int CAP>
static void *lane_main(void *arg) {
    auto *j = (Job<CAP> *)arg;
    WarpEmu g{j->lane, j->ctx};
    typedef Blossom<WarpEmu, CAP> BL;
    BL &bl = *BL::place(j->mem);
    if (j->lane == 0) {
        bl.bind(j->mem, j->n, 32);
        memset(bl.W(), 0, sizeof(int64_t) * bl.ws() * (j->n + 1));
        fill(bl, *j);
    }
    g.wsync();
    bl.prepare(g);
    j->pairs = bl.solve_w(g);
    return nullptr;
}

static int g_blossom_cases = 0, g_imperfect = 0;
template <int CAP>
static int run_case(int n, const std::vector<int64_t> &w, int id) {
    // sequential
    typedef Blossom<HostGroup, CAP> BS;
    std::vector<uint8_t> m1(BS::bytes_needed(n) + 64, 0), m2(Blossom<WarpEmu, CAP>::bytes_needed(n) + 64, 0);
    HostGroup hg;
    BS &bs = *BS::place(m1.data());
    bs.bind(m1.data(), n, 1);
    for (int u = 0; u <= n; u++)
        for (int v = 0; v <= n; v++) bs.W()[u * bs.ws() + v] = w[u * (n + 1) + v];
    bs.prepare(hg);
    int p1 = bs.solve(hg);
    if (bs.n_x > n) g_blossom_cases++;
    if (p1 * 2 != n) g_imperfect++;
    // warp
    WarpCtx ctx;
    pthread_barrier_init(&ctx.bar, nullptr, 32);
    std::vector<Job<CAP>> jobs(32);
    pthread_t th[32];
    for (int l = 0; l < 32; l++) {
        jobs[l] = Job<CAP>{n, w, m2.data(), &ctx, l, -1};
        pthread_create(&th[l], nullptr, lane_main<CAP>, &jobs[l]);
    }
    for (int l = 0; l < 32; l++) pthread_join(th[l], nullptr);
    pthread_barrier_destroy(&ctx.bar);
    auto &bw = *Blossom<WarpEmu, CAP>::place(m2.data());
    int bad = 0;
    for (int l = 0; l < 32; l++)
        if (jobs[l].pairs != p1) bad++;
    if (bs.broken != bw.broken) bad++;
    for (int u = 1; u <= n; u++) {
        if (bs.match()[u] != bw.match()[u]) bad++;
        if (bs.lab()[u] != bw.lab()[u]) bad++;
    }
    if (bad) {
        printf("case %d (CAP %d, n %d): MISMATCH pairs seq %d warp %d broken %d/%d\n", id, CAP, n, p1, jobs[0].pairs, (int)bs.broken, (int)bw.broken);
        for (int u = 1; u <= n; u++)
            printf("  u %d: match %d / %d   lab %lld / %lld\n", u, bs.match()[u], bw.match()[u], (long long)bs.lab()[u], (long long)bw.lab()[u]);
    }
    return bad ? 1 : 0;
}

int main(int argc, char **argv) {
    int cases = argc > 1 ? atoi(argv[1]) : 300;
    uint64_t seed = argc > 2 ? strtoull(argv[2], nullptr, 10) : 1;
    const int big_n = argc > 3 ? atoi(argv[3]) : 90;
    std::mt19937_64 rng(seed);
    int failures = 0;
    const int64_t BIG = int64_t{1} << 52;
    for (int id = 0; id < cases; id++) {
        // fixed layouts of 16 / 32 vertices, and the run-time layout (CAP 0) with up to `big_n` vertices
        const int cap = (id % 3 == 2) ? 0 : ((id & 1) ? 32 : 16);
        int n = cap ? 2 + (int)(rng() % (cap - 1)) : 33 + (int)(rng() % (big_n - 32));
        int style = (int)(rng() % 4);
        // points on a line / small integer distances produce many ties and blossoms; odd n and sparse graphs included
        std::vector<int64_t> w((n + 1) * (n + 1), 0);
        std::vector<int64_t> pos(n + 1);
        for (int u = 1; u <= n; u++) pos[u] = (int64_t)(rng() % (style == 0 ? 8 : 60));
        double density = style == 3 ? 0.35 : (style == 2 ? 0.6 : 1.0);
        for (int u = 1; u <= n; u++)
            for (int v = u + 1; v <= n; v++) {
                if ((double)(rng() % 1000) / 1000.0 >= density) continue;
                int64_t d = style == 1 ? 2 * (int64_t)(1 + rng() % 5) : 2 * (1 + llabs(pos[u] - pos[v]) + (int64_t)(rng() % 3));
                w[u * (n + 1) + v] = w[v * (n + 1) + u] = BIG - d;
            }
        failures += cap == 16 ? run_case<16>(n, w, id) : (cap == 32 ? run_case<32>(n, w, id) : run_case<0>(n, w, id));
    }
    printf("%d cases (%d formed blossoms, %d without a perfect matching), %d failures\n", cases, g_blossom_cases, g_imperfect, failures);
    return failures ? 1 : 0;
}
