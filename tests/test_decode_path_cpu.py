"""CPU tier: the device decode logic (host simulation build, group of one thread) against the oracle, the oracle's
matcher against brute force, and the RNG-free single-error property of the reference's integration test."""
import itertools

import numpy as np
import pytest

import oracle
from dem_util import read_dem, NumpyDemSampler
from sim_util import SimDecoder

SMALL = [
    ('phenom_color_code_d5_r5_p1000', 6000),
    ('phenom_color_code_d5_r5_p1000_with_ignored', 3000),
    ('midout_color_code_d5_r10_p1000', 1500),
    ('superdense_color_code_d5_r20_p1000', 400),
    ('color2surface_d7_phenom_r7_p100', 300),
    ('color2surface_d5_transit_p100', 1500),
    ('surface_code_d5_r5_p1000', 1500),
    ('rep_code_d9_transit_p10', 1500),
    ('rep_code_rg_d9_transit_p10', 1500),
    ('rep_code_rbrrr_d9_transit_p10', 1500),
]


@pytest.mark.parametrize('name,shots', SMALL)
def test_device_logic_matches_oracle(built, name, shots):
    text = read_dem(name)
    dets, _ = NumpyDemSampler(text).sample(shots, seed=11)
    sim = SimDecoder(text)
    obs, w, st, rd = sim.decode(dets)
    o_obs, o_w = oracle.OracleDecoder(text).decode_batch(dets)
    assert (st == 0).all()
    assert (w == o_w).all(), 'integer matching weights must be bit exact'
    assert int((obs != o_obs).sum()) <= max(1, shots // 300)


@pytest.mark.parametrize('stride', [0, 32, 64, 512])
@pytest.mark.parametrize('name,shots', [('midout_color_code_d5_r10_p1000', 600), ('color2surface_d7_phenom_r7_p100', 200),
                                        ('rep_code_rg_d9_transit_p10', 400)])
def test_near_table_sizes_give_identical_results(built, monkeypatch, name, shots, stride):
    """Short near-table rows (and none at all) send ball queries and paths to the Dijkstra search: same answers."""
    text = read_dem(name)
    dets, _ = NumpyDemSampler(text).sample(shots, seed=4)
    full = SimDecoder(text).decode(dets, want_edges=False)
    monkeypatch.setenv('CB200_NEAR_STRIDE', str(stride))
    short = SimDecoder(text).decode(dets)
    monkeypatch.delenv('CB200_NEAR_STRIDE')
    assert (short[2] == 0).all() and (full[2] == 0).all()
    assert (short[0] == full[0]).all() and (short[1] == full[1]).all()
    o_obs, o_w = oracle.OracleDecoder(text).decode_batch(dets)
    assert (short[1] == o_w).all()
    one = SimDecoder(text)
    for row in dets[:40]:
        a = one.decode(row, want_edges=True)
        monkeypatch.setenv('CB200_NEAR_STRIDE', str(stride))
        b = SimDecoder(text).decode(row, want_edges=True)
        monkeypatch.delenv('CB200_NEAR_STRIDE')
        assert (a[4] == b[4]).all(), 'matched edge lists must not depend on the near-table size'


@pytest.mark.parametrize('pool_kb', [0, 3, 8])
def test_blossom_layout_classes_agree(built, monkeypatch, pool_kb):
    """Fixed-layout (8 / 16 / 32 vertices, shared pool) and run-time sized blossoms give the same matching and duals."""
    text = read_dem('color2surface_d7_phenom_r7_p100')   # p = 1 %: dense syndromes, blossoms of every size class
    dets, _ = NumpyDemSampler(text).sample(250, seed=9)
    full = SimDecoder(text).decode(dets)
    monkeypatch.setenv('CBSIM_POOL_KB', str(pool_kb))
    other = SimDecoder(text).decode(dets)
    assert (full[2] == 0).all() and (other[2] == 0).all()
    assert (other[0] == full[0]).all() and (other[1] == full[1]).all() and (other[3] == full[3]).all()


def test_table_cache_round_trip(built, monkeypatch, tmp_path):
    """CHROMOBIUS_B200_CACHE: tables loaded from disk equal freshly compiled ones; damaged files are ignored and rewritten."""
    import os
    text = read_dem('midout_color_code_d5_r10_p1000')
    dets, _ = NumpyDemSampler(text).sample(300, seed=6)
    fresh = SimDecoder(text)
    want = fresh.decode(dets)
    monkeypatch.setenv('CHROMOBIUS_B200_CACHE', str(tmp_path))
    first = SimDecoder(text)       # compiles and writes
    files = os.listdir(tmp_path)
    assert len(files) == 1 and files[0].startswith('cb200_') and files[0].endswith('.tables')
    second = SimDecoder(text)      # loads
    for d in (first, second):
        got = d.decode(dets)
        assert all((a == b).all() for a, b in zip(got, want))
        assert d.describe() == fresh.describe() and d.mobius_dem() == fresh.mobius_dem()
        assert all((a == b).all() for a, b in zip(d.graph(), fresh.graph()))
    path = os.path.join(tmp_path, files[0])
    size = os.path.getsize(path)
    with open(path, 'r+b') as f:   # truncate: the loader must notice and recompile
        f.truncate(size // 2)
    third = SimDecoder(text)
    assert all((a == b).all() for a, b in zip(third.decode(dets), want))
    assert os.path.getsize(path) == size
    other = SimDecoder(read_dem('surface_code_d5_r5_p1000'))   # a different DEM gets its own file
    assert len(os.listdir(tmp_path)) == 2 and other.num_detectors != first.num_detectors


def test_device_logic_survives_tiny_scratch(built):
    """Scratch overflow is reported per shot (status 1), never silently wrong."""
    text = read_dem('midout_color_code_d5_r10_p1000')
    dets, _ = NumpyDemSampler(text).sample(300, seed=2)
    sim = SimDecoder(text)
    good = sim.decode(dets)
    obs, w, st, rd = sim.decode(dets, ht_log2=4, dyn_bytes=4096)
    assert set(np.unique(st)) <= {0, 1}
    ok = st == 0
    assert (obs[ok] == good[0][ok]).all() and (w[ok] == good[1][ok]).all()
    assert (st == 1).any()


def all_pairs_dist(row_ptr, nbr, wt, sources):
    import heapq
    out = {}
    for s in sources:
        dist = {s: 0}
        pq = [(0, s)]
        while pq:
            d, u = heapq.heappop(pq)
            if d != dist[u]:
                continue
            for k in range(row_ptr[u], row_ptr[u + 1]):
                v = int(nbr[k])
                nd = d + int(wt[k])
                if v not in dist or nd < dist[v]:
                    dist[v] = nd
                    heapq.heappush(pq, (nd, v))
        out[s] = dist
    return out


def brute_force_mwpm(events, dist):
    best = [None]

    def rec(rest, acc):
        if best[0] is not None and acc >= best[0]:
            return
        if not rest:
            best[0] = acc
            return
        a = rest[0]
        for k in range(1, len(rest)):
            b = rest[k]
            d = dist[a].get(b)
            if d is None:
                continue
            rec(rest[1:k] + rest[k + 1:], acc + d)

    rec(list(events), 0)
    return best[0]


@pytest.mark.parametrize('name', ['midout_color_code_d5_r10_p1000', 'color2surface_d5_transit_p100', 'surface_code_d5_r5_p1000'])
def test_matching_weight_against_brute_force(built, name):
    """Exhaustive pairing over exact all-pairs distances pins the integer weight for shots with <= 10 matcher events."""
    text = read_dem(name)
    od = oracle.OracleDecoder(text)
    sim = SimDecoder(text)
    row_ptr, nbr, wt = od.graph()
    r2, n2, w2 = sim.graph()
    assert (row_ptr == r2).all() and (nbr == n2).all() and (wt == w2).all(), 'oracle and product build the same matcher graph'
    smp = NumpyDemSampler(text)
    dets, _ = smp.sample(3000, seed=7)
    nc = od.node_colors()
    checked = 0
    for row in dets:
        bits = np.unpackbits(row, bitorder='little')[:od.num_detectors]
        ds = [d for d in np.nonzero(bits)[0] if not nc[d, 2]]
        if not (2 <= len(ds) <= 5):
            continue
        events = [2 * int(d) + s for d in ds for s in (0, 1)]
        dist = all_pairs_dist(row_ptr, nbr, wt, events)
        want = brute_force_mwpm(events, dist)
        got_o = od.decode_batch(row[None, :])[1][0]
        got_s = sim.decode(row[None, :])[1][0]
        assert want == got_o == got_s
        checked += 1
        if checked >= 120:
            break
    assert checked >= 20


@pytest.mark.parametrize('name', [n for n, _ in SMALL] + ['midout_color_code_d9_r36_p1000'])
def test_single_error_resilience(built, name):
    """decode/decoder_integration.test.cc:44-120: forcing exactly DEM error k (k < num_detectors, as in the reference's
    loop bound) must be decoded to that error's observable mask. RNG-free known-answer test."""
    text = read_dem(name)
    smp = NumpyDemSampler(text)
    n = min(smp.num_dets, len(smp.p))
    if name.startswith('midout_color_code_d9'):
        ks = list(range(0, n, 7))
    else:
        ks = list(range(n))
    rows = np.zeros((len(ks), smp.nbytes), np.uint8)
    want = np.zeros(len(ks), np.uint64)
    for i, k in enumerate(ks):
        rows[i], want[i] = smp.single_error_shot(k)
    sim = SimDecoder(text)
    obs, w, st, rd = sim.decode(rows)
    assert (st == 0).all()
    assert (obs == want).all()
    o_obs, _ = oracle.OracleDecoder(text).decode_batch(rows)
    assert (o_obs == want).all()


def test_lift_is_pinned_on_oracle_edges(built):
    """Stages 3-4 of the device logic, fed the device's own matched edges, agree with the oracle's Euler tour + discharge."""
    text = read_dem('superdense_color_code_d5_r20_p1000')
    od = oracle.OracleDecoder(text)
    sim = SimDecoder(text)
    dets, _ = NumpyDemSampler(text).sample(120, seed=4)
    for row in dets:
        obs, w, st, rd, edges = sim.decode(row, want_edges=True)
        if len(edges):
            assert od.lift_edges(row, edges) == int(obs[0])


def test_wide_blossom_layout(built, tmp_path):
    """Components beyond 250 events switch the blossom's edge matrix to u16 endpoints. Force that layout for every
    component above 5 events (test build with -DCB_NARROW_COMPONENT=5) and require identical results."""
    import ctypes
    import os
    import subprocess
    import sim_util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / 'libcb200_sim_wide.so')
    subprocess.check_call(['g++', '-O1', '-std=c++17', '-fPIC', '-Wno-sign-compare', '-DCB_NARROW_COMPONENT=5', '-shared', '-o', so,
                           os.path.join(root, 'tests', 'sim', 'host_sim.cc'),
                           os.path.join(root, 'chromobius_b200', 'csrc', 'host_compile.cc')])
    text = read_dem('superdense_color_code_d5_r20_p1000')
    dets, _ = NumpyDemSampler(text).sample(300, seed=8)
    want = SimDecoder(text).decode(dets)
    saved = sim_util._lib
    try:
        sim_util._lib = None
        real_build = built.build_sim
        built.build_sim = lambda force=False: so
        got = SimDecoder(text).decode(dets)
    finally:
        built.build_sim = real_build
        sim_util._lib = saved
    assert (got[2] == 0).all()
    assert (got[0] == want[0]).all() and (got[1] == want[1]).all()


def test_plain_cycle_lift_equals_general_lift_in_the_host_simulation(built, monkeypatch):
    """lift_shot_plain (cycles straight from the pair list when the shot has no junction) against the general half-edge
    machinery of lift_shot, on the same device code compiled for the host."""
    from sim_util import SimDecoder
    from reference_pins import load_circuit_shots
    for name, n in [('superdense_color_code_d5_r20_p1000', 1500), ('color2surface_d7_phenom_r7_p100', 800),
                    ('toric_superdense_color_code_epr_d12_r5_p1000', 800), ('midout488_color_code_d9_r33_p1000', 400)]:
        dets = load_circuit_shots(name)[0][:n]
        sim = SimDecoder(read_dem(name))
        obs_a, w_a, st_a, _ = sim.decode(dets)
        monkeypatch.setenv('CB200_NO_PLAIN_LIFT', '1')
        obs_b, w_b, st_b, _ = sim.decode(dets)
        monkeypatch.delenv('CB200_NO_PLAIN_LIFT')
        assert st_a.max() == 0 and st_b.max() == 0
        assert (obs_a == obs_b).all() and (w_a == w_b).all()


def chain_dem(n, p=0.05):
    """An RG repetition-code chain of n detectors (the shape of test_data/rep_code_rg_d9_transit_p10.stim's model):
    with every detector fired, each Moebius subgraph becomes ONE matching component of n events."""
    lines = ['error(%g) D0 L0' % p]
    lines += ['error(%g) D%d D%d' % (p, k, k + 1) for k in range(n - 1)]
    lines.append('error(%g) D%d' % (p, n - 1))
    lines += ['detector(%g, 0, 0, %d) D%d' % (k + 0.5, 3 + (k % 2), k) for k in range(n)]
    return '\n'.join(lines) + '\n'


def all_fired(n):
    dets = np.zeros((1, (n + 7) // 8), np.uint8)
    for k in range(n):
        dets[0, k >> 3] |= 1 << (k & 7)
    return dets


def test_component_size_limit_is_reported_not_mis_decoded(built):
    """MAX_COMPONENT (decode_core.h): a matching component of up to 2000 events is solved exactly; beyond that the shot
    is reported as `component_too_large` (status 6) -- a documented limit the reference does not have (its sparse blossom
    has no dense per-component matrix). Just below the limit the answer equals the oracle's."""
    ok_n, big_n = 1900, 2100
    text = chain_dem(ok_n)
    o = oracle.OracleDecoder(text)
    o.set_matcher_mode(3)
    o_obs, o_w = o.decode_batch(all_fired(ok_n))
    assert o.matcher_stats()['max_component'] == ok_n
    obs, w, st, _ = SimDecoder(text).decode(all_fired(ok_n), dyn_bytes=512 << 20)
    assert st[0] == 0 and obs[0] == o_obs[0] and w[0] == o_w[0]
    obs, w, st, _ = SimDecoder(chain_dem(big_n)).decode(all_fired(big_n), dyn_bytes=512 << 20)
    assert st[0] == 6
