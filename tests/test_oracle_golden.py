"""Pins the CPU oracle AND the product's host table compiler against the reference's own known-answer tests.

Every expected value below is restated from a test in /root/reference/src/chromobius (file:line in each test).
Both implementations are checked: `oracle.OracleDecoder` (oracle/chromobius_oracle.cc) and the host compile pass of the
product (chromobius_b200/csrc/host_compile.cc) together with the host simulation of the device decode path
(tests/sim/host_sim.cc).
"""
import re

import numpy as np
import pytest

import oracle
from sim_util import SimDecoder

TWO_DET_DEM = """
    error(0.1) D0 L0
    error(0.1) D0 D1 L1
    error(0.1) D1 L2
    detector(0, 0, 0, 0) D0
    detector(0, 0, 0, 1) D1
"""

D5_X_ONLY_DEM = """
    error(0.1) D0 L0 L1
    error(0.1) D0 D2 L0
    error(0.1) D2 D3 L0
    error(0.1) D3 D7 L0
    error(0.1) D7 L0
    error(0.1) D0 D1
    error(0.1) D0 D1 D2 L1
    error(0.1) D1 D2 D4 L1
    error(0.1) D1 D4
    error(0.1) D4 D6
    error(0.1) D6
    error(0.1) D6 D8
    error(0.1) D4 D5 D8 L1
    error(0.1) D2 D4 D5
    error(0.1) D3 D5 D7
    error(0.1) D5 D7
    error(0.1) D5 D8 L1
    error(0.1) D2 D3 D5
    error(0.1) D4 D6 D8
    detector(+0, 0, 0, 0) D0
    detector(+0, 2, 0, 1) D2
    detector(+0, 4, 0, 2) D5
    detector(+1, 1, 0, 2) D1
    detector(+1, 3, 0, 0) D4
    detector(+1, 5, 0, 1) D8
    detector(+2, 4, 0, 2) D6
    detector(-1, 3, 0, 0) D3
    detector(-1, 5, 0, 1) D7
"""

IMPLS = ['oracle', 'product_host']


def make(impl, dem):
    return oracle.OracleDecoder(dem) if impl == 'oracle' else SimDecoder(dem)


def decode_masks(impl, dec, dets):
    dets = np.ascontiguousarray(dets, np.uint8)
    if impl == 'oracle':
        return dec.decode_batch(dets)[0]
    obs, w, st, rd = dec.decode(dets)
    assert (st == 0).all()
    return obs


def parse_mobius_errors(text):
    """-> list of (p, frozenset of frozenset edges) from `error(p) Da Db ^ Dc Dd` lines."""
    out = []
    for line in text.split('\n'):
        line = line.strip()
        m = re.match(r'^error\(([^)]*)\)\s*(.*)$', line)
        if not m:
            continue
        comps = [tuple(int(t[1:]) for t in c.split()) for c in m.group(2).split('^')]
        out.append((float(m.group(1)), [tuple(c) for c in comps]))
    return out


def section(text, start, end):
    a = text.index(start)
    b = text.index(end, a)
    return text[a:b]


@pytest.mark.parametrize('impl', IMPLS)
def test_describe_decoder_golden(built, impl):
    # commands/main_describe_decoder.test.cc:21-73
    d = make(impl, TWO_DET_DEM).describe()
    assert section(d, '.charge_graph', '.rgb_reps') == """.charge_graph=ChargeGraph{.nodes={
    ChargeGraphNode{.neighbors={{0,0}, {1,2}, {BOUNDARY_NODE,1}}}, // node 0
    ChargeGraphNode{.neighbors={{0,2}, {1,0}, {BOUNDARY_NODE,4}}}, // node 1
}}

"""
    assert section(d, '.rgb_reps', '.drag_graph') == """.rgb_reps={
    RgbEdge{.red_node=0, .green_node=1, .blue_node=BOUNDARY_NODE, .obs_flip=2, .charge_flip=B} // rep 0
    RgbEdge{.red_node=0, .green_node=1, .blue_node=BOUNDARY_NODE, .obs_flip=2, .charge_flip=B} // rep 1
}

"""
    assert section(d, '.drag_graph', '.mobius_dem') == """.drag_graph=DragGraph{.mmm={
    NEUTRAL@0:NEUTRAL@0 = 0
    NEUTRAL@0:R@0 = 1
    R@0:NEUTRAL@0 = 1
    NEUTRAL@0:NEUTRAL@1 = 0
    R@0:R@1 = 0
    R@0:G@1 = 2
    G@0:G@1 = 0
    NEUTRAL@1:NEUTRAL@0 = 0
    R@1:R@0 = 0
    G@1:R@0 = 2
    G@1:G@0 = 0
    NEUTRAL@1:NEUTRAL@1 = 0
    NEUTRAL@1:G@1 = 4
    G@1:NEUTRAL@1 = 4
}}

"""
    errs = parse_mobius_errors(d[d.index('.mobius_dem'):])
    assert [(round(p, 12), e) for p, e in errs] == [
        (0.01, [(0, 1)]),
        (0.1, [(0, 2), (1, 3)]),
        (0.01, [(2, 3)]),
    ]


@pytest.mark.parametrize('impl', IMPLS)
def test_main_predict_golden(built, impl):
    # commands/main_predict.test.cc:22-44: shots {}, {D0}, {D1}, {D0 D1} -> {}, L0, L2, L1
    dec = make(impl, TWO_DET_DEM)
    dets = np.array([[0b00], [0b01], [0b10], [0b11]], np.uint8)
    assert list(decode_masks(impl, dec, dets)) == [0, 1, 4, 2]


@pytest.mark.parametrize('impl', IMPLS)
def test_main_benchmark_golden(built, impl):
    # commands/main_benchmark.test.cc:22-58: actual observables L0, L0, L2, L1 -> one mistake (the first shot)
    dec = make(impl, TWO_DET_DEM)
    dets = np.array([[0b00], [0b01], [0b10], [0b11]], np.uint8)
    actual = np.array([1, 1, 4, 2], np.uint64)
    pred = decode_masks(impl, dec, dets)
    assert int((pred != actual).sum()) == 1
    assert int(np.unpackbits(dets).sum()) == 4  # num_detection_events, detection_fraction = 4 / (4 * 2)


def mobius_errors_of(impl, dec):
    if impl == 'oracle':
        d = dec.describe()
        return parse_mobius_errors(d[d.index('.mobius_dem'):])
    return parse_mobius_errors(dec.mobius_dem())


def norm_errors(errs):
    return sorted((round(p, 9), tuple(sorted(tuple(sorted(e)) for e in comps))) for p, comps in errs)


@pytest.mark.parametrize('impl', IMPLS)
def test_d5_x_only_tables_golden(built, impl):
    # decode/decoder.test.cc:23-102
    dec = make(impl, D5_X_ONLY_DEM)
    d = dec.describe()
    reps = re.findall(r'RgbEdge\{\.red_node=(\w+), \.green_node=(\w+), \.blue_node=(\w+), \.obs_flip=(\d+)', d)
    assert [(int(r), int(g), int(b), int(o)) for r, g, b, o in reps] == [
        (0, 2, 1, 2), (0, 2, 1, 2), (0, 2, 1, 2), (3, 2, 5, 0), (4, 2, 1, 2), (3, 2, 5, 0), (4, 8, 6, 0), (3, 7, 5, 0),
        (4, 8, 5, 2)]
    if impl == 'oracle':
        nc = dec.node_colors()
        R, G, B = 1, 2, 3
        assert list(nc[:, 0]) == [R, B, G, R, R, B, B, G, G]
        assert list(nc[:, 1]) == [1] * 9
    expected = parse_mobius_errors("""
        error(0.01) D0 D1
        error(0.1) D0 D4 ^ D1 D5
        error(0.1) D4 D6 ^ D5 D7
        error(0.1) D6 D14 ^ D7 D15
        error(0.01) D14 D15
        error(0.1) D0 D3 ^ D1 D2
        error(0.1) D1 D5 ^ D2 D4 ^ D0 D3
        error(0.1) D5 D9 ^ D2 D4 ^ D3 D8
        error(0.1) D2 D9 ^ D3 D8
        error(0.1) D8 D13 ^ D9 D12
        error(0.01) D12 D13
        error(0.1) D12 D16 ^ D13 D17
        error(0.1) D9 D17 ^ D10 D16 ^ D8 D11
        error(0.1) D5 D9 ^ D4 D10 ^ D8 D11
        error(0.1) D7 D15 ^ D10 D14 ^ D6 D11
        error(0.1) D10 D14 ^ D11 D15
        error(0.1) D10 D16 ^ D11 D17
        error(0.1) D5 D7 ^ D4 D10 ^ D6 D11
        error(0.1) D9 D17 ^ D12 D16 ^ D8 D13
    """)
    got = mobius_errors_of(impl, dec)
    # stim's approx_equals compares instruction by instruction, in order
    assert [(round(p, 9), [tuple(sorted(e)) for e in c]) for p, c in got] == [
        (round(p, 9), [tuple(sorted(e)) for e in c]) for p, c in expected]


REPEAT_DEM = """
    error(0.125) D0 D1 D2
    error(0.0625) D3 D4 D5
    error(0.0625) D0 D1 D2 D3 D4 D5
    error(0.25) D0 L1
    detector(0, 0, 0, 0) D0
    detector(0, 0, 0, 1) D1
    detector(0, 0, 0, 2) D2
    detector(0, 0, 0, 3) D3
    repeat 2 {
        detector(0, 0, 0, 4) D4
        shift_detectors(0, 0, 0, 1) 1
    }
"""

EXPECTED_REPEAT_MOBIUS = """
    error(0.125) D1 D3 ^ D2 D4 ^ D0 D5
    error(0.0625) D7 D9 ^ D8 D10 ^ D6 D11
    error(0.0625) D1 D3 ^ D2 D4 ^ D0 D5 ^ D7 D9 ^ D8 D10 ^ D6 D11
    error(0.0625) D0 D1
"""


@pytest.mark.parametrize('impl', IMPLS)
def test_mobius_dem_with_repeat_and_shift(built, impl):
    # decode/decoder.test.cc:104-140
    dec = make(impl, REPEAT_DEM)
    assert dec.num_detectors == 6
    assert norm_errors(mobius_errors_of(impl, dec)) == norm_errors(parse_mobius_errors(EXPECTED_REPEAT_MOBIUS))


@pytest.mark.parametrize('impl', IMPLS)
def test_ignored_detectors_are_dropped_and_dem_is_padded(built, impl):
    # decode/decoder.test.cc:142-184: detectors annotated -1 are ignored; the Moebius DEM is padded to 2 * 24 detectors
    dem = REPEAT_DEM.replace('repeat 2 {', """error(0.125) D20 D21 D22 D23
    detector(0, 0, 0, -1) D20
    detector(0, 0, 0, -1) D21
    detector(0, 0, 0, -1) D22
    detector(0, 0, 0, -1) D23
    repeat 2 {""")
    dec = make(impl, dem)
    assert dec.num_detectors == 24
    assert norm_errors(mobius_errors_of(impl, dec)) == norm_errors(parse_mobius_errors(EXPECTED_REPEAT_MOBIUS))
    row_ptr = dec.graph()[0]
    assert len(row_ptr) == 2 * 24 + 1  # "detector D47"
    # events on ignored detectors are not decoded at all
    dets = np.zeros((1, 3), np.uint8)
    dets[0, 2] = 0b11110000  # D20..D23
    assert list(decode_masks(impl, dec, dets)) == [0]


@pytest.mark.parametrize('impl', IMPLS)
def test_mobius_edge_rule_goldens(built, impl):
    # datatypes/atomic_error.test.cc:109-244, read through one-error DEMs: node 2d+s, subgraph offsets per colour pair
    def edges_for(dets_colors):
        lines = ['error(0.1) ' + ' '.join('D%d' % d for d, _ in dets_colors)]
        # make every single detector atomic too so that pairs/triplets are not the only errors (keeps reps defined)
        for d, c in dets_colors:
            lines.append('detector(0, 0, 0, %d) D%d' % (c, d))
        dec = make(impl, '\n'.join(lines))
        errs = mobius_errors_of(impl, dec)
        assert len(errs) == 1
        return sorted(tuple(sorted(e)) for e in errs[0][1])

    R, G, B = 0, 1, 2
    assert edges_for([(2, R)]) == [(4, 5)]
    assert edges_for([(2, R), (3, R)]) == [(4, 6), (5, 7)]
    assert edges_for([(2, G), (3, G)]) == [(4, 6), (5, 7)]
    # R^G = B != G -> straight; R^B = G -> crossed; G^B = R -> straight
    assert edges_for([(2, R), (3, G)]) == [(4, 6), (5, 7)]
    assert edges_for([(2, R), (3, B)]) == [(4, 7), (5, 6)]
    assert edges_for([(2, G), (3, B)]) == [(4, 6), (5, 7)]
    # RGB triplet: NotBlue (r+1, g+1), NotRed (g+0, b+0), NotGreen (r+0, b+1)
    assert edges_for([(2, R), (3, G), (5, B)]) == sorted([(5, 7), (6, 10), (4, 11)])


def test_euler_tour_goldens(built):
    # graph/euler_tours.test.cc:21-102 (exact cycle node order)
    assert oracle.euler_tours(10, []) == []
    with pytest.raises(ValueError):
        oracle.euler_tours(10, [1, 2])
    assert oracle.euler_tours(10, [1, 2, 2, 1]) == [[1, 2]]
    assert oracle.euler_tours(10, [1, 2, 3, 1, 2, 3]) == [[1, 2, 3]]
    assert oracle.euler_tours(10, [1, 2, 4, 5, 2, 1, 5, 6, 6, 4]) == [[1, 2], [4, 5, 6]]
    assert oracle.euler_tours(10, [1, 2, 2, 1, 2, 3, 3, 2, 3, 4, 4, 3, 2, 5, 5, 2]) == [[3, 2, 5, 2, 1, 2, 3, 4]]


def test_charge_graph_goldens(built):
    # graph/charge_graph.test.cc:54-101, through DEMs whose atomic errors are the listed ones
    def graph(dem):
        d = oracle.OracleDecoder(dem).describe()
        rows = re.findall(r'ChargeGraphNode\{\.neighbors=\{(.*)\}\}, // node', d)
        out = []
        for r in rows:
            out.append([(a if a == 'BOUNDARY_NODE' else int(a), int(b)) for a, b in re.findall(r'\{(\w+),(\d+)\}', r)])
        return out

    dets = 'detector(0,0,0,0) D0\ndetector(0,0,0,1) D1\ndetector(0,0,0,2) D2\n'
    assert graph(dets) == [[(0, 0)], [(1, 0)], [(2, 0)]]
    assert graph(dets + 'error(0.1) D1 L0') == [[(0, 0)], [(1, 0), ('BOUNDARY_NODE', 1)], [(2, 0)]]
    assert graph(dets + 'error(0.1) D1 D2 L2') == [[(0, 0)], [(1, 0), (2, 4)], [(1, 4), (2, 0)]]
    assert graph(dets + 'error(0.1) D0 D1 D2 L2') == [[(0, 0)], [(1, 0)], [(2, 0)]]
    dets4 = dets + 'detector(0,0,0,0) D3\n'
    assert graph(dets4 + 'error(0.1) D0 D1 D2 L2\nerror(0.1) D1 D2 D3 L1') == [
        [(0, 0), (3, 6)], [(1, 0)], [(2, 0)], [(0, 6), (3, 0)]]


def test_choose_rgb_reps_golden(built):
    # graph/choose_rgb_reps.test.cc:21-43
    dem = """
        detector(0,0,0,0) D0
        detector(0,0,0,1) D1
        detector(0,0,0,2) D2
        detector(0,0,0,0) D3
        error(0.1) D0 D1 D2 L0
        error(0.1) D2 D3 L1
    """
    reps = oracle.OracleDecoder(dem).rgb_reps()
    NONE = 0xFFFFFFFF
    assert [tuple(int(x) for x in r) for r in reps] == [
        (0, 1, 2, 1, 0, 0), (0, 1, 2, 1, 0, 0), (0, 1, 2, 1, 0, 0), (3, NONE, 2, 2, 0, 2)]


@pytest.mark.parametrize('impl', IMPLS)
def test_compile_errors(built, impl):
    # decode/decoder_integration.test.cc:215-219 and pybind/chromobius_pybind_test.py:36-47 ('4 coordinates')
    with pytest.raises(ValueError, match='4 coordinates'):
        make(impl, 'error(0.1) D0\ndetector D0')
    with pytest.raises(ValueError, match='4 coordinates'):
        make(impl, 'error(0.1) D0\ndetector(0, 0, 0, 0.5) D0')
    with pytest.raises(ValueError, match='4 coordinates'):
        make(impl, 'error(0.1) D0\ndetector(0, 0, 0, 6) D0')
    dec = make(impl, '')  # an empty DEM compiles (pybind/chromobius_pybind_test.py:79-81)
    assert dec.num_detectors == 0


@pytest.mark.parametrize('impl', IMPLS)
def test_fix_check_event_lists_decode(built, impl):
    # decode/decoder_integration.test.cc:170-176: these detection-event lists must decode without failing
    from dem_util import read_dem
    dec = make(impl, read_dem('fix_check_1'))
    assert dec.num_detectors == 9000
    for events in ([8687, 8736, 8737, 8763, 8945, 8720], [8915, 8914, 8890]):
        row = np.zeros((1, (9000 + 7) // 8), np.uint8)
        for d in events:
            row[0, d >> 3] |= 1 << (d & 7)
        decode_masks(impl, dec, row)
