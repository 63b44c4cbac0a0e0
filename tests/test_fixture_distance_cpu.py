"""CPU tier: the fixture models under tests/golden/dems were written by this repo's own circuit analyser
(tools/circuit_to_dem.py); besides the forward fault enumeration (tests/test_reference_pins.py) they are pinned here by a
number the reference itself states for the circuits they come from: its generator's test requires
`len(noisy.shortest_graphlike_error())` to be base_width // 2 + 1 for the 6.6.6 middle-out circuits and
base_width // 4 + 1 for the 4.8.8 ones (src/clorco/color_code/_midout_planar_color_code_circuits_test.py:77-85, uniform
depolarizing noise) -- the 4.8.8 circuit at d = 9 has graph-like distance 3, which is why its decoding is dominated by
implementation-defined choices."""
import collections

import pytest

from dem_util import read_dem, flatten_dem


def shortest_graphlike_error(text):
    """Fewest graph-like mechanisms (one or two detectors) without symptoms that flip observable 0: the shortest path
    from (v, even) to (v, odd) in the detector graph doubled by observable parity, over all v (boundary included)."""
    _, _, errors, _ = flatten_dem(text)
    adj = collections.defaultdict(list)
    for _, groups in errors:
        dets, obs = [], 0
        for d, o in groups:
            dets += list(d)
            obs ^= int(0 in o)
        if len(dets) == 1:
            a, b = dets[0], -1
        elif len(dets) == 2:
            a, b = dets
        else:
            continue
        adj[a].append((b, obs))
        adj[b].append((a, obs))
    best = None
    for s in list(adj):
        dist = {(s, 0): 0}
        queue = collections.deque([(s, 0)])
        while queue:
            v, par = queue.popleft()
            dv = dist[(v, par)]
            if best is not None and dv + 1 >= best:
                break
            for w, m in adj[v]:
                st = (w, par ^ m)
                if st not in dist:
                    dist[st] = dv + 1
                    queue.append(st)
        if (s, 1) in dist and (best is None or dist[(s, 1)] < best):
            best = dist[(s, 1)]
    return best


@pytest.mark.parametrize('name,expected', [
    ('midout_color_code_d5_r10_p1000', 5 // 2 + 1),
    ('midout_color_code_d9_r36_p1000', 9 // 2 + 1),
    ('midout488_color_code_d9_r33_p1000', 9 // 4 + 1),
])
def test_fixture_dem_has_the_graphlike_distance_the_reference_generator_requires(name, expected):
    assert shortest_graphlike_error(read_dem(name)) == expected
