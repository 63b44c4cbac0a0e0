"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded shots."""
import os

import numpy as np
import pytest

from dem_util import read_dem, NumpyDemSampler

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    # (dem, shots, seed)
    ('phenom_color_code_d5_r5_p1000', 20000, 0),
    ('phenom_color_code_d5_r5_p1000_with_ignored', 4000, 1),
    ('midout_color_code_d5_r10_p1000', 4000, 0),
    ('superdense_color_code_d5_r20_p1000', 2000, 0),
    ('color2surface_d7_phenom_r7_p100', 2000, 0),
    ('color2surface_d5_transit_p100', 2000, 0),
    ('surface_code_d5_r5_p1000', 3000, 0),
    ('rep_code_rg_d9_transit_p10', 3000, 0),
    ('midout_color_code_d9_r36_p1000', 600, 0),
    ('midout488_color_code_d9_r33_p1000', 600, 0),
]


def masks(pred):
    out = np.zeros(pred.shape[0], np.uint64)
    for k in range(pred.shape[1]):
        out |= pred[:, k].astype(np.uint64) << np.uint64(8 * k)
    return out


@pytest.fixture(scope='module')
def libs(built):
    import chromobius_b200
    import oracle
    chromobius_b200.load_library()
    return chromobius_b200, oracle


@pytest.mark.parametrize('group', [32, 128, 256])
@pytest.mark.parametrize('name,shots,seed', CASES)
def test_batch_matches_oracle(libs, name, shots, seed, group):
    cb, oracle = libs
    text = read_dem(name)
    dec = cb._compile_text(text, device=0)
    dec.set_tuning(group_threads=group)
    dets, _ = NumpyDemSampler(text).sample(shots, seed)
    pred, wint = dec.decode_with_integer_weights(dets)
    o_obs, o_w = oracle.OracleDecoder(text).decode_batch(dets)
    assert dec.last_stats.group_threads == group
    assert (wint == o_w).all(), 'integer matching weights must be bit exact on every shot'
    diff = int((masks(pred) != o_obs).sum())
    # Equal-weight matchings may legitimately lift to different predictions; they are counted, not hidden.
    assert diff <= max(2, shots // 200), 'tie-break divergences: %d of %d' % (diff, shots)


@pytest.mark.parametrize('name', ['midout_color_code_d5_r10_p1000', 'superdense_color_code_d5_r20_p1000'])
def test_lift_of_gpu_edges_matches_oracle_lift(libs, name):
    """Stages 3-4 in isolation: the oracle's Euler tour + discharge on the GPU's own matched edges gives the GPU's answer."""
    cb, oracle = libs
    text = read_dem(name)
    dec = cb._compile_text(text, device=0)
    od = oracle.OracleDecoder(text)
    dets, _ = NumpyDemSampler(text).sample(200, 3)
    for row in dets[:60]:
        edges, obs, w = dec.decode_shot_edges(row)
        if len(edges) == 0:
            assert obs == 0 and w == 0
            continue
        assert od.lift_edges(row, edges) == obs
        _, o_obs, o_w = od.decode_edges(row)
        assert o_w == w


@pytest.mark.parametrize('name,shots,group', [('midout_color_code_d9_r36_p1000', 6000, 128),
                                              ('midout_color_code_d9_r36_p1000', 3000, 256),
                                              ('superdense_color_code_d5_r20_p1000', 6000, 32)])
def test_out_of_scratch_retries_give_the_same_answers(libs, monkeypatch, name, shots, group):
    """A shot that runs out of scratch is decoded again with larger scratch: inside the kernel in a spare arena, or, when
    none is free (or CB200_NO_SPARES), by a retry launch. Both must return what a roomy first attempt returns."""
    cb, oracle = libs
    text = read_dem(name)
    dets, _ = NumpyDemSampler(text).sample(shots, 5)
    dec = cb._compile_text(text, device=0)
    dec.set_tuning(group_threads=group)
    pred0, w0 = dec.decode_with_integer_weights(dets)
    assert dec.last_stats.num_retried_shots == 0
    results = {}
    for spares in (True, False):
        monkeypatch.setenv('CB200_TEST_SCRATCH_DIV', '64')
        if not spares:
            monkeypatch.setenv('CB200_NO_SPARES', '1')
        dec2 = cb._compile_text(text, device=0)
        dec2.set_tuning(group_threads=group)
        pred, w = dec2.decode_with_integer_weights(dets)
        st = dec2.last_stats
        monkeypatch.delenv('CB200_TEST_SCRATCH_DIV')
        monkeypatch.delenv('CB200_NO_SPARES', raising=False)
        assert st.num_retried_shots > 0, 'the test did not exercise the retry path'
        assert st.num_failed_shots == 0
        assert (w == w0).all() and (pred == pred0).all()
        results[spares] = st.num_retried_shots
    assert results[True] > 0 and results[False] > 0


@pytest.mark.parametrize('name,shots,chunk', [('midout_color_code_d9_r36_p1000', 20000, 1500),
                                              ('superdense_color_code_d5_r20_p1000', 50000, 4096)])
def test_streaming_host_input_gives_the_same_answers(libs, monkeypatch, name, shots, chunk):
    """A large host batch is copied in chunks while the decode kernel already runs (decode_impl: streaming input); the
    answers must not depend on it, also when the chunks that arrive later hold heavier shots than the first chunk the
    arena was sized from (those take the scratch-overflow retry path)."""
    cb, _ = libs
    text = read_dem(name)
    dets, _ = NumpyDemSampler(text).sample(shots, 12)
    # light shots first, the heaviest last: the first chunk under-estimates the batch
    order = np.argsort(np.unpackbits(dets, axis=1).sum(axis=1), kind='stable')
    dets = np.ascontiguousarray(dets[order])
    for k in range(1, 41):  # ... and forty shots far beyond the head room: the XOR of five heavy shots each
        dets[-k] = dets[-k] ^ dets[-k - 50] ^ dets[-k - 100] ^ dets[-k - 150] ^ dets[-k - 200]
    dec = cb._compile_text(text, device=0)
    monkeypatch.setenv('CB200_NO_PIPELINE', '1')
    pred0, w0 = dec.decode_with_integer_weights(dets)
    monkeypatch.delenv('CB200_NO_PIPELINE')
    monkeypatch.setenv('CB200_PIPELINE_CHUNK', str(chunk))
    dec2 = cb._compile_text(text, device=0)
    pred1, w1 = dec2.decode_with_integer_weights(dets)
    assert (pred1 == pred0).all() and (w1 == w0).all()
    assert dec2.last_stats.num_detection_events == dec.last_stats.num_detection_events
    assert dec2.last_stats.max_events_per_shot == dec.last_stats.max_events_per_shot
    assert dec2.last_stats.num_retried_shots > 0, 'the sorted batch should have outgrown the first chunk'
    # strided host rows stream too
    wide = np.zeros((shots, dets.shape[1] + 3), np.uint8)
    wide[:, :dets.shape[1]] = dets
    assert (dec2.predict_obs_flips_from_dets_bit_packed(wide[:, :dets.shape[1]]) == pred0).all()


def test_streaming_input_is_off_when_launches_are_serialised():
    """CUDA_LAUNCH_BLOCKING=1 (or an injected profiler) lets a launch finish before the next copy starts: a persistent
    kernel waiting for streamed chunks would never end. The library must fall back to copy-then-decode."""
    import subprocess
    import sys
    code = (
        "import os, sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import chromobius_b200 as cb\n"
        "from dem_util import read_dem, NumpyDemSampler\n"
        "text = read_dem('midout_color_code_d9_r36_p1000')\n"
        "dets, _ = NumpyDemSampler(text).sample(6000, 3)\n"
        "os.environ['CB200_PIPELINE_CHUNK'] = '1000'\n"
        "a = cb._compile_text(text, device=0).decode_with_integer_weights(dets)\n"
        "os.environ['CUDA_LAUNCH_BLOCKING'] = '1'\n"
        "b = cb._compile_text(text, device=0).decode_with_integer_weights(dets)\n"
        "assert (a[0] == b[0]).all() and (a[1] == b[1]).all()\n"
        "print('ok')\n") % (ROOT, os.path.join(ROOT, 'tests'))
    env = dict(os.environ)
    env['CUDA_LAUNCH_BLOCKING'] = '1'
    env['PYTHONPATH'] = os.path.join(ROOT, 'tools') + os.pathsep + env.get('PYTHONPATH', '')
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stderr[-2000:]


def test_plain_cycle_lift_equals_general_lift(libs, monkeypatch):
    cb, _ = libs
    for name, shots in [('midout_color_code_d9_r36_p1000', 20000), ('color2surface_d7_phenom_r7_p100', 20000),
                        ('toric_superdense_color_code_epr_d12_r5_p1000', 8000)]:
        text = read_dem(name)
        dets, _ = NumpyDemSampler(text).sample(shots, 8)
        a = cb._compile_text(text, device=0).decode_with_integer_weights(dets)
        monkeypatch.setenv('CB200_NO_PLAIN_LIFT', '1')
        b = cb._compile_text(text, device=0).decode_with_integer_weights(dets)
        monkeypatch.delenv('CB200_NO_PLAIN_LIFT')
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


def test_component_size_limit_on_the_gpu(libs):
    """The giant-component shots of tests/test_decode_path_cpu.py through the CUDA path: 1900 events in one component
    decode to the oracle's answer (after the scratch retries enlarge the dense blossom's memory); 2100 fail the call
    with the documented message instead of returning a wrong answer."""
    from test_decode_path_cpu import all_fired, chain_dem
    cb, oracle = libs
    text = chain_dem(1900)
    dec = cb._compile_text(text, device=0)
    pred, w = dec.decode_with_integer_weights(all_fired(1900))
    od = oracle.OracleDecoder(text)
    od.set_matcher_mode(3)
    o_obs, o_w = od.decode_batch(all_fired(1900))
    assert w[0] == o_w[0] and masks(pred)[0] == o_obs[0]
    dec2 = cb._compile_text(chain_dem(2100), device=0)
    with pytest.raises(ValueError, match='exceeded the supported size'):
        dec2.predict_obs_flips_from_dets_bit_packed(all_fired(2100))
    assert dec2.last_stats.first_failed_status == 6
