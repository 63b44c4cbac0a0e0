"""Parity with the REFERENCE (not with the oracle's own conventions) on the reference's statistical pins.

* The oracle decodes circuit-level samples of the 13 circuits of decoder_integration.test.cc:144-156 and its mistake
  counts are compared with the reference's pinned counts as independent binomial samples (tests/reference_pins.py).
* The product's compile step must build the same tables as the oracle on every fixture DEM, including the entries that
  depend on the iteration order of the reference's std::unordered_map (graph/charge_graph.h:52, graph/drag_graph.cc:87).
* The fixture DEMs are checked against a forward fault enumeration of the reference's circuits (when /root/reference
  is present), so circuit-level shots + analyser-made DEM do not form a closed loop.
"""
import ctypes
import os

import numpy as np
import pytest

import oracle
from dem_util import flatten_dem, read_dem
from reference_pins import KNOWN_OUTSIDE_BAND, PIN_SHOTS, PINS, count_mistakes, load_circuit_shots, within_band
from sim_util import SimDecoder, lib as sim_lib

REFERENCE_DATA = '/root/reference/test_data'


def _visit_order(keys, which):
    out = np.zeros(len(keys), np.uint32)
    sim_lib().cbsim_container_visit_order(keys.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(len(keys)), which,
                                          out.ctypes.data_as(ctypes.c_void_p))
    return out


def test_container_visit_order_restatement_matches_libstdcxx():
    """host_compile.cc restates libstdc++'s unordered_map iteration order; compare with the real container."""
    rng = np.random.default_rng(7)
    for trial in range(600):
        n = int(rng.integers(1, 150)) if trial < 570 else int(rng.integers(1000, 12000))
        kind = trial % 3
        if kind == 0:
            keys = rng.choice(20000, size=n, replace=False).astype(np.uint32)
        elif kind == 1:
            keys = rng.choice(2 ** 32 - 1, size=n, replace=False).astype(np.uint32)
        else:
            keys = rng.choice(20000, size=n, replace=False).astype(np.uint32)
            keys[rng.integers(0, n)] = 0xFFFFFFFF  # BOUNDARY_NODE is a key of the reference's maps
        assert (_visit_order(keys, 0) == _visit_order(keys, 1)).all(), (trial, n)


@pytest.mark.parametrize('name', [n for n, _ in PINS] + ['fix_check_1', 'hard_config', 'detector_reorder_superdense_round',
                                                         'detector_unreorder_superdense_round'])
def test_product_tables_equal_oracle_tables(name):
    """describe_decoder dumps (charge graph, rgb reps, every drag entry, Moebius DEM) of the two compile steps agree,
    in particular on the circuits where the drag table depends on the container's visit order."""
    text = read_dem(name)
    o = oracle.OracleDecoder(text)
    s = SimDecoder(text)
    assert s.describe() == o.describe()
    if name in ('toric_superdense_color_code_epr_d12_r5_p1000', 'midout488_color_code_d9_r33_p1000',
                'superdense_color_code_d5_r20_p1000'):
        assert o.stats()['drag_ambiguities'] > 0  # these are the order-dependent ones


@pytest.mark.parametrize('name,pin', PINS)
def test_oracle_mistakes_within_band_of_reference_pin(name, pin):
    dets, truth = load_circuit_shots(name)
    dec = oracle.OracleDecoder(read_dem(name))
    assert dets.shape == (PIN_SHOTS, dec.num_detector_bytes)
    pred, _ = dec.decode_batch(dets)
    mistakes = count_mistakes(pred, truth, dec.num_observables)
    ok = within_band(mistakes, len(dets), pin)
    print('%s: %d / %d here, reference pin %d / %d' % (name, mistakes, len(dets), pin, PIN_SHOTS))
    if name in KNOWN_OUTSIDE_BAND:
        if not ok:
            pytest.xfail('%s: %d mistakes vs pin %d -- open parity gap, see DESIGN.md section 6' % (name, mistakes, pin))
        return
    assert ok, (name, mistakes, pin)


@pytest.mark.parametrize('name,rate,errors,shots', [
    # assets/stats.csv:64 and :324 (same circuit families and sizes as the test_data fixtures)
    ('midout_color_code_d9_r36_p1000', 6.00e-3, 1038, 172968),
    ('superdense_color_code_d5_r20_p1000', 3.57e-2, 1031, 28865),
])
def test_oracle_rate_within_band_of_published_rate(name, rate, errors, shots):
    dets, truth = load_circuit_shots(name)
    dec = oracle.OracleDecoder(read_dem(name))
    pred, _ = dec.decode_batch(dets)
    mistakes = count_mistakes(pred, truth, dec.num_observables)
    assert within_band(mistakes, len(dets), errors, shots), (name, mistakes / len(dets), rate)


@pytest.mark.skipif(not os.path.isdir(REFERENCE_DATA), reason='needs the reference circuits (this container only)')
@pytest.mark.parametrize('name', ['superdense_color_code_d5_r20_p1000', 'midout488_color_code_d9_r33_p1000',
                                  'color2surface_d7_phenom_r7_p100', 'rep_code_rg_d9_transit_p10',
                                  'phenom_color_code_d5_r5_p1000_with_ignored'])
def test_fixture_dem_equals_forward_fault_enumeration(name):
    from fault_enumeration import dem_mechanisms, enumerate_faults
    with open(os.path.join(REFERENCE_DATA, name + '.stim')) as f:
        mech, _ = enumerate_faults(f.read())
    _, _, errors, _ = flatten_dem(read_dem(name))
    dem = dem_mechanisms(errors)
    assert set(mech) == set(dem)
    worst = max(abs(mech[k] - dem[k]) / dem[k] for k in dem)
    assert worst < 1e-9


@pytest.mark.skipif(not os.path.isdir(REFERENCE_DATA), reason='needs the reference circuits (this container only)')
def test_committed_circuit_shots_are_reproducible():
    from circuit_sampler import sample_circuit
    name = 'midout_color_code_d5_r10_p1000'
    with open(os.path.join(REFERENCE_DATA, name + '.stim')) as f:
        dets, obs, _, _ = sample_circuit(f.read(), PIN_SHOTS, seed=0)
    g_dets, g_obs = load_circuit_shots(name)
    assert (dets == g_dets).all() and (obs == g_obs).all()
