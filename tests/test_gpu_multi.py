"""GPU: one decoder replicated over several GPUs behind ONE call (cb200_compile_multi / cb200_multi_decode_batch).
With a single visible GPU the same device is used twice, which still exercises the fan-out (two replicas, two host
threads, two streams, results written into disjoint ranges of the caller's buffers); boxes with >= 2 GPUs use them."""
import os
import sys

import numpy as np
import pytest

from dem_util import read_dem, NumpyDemSampler

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    n = torch.cuda.device_count()
    return list(range(min(n, 8))) if n >= 2 else [0, 0]


@pytest.mark.parametrize('name,shots', [('midout_color_code_d9_r36_p1000', 30001), ('phenom_color_code_d5_r5_p1000', 7)])
def test_multi_gpu_decoder_equals_single_gpu_decoder(built, name, shots):
    import chromobius_b200 as cb
    text = read_dem(name)
    dets, _ = NumpyDemSampler(text).sample(shots, seed=4)
    single = cb._compile_text(text, device=0)
    multi = cb._compile_text(text, device=_devices())
    assert isinstance(multi, cb.MultiGpuCompiledDecoder) and multi.num_devices == len(_devices())
    p1, w1 = single.predict_weighted_obs_flips_from_dets_bit_packed(dets)
    p2, w2 = multi.predict_weighted_obs_flips_from_dets_bit_packed(dets)
    assert (p1 == p2).all() and (w1 == w2).all()
    assert multi.last_stats.num_shots == shots
    q1, i1 = single.decode_with_integer_weights(dets)
    q2, i2 = multi.decode_with_integer_weights(dets)
    assert (q1 == q2).all() and (i1 == i2).all()
    # fewer shots than devices, and an empty batch
    assert (multi.predict_obs_flips_from_dets_bit_packed(dets[:1]) == p1[:1]).all()
    assert multi.predict_obs_flips_from_dets_bit_packed(dets[:0]).shape == (0, single.num_observable_bytes)


def test_multi_gpu_through_the_pybind_module(built):
    path = built.build_pybind()
    sys.path.insert(0, os.path.dirname(path))
    import chromobius
    from test_pybind_module import DetectorErrorModel
    text = read_dem('superdense_color_code_d5_r20_p1000')
    dets, _ = NumpyDemSampler(text).sample(5000, seed=2)
    a = chromobius.compile_decoder_for_dem(DetectorErrorModel(text))
    b = chromobius.compile_decoder_for_dem(DetectorErrorModel(text), device=_devices())
    c = chromobius.compile_decoder_for_dem(DetectorErrorModel(text), device='all')
    pa = a.predict_obs_flips_from_dets_bit_packed(dets)
    assert np.array_equal(pa, b.predict_obs_flips_from_dets_bit_packed(dets))
    assert np.array_equal(pa, c.predict_obs_flips_from_dets_bit_packed(dets))


def test_a_failing_shot_fails_the_multi_gpu_call_with_the_batch_index(built):
    import chromobius_b200 as cb
    # detectors 0 and 1 are joined by one error only; D2 is isolated: a shot with D2 alone has no perfect matching
    text = 'error(0.1) D0 D1\nerror(0.1) D2 D3\ndetector(0,0,0,0) D0\ndetector(0,0,0,1) D1\ndetector(0,0,0,3) D2\ndetector(0,0,0,4) D3\n'
    multi = cb._compile_text(text, device=[0, 0])
    dets = np.zeros((10, 1), np.uint8)
    dets[7, 0] = 0b0101  # D0 and D2: neither has its partner
    with pytest.raises(ValueError):
        multi.predict_obs_flips_from_dets_bit_packed(dets)
    assert multi.last_stats.first_failed_shot == 7
