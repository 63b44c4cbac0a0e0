"""GPU: the on-device DEM sampler produces the distribution of the independent-mechanism model."""
import numpy as np
import pytest

from dem_util import read_dem, NumpyDemSampler

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,shots', [('midout_color_code_d9_r36_p1000', 40000), ('phenom_color_code_d5_r5_p1000', 200000)])
def test_sampler_rates(name, shots):
    import chromobius_b200 as cb
    text = read_dem(name)
    smp = cb.DeviceDemSampler(text, device=0)
    dets, obs = smp.sample(shots, seed=3)
    bits = np.unpackbits(dets, axis=1, bitorder='little')[:, :smp.num_detectors]
    ref = NumpyDemSampler(text)
    # per-detector firing rates against the exact marginals of the model: P(det) = (1 - prod(1 - 2 p_e)) / 2
    prod = np.ones(ref.num_dets)
    for p, d in zip(ref.p, ref.dets):
        prod[d] *= 1 - 2 * p
    want = (1 - prod) / 2
    got = bits.mean(axis=0)
    sigma = np.sqrt(want * (1 - want) / shots) + 1e-9
    z = np.abs(got - want) / sigma
    assert z.max() < 6.0, 'per-detector rate off by %.1f sigma' % z.max()
    assert abs(bits.sum() / shots - want.sum()) < 6 * np.sqrt(want.sum() / shots) + 1e-3
    # observable marginals
    for k in range(smp.num_observables):
        prod_o = 1.0
        for p, o in zip(ref.p, ref.obs):
            if (int(o) >> k) & 1:
                prod_o *= 1 - 2 * p
        w = (1 - prod_o) / 2
        g = float(((obs >> np.uint64(k)) & np.uint64(1)).mean())
        assert abs(g - w) < 6 * np.sqrt(w * (1 - w) / shots) + 1e-9
    # reproducible, and different seeds differ
    again, _ = smp.sample(1000, seed=3)
    assert (again == dets[:1000]).all()
    other, _ = smp.sample(1000, seed=4)
    assert (other != dets[:1000]).any()


def test_sampled_shots_decode_consistently():
    """Sampler -> decoder on the device: pad bits are zero, shots decode, logical error rate is plausible."""
    import chromobius_b200 as cb
    text = read_dem('midout_color_code_d9_r36_p1000')
    smp = cb.DeviceDemSampler(text, device=0)
    dec = cb._compile_text(text, device=0)
    dets, obs = smp.sample(20000, seed=9)
    pred = dec.predict_obs_flips_from_dets_bit_packed(dets)
    mistakes = int((pred[:, 0].astype(np.uint64) != obs).sum())
    # published logical error rate 6.0e-3 per shot (assets/stats.csv:64); 8192-shot test count 45 (decoder_integration.test.cc:148)
    assert 60 <= mistakes <= 190, mistakes


def test_shots_are_independent_of_each_other():
    """Streams of different shots must not be shifted copies (round-1 bug: draw j of shot s + 1 was draw j + 1 of shot s,
    so the first gap of one shot equalled the second gap of the previous one). One probability class, 32 x 200 errors
    `error(p) Dk`: lane 0 of the sampling warp covers detectors 0..199."""
    import chromobius_b200 as cb
    n, per_lane, p = 32 * 200, 200, 0.05
    text = '\n'.join('error(%g) D%d' % (p, k) for k in range(n)) + '\n' + \
           '\n'.join('detector(0, 0, 0, %d) D%d' % (k % 3, k) for k in range(n)) + '\n'
    smp = cb.DeviceDemSampler(text, device=0)
    shots = 20000
    dets, _ = smp.sample(shots, seed=1)
    bits = np.unpackbits(dets, axis=1, bitorder='little')[:, :per_lane]
    first = np.full(shots, -1)
    second_gap = np.full(shots, -1)
    for s in range(shots):
        hits = np.nonzero(bits[s])[0]
        if len(hits) >= 1:
            first[s] = hits[0]
        if len(hits) >= 2:
            second_gap[s] = hits[1] - hits[0] - 1
    ok = (first[1:] >= 0) & (second_gap[:-1] >= 0)
    a, b = first[1:][ok].astype(float), second_gap[:-1][ok].astype(float)
    assert len(a) > 10000
    corr = np.corrcoef(a, b)[0, 1]
    assert abs(corr) < 0.05, 'first gap of shot s+1 is correlated with the second gap of shot s: %.3f' % corr
    # and rows of neighbouring shots agree no more often than rows of distant shots
    near = (bits[1:] & bits[:-1]).sum() / (shots - 1)
    far = (bits[100:] & bits[:-100]).sum() / (shots - 100)
    assert abs(near - far) < 0.05
