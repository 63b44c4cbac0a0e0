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
