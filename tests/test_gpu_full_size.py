"""GPU: BASELINE.json-sized batches checked through size-independent properties, plus the oracle on a sample of them.

Properties (all exact, no tolerance -- the path is integer / byte work):
  * batch independence: decoding a batch equals decoding its halves / a permutation of it
  * determinism: decoding twice gives identical bytes
  * strided (padded-row) input decodes like contiguous input
  * empty shots predict 0 with weight 0; weights are >= 0 and even; pad bits and ignored detectors do not matter
  * single-error known answers at full DEM size (decode/decoder_integration.test.cc:44-120, extended to every k-th error)
  * host API == device-resident API
"""
import numpy as np
import pytest

from dem_util import read_dem, NumpyDemSampler

pytestmark = pytest.mark.gpu

FULL = [
    # (dem, shots at BASELINE size, oracle sample)
    ('phenom_color_code_d5_r5_p1000', 100000, 3000),
    ('midout_color_code_d9_r36_p1000', 1000000, 300),
    ('superdense_color_code_d5_r20_p1000', 1000000, 300),
    ('midout488_color_code_d9_r33_p1000', 1000000, 300),
    ('color2surface_d7_phenom_r7_p100', 1000000, 300),
    ('midout_color_code_d25_r100_p1000', 40000, 40),   # per-GPU slice of the 10M-shot config is bench.py's job
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


@pytest.mark.parametrize('name,shots,n_oracle', FULL)
def test_full_size_properties(libs, name, shots, n_oracle):
    cb, oracle = libs
    text = read_dem(name)
    dec = cb._compile_text(text, device=0)
    smp = cb.DeviceDemSampler(text, device=0)
    dets, actual = smp.sample(shots, seed=77)
    pred, w = dec.decode_with_integer_weights(dets)
    m = masks(pred)
    # determinism
    pred2, w2 = dec.decode_with_integer_weights(dets)
    assert (pred2 == pred).all() and (w2 == w).all()
    # weights: non-negative, even; empty shots -> prediction 0, weight 0
    assert (w >= 0).all() and (w % 2 == 0).all()
    empty = ~dets.any(axis=1)
    assert (m[empty] == 0).all() and (w[empty] == 0).all()
    if name.startswith('phenom'):
        assert empty.mean() > 0.5
    # batch independence: halves and a permutation (on a slice to keep the test fast)
    n = min(shots, 200000)
    half = n // 2
    pa, wa = dec.decode_with_integer_weights(dets[:half])
    pb, wb = dec.decode_with_integer_weights(dets[half:n])
    assert (np.concatenate([pa, pb]) == pred[:n]).all() and (np.concatenate([wa, wb]) == w[:n]).all()
    perm = np.random.default_rng(5).permutation(n)
    pp, wp = dec.decode_with_integer_weights(dets[:n][perm])
    assert (pp == pred[:n][perm]).all() and (wp == w[:n][perm]).all()
    # strided rows through the reference-shaped API
    k = min(n, 50000)
    wide = np.zeros((k, dets.shape[1] + 5), np.uint8)
    wide[:, :dets.shape[1]] = dets[:k]
    wide[:, dets.shape[1]:] = 0xFF  # garbage beyond the row must never be read as detection events
    assert (dec.predict_obs_flips_from_dets_bit_packed(wide[:, :dets.shape[1]]) == pred[:k]).all()
    # float weights = integer weights / normalising constant (decode/pymatcher.cc:40)
    pf, wf = dec.predict_weighted_obs_flips_from_dets_bit_packed(dets[:k])
    assert (pf == pred[:k]).all()
    assert np.allclose(wf, (w[:k] / dec.normalising_constant).astype(np.float32), rtol=1e-6, atol=0)
    # pad bits of the last byte are masked
    if dec.num_detectors % 8:
        padded = dets[:1000].copy()
        padded[:, -1] |= np.uint8((0xFF << (dec.num_detectors % 8)) & 0xFF)
        assert (dec.predict_obs_flips_from_dets_bit_packed(padded) == pred[:1000]).all()
    # the oracle on a sample: weights bit exact, predictions equal up to counted tie-break divergences
    od = oracle.OracleDecoder(text)
    o_obs, o_w = od.decode_batch(dets[:n_oracle])
    assert (w[:n_oracle] == o_w).all()
    assert int((m[:n_oracle] != o_obs).sum()) <= max(1, n_oracle // 100)
    # logical error rate sanity against the sampled ground truth
    rate = float((m != actual).mean())
    assert rate < 0.2


@pytest.mark.parametrize('name,step', [('midout_color_code_d9_r36_p1000', 1), ('midout_color_code_d25_r100_p1000', 23)])
def test_single_error_known_answers_full_dem(libs, name, step):
    """Every (step-th) error mechanism of the DEM, fired alone, must be decoded to its own observable mask."""
    cb, _ = libs
    text = read_dem(name)
    smp = NumpyDemSampler(text)
    dec = cb._compile_text(text, device=0)
    ks = np.arange(0, len(smp.p), step)
    rows = np.zeros((len(ks), smp.nbytes), np.uint8)
    want = np.zeros(len(ks), np.uint64)
    for i, k in enumerate(ks):
        d = smp.dets[k]
        np.bitwise_or.at(rows[i], d >> 3, (1 << (d & 7)).astype(np.uint8))
        want[i] = smp.obs[k]
    pred = dec.predict_obs_flips_from_dets_bit_packed(rows)
    bad = np.nonzero(masks(pred) != want)[0]
    # Errors that chromobius cannot decompose are dropped from the Moebius DEM (decoder.h:40-62 default options), so a
    # handful of exotic mechanisms may legitimately decode to a different class; the reference test only probes k < D.
    assert len(bad) <= len(ks) // 200, (len(bad), len(ks))
    assert (masks(pred)[ks < smp.num_dets] == want[ks < smp.num_dets]).all()


def test_device_resident_api_matches_host_api(libs):
    import torch
    cb, _ = libs
    text = read_dem('midout_color_code_d9_r36_p1000')
    dec = cb._compile_text(text, device=0)
    smp = cb.DeviceDemSampler(text, device=0)
    n, nb, ob = 50000, dec.num_detector_bytes, dec.num_observable_bytes
    d_dets = torch.empty((n, nb), dtype=torch.uint8, device='cuda:0')
    d_obs = torch.empty(n, dtype=torch.int64, device='cuda:0')
    smp.sample_device(n, d_dets.data_ptr(), nb, d_obs.data_ptr(), seed=5)
    d_pred = torch.empty((n, ob), dtype=torch.uint8, device='cuda:0')
    d_w = torch.empty(n, dtype=torch.int64, device='cuda:0')
    torch.cuda.synchronize()
    st = dec.decode_device(d_dets.data_ptr(), n, nb, d_pred.data_ptr(), weight_int_ptr=d_w.data_ptr())
    torch.cuda.synchronize()
    assert st.kernel_launches >= 3
    host_pred, host_w = dec.decode_with_integer_weights(d_dets.cpu().numpy())
    assert (d_pred.cpu().numpy() == host_pred).all() and (d_w.cpu().numpy() == host_w).all()


def test_one_gpu_share_of_the_headline_config_with_and_without_spare_arenas(libs, monkeypatch):
    """bench.py's step: 1.25 M shots of midout d25 r100 resident in HBM (one GPU's share of the 10 M-shot configuration).
    The few shots of such a batch that run out of first-level scratch -- the heaviest ones -- must decode to the same
    weights and predictions whether the kernel retries them in a spare arena or a retry launch does; and the
    oracle agrees on those very shots."""
    import torch
    cb, oracle = libs
    name, S = 'midout_color_code_d25_r100_p1000', 1250000
    text = read_dem(name)
    dec = cb._compile_text(text, device=0)
    smp = cb.DeviceDemSampler(text, device=0)
    nb, ob = dec.num_detector_bytes, dec.num_observable_bytes
    dev = torch.device('cuda', 0)
    d_dets = torch.empty((S, nb), dtype=torch.uint8, device=dev)
    d_actual = torch.empty(S, dtype=torch.int64, device=dev)
    smp.sample_device(S, d_dets.data_ptr(), nb, d_actual.data_ptr(), seed=1000)
    torch.cuda.synchronize()
    out = {}
    for spares in (True, False):
        if not spares:
            monkeypatch.setenv('CB200_NO_SPARES', '1')
        d_pred = torch.empty((S, ob), dtype=torch.uint8, device=dev)
        d_w = torch.empty(S, dtype=torch.int64, device=dev)
        st = dec.decode_device(d_dets.data_ptr(), S, nb, d_pred.data_ptr(), weight_int_ptr=d_w.data_ptr())
        monkeypatch.delenv('CB200_NO_SPARES', raising=False)
        assert st.num_failed_shots == 0
        out[spares] = (d_pred, d_w, st.num_retried_shots, st.kernel_launches)
    assert out[True][2] > 0 and out[True][2] == out[False][2], 'this batch is known to hold shots that need the second scratch level'
    assert bool((out[True][0] == out[False][0]).all()) and bool((out[True][1] == out[False][1]).all())
    # the heaviest shots (largest weights) against the oracle
    w = out[True][1]
    top = torch.topk(w, 12).indices.cpu().numpy()
    rows = d_dets[torch.from_numpy(top).to(dev)].cpu().numpy()
    o_obs, o_w = oracle.OracleDecoder(text).decode_batch(rows)
    assert (o_w == w.cpu().numpy()[top]).all()
    mistakes = int((out[True][0][:, 0].to(torch.int64) != d_actual).sum().item())
    assert mistakes < 60  # logical error rate of this configuration is ~2.5e-6
