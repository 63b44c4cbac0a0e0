"""The pybind11 module literally named `chromobius` (chromobius_b200/csrc/chromobius_pybind.cc over the C ABI).

Ports /root/reference/src/chromobius/pybind/chromobius_pybind_test.py:22-81 and sinter_compat_pybind_test.py:18-21.
stim is not in this image: the DEM object is a stand-in whose type name reads like stim's (the reference only looks at
`str(type(dem))` and `str(dem)`, pybind/chromobius.pybind.cc:33-38), the DEM text comes from tools/circuit_to_dem.py and
the shots from tests/circuit_sampler.py.
"""
import os
import pickle
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))


class _StimLikeDem:
    def __init__(self, text):
        self._text = text

    def __str__(self):
        return self._text


# str(type(x)) == "<class 'stim._stand_in.DetectorErrorModel'>"
DetectorErrorModel = type('DetectorErrorModel', (_StimLikeDem,), {'__module__': 'stim._stand_in'})

COLOR_REP_CODE = """
    X_ERROR(0.1) 0 1 2 3 4 5 6 7 8
    M 0 1 2 3 4 5 6 7 8
    DETECTOR(0, 0, 0, 0) rec[-9] rec[-8] rec[-7]
    DETECTOR(1, 0, 0, 1) rec[-8] rec[-7] rec[-6]
    DETECTOR(2, 0, 0, 2) rec[-7] rec[-6] rec[-5]
    DETECTOR(3, 0, 0, 0) rec[-6] rec[-5] rec[-4]
    DETECTOR(4, 0, 0, 1) rec[-5] rec[-4] rec[-3]
    DETECTOR(5, 0, 0, 2) rec[-4] rec[-3] rec[-2]
    DETECTOR(6, 0, 0, 0) rec[-3] rec[-2] rec[-1]
    DETECTOR(7, 0, 0, 1) rec[-2] rec[-1]
    OBSERVABLE_INCLUDE(0) rec[-1]
    OBSERVABLE_INCLUDE(1) rec[-2]
    OBSERVABLE_INCLUDE(2) rec[-3]
    OBSERVABLE_INCLUDE(3) rec[-4]
    OBSERVABLE_INCLUDE(4) rec[-5]
    OBSERVABLE_INCLUDE(5) rec[-6]
    OBSERVABLE_INCLUDE(6) rec[-7]
    OBSERVABLE_INCLUDE(7) rec[-8]
    OBSERVABLE_INCLUDE(8) rec[-9]
"""


@pytest.fixture(scope='module')
def chromobius(built):
    built.build_cuda()
    path = built.build_pybind()
    sys.path.insert(0, os.path.dirname(path))
    import chromobius
    assert os.path.samefile(chromobius.__file__, path)
    return chromobius


def test_version(chromobius):
    assert '.' in chromobius.__version__
    assert chromobius.__version__ is not None


def test_surface(chromobius):
    for name in ('compile_decoder_for_dem', 'CompiledDecoder', 'sinter_decoders', '_ChromobiusSinterDecoder',
                 '_ChromobiusSinterCompiledDecoder'):
        assert hasattr(chromobius, name)
    for name in ('from_dem', 'predict_obs_flips_from_dets_bit_packed', 'predict_weighted_obs_flips_from_dets_bit_packed'):
        assert hasattr(chromobius.CompiledDecoder, name)
    d = chromobius.sinter_decoders()
    assert list(d) == ['chromobius']
    dec = d['chromobius']
    # sinter_compat_pybind_test.py:18-21
    assert hasattr(dec, 'compile_decoder_for_dem') and hasattr(dec, 'decode_via_files')
    assert pickle.loads(pickle.dumps(dec)) == dec
    assert not (chromobius._ChromobiusSinterDecoder(0) != dec)


def test_errors(chromobius):
    with pytest.raises(ValueError, match='must be a stim.DetectorErrorModel'):
        chromobius.compile_decoder_for_dem(object())
    with pytest.raises(ValueError, match='4 coordinates'):
        chromobius.compile_decoder_for_dem(DetectorErrorModel("""
            error(0.1) D0
        """))


def _rep_code_dem_and_shots(shots):
    from circuit_sampler import sample_circuit
    from circuit_to_dem import circuit_text_to_dem_text
    dem = circuit_text_to_dem_text(COLOR_REP_CODE, fold=False)
    dets, obs, nd, no = sample_circuit(COLOR_REP_CODE, shots, seed=11)
    assert (nd, no) == (8, 9)
    obs_bytes = np.zeros((shots, 2), np.uint8)
    obs_bytes[:, 0] = obs & np.uint64(0xFF)
    obs_bytes[:, 1] = (obs >> np.uint64(8)) & np.uint64(0xFF)
    return dem, dets, obs_bytes


@pytest.mark.gpu
def test_decoding(chromobius):
    dem, dets, obs = _rep_code_dem_and_shots(1024)
    decoder = chromobius.compile_decoder_for_dem(DetectorErrorModel(dem))
    decoder2 = chromobius.CompiledDecoder.from_dem(DetectorErrorModel(dem))
    predictions = decoder.predict_obs_flips_from_dets_bit_packed(dets)
    predictions2 = decoder2.predict_obs_flips_from_dets_bit_packed(dets)
    assert predictions.dtype == np.uint8 and predictions.shape == (1024, 2)
    assert np.array_equal(predictions, predictions2)
    mistakes = np.count_nonzero(np.any(predictions != obs, axis=1))
    assert mistakes < 100
    # the weighted variant, the 1-D form and a strided (sliced) input
    p3, w3 = decoder.predict_weighted_obs_flips_from_dets_bit_packed(dets)
    assert np.array_equal(p3, predictions) and w3.dtype == np.float32 and w3.shape == (1024,)
    one = decoder.predict_obs_flips_from_dets_bit_packed(dets[5])
    assert one.shape == (2,) and np.array_equal(one, predictions[5])
    wide = np.zeros((1024, 4), np.uint8)
    wide[:, :1] = dets
    assert np.array_equal(decoder.predict_obs_flips_from_dets_bit_packed(wide[:, :1]), predictions)
    with pytest.raises(ValueError, match="dets.dtype wasn't np.uint8"):
        decoder.predict_obs_flips_from_dets_bit_packed(dets.astype(np.int32))
    with pytest.raises(ValueError, match='num_detector_bytes=1'):
        decoder.predict_obs_flips_from_dets_bit_packed(np.zeros((3, 2), np.uint8))
    # same answers as the ctypes mirror
    import chromobius_b200
    mirror = chromobius_b200._compile_text(dem, device=0)
    assert np.array_equal(mirror.predict_obs_flips_from_dets_bit_packed(dets), predictions)


@pytest.mark.gpu
def test_empty(chromobius):
    assert chromobius.compile_decoder_for_dem(DetectorErrorModel('')) is not None


@pytest.mark.gpu
def test_sinter_paths_and_decode_via_files(chromobius, tmp_path):
    """sinter_compat.pybind.cc:30-72 (decode_shots_bit_packed) and :96-132 (decode_via_files: b8 in, b8 out)."""
    dem, dets, obs = _rep_code_dem_and_shots(777)
    sd = chromobius.sinter_decoders()['chromobius']
    compiled = sd.compile_decoder_for_dem(dem=dem)  # any object with a DEM str(): no type check on this path
    pred = compiled.decode_shots_bit_packed(bit_packed_detection_event_data=dets)
    assert pred.shape == (777, 2)
    with pytest.raises(ValueError, match='ndim'):
        compiled.decode_shots_bit_packed(bit_packed_detection_event_data=dets[0])
    dem_path, in_path, out_path = tmp_path / 'm.dem', tmp_path / 'dets.b8', tmp_path / 'pred.b8'
    dem_path.write_text(dem)
    dets.tofile(str(in_path))
    sd.decode_via_files(num_shots=777, num_dets=8, num_obs=9, dem_path=dem_path, dets_b8_in_path=in_path,
                        obs_predictions_b8_out_path=out_path, tmp_dir=tmp_path)
    got = np.fromfile(str(out_path), np.uint8).reshape(777, 2)
    assert np.array_equal(got, pred)
    # the ctypes mirror's decode_via_files writes the same bytes
    import chromobius_b200
    out2 = tmp_path / 'pred2.b8'
    chromobius_b200.sinter_decoders()['chromobius'].decode_via_files(
        num_shots=777, num_dets=8, num_obs=9, dem_path=dem_path, dets_b8_in_path=in_path,
        obs_predictions_b8_out_path=out2, tmp_dir=tmp_path)
    assert out2.read_bytes() == out_path.read_bytes()
