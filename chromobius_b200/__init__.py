"""chromobius_b200: B200-native drop-in for chromobius's batched decode path.

Mirrors the Python surface of the reference module `chromobius`
(/root/reference/src/chromobius/pybind/chromobius.pybind.cc:120-440,
/root/reference/src/chromobius/pybind/sinter_compat.pybind.cc:149-241, stub doc/chromobius.pyi):

    compile_decoder_for_dem(dem) -> CompiledDecoder
    CompiledDecoder.from_dem(dem)
    CompiledDecoder.predict_obs_flips_from_dets_bit_packed(dets)
    CompiledDecoder.predict_weighted_obs_flips_from_dets_bit_packed(dets)
    sinter_decoders() -> {'chromobius': <sinter.Decoder compatible object>}
    __version__

Everything below the argument checks runs in libchromobius_b200.so (hand-written CUDA for sm_100a behind the C ABI in
include/chromobius_b200.h). There is no CPU fallback: importing works anywhere, but compiling a decoder without the
library or without a GPU raises.
"""
import ctypes
import os

import numpy as np

__version__ = '1.2.dev0+b200'

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get('CHROMOBIUS_B200_LIB') or os.path.join(_HERE, 'libchromobius_b200.so')  # env: A/B builds
_lib = None

CB200_IO_HOST = 0
CB200_IO_DEVICE = 1

_STATUS_NAMES = {0: 'ok', 1: 'scratch_overflow', 2: 'no_perfect_matching', 3: 'euler_failed', 4: 'discharge_failed',
                 5: 'round_limit', 6: 'component_too_large'}


class BatchStats(ctypes.Structure):
    """cb200_batch_stats (include/chromobius_b200.h)."""
    _fields_ = [
        ('num_shots', ctypes.c_uint64),
        ('num_detection_events', ctypes.c_uint64),
        ('num_retried_shots', ctypes.c_uint64),
        ('num_failed_shots', ctypes.c_uint64),
        ('first_failed_shot', ctypes.c_int64),
        ('first_failed_status', ctypes.c_uint32),
        ('max_events_per_shot', ctypes.c_uint32),
        ('group_threads', ctypes.c_uint32),
        ('num_groups', ctypes.c_uint32),
        ('kernel_launches', ctypes.c_uint64),
        ('ms_h2d', ctypes.c_float),
        ('ms_count', ctypes.c_float),
        ('ms_decode', ctypes.c_float),
        ('ms_d2h', ctypes.c_float),
        ('ms_total', ctypes.c_float),
        ('matcher_rounds', ctypes.c_uint64),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


def library_path():
    return _LIB_PATH


def load_library():
    """Loads libchromobius_b200.so; raises ImportError when it has not been built (see __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            'chromobius_b200: %s is missing. Build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(nvcc, sm_100a). There is no CPU fallback.' % _LIB_PATH)
    L = ctypes.CDLL(_LIB_PATH)
    vp, u64, i64, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint32
    L.cb200_last_error.restype = ctypes.c_char_p
    L.cb200_compile.restype = ctypes.c_int
    L.cb200_compile.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(vp)]
    L.cb200_free.argtypes = [vp]
    L.cb200_free.restype = None
    for name in ['cb200_num_detectors', 'cb200_num_observables', 'cb200_num_detector_bytes', 'cb200_num_observable_bytes',
                 'cb200_table_bytes', 'cb200_graph_half_edges']:
        getattr(L, name).restype = u64
        getattr(L, name).argtypes = [vp]
    L.cb200_normalising_constant.restype = ctypes.c_double
    L.cb200_normalising_constant.argtypes = [vp]
    L.cb200_describe.restype = vp
    L.cb200_describe.argtypes = [vp]
    L.cb200_free_string.argtypes = [vp]
    L.cb200_free_string.restype = None
    L.cb200_decode_batch.restype = ctypes.c_int
    L.cb200_decode_batch.argtypes = [vp, vp, u64, u64, vp, vp, vp, u32, ctypes.POINTER(BatchStats)]
    L.cb200_decode_shot_edges.restype = ctypes.c_int
    L.cb200_decode_shot_edges.argtypes = [vp, vp, vp, u64, ctypes.POINTER(u64), ctypes.POINTER(u64), ctypes.POINTER(i64)]
    L.cb200_set_tuning.restype = ctypes.c_int
    L.cb200_set_tuning.argtypes = [vp, u32, u32]
    L.cb200_set_groups_per_sm.restype = ctypes.c_int
    L.cb200_set_groups_per_sm.argtypes = [vp, u32]
    L.cb200_profile_phases.restype = ctypes.c_int
    L.cb200_profile_phases.argtypes = [vp, ctypes.c_int, vp]
    L.cb200_compile_multi.restype = ctypes.c_int
    L.cb200_compile_multi.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(vp)]
    L.cb200_multi_free.argtypes = [vp]
    L.cb200_multi_free.restype = None
    L.cb200_multi_num_devices.restype = ctypes.c_int
    L.cb200_multi_num_devices.argtypes = [vp]
    L.cb200_multi_replica.restype = vp
    L.cb200_multi_replica.argtypes = [vp, ctypes.c_int]
    L.cb200_multi_decode_batch.restype = ctypes.c_int
    L.cb200_multi_decode_batch.argtypes = [vp, vp, u64, u64, vp, vp, vp, ctypes.POINTER(BatchStats)]
    L.cb200_sampler_last_error.restype = ctypes.c_char_p
    L.cb200_sampler_create.restype = ctypes.c_int
    L.cb200_sampler_create.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(vp)]
    L.cb200_sampler_free.argtypes = [vp]
    L.cb200_sampler_free.restype = None
    L.cb200_sampler_num_detectors.restype = u64
    L.cb200_sampler_num_detectors.argtypes = [vp]
    L.cb200_sampler_num_observables.restype = u64
    L.cb200_sampler_num_observables.argtypes = [vp]
    L.cb200_sampler_expected_events_per_shot.restype = ctypes.c_double
    L.cb200_sampler_expected_events_per_shot.argtypes = [vp]
    L.cb200_sampler_sample.restype = ctypes.c_int
    L.cb200_sampler_sample.argtypes = [vp, u64, u64, u64, vp, u64, vp, u32]
    _lib = L
    return L


def _raise_last(L):
    raise ValueError(L.cb200_last_error().decode('utf-8', 'replace'))


class CompiledDecoder:
    """A chromobius decoder that has been configured for a specific decoding task, resident on one B200.

    Reference: CompiledDecoder in pybind/chromobius.pybind.cc:26-118.
    """

    def __init__(self, _handle, _lib_ref):
        self._h = _handle
        self._L = _lib_ref
        self.num_detectors = int(_lib_ref.cb200_num_detectors(_handle))
        self.num_observables = int(_lib_ref.cb200_num_observables(_handle))
        self.num_detector_bytes = int(_lib_ref.cb200_num_detector_bytes(_handle))
        self.num_observable_bytes = int(_lib_ref.cb200_num_observable_bytes(_handle))
        self.last_stats = None

    def __del__(self):
        try:
            if self._h:
                self._L.cb200_free(self._h)
                self._h = None
        except Exception:
            pass

    @staticmethod
    def from_dem(dem, *, device=None):
        """Compiles a decoder for a stim.DetectorErrorModel (pybind/chromobius.pybind.cc:32-47).

        The DEM crosses the boundary as text (`str(dem)`), exactly as in the reference. `device`: a GPU index (default:
        CHROMOBIUS_B200_DEVICE / LOCAL_RANK / 0), a list of indices, or 'all' -- the latter two return a decoder that
        splits every batch over those GPUs (MultiGpuCompiledDecoder).
        """
        type_name = str(type(dem))
        if 'stim.' not in type_name or '.DetectorErrorModel' not in type_name:
            raise ValueError('dem must be a stim.DetectorErrorModel.')
        return _compile_text(str(dem), device)

    # -- the hot path ------------------------------------------------------------------------------------------
    def _decode_host(self, dets_ptr, num_shots, stride, obs_ptr, weight_ptr, weight_int_ptr, stats):
        return self._L.cb200_decode_batch(self._h, dets_ptr, num_shots, stride, obs_ptr, weight_ptr, weight_int_ptr,
                                          CB200_IO_HOST, ctypes.byref(stats))

    def _predict(self, dets, include_weight):
        # argument checks and messages follow pybind/chromobius.pybind.cc:49-93
        if not isinstance(dets, np.ndarray) or dets.dtype != np.uint8:
            raise ValueError("Expected bit packed detection event data, but dets.dtype wasn't np.uint8.")
        if dets.ndim == 2:
            num_shots = dets.shape[0]
            if dets.strides[1] != 1 and dets.shape[1] > 1:
                raise ValueError(
                    "Bit packed shot data must be contiguous in memory, but dets.stride[1] wasn't equal to 1.\n"
                    'It was %d.' % dets.strides[1])
            det_shape = dets.shape[1]
            stride = dets.strides[0]
            out_shape = (num_shots, self.num_observable_bytes)
            w_shape = (num_shots,)
        elif dets.ndim == 1:
            num_shots = 1
            det_shape = dets.shape[0]
            if dets.strides[0] != 1 and dets.shape[0] > 1:
                dets = np.ascontiguousarray(dets)
            stride = self.num_detector_bytes
            out_shape = (self.num_observable_bytes,)
            w_shape = ()
        else:
            raise ValueError('dets.shape not in [1, 2]')
        if det_shape != self.num_detector_bytes:
            raise ValueError(
                'Expected dets.shape[1]=%d == num_detector_bytes=%d because dets.shape == 2 and dets.dtype==np.uint8 '
                'indicating bit packed shots.' % (det_shape, self.num_detector_bytes))
        if stride < self.num_detector_bytes or stride < 0:
            dets = np.ascontiguousarray(dets)  # broadcast / negative strides: the reference reads them, we copy first
            stride = self.num_detector_bytes
        result = np.empty(out_shape, np.uint8)
        weights = np.empty(w_shape, np.float32) if include_weight else None
        stats = BatchStats()
        rc = self._decode_host(dets.ctypes.data_as(ctypes.c_void_p), num_shots, stride, result.ctypes.data_as(ctypes.c_void_p),
                               weights.ctypes.data_as(ctypes.c_void_p) if include_weight else None, None, stats)
        self.last_stats = stats
        if rc != 0:
            _raise_last(self._L)
        if include_weight:
            return result, weights
        return result

    def predict_obs_flips_from_dets_bit_packed(self, dets):
        """Predicts observable flips from bit packed detection event data.

        dets: uint8 array, shape (num_shots, ceil(num_detectors / 8)) or (ceil(num_detectors / 8),); bit d of a shot
        is (dets[shot, d // 8] >> (d % 8)) & 1. Returns uint8 (num_shots, ceil(num_observables / 8)) (or 1-D for a
        1-D input). Reference: pybind/chromobius.pybind.cc:153-158.
        """
        return self._predict(dets, False)

    def predict_weighted_obs_flips_from_dets_bit_packed(self, dets):
        """Same as predict_obs_flips_from_dets_bit_packed plus the float32 matching weight of every shot.

        Reference: pybind/chromobius.pybind.cc:236-241, decode/pymatcher.cc:32-41.
        """
        return self._predict(dets, True)

    # -- extras used by bench / tests (not part of the reference surface) ------------------------------------------
    def decode_device(self, dets_ptr, num_shots, row_stride, obs_ptr, weight_int_ptr=None, weight_ptr=None):
        """Device-resident decode: all pointers are device addresses on this decoder's GPU."""
        stats = BatchStats()
        rc = self._L.cb200_decode_batch(self._h, ctypes.c_void_p(dets_ptr), num_shots, row_stride, ctypes.c_void_p(obs_ptr),
                                        ctypes.c_void_p(weight_ptr) if weight_ptr else None,
                                        ctypes.c_void_p(weight_int_ptr) if weight_int_ptr else None, CB200_IO_DEVICE,
                                        ctypes.byref(stats))
        self.last_stats = stats
        if rc != 0:
            _raise_last(self._L)
        return stats

    def decode_with_integer_weights(self, dets):
        """Host batch decode that also returns the exact int64 matching weights (parity hook)."""
        dets = np.ascontiguousarray(dets, dtype=np.uint8)
        if dets.ndim == 1:
            dets = dets[None, :]
        n = dets.shape[0]
        if dets.shape[1] != self.num_detector_bytes:
            raise ValueError('dets.shape[1] != num_detector_bytes')
        result = np.empty((n, self.num_observable_bytes), np.uint8)
        wint = np.empty(n, np.int64)
        stats = BatchStats()
        rc = self._decode_host(dets.ctypes.data_as(ctypes.c_void_p), n, dets.shape[1], result.ctypes.data_as(ctypes.c_void_p),
                               None, wint.ctypes.data_as(ctypes.c_void_p), stats)
        self.last_stats = stats
        if rc != 0:
            _raise_last(self._L)
        return result, wint

    def decode_shot_edges(self, dets_row):
        """One shot -> (matched graph edges int64[k, 2], observable mask, integer weight)."""
        dets_row = np.ascontiguousarray(dets_row, dtype=np.uint8)
        cap = 4096
        while True:
            edges = np.zeros(cap, np.int64)
            need = ctypes.c_uint64(0)
            obs = ctypes.c_uint64(0)
            w = ctypes.c_int64(0)
            rc = self._L.cb200_decode_shot_edges(self._h, dets_row.ctypes.data_as(ctypes.c_void_p),
                                                 edges.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(need),
                                                 ctypes.byref(obs), ctypes.byref(w))
            if rc != 0:
                _raise_last(self._L)
            if need.value <= cap:
                return edges[:need.value].reshape(-1, 2), int(obs.value), int(w.value)
            cap = int(need.value)

    def set_tuning(self, group_threads=0, ht_log2=0, groups_per_sm=0):
        if self._L.cb200_set_tuning(self._h, group_threads, ht_log2) != 0:
            _raise_last(self._L)
        self._L.cb200_set_groups_per_sm(self._h, groups_per_sm)

    def profile_phases(self, enable=True):
        """Enables phase profiling for later calls; returns the cycle counters of the last profiled call as a dict."""
        buf = (ctypes.c_ulonglong * 24)()
        self._L.cb200_profile_phases(self._h, 1 if enable else 0, buf)
        names = ['ingest', 'explore_first', 'components', 'solve', 'explore_regrow', 'paths', 'lift', 'shots',
                 'blossom_le8', 'blossom_le16', 'blossom_le32', 'blossom_gt32', 'n_le8', 'n_le16', 'n_le32', 'n_gt32',
                 'solve_closed_form', 'solve_warp_blossoms', 'solve_block_blossoms', 'ball_pops', 'ball_cycles', 'balls', 'ball_nodes', 'near_queries']
        return dict(zip(names, [int(x) for x in buf]))

    def describe(self):
        p = self._L.cb200_describe(self._h)
        try:
            return ctypes.string_at(p).decode()
        finally:
            self._L.cb200_free_string(p)

    @property
    def normalising_constant(self):
        return float(self._L.cb200_normalising_constant(self._h))

    @property
    def table_bytes(self):
        return int(self._L.cb200_table_bytes(self._h))


class MultiGpuCompiledDecoder(CompiledDecoder):
    """One decoder replicated over several GPUs of the box (cb200_compile_multi): the same two predict methods, one call,
    the batch split into contiguous row ranges that the devices decode side by side. Host buffers only."""

    def __init__(self, _multi_handle, _lib_ref):
        self._m = _multi_handle
        self.num_devices = int(_lib_ref.cb200_multi_num_devices(_multi_handle))
        # the first replica answers the queries; it is owned by the multi handle (see __del__)
        CompiledDecoder.__init__(self, ctypes.c_void_p(_lib_ref.cb200_multi_replica(_multi_handle, 0)), _lib_ref)

    def __del__(self):
        try:
            if self._m:
                self._L.cb200_multi_free(self._m)
                self._m = None
                self._h = None
        except Exception:
            pass

    def _decode_host(self, dets_ptr, num_shots, stride, obs_ptr, weight_ptr, weight_int_ptr, stats):
        return self._L.cb200_multi_decode_batch(self._m, dets_ptr, num_shots, stride, obs_ptr, weight_ptr, weight_int_ptr,
                                                ctypes.byref(stats))

    def decode_device(self, *a, **k):
        raise ValueError('a multi-GPU decoder takes host buffers; use the per-device decoders for device-resident batches')

    def decode_shot_edges(self, dets_row):
        raise ValueError('decode_shot_edges is a single-GPU debugging hook')

    def set_tuning(self, group_threads=0, ht_log2=0, groups_per_sm=0):
        for k in range(self.num_devices):
            h = ctypes.c_void_p(self._L.cb200_multi_replica(self._m, k))
            if self._L.cb200_set_tuning(h, group_threads, ht_log2) != 0:
                _raise_last(self._L)
            self._L.cb200_set_groups_per_sm(h, groups_per_sm)


def _default_device():
    for key in ('CHROMOBIUS_B200_DEVICE', 'LOCAL_RANK'):
        v = os.environ.get(key)
        if v is not None and v.strip().lstrip('-').isdigit():
            return int(v)
    return 0


def _compile_text(dem_text, device=None):
    L = load_library()
    if isinstance(device, (list, tuple)) or device == 'all':
        devs = [] if device == 'all' else [int(d) for d in device]
        arr = (ctypes.c_int * len(devs))(*devs)
        handle = ctypes.c_void_p()
        rc = L.cb200_compile_multi(dem_text.encode('utf-8'), arr if devs else None, len(devs), ctypes.byref(handle))
        if rc != 0:
            if rc == -2:
                raise RuntimeError(L.cb200_last_error().decode('utf-8', 'replace'))
            _raise_last(L)
        return MultiGpuCompiledDecoder(handle, L)
    handle = ctypes.c_void_p()
    rc = L.cb200_compile(dem_text.encode('utf-8'), _default_device() if device is None else int(device), ctypes.byref(handle))
    if rc != 0:
        if rc == -2:
            raise RuntimeError(L.cb200_last_error().decode('utf-8', 'replace'))
        _raise_last(L)
    return CompiledDecoder(handle, L)


def compile_decoder_for_dem(dem, *, device=None):
    """Compiles a decoder for a stim detector error model (pybind/chromobius.pybind.cc:323-326)."""
    return CompiledDecoder.from_dem(dem, device=device)


# -- sinter adapter (pybind/sinter_compat.pybind.cc) ---------------------------------------------------------------
class _ChromobiusSinterCompiledDecoder:
    """sinter.CompiledDecoder compatible wrapper (pybind/sinter_compat.pybind.cc:23-72)."""

    def __init__(self, decoder):
        self._decoder = decoder

    def decode_shots_bit_packed(self, *, bit_packed_detection_event_data):
        d = bit_packed_detection_event_data
        if d.ndim != 2:
            raise ValueError('bit_packed_detection_event_data.ndim() != 2')
        if d.strides[1] != 1 and d.shape[1] > 1:
            raise ValueError('bit_packed_detection_event_data.strides(1) != 1')
        if d.shape[1] != self._decoder.num_detector_bytes:
            raise ValueError('bit_packed_detection_event_data.shape(1) != num_detector_bytes')
        return self._decoder.predict_obs_flips_from_dets_bit_packed(d)


class _ChromobiusSinterDecoder:
    """sinter.Decoder compatible object (pybind/sinter_compat.pybind.cc:74-147, 149-241)."""

    def __init__(self, sub_decoder=0):
        self._sub_decoder = int(sub_decoder) & 0xFF

    def __eq__(self, other):
        return isinstance(other, _ChromobiusSinterDecoder)

    def __ne__(self, other):
        return not self == other

    def __hash__(self):
        return hash('_ChromobiusSinterDecoder')

    def __getstate__(self):
        return self._sub_decoder

    def __setstate__(self, state):
        self._sub_decoder = int(state)

    def compile_decoder_for_dem(self, *, dem):
        # The sinter path skips the type check (sinter_compat.pybind.cc:134-136).
        # With CHROMOBIUS_B200_SERVICE set, the worker shares one GPU-owning process (chromobius_b200/service.py).
        from . import service
        address = service.service_address()
        if address:
            return service.RemoteCompiledDecoder(address, str(dem))
        return _ChromobiusSinterCompiledDecoder(_compile_text(str(dem)))

    def decode_via_files(self, *, num_shots, num_dets, num_obs, dem_path, dets_b8_in_path, obs_predictions_b8_out_path,
                         tmp_dir=None):
        """b8 in, b8 out through files (sinter_compat.pybind.cc:96-132), decoded as one GPU batch."""
        from . import service
        with open(str(dem_path)) as f:
            address = service.service_address()
            decoder = service.RemoteCompiledDecoder(address, f.read()) if address else _compile_text(f.read())
        nb = (int(num_dets) + 7) // 8
        ob = (int(num_obs) + 7) // 8
        raw = np.fromfile(str(dets_b8_in_path), dtype=np.uint8, count=int(num_shots) * nb)
        if raw.size != int(num_shots) * nb:
            raise ValueError('dets_b8_in_path holds fewer than num_shots records')
        dets = raw.reshape(int(num_shots), nb)
        if nb != decoder.num_detector_bytes:
            raise ValueError('num_dets does not match the detector error model')
        pred = decoder.predict_obs_flips_from_dets_bit_packed(dets)
        out = np.zeros((int(num_shots), ob), np.uint8)
        k = min(ob, pred.shape[1])
        out[:, :k] = pred[:, :k]
        if num_obs % 8:
            out[:, -1] &= (1 << (num_obs % 8)) - 1
        out.tofile(str(obs_predictions_b8_out_path))


def sinter_decoders():
    """A dictionary describing chromobius to sinter (pybind/sinter_compat.pybind.cc:224-241)."""
    return {'chromobius': _ChromobiusSinterDecoder(0)}


class DeviceDemSampler:
    """On-device DEM sampler (the stim::DemSampler role in front of the decode path; include/chromobius_b200.h)."""

    def __init__(self, dem_text, device=None):
        L = load_library()
        h = ctypes.c_void_p()
        rc = L.cb200_sampler_create(str(dem_text).encode(), _default_device() if device is None else int(device), ctypes.byref(h))
        if rc != 0:
            msg = L.cb200_sampler_last_error().decode('utf-8', 'replace')
            raise (RuntimeError if rc == -2 else ValueError)(msg)
        self._h, self._L = h, L
        self.num_detectors = int(L.cb200_sampler_num_detectors(h))
        self.num_observables = int(L.cb200_sampler_num_observables(h))
        self.num_detector_bytes = (self.num_detectors + 7) // 8
        self.expected_events_per_shot = float(L.cb200_sampler_expected_events_per_shot(h))

    def __del__(self):
        try:
            if self._h:
                self._L.cb200_sampler_free(self._h)
                self._h = None
        except Exception:
            pass

    def sample(self, num_shots, seed=0, first_shot=0):
        """-> (dets uint8[num_shots, num_detector_bytes], obs uint64[num_shots]) on the host."""
        dets = np.zeros((num_shots, self.num_detector_bytes), np.uint8)
        obs = np.zeros(num_shots, np.uint64)
        rc = self._L.cb200_sampler_sample(self._h, num_shots, seed, first_shot, dets.ctypes.data_as(ctypes.c_void_p),
                                          self.num_detector_bytes, obs.ctypes.data_as(ctypes.c_void_p), CB200_IO_HOST)
        if rc != 0:
            raise RuntimeError(self._L.cb200_sampler_last_error().decode('utf-8', 'replace'))
        return dets, obs

    def sample_device(self, num_shots, dets_ptr, row_stride, obs_ptr=None, seed=0, first_shot=0):
        """Writes shots straight into device memory (addresses as integers)."""
        rc = self._L.cb200_sampler_sample(self._h, num_shots, seed, first_shot, ctypes.c_void_p(dets_ptr), row_stride,
                                          ctypes.c_void_p(obs_ptr) if obs_ptr else None, CB200_IO_DEVICE)
        if rc != 0:
            raise RuntimeError(self._L.cb200_sampler_last_error().decode('utf-8', 'replace'))
