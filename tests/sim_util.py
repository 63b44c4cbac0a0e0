"""ctypes wrapper of the host simulation build (tests/sim/host_sim.cc). Test infrastructure only."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        import __graft_entry__ as g
        L = ctypes.CDLL(g.build_sim())
        vp, u64 = ctypes.c_void_p, ctypes.c_uint64
        L.cbsim_compile.restype = vp
        L.cbsim_compile.argtypes = [ctypes.c_char_p]
        L.cbsim_free.argtypes = [vp]
        L.cbsim_last_error.restype = ctypes.c_char_p
        for n in ['cbsim_num_detectors', 'cbsim_num_observables', 'cbsim_graph_half_edges']:
            getattr(L, n).restype = u64
            getattr(L, n).argtypes = [vp]
        L.cbsim_normalising_constant.restype = ctypes.c_double
        L.cbsim_normalising_constant.argtypes = [vp]
        for n in ['cbsim_describe', 'cbsim_mobius_dem']:
            getattr(L, n).restype = vp
            getattr(L, n).argtypes = [vp]
        L.cbsim_free_string.argtypes = [vp]
        L.cbsim_graph_export.argtypes = [vp, vp, vp, vp]
        L.cbsim_decode.argtypes = [vp, vp, u64, u64, vp, vp, vp, vp, ctypes.c_uint32, ctypes.c_uint32, u64, vp, u64, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class SimDecoder:
    def __init__(self, dem_text):
        L = lib()
        self._h = L.cbsim_compile(str(dem_text).encode())
        if not self._h:
            raise ValueError(L.cbsim_last_error().decode())
        self.num_detectors = L.cbsim_num_detectors(self._h)
        self.num_observables = L.cbsim_num_observables(self._h)
        self.num_detector_bytes = (self.num_detectors + 7) // 8
        self.normalising_constant = L.cbsim_normalising_constant(self._h)

    def __del__(self):
        try:
            lib().cbsim_free(self._h)
        except Exception:
            pass

    def _str(self, fn):
        L = lib()
        p = fn(self._h)
        s = ctypes.string_at(p).decode()
        L.cbsim_free_string(p)
        return s

    def describe(self):
        return self._str(lib().cbsim_describe)

    def mobius_dem(self):
        return self._str(lib().cbsim_mobius_dem)

    def graph(self):
        L = lib()
        m = L.cbsim_graph_half_edges(self._h)
        row_ptr = np.zeros(2 * self.num_detectors + 1, np.uint32)
        nbr = np.zeros(m, np.uint32)
        wt = np.zeros(m, np.int32)
        L.cbsim_graph_export(self._h, _p(row_ptr), _p(nbr), _p(wt))
        return row_ptr, nbr, wt

    def decode(self, dets, ht_log2=12, dyn_bytes=64 << 20, want_edges=False, sh_ht_log2=5):
        dets = np.ascontiguousarray(dets, dtype=np.uint8)
        if dets.ndim == 1:
            dets = dets[None, :]
        n = dets.shape[0]
        obs = np.zeros(n, np.uint64)
        w = np.zeros(n, np.int64)
        st = np.zeros(n, np.uint32)
        rd = np.zeros(n, np.uint32)
        edges = np.zeros(1 << 16, np.int64) if want_edges else None
        ne = ctypes.c_uint64(0)
        lib().cbsim_decode(self._h, _p(dets), n, dets.shape[1], _p(obs), _p(w), _p(st), _p(rd), ht_log2, sh_ht_log2, dyn_bytes,
                           _p(edges) if want_edges else None, len(edges) if want_edges else 0,
                           ctypes.byref(ne) if want_edges else None)
        if want_edges:
            return obs, w, st, rd, edges[:ne.value].reshape(-1, 2)
        return obs, w, st, rd
