"""ctypes wrapper around the CPU oracle (oracle/chromobius_oracle.cc).

TEST INFRASTRUCTURE ONLY: may be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs, never from the product package `chromobius_b200`.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libchromobius_oracle.so')
_lib = None


def build(force=False):
    src = os.path.join(_HERE, 'chromobius_oracle.cc')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        vp, u64, i64 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int64
        L.orc_last_error.restype = ctypes.c_char_p
        L.orc_compile.restype = vp
        L.orc_compile.argtypes = [ctypes.c_char_p]
        L.orc_clone.restype = vp
        L.orc_clone.argtypes = [vp]
        L.orc_free.argtypes = [vp]
        for name in ['orc_num_detectors', 'orc_num_observables', 'orc_drag_ambiguities', 'orc_radius_retries',
                     'orc_graph_num_half_edges', 'orc_drag_size', 'orc_num_atomic_errors', 'orc_num_mobius_errors']:
            getattr(L, name).restype = u64
            getattr(L, name).argtypes = [vp]
        L.orc_normalising_constant.restype = ctypes.c_double
        L.orc_normalising_constant.argtypes = [vp]
        L.orc_describe.restype = ctypes.c_void_p
        L.orc_describe.argtypes = [vp]
        L.orc_free_str.argtypes = [ctypes.c_void_p]
        L.orc_graph_export.argtypes = [vp, vp, vp, vp]
        L.orc_node_colors.argtypes = [vp, vp]
        L.orc_rgb_reps.argtypes = [vp, vp]
        L.orc_drag_export.argtypes = [vp, vp, vp]
        L.orc_decode_batch.restype = i64
        L.orc_decode_batch.argtypes = [vp, vp, u64, u64, vp, vp]
        L.orc_decode_edges.restype = i64
        L.orc_decode_edges.argtypes = [vp, vp, vp, u64, vp, vp]
        L.orc_euler_tours.restype = i64
        L.orc_euler_tours.argtypes = [u64, vp, u64, vp, u64]
        L.orc_lift_edges.restype = i64
        L.orc_lift_edges.argtypes = [vp, vp, vp, u64, vp]
        L.orc_set_matcher_mode.argtypes = [vp, ctypes.c_int]
        L.orc_matcher_stats.argtypes = [vp, vp, ctypes.c_int]
        L.orc_set_odd_policy.argtypes = [vp, ctypes.c_int]
        L.orc_set_growth_extra.argtypes = [vp, ctypes.c_int]
        L.orc_match_events.restype = i64
        L.orc_match_events.argtypes = [vp, vp, u64, vp, vp]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleDecoder:
    """Restated CPU chromobius decoder (see chromobius_oracle.cc)."""

    def __init__(self, dem_text, _handle=None):
        L = lib()
        if _handle is None:
            _handle = L.orc_compile(str(dem_text).encode())
            if not _handle:
                raise ValueError(L.orc_last_error().decode())
        self._h = _handle
        self.num_detectors = L.orc_num_detectors(self._h)
        self.num_observables = L.orc_num_observables(self._h)
        self.num_detector_bytes = (self.num_detectors + 7) // 8
        self.num_observable_bytes = (self.num_observables + 7) // 8
        self.normalising_constant = L.orc_normalising_constant(self._h)

    def set_matcher_mode(self, mode):
        """0: dense all-pairs matcher up to 160 events, certified-local (lazy merging) beyond; 1: certified-local with eager
        merging; 2: always dense; 3: always certified-local with lazy merging (the algorithm the kernels run)."""
        lib().orc_set_matcher_mode(self._h, int(mode))

    def set_growth_extra(self, extra):
        """Events an open-ended growth request of the certified-local matcher asks for beyond the one it needs. 1 (default)
        is what the kernels do (fewer matcher rounds); 0 is faster on a CPU. Weights are exact either way; among
        equal-weight matchings the choice can differ."""
        lib().orc_set_growth_extra(self._h, int(extra))

    def set_odd_policy(self, policy):
        lib().orc_set_odd_policy(self._h, int(policy))

    def matcher_stats(self, reset=True):
        out = np.zeros(7, np.uint64)
        lib().orc_matcher_stats(self._h, _ptr(out), 1 if reset else 0)
        return dict(blossoms_le8=int(out[0]), blossoms_le16=int(out[1]), blossoms_le32=int(out[2]), blossoms_gt32=int(out[3]),
                    rounds=int(out[4]), closed_form_pairs=int(out[5]), max_component=int(out[6]))

    def clone(self):
        return OracleDecoder(None, _handle=lib().orc_clone(self._h))

    def __del__(self):
        try:
            if self._h:
                lib().orc_free(self._h)
                self._h = None
        except Exception:
            pass

    def describe(self):
        L = lib()
        p = L.orc_describe(self._h)
        s = ctypes.string_at(p).decode()
        L.orc_free_str(p)
        return s

    def graph(self):
        L = lib()
        n = self.num_detectors * 2
        m = L.orc_graph_num_half_edges(self._h)
        row_ptr = np.zeros(n + 1, np.uint32)
        nbr = np.zeros(m, np.uint32)
        wt = np.zeros(m, np.int32)
        L.orc_graph_export(self._h, _ptr(row_ptr), _ptr(nbr), _ptr(wt))
        return row_ptr, nbr, wt

    def node_colors(self):
        out = np.zeros((self.num_detectors, 3), np.uint8)
        lib().orc_node_colors(self._h, _ptr(out))
        return out

    def rgb_reps(self):
        out = np.zeros((self.num_detectors, 6), np.uint32)
        lib().orc_rgb_reps(self._h, _ptr(out))
        return out

    def drag(self):
        L = lib()
        n = L.orc_drag_size(self._h)
        keys = np.zeros((n, 4), np.uint32)
        vals = np.zeros(n, np.uint64)
        L.orc_drag_export(self._h, _ptr(keys), _ptr(vals))
        return keys, vals

    def stats(self):
        L = lib()
        return dict(atomic_errors=L.orc_num_atomic_errors(self._h), mobius_errors=L.orc_num_mobius_errors(self._h),
                    drag_entries=L.orc_drag_size(self._h), drag_ambiguities=L.orc_drag_ambiguities(self._h),
                    half_edges=L.orc_graph_num_half_edges(self._h), radius_retries=L.orc_radius_retries(self._h))

    def decode_batch(self, dets):
        """dets: uint8[shots, num_detector_bytes] -> (obs_masks uint64[shots], weights int64[shots])."""
        dets = np.ascontiguousarray(dets, dtype=np.uint8)
        if dets.ndim == 1:
            dets = dets[None, :]
        assert dets.shape[1] == self.num_detector_bytes
        n = dets.shape[0]
        obs = np.zeros(n, np.uint64)
        wt = np.zeros(n, np.int64)
        r = lib().orc_decode_batch(self._h, _ptr(dets), n, dets.shape[1], _ptr(obs), _ptr(wt))
        if r != 0:
            raise ValueError('shot %d: %s' % (-r - 1, lib().orc_last_error().decode()))
        return obs, wt

    def predict_obs_flips_from_dets_bit_packed(self, dets):
        obs, _ = self.decode_batch(dets)
        out = np.zeros((len(obs), self.num_observable_bytes), np.uint8)
        for k in range(self.num_observable_bytes):
            out[:, k] = (obs >> np.uint64(8 * k)) & np.uint64(0xFF)
        return out

    def decode_edges(self, dets_row):
        dets_row = np.ascontiguousarray(dets_row, dtype=np.uint8)
        cap = 1 << 16
        while True:
            edges = np.zeros(cap, np.int64)
            obs = np.zeros(1, np.uint64)
            wt = np.zeros(1, np.int64)
            r = lib().orc_decode_edges(self._h, _ptr(dets_row), _ptr(edges), cap, _ptr(obs), _ptr(wt))
            if r < 0:
                raise ValueError(lib().orc_last_error().decode())
            if r <= cap:
                return edges[:r].reshape(-1, 2), int(obs[0]), int(wt[0])
            cap = int(r)

    def lift_edges(self, dets_row, edges):
        dets_row = np.ascontiguousarray(dets_row, dtype=np.uint8)
        edges = np.ascontiguousarray(edges, dtype=np.int64).reshape(-1)
        obs = np.zeros(1, np.uint64)
        r = lib().orc_lift_edges(self._h, _ptr(dets_row), _ptr(edges), len(edges), _ptr(obs))
        if r != 0:
            raise ValueError('lift failed')
        return int(obs[0])

    def match_events(self, events):
        events = np.ascontiguousarray(events, dtype=np.uint64)
        partner = np.full(len(events), -1, np.int32)
        wt = np.zeros(1, np.int64)
        r = lib().orc_match_events(self._h, _ptr(events), len(events), _ptr(partner), _ptr(wt))
        if r != 0:
            raise ValueError(lib().orc_last_error().decode())
        return partner, int(wt[0])


def euler_tours(num_nodes, edges):
    edges = np.ascontiguousarray(edges, dtype=np.int64).reshape(-1)
    out = np.zeros(4 * len(edges) + 16, np.int64)
    r = lib().orc_euler_tours(num_nodes, _ptr(edges), len(edges), _ptr(out), len(out))
    if r < 0:
        raise ValueError(lib().orc_last_error().decode())
    cycles = []
    k = 0
    while k < r:
        n = int(out[k])
        cycles.append([int(v) for v in out[k + 1:k + 1 + n]])
        k += 1 + n
    return cycles
