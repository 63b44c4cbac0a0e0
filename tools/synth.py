"""Synthetic syndrome generation for bench.py and the tests: a vectorised numpy DEM sampler.

Every error mechanism of the flattened DEM fires independently with its probability; a shot is the XOR of the
detector sets (and observable masks) of the mechanisms that fired -- the distribution stim's DemSampler draws from
(/root/reference/src/chromobius/decode/decoder_integration.test.cc:55-68 uses it). stim itself is not in this image.
Bench / test tooling, not product code.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
from dem_util import flatten_dem  # noqa: E402


class FastDemSampler:
    def __init__(self, dem_text):
        self.num_dets, self.num_obs, errors, _ = flatten_dem(dem_text)
        self.nbytes = (self.num_dets + 7) // 8
        p = []
        ptr = [0]
        flat = []
        obs = []
        for prob, groups in errors:
            ds = {}
            om = 0
            for dets, ob in groups:
                for d in dets:
                    ds[d] = ds.get(d, 0) ^ 1
                for o in ob:
                    om ^= 1 << o
            flat.extend(sorted(d for d, v in ds.items() if v))
            ptr.append(len(flat))
            obs.append(om)
            p.append(prob)
        self.p = np.array(p, np.float64)
        self.ptr = np.array(ptr, np.int64)
        self.flat = np.array(flat, np.int64)
        self.obs = np.array(obs, np.uint64)
        self.expected_events = float((self.p * np.diff(self.ptr)).sum())

    def sample(self, shots, seed, chunk=1 << 16):
        """Returns (dets uint8[shots, nbytes] bit packed little-endian, obs uint64[shots])."""
        rng = np.random.default_rng(seed)
        out = np.zeros((shots, self.nbytes), np.uint8)
        obs = np.zeros(shots, np.uint64)
        nbits = self.nbytes * 8
        uniq_p = np.unique(self.p)
        groups = [(pp, np.nonzero(self.p == pp)[0]) for pp in uniq_p]
        for s0 in range(0, shots, chunk):
            n = min(chunk, shots - s0)
            shot_idx = []
            err_idx = []
            for pp, idx in groups:
                total = len(idx) * n
                k = rng.binomial(total, pp)
                if k == 0:
                    continue
                pos = rng.choice(total, size=k, replace=False)
                err_idx.append(idx[pos % len(idx)])
                shot_idx.append(pos // len(idx))
            if not err_idx:
                continue
            err = np.concatenate(err_idx)
            sh = np.concatenate(shot_idx)
            # observables
            o = self.obs[err]
            nz = o != 0
            if nz.any():
                np.bitwise_xor.at(obs, s0 + sh[nz], o[nz])
            # detectors: expand (shot, error) -> (shot, det) and keep the odd-parity ones
            cnt = self.ptr[err + 1] - self.ptr[err]
            starts = np.repeat(self.ptr[err], cnt)
            within = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
            det = self.flat[starts + within]
            key = np.repeat(sh, cnt) * nbits + det
            u, c = np.unique(key, return_counts=True)
            u = u[(c & 1) == 1]
            bits = np.zeros(n * nbits, np.uint8)
            bits[u] = 1
            out[s0:s0 + n] = np.packbits(bits.reshape(n, nbits), axis=1, bitorder='little')
        return out, obs
