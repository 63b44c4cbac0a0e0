"""Test helpers: a tiny independent DEM text parser/flattener and a numpy DEM sampler.

Independent of both the oracle and the product so it can arbitrate between them.
"""
import os
import re

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
DEM_DIR = os.path.join(GOLDEN_DIR, 'dems')


def dem_path(name):
    return os.path.join(DEM_DIR, name + '.dem')


def read_dem(name):
    with open(dem_path(name)) as f:
        return f.read()


def flatten_dem(text):
    """Returns (num_dets, num_obs, errors, detectors) with errors = list of (p, [groups]) where each group is
    (dets tuple, obs tuple) ('^'-separated), and detectors = {abs id: coords tuple}."""
    lines = [l.split('#')[0].strip() for l in text.split('\n')]
    lines = [l for l in lines if l]
    pos = 0

    def parse_block():
        nonlocal pos
        out = []
        while pos < len(lines):
            line = lines[pos]
            pos += 1
            if line == '}':
                return out
            m = re.match(r'^repeat\s+(\d+)\s*\{$', line, re.I)
            if m:
                out.append(('repeat', int(m.group(1)), parse_block()))
                continue
            m = re.match(r'^([a-zA-Z_]+)(?:\(([^)]*)\))?\s*(.*)$', line)
            name = m.group(1).lower()
            args = tuple(float(a) for a in m.group(2).split(',')) if m.group(2) else ()
            out.append((name, args, m.group(3).split()))
        return out

    tree = parse_block()
    errors = []
    detectors = {}
    state = {'off': 0, 'coords': [], 'maxd': -1, 'maxo': -1}

    def run(block):
        for ins in block:
            if ins[0] == 'repeat':
                for _ in range(ins[1]):
                    run(ins[2])
            elif ins[0] == 'error':
                groups = []
                dets, obs = [], []
                for t in ins[2]:
                    if t == '^':
                        groups.append((tuple(dets), tuple(obs)))
                        dets, obs = [], []
                    elif t[0] in 'Dd':
                        d = int(t[1:]) + state['off']
                        dets.append(d)
                        state['maxd'] = max(state['maxd'], d)
                    elif t[0] in 'Ll':
                        obs.append(int(t[1:]))
                        state['maxo'] = max(state['maxo'], int(t[1:]))
                groups.append((tuple(dets), tuple(obs)))
                errors.append((ins[1][0], groups))
            elif ins[0] == 'detector':
                c = list(ins[1])
                for i in range(min(len(c), len(state['coords']))):
                    c[i] += state['coords'][i]
                for t in ins[2]:
                    d = int(t[1:]) + state['off']
                    detectors[d] = tuple(c)
                    state['maxd'] = max(state['maxd'], d)
            elif ins[0] == 'shift_detectors':
                if ins[2]:
                    state['off'] += int(ins[2][0])
                while len(state['coords']) < len(ins[1]):
                    state['coords'].append(0.0)
                for i, a in enumerate(ins[1]):
                    state['coords'][i] += a
            elif ins[0] == 'logical_observable':
                for t in ins[2]:
                    state['maxo'] = max(state['maxo'], int(t[1:]))
            else:
                raise ValueError('bad instruction ' + ins[0])

    run(tree)
    return state['maxd'] + 1, state['maxo'] + 1, errors, detectors


class NumpyDemSampler:
    """Independent-error DEM sampler (each error fires with its probability; XOR symptoms)."""

    def __init__(self, dem_text):
        self.num_dets, self.num_obs, errors, _ = flatten_dem(dem_text)
        self.p = np.array([e[0] for e in errors], np.float64)
        self.dets = []
        self.obs = []
        for p, groups in errors:
            ds = {}
            om = 0
            for dets, obs in groups:
                for d in dets:
                    ds[d] = ds.get(d, 0) ^ 1
                for o in obs:
                    om ^= 1 << o
            self.dets.append(np.array(sorted(d for d, v in ds.items() if v), np.int64))
            self.obs.append(om)
        self.obs = np.array(self.obs, np.uint64)
        self.nbytes = (self.num_dets + 7) // 8

    def sample(self, shots, seed):
        """Returns (dets uint8[shots, nbytes] bit-packed little-endian, obs uint64[shots])."""
        rng = np.random.default_rng(seed)
        bits = np.zeros((shots, self.nbytes * 8), np.uint8)
        obs = np.zeros(shots, np.uint64)
        # group errors by probability and use binomial counts + uniform positions
        for p in np.unique(self.p):
            idx = np.nonzero(self.p == p)[0]
            total = len(idx) * shots
            k = rng.binomial(total, p)
            if k == 0:
                continue
            pos = rng.choice(total, size=k, replace=False) if total < (1 << 62) else None
            err = idx[pos % len(idx)]
            shot = pos // len(idx)
            for e, s in zip(err, shot):
                bits[s, self.dets[e]] ^= 1
                obs[s] ^= self.obs[e]
        packed = np.packbits(bits, axis=1, bitorder='little')
        return packed, obs

    def single_error_shot(self, k):
        bits = np.zeros(self.nbytes * 8, np.uint8)
        bits[self.dets[k]] ^= 1
        return np.packbits(bits, bitorder='little'), int(self.obs[k])
