"""Forward single-fault enumeration of a stim circuit with the frame simulator of tests/circuit_sampler.py.
TEST INFRASTRUCTURE ONLY.

Every noise location x every Pauli component is injected alone, in its own shot of an otherwise noiseless run; the
(detector set, observable set) it produces, with the component's probability, is one error mechanism. Mechanisms with
equal symptoms combine as p <- p1 (1 - p2) + p2 (1 - p1). This is what stim's ErrorAnalyzer computes backwards
(/root/reference/src/chromobius/decode/decoder_integration.test.cc:26-27 feeds its output to the decoder), computed
forwards and independently of tools/circuit_to_dem.py, which wrote the fixture DEMs under tests/golden/dems.
"""
import math

import numpy as np

import circuit_sampler as cs


def count_faults(block):
    n = 0
    for ins in block:
        if ins[0] == 'REPEAT':
            n += ins[1] * count_faults(ins[2])
        elif ins[0] in ('X_ERROR', 'Z_ERROR'):
            n += len(ins[2])
        elif ins[0] == 'DEPOLARIZE1':
            n += 3 * len(ins[2])
        elif ins[0] == 'DEPOLARIZE2':
            n += 15 * (len(ins[2]) // 2)
        elif ins[0] in ('M', 'MX', 'MPP') and ins[1] and ins[1][0] > 0:
            n += len(ins[2])
    return n

class FaultSim(cs._Sim):
    def __init__(self, nq, lo, hi):
        super().__init__(nq, hi - lo, None)
        self.lo, self.hi = lo, hi
        self.next_fault = 0
        self.q_of_fault = {}
    def take(self, k, q):
        """allocate k fault ids; returns list of (local shot or None)."""
        out = []
        for i in range(k):
            f = self.next_fault + i
            if self.lo <= f < self.hi:
                out.append(f - self.lo)
                self.q_of_fault[f] = q
            else:
                out.append(None)
        self.next_fault += k
        return out
    def run(self, block):
        x, z = self.x, self.z
        for ins in block:
            name = ins[0]
            if name == 'REPEAT':
                for _ in range(ins[1]):
                    self.run(ins[2])
                continue
            args, targets = ins[1], ins[2]
            if name in ('X_ERROR', 'Z_ERROR'):
                arr = x if name == 'X_ERROR' else z
                for t, s in zip(targets, self.take(len(targets), args[0])):
                    if s is not None: arr[int(t), s] ^= 1
            elif name == 'DEPOLARIZE1':
                q = 0.5 - 0.5 * math.sqrt(1 - 4 * args[0] / 3)
                for t in targets:
                    sl = self.take(3, q)
                    for kind, s in zip((1, 2, 3), sl):
                        if s is None: continue
                        if kind != 3: x[int(t), s] ^= 1
                        if kind != 1: z[int(t), s] ^= 1
            elif name == 'DEPOLARIZE2':
                q = 0.5 - 0.5 * (1 - 16 * args[0] / 15) ** 0.125
                for i in range(0, len(targets), 2):
                    a, b = int(targets[i]), int(targets[i + 1])
                    sl = self.take(15, q)
                    for kind, s in zip(range(1, 16), sl):
                        if s is None: continue
                        if kind & 1: x[a, s] ^= 1
                        if kind & 2: z[a, s] ^= 1
                        if kind & 4: x[b, s] ^= 1
                        if kind & 8: z[b, s] ^= 1
            elif name in ('M', 'MX'):
                src = x if name == 'M' else z
                qs = [int(t) for t in targets]
                res = src[qs].copy()
                if args and args[0] > 0:
                    for i, s in enumerate(self.take(len(qs), args[0])):
                        if s is not None: res[i, s] ^= 1
                self.rec.extend(res)
            elif name == 'MPP':
                sl = self.take(len(targets), args[0]) if args and args[0] > 0 else [None] * len(targets)
                for i, t in enumerate(targets):
                    r = np.zeros(self.shots, np.uint8)
                    for f in t.split('*'):
                        qb = int(f[1:])
                        if f[0] in 'XY': r ^= z[qb]
                        if f[0] in 'ZY': r ^= x[qb]
                    if sl[i] is not None: r[sl[i]] ^= 1
                    self.rec.append(r)
            else:
                super().run([ins])

def enumerate_faults(text, chunk=20000):
    tree = cs.parse_circuit(text)
    nq = cs._num_qubits(tree)
    total = count_faults(tree)
    mech = {}
    for lo in range(0, total, chunk):
        hi = min(total, lo + chunk)
        sim = FaultSim(nq, lo, hi)
        sim.run(tree)
        dets = np.stack(sim.dets, axis=1) if sim.dets else np.zeros((hi - lo, 0), np.uint8)
        for s in range(hi - lo):
            ds = tuple(np.nonzero(dets[s])[0].tolist())
            ob = tuple(k for k, r in sorted(sim.obs.items()) if r[s])
            if not ds and not ob: continue
            q = sim.q_of_fault[lo + s]
            key = (ds, ob)
            old = mech.get(key, 0.0)
            mech[key] = old * (1 - q) + q * (1 - old)
    return mech, total


def dem_mechanisms(flat_errors):
    """{(dets, obs): p} of a flattened DEM (tests/dem_util.flatten_dem errors), '^' groups XOR-ed together."""
    dem = {}
    for p, groups in flat_errors:
        ds, ob = set(), set()
        for d, o in groups:
            ds ^= set(d)
            ob ^= set(o)
        key = (tuple(sorted(ds)), tuple(sorted(ob)))
        old = dem.get(key, 0.0)
        dem[key] = old * (1 - p) + p * (1 - old)
    return dem
