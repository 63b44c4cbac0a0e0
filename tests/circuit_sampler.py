"""Circuit-level Pauli-frame sampler (numpy). TEST INFRASTRUCTURE ONLY.

The reference's integration test samples detection events from the *circuit*
(/root/reference/src/chromobius/decode/decoder_integration.test.cc:24-42:
`stim::sample_batch_detection_events<64>(src_circuit, shots, rng)`), not from the DEM.  stim is not in this
image, so this file restates a Pauli-frame simulator for the gate set of the reference's `test_data/*.stim`
(SURVEY.md App. C.3):

    QUBIT_COORDS R RX M MX MPP CX TICK X_ERROR Z_ERROR DEPOLARIZE1 DEPOLARIZE2
    DETECTOR OBSERVABLE_INCLUDE SHIFT_COORDS REPEAT      (+ the (p) result-flip argument of M / MX / MPP)

It is independent of tools/circuit_to_dem.py (forward frames per shot here, backward sensitivity sets there), of
the oracle and of the product, so shots drawn here together with a DEM made there close no loop: a wrong analyser
shows up as a wrong logical error rate against the reference's pins.

Frame semantics (x/z = "the shot differs from the noiseless run by an X / Z on this qubit"):
    CX c t : x[t] ^= x[c] ; z[c] ^= z[t]          R : x = z = 0        RX : x = z = 0
    M q -> x[q]      MX q -> z[q]      MPP -> XOR over factors (X: z, Z: x, Y: x^z)      (p): result ^= Bernoulli(p)
    X_ERROR(p): x ^= B(p)    Z_ERROR(p): z ^= B(p)
    DEPOLARIZE1(p): with probability p one of X, Y, Z uniformly; DEPOLARIZE2(p): one of the 15 two-qubit Paulis
The random Z (X) kick stim applies after M / R (MX / RX) only matters for non-deterministic measurements; every
detector and observable of these circuits is deterministic, so it is left out.
"""
import re

import numpy as np


def parse_circuit(text):
    lines = text.split('\n')
    pos = 0

    def parse_block():
        nonlocal pos
        out = []
        while pos < len(lines):
            line = lines[pos].split('#')[0].strip()
            pos += 1
            if not line:
                continue
            if line == '}':
                return out
            m = re.match(r'^REPEAT\s+(\d+)\s*\{$', line)
            if m:
                out.append(('REPEAT', int(m.group(1)), parse_block()))
                continue
            m = re.match(r'^([A-Z_0-9]+)(?:\(([^)]*)\))?\s*(.*)$', line)
            args = tuple(float(a) for a in m.group(2).split(',')) if m.group(2) else ()
            out.append((m.group(1), args, m.group(3).split()))
        return out

    return parse_block()


def _num_qubits(block):
    n = 0
    for ins in block:
        if ins[0] == 'REPEAT':
            n = max(n, _num_qubits(ins[2]))
        elif ins[0] in ('DETECTOR', 'OBSERVABLE_INCLUDE', 'SHIFT_COORDS', 'TICK'):
            continue
        elif ins[0] == 'MPP':
            for t in ins[2]:
                for f in t.split('*'):
                    n = max(n, int(f[1:]) + 1)
        else:
            for t in ins[2]:
                n = max(n, int(t) + 1)
    return n


class _Sim:
    def __init__(self, nq, shots, rng):
        self.x = np.zeros((nq, shots), np.uint8)
        self.z = np.zeros((nq, shots), np.uint8)
        self.shots = shots
        self.rng = rng
        self.rec = []
        self.dets = []
        self.obs = {}

    def bern(self, n, p):
        """uint8[n, shots] of independent Bernoulli(p) bits, drawn sparsely (p is small)."""
        out = np.zeros(n * self.shots, np.uint8)
        if p <= 0:
            return out.reshape(n, self.shots)
        if p > 0.2:
            return (self.rng.random((n, self.shots)) < p).astype(np.uint8)
        k = self.rng.binomial(n * self.shots, p)
        if k:
            out[self.rng.choice(n * self.shots, size=k, replace=False)] = 1
        return out.reshape(n, self.shots)

    def run(self, block):
        x, z = self.x, self.z
        for ins in block:
            name = ins[0]
            if name == 'REPEAT':
                for _ in range(ins[1]):
                    self.run(ins[2])
                continue
            args, targets = ins[1], ins[2]
            if name == 'CX':
                q = np.array([int(t) for t in targets], np.int64)
                c, t = q[0::2], q[1::2]
                if len(set(q.tolist())) == len(q):
                    x[t] ^= x[c]
                    z[c] ^= z[t]
                else:
                    for ci, ti in zip(c, t):
                        x[ti] ^= x[ci]
                        z[ci] ^= z[ti]
            elif name in ('R', 'RX'):
                q = [int(t) for t in targets]
                x[q] = 0
                z[q] = 0
            elif name in ('M', 'MX'):
                src = x if name == 'M' else z
                q = [int(t) for t in targets]
                res = src[q].copy()
                if args and args[0] > 0:
                    res ^= self.bern(len(q), args[0])
                self.rec.extend(res)
            elif name == 'MPP':
                flips = self.bern(len(targets), args[0]) if args and args[0] > 0 else None
                for i, t in enumerate(targets):
                    r = np.zeros(self.shots, np.uint8)
                    for f in t.split('*'):
                        qb = int(f[1:])
                        if f[0] in 'XY':
                            r ^= z[qb]
                        if f[0] in 'ZY':
                            r ^= x[qb]
                    if flips is not None:
                        r ^= flips[i]
                    self.rec.append(r)
            elif name == 'X_ERROR':
                q = [int(t) for t in targets]
                x[q] ^= self.bern(len(q), args[0])
            elif name == 'Z_ERROR':
                q = [int(t) for t in targets]
                z[q] ^= self.bern(len(q), args[0])
            elif name == 'DEPOLARIZE1':
                q = np.array([int(t) for t in targets], np.int64)
                hit = self.bern(len(q), args[0])
                qi, si = np.nonzero(hit)
                kind = self.rng.integers(1, 4, size=len(qi))  # 1 = X, 2 = Y, 3 = Z
                np.bitwise_xor.at(x, (q[qi], si), (kind != 3).astype(np.uint8))
                np.bitwise_xor.at(z, (q[qi], si), (kind != 1).astype(np.uint8))
            elif name == 'DEPOLARIZE2':
                q = np.array([int(t) for t in targets], np.int64)
                a, b = q[0::2], q[1::2]
                hit = self.bern(len(a), args[0])
                pi, si = np.nonzero(hit)
                kind = self.rng.integers(1, 16, size=len(pi))  # bits: xa, za, xb, zb
                np.bitwise_xor.at(x, (a[pi], si), (kind & 1).astype(np.uint8))
                np.bitwise_xor.at(z, (a[pi], si), ((kind >> 1) & 1).astype(np.uint8))
                np.bitwise_xor.at(x, (b[pi], si), ((kind >> 2) & 1).astype(np.uint8))
                np.bitwise_xor.at(z, (b[pi], si), ((kind >> 3) & 1).astype(np.uint8))
            elif name == 'DETECTOR':
                r = np.zeros(self.shots, np.uint8)
                for t in targets:
                    r ^= self.rec[int(t[4:-1])]
                self.dets.append(r)
            elif name == 'OBSERVABLE_INCLUDE':
                k = int(args[0])
                r = self.obs.get(k)
                if r is None:
                    r = np.zeros(self.shots, np.uint8)
                for t in targets:
                    r = r ^ self.rec[int(t[4:-1])]
                self.obs[k] = r
            elif name in ('TICK', 'QUBIT_COORDS', 'SHIFT_COORDS'):
                pass
            else:
                raise ValueError('unsupported instruction ' + name)


def sample_circuit(text, shots, seed):
    """Returns (dets uint8[shots, ceil(D/8)] bit-packed little-endian, obs uint64[shots], num_dets, num_obs)."""
    tree = parse_circuit(text)
    sim = _Sim(_num_qubits(tree), shots, np.random.default_rng(seed))
    sim.run(tree)
    nd = len(sim.dets)
    no = (max(sim.obs) + 1) if sim.obs else 0
    bits = np.zeros((shots, ((nd + 7) // 8) * 8), np.uint8)
    if nd:
        bits[:, :nd] = np.stack(sim.dets, axis=1)
    dets = np.packbits(bits, axis=1, bitorder='little')
    obs = np.zeros(shots, np.uint64)
    for k, r in sim.obs.items():
        obs |= r.astype(np.uint64) << np.uint64(k)
    return dets, obs, nd, no
