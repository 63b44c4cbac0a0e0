#!/usr/bin/env python3
"""Generates tests/golden/circuit_shots/<circuit>.npz: 8192 circuit-level (Pauli-frame) samples of every circuit the
reference's integration test decodes (/root/reference/src/chromobius/decode/decoder_integration.test.cc:139-163), with the
true observable flips. Needs /root/reference/test_data (this container only); the committed files travel to the GPU box.

    dets: uint8[8192, ceil(D/8)] bit packed (little-endian bit order), obs: uint64[8192], seed 0 of tests/circuit_sampler.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from circuit_sampler import sample_circuit  # noqa: E402
from reference_pins import PINS  # noqa: E402

SHOTS = 8192


def main():
    out_dir = os.path.join(ROOT, 'tests', 'golden', 'circuit_shots')
    os.makedirs(out_dir, exist_ok=True)
    for name, _ in PINS:
        with open('/root/reference/test_data/%s.stim' % name) as f:
            dets, obs, nd, no = sample_circuit(f.read(), SHOTS, seed=0)
        np.savez_compressed(os.path.join(out_dir, name + '.npz'), dets=dets, obs=obs, num_dets=nd, num_obs=no)
        print(name, dets.shape, 'events/shot %.2f' % (np.unpackbits(dets, axis=1).sum() / SHOTS))


if __name__ == '__main__':
    main()
