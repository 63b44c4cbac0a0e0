#!/usr/bin/env python3
"""Prints the reference-pin table of DESIGN.md section 6: for the 13 circuits of decoder_integration.test.cc:144-156,
mistakes per 8192 circuit-sampled shots of the oracle with the independent dense matcher (mode 2) and with the kernels'
certified-local matcher (mode 3; the CUDA path is bit-exact against it, tests/test_gpu_reference_pins.py)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle  # noqa: E402
from dem_util import read_dem  # noqa: E402
from reference_pins import PINS, count_mistakes, load_circuit_shots, within_band  # noqa: E402

print('| circuit | events/shot | reference pin / 8192 | oracle, dense matcher | oracle, kernel matcher = CUDA path | weight diffs | in band |')
print('|---|---|---|---|---|---|---|')
for name, pin in PINS:
    dets, truth = load_circuit_shots(name)
    text = read_dem(name)
    res = {}
    for mode in (2, 3):
        dec = oracle.OracleDecoder(text)
        dec.set_matcher_mode(mode)
        obs, w = dec.decode_batch(dets)
        res[mode] = (count_mistakes(obs, truth, dec.num_observables), w)
    import numpy as np
    ev = np.unpackbits(dets, axis=1).sum() / len(dets)
    print('| %s | %.2f | %d | %d | %d | %d | %s |' % (name, ev, pin, res[2][0], res[3][0], int((res[2][1] != res[3][1]).sum()),
                                                   'yes' if within_band(res[3][0], len(dets), pin) else '**no**'), flush=True)
