#!/usr/bin/env python3
"""G_shot: the graph-side bytes any exact matcher must read per shot,

    G_shot = sum over the shot's event nodes of row_bytes(node) + 8 B x (matched path edges),

from the CPU oracle's counters on a sample of DEM-sampled shots; frozen to profiles/r02_g_shot.json, which bench.py reads
for its `roofline_l2` object. row_bytes(node) = the node's adjacency row as the kernels store it (AdjEntry = 8 B; one
128-byte row per node at d25 r100, chromobius_b200/csrc/tables.h) -- each event's row read once, each path edge once.

    python tools/g_shot.py [workload] [shots]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import oracle  # noqa: E402
from dem_util import read_dem  # noqa: E402
from synth import FastDemSampler  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'midout_color_code_d25_r100_p1000'
    shots = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    text = read_dem(name)
    dec = oracle.OracleDecoder(text)
    dec.set_matcher_mode(3)
    row_ptr, nbr, wt = dec.graph()
    deg = np.diff(row_ptr.astype(np.int64))
    stride = 1
    while stride < deg.max():
        stride *= 2
    dets, _ = FastDemSampler(text).sample(shots, seed=7)
    ev_nodes, edges, row_csr = 0, 0, 0
    for row in dets:
        d = np.nonzero(np.unpackbits(row, bitorder='little')[:dec.num_detectors])[0]
        nodes = np.concatenate([2 * d, 2 * d + 1])
        ev_nodes += len(nodes)
        row_csr += int(deg[nodes].sum()) * 8
        e, _, _ = dec.decode_edges(row)
        edges += len(e)
    out = {
        'workload': name, 'shots_sampled': shots, 'event_nodes_per_shot': ev_nodes / shots,
        'matched_path_edges_per_shot': edges / shots, 'max_degree': int(deg.max()), 'ell_row_bytes': stride * 8,
        'g_shot_bytes_ell_rows': (ev_nodes * stride * 8 + 8 * edges) / shots,
        'g_shot_bytes_csr_rows': (row_csr + 8 * edges) / shots,
        'definition': 'sum over event nodes of row_bytes(node) + 8 B per matched path edge (SURVEY.md 8d); oracle mode 3',
    }
    path = os.path.join(ROOT, 'profiles', 'r02_g_shot.json')
    try:
        with open(path) as f:
            allw = json.load(f)
    except Exception:
        allw = {}
    allw[name] = out
    with open(path, 'w') as f:
        json.dump(allw, f, indent=1, sort_keys=True)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
