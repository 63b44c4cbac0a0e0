#!/usr/bin/env python3
"""e2e of ONE Python call on N GPUs of the box: CompiledDecoder.predict_obs_flips_from_dets_bit_packed of a decoder
compiled with device=[0..N-1] (cb200_compile_multi: contiguous row ranges, one host thread and stream per device).
Host (pinned) input of N x 1.25 M shots of the headline workload -- at N = 8 exactly BASELINE.json's 10 M-shot configuration.

    python tools/multi_api_bench.py N [shots per GPU] [steps]
Prints one JSON line. Not the driver's bench (that is bench.py under torchrun); kept under profiles/ as evidence.
"""
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import chromobius_b200 as cb  # noqa: E402
from dem_util import read_dem  # noqa: E402


def main():
    n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    per_gpu = int(sys.argv[2]) if len(sys.argv) > 2 else 1250000
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    name = 'midout_color_code_d25_r100_p1000'
    text = read_dem(name)
    S = n_gpus * per_gpu
    smp = cb.DeviceDemSampler(text, device=0)
    nb = smp.num_detector_bytes
    host = torch.empty((S, nb), dtype=torch.uint8).pin_memory()
    d = torch.empty((per_gpu, nb), dtype=torch.uint8, device='cuda:0')
    for k in range(n_gpus):
        smp.sample_device(per_gpu, d.data_ptr(), nb, None, seed=1000 + k)
        torch.cuda.synchronize()
        host[k * per_gpu:(k + 1) * per_gpu].copy_(d)
    del d
    dets = host.numpy()
    single = cb._compile_text(text, device=0)
    multi = cb._compile_text(text, device=list(range(n_gpus)))
    ref = single.predict_obs_flips_from_dets_bit_packed(dets[:65536])
    for _ in range(2):
        pred = multi.predict_obs_flips_from_dets_bit_packed(dets)
    t0 = time.perf_counter()
    for _ in range(steps):
        pred = multi.predict_obs_flips_from_dets_bit_packed(dets)
    secs = time.perf_counter() - t0
    ok = bool((pred[:65536] == ref).all())
    print(json.dumps({'metric': 'decoded shots/s, midout d25 r100 p=0.1%, one API call on N GPUs', 'n_gpus': n_gpus,
                      'shots_per_call': S, 'steps': steps, 'e2e_value': S * steps / secs, 'unit': 'shots/s',
                      'h2d_bytes_per_call': S * nb, 'matches_single_gpu_decoder': ok,
                      'stats': multi.last_stats.as_dict()}))


if __name__ == '__main__':
    main()
