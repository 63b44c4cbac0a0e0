#!/usr/bin/env python3
"""Per-config table for every circuit BASELINE.json names: GPU shots/s (device-resident and through the host API), the
CPU oracle on all host cores beside it, and parity of a sample against the oracle (integer weights, predictions).

usage (on a GPU box): python tools/bench_configs.py [out.json [substring of the config names to run]]
Shots are drawn by the device sampler (seed 0) at the shot counts of BASELINE.json.configs; the d25 r100 row uses one
bench.py-sized batch per step. The oracle is used as the checker and as the timed CPU baseline only.
"""
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import bench
import chromobius_b200 as cb
import oracle
from dem_util import read_dem

CONFIGS = [
    # (fixture, shots, shots per GPU call, oracle parity sample, CPU-timing sample per core)
    ('phenom_color_code_d5_r5_p1000', 100_000, 100_000, 20_000, 100_000),
    ('midout_color_code_d9_r36_p1000', 1_000_000, 1_000_000, 4_000, 12_000),
    ('superdense_color_code_d5_r20_p1000', 1_000_000, 1_000_000, 8_000, 30_000),
    ('midout488_color_code_d9_r33_p1000', 1_000_000, 1_000_000, 4_000, 16_000),
    ('color2surface_d7_phenom_r7_p100', 1_000_000, 1_000_000, 4_000, 16_000),
    ('midout_color_code_d25_r100_p1000', 262_144, 65_536, 500, 1_500),
]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    dev = torch.device('cuda', 0)
    cores = bench.host_cores()
    rows = []
    only = sys.argv[2] if len(sys.argv) > 2 else ''
    for name, shots, per_call, n_check, cpu_per_core in CONFIGS:
        if only not in name:
            continue
        text = read_dem(name)
        t0 = time.time()
        dec = cb._compile_text(text, device=0)
        compile_s = time.time() - t0
        smp = cb.DeviceDemSampler(text, device=0)
        nb, ob = dec.num_detector_bytes, dec.num_observable_bytes
        d_dets = torch.empty((per_call, nb), dtype=torch.uint8, device=dev)
        d_act = torch.empty(per_call, dtype=torch.int64, device=dev)
        d_pred = torch.empty((per_call, ob), dtype=torch.uint8, device=dev)
        d_w = torch.empty(per_call, dtype=torch.int64, device=dev)
        calls = shots // per_call
        ms_dev = 0.0
        events = 0
        mistakes = 0
        group = 0
        first = None
        for it in range(calls + 1):  # call 0 is the warm-up (arena set-up, clocks)
            smp.sample_device(per_call, d_dets.data_ptr(), nb, d_act.data_ptr(), seed=0, first_shot=max(it - 1, 0) * per_call)
            torch.cuda.synchronize()
            st = dec.decode_device(d_dets.data_ptr(), per_call, nb, d_pred.data_ptr(), weight_int_ptr=d_w.data_ptr())
            torch.cuda.synchronize()
            if it == 0:
                continue
            ms_dev += st.ms_total
            events += st.num_detection_events
            group = st.group_threads
            pred64 = d_pred[:, 0].to(torch.int64)
            for k in range(1, ob):
                pred64 |= d_pred[:, k].to(torch.int64) << (8 * k)
            mistakes += int((pred64 != d_act).sum().item())
            if first is None:
                first = (d_dets[:max(n_check, 1)].cpu().numpy().copy(), pred64[:n_check].cpu().numpy().astype(np.uint64),
                         d_w[:n_check].cpu().numpy().copy())
        gpu_rate = calls * per_call / (ms_dev * 1e-3)
        # host API (pinned host input, host output): the first full-size call (it allocates the staging buffers and
        # sizes the arena) and the one after it
        h = torch.empty((per_call, nb), dtype=torch.uint8).pin_memory()
        h.copy_(d_dets)
        hn = h.numpy()
        dec.predict_obs_flips_from_dets_bit_packed(hn[:1024])
        t0 = time.perf_counter()
        dec.predict_obs_flips_from_dets_bit_packed(hn)
        e2e_first_rate = per_call / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        dec.predict_obs_flips_from_dets_bit_packed(hn)
        e2e_rate = per_call / (time.perf_counter() - t0)
        # parity sample against the oracle
        o_obs, o_w = oracle.OracleDecoder(text).decode_batch(first[0])
        w_diff = int((o_w != first[2]).sum())
        p_diff = int((o_obs != first[1]).sum())
        # CPU oracle on all cores, bounded sample
        n_cpu = min(cores * cpu_per_core, per_call)
        cpu_rate, cpu_s = bench.cpu_oracle_throughput(text, np.ascontiguousarray(hn[:n_cpu]), cores)
        row = {'config': name, 'detectors': int(dec.num_detectors), 'shots': calls * per_call, 'group_threads': int(group),
               'events_per_shot': events / (calls * per_call), 'gpu_shots_per_s': gpu_rate, 'gpu_us_per_shot': 1e6 / gpu_rate,
               'host_api_shots_per_s': e2e_rate, 'host_api_first_call_shots_per_s': e2e_first_rate, 'cpu_oracle_shots_per_s': cpu_rate, 'cpu_cores': cores,
               'cpu_sample_shots': int(n_cpu), 'cpu_sample_s': cpu_s, 'gpu_over_cpu': gpu_rate / cpu_rate,
               'logical_error_rate': mistakes / (calls * per_call), 'oracle_checked_shots': int(len(first[0])),
               'weight_diffs': w_diff, 'prediction_diffs': p_diff, 'compile_s': compile_s, 'table_bytes': int(dec.table_bytes)}
        rows.append(row)
        print(json.dumps(row), flush=True)
        del dec, smp
    print()
    print('%-40s %7s %9s %6s %12s %12s %12s %7s %9s %s' % ('config', 'D', 'shots', 'group', 'GPU shots/s', 'host API', 'CPU %dc' % cores,
                                                        'x CPU', 'LER', 'oracle sample: weight/pred diffs'))
    for r in rows:
        print('%-40s %7d %9d %6d %12.0f %12.0f %12.0f %7.1f %9.5f %d: %d/%d' % (
            r['config'], r['detectors'], r['shots'], r['group_threads'], r['gpu_shots_per_s'], r['host_api_shots_per_s'],
            r['cpu_oracle_shots_per_s'], r['gpu_over_cpu'], r['logical_error_rate'], r['oracle_checked_shots'],
            r['weight_diffs'], r['prediction_diffs']))
    if out_path:
        with open(out_path, 'w') as f:
            json.dump(rows, f, indent=1)


if __name__ == '__main__':
    main()
