#!/usr/bin/env python3
"""bench.py -- decoded shots/s on the midout d25 r100 p=0.1% workload (BASELINE.json's metric), 1..8 B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU oracle on the box's host cores (reference arm)

A step = one pass of the decode path over one batch of `--shots` synthetic shots per GPU (weak scaling: every GPU holds a
replica of the decoder tables and its own batch). The default, 1 250 000 shots per GPU per step, is one GPU's share of
BASELINE.json's headline configuration (10 M shots sharded across 8 B200), so `--gpus 8` is exactly that configuration.
Shots are drawn on the device by the DEM sampler (independent mechanisms of the midout_color_code_d25_r100_p1000
detector error model); the batch (3.85 GB per GPU) is far larger than L2. The batch size matters: per-shot decode time
has a long tail (mean 1.5 ms, the longest shot of 65 536 takes 60-100 ms), and the persistent kernel ends when its
last shot does -- a 65 536-shot batch loses ~15 % to that tail (`--shots 65536` reproduces the earlier rounds' figure).

  value : shots/s with the batch already resident in HBM, timed with the CUDA events the library records on its own
          stream (sum over the K steps, max over ranks).
  e2e   : shots/s through the public API (CompiledDecoder.predict_obs_flips_from_dets_bit_packed) with pinned HOST input
          and host output, H2D and D2H inside the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

WORKLOAD = 'midout_color_code_d25_r100_p1000'
METRIC = 'decoded shots/s, midout d25 r100 p=0.1%'


def read_dem(name):
    with open(os.path.join(ROOT, 'tests', 'golden', 'dems', name + '.dem')) as f:
        return f.read()


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def measure_l2_copy_peak(torch, dev):
    """GB/s (read + write) of a copy whose source and destination both stay in the 126 MB L2."""
    try:
        n = 24 << 20
        a = torch.empty(n, dtype=torch.uint8, device=dev)
        b = torch.empty(n, dtype=torch.uint8, device=dev)
        for _ in range(5):
            b.copy_(a)
        best = None
        for _ in range(20):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                b.copy_(a)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            best = ms if best is None else min(best, ms)
        return 2 * n / (best * 1e-3) / 1e9
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        # ONE long-lived nvidia-smi per rank (-lms 200), as in the recipe: spawning a fresh nvidia-smi per sample costs
        # about a CPU-second each on an 8-GPU box and starves the ranks' launch threads.
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        for ln in (out or '').splitlines():
            if ln.strip():
                self.rows.append([x.strip() for x in ln.split(',')])

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            for k, nme in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith('active'):
                    reasons.add(nme)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.rows)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    dem_text, dets = args
    import oracle
    dec = oracle.OracleDecoder(dem_text)
    dec.set_matcher_mode(3)
    # the fastest CPU configuration of the restated decoder: no extra event per growth request (the kernels ask for one
    # to save matcher rounds, which costs a CPU 5-7 %). Weights are exact either way.
    dec.set_growth_extra(0)
    t0 = time.monotonic()
    obs, w = dec.decode_batch(dets)
    return time.monotonic() - t0, obs, w


def cpu_oracle_throughput(dem_text, dets, cores, want_results=False):
    """Decodes `dets` split over `cores` processes; the clock covers the decode call only, as in the reference's
    tools/gen_timing_stats:26-50. Returns shots/s = total shots / slowest worker (and the oracle's answers on request)."""
    import numpy as np
    chunks = [c for c in np.array_split(dets, cores) if len(c)]
    ctx = mp.get_context('fork')
    with ctx.Pool(len(chunks)) as pool:
        res = pool.map(_cpu_worker, [(dem_text, c) for c in chunks])
    slowest = max(r[0] for r in res)
    if want_results:
        return len(dets) / slowest, slowest, np.concatenate([r[1] for r in res]), np.concatenate([r[2] for r in res])
    return len(dets) / slowest, slowest


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def sample_host_shots(dem_text, shots, seed):
    """Shots for the reference arm: the numpy DEM sampler (tools/synth.py). The reference arm maps no product code."""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    from synth import FastDemSampler
    return FastDemSampler(dem_text).sample(shots, seed, chunk=4096)[0]


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    import __graft_entry__ as g
    g.build_oracle()
    dem_text = read_dem(args.workload)
    cores = host_cores()
    per_core = args.cpu_shots_per_core
    dets = sample_host_shots(dem_text, cores * per_core, seed=1)
    for _ in range(max(args.warmup, 0) and 1):
        cpu_oracle_throughput(dem_text, dets[:cores * 8], cores)
    t_total = 0.0
    n_total = 0
    for _ in range(args.steps):
        rate, secs = cpu_oracle_throughput(dem_text, dets, cores)
        t_total += secs
        n_total += len(dets)
    value = n_total / t_total
    sample = '%d shots per step (%d per core) of %s, %d steps' % (len(dets), per_core, args.workload, args.steps)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'shots/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * t_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'int64', 'data': 'synthetic',
        'config': {'workload': args.workload, 'shots_per_step': len(dets), 'host_cores': cores,
                   'note': 'restated CPU oracle (certified-local exact MWPM), not the PyMatching binary: stim/pymatching are not installable here; '
                           'reference perf goal for this config is 305 shots/s/core (src/chromobius/decode/decoder.perf.cc:178,199)'},
        'cpu_baseline': {'value': value, 'unit': 'shots/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'shots/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import chromobius_b200 as cb

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device; chromobius_b200 has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dem_text = read_dem(args.workload)
    t0 = time.time()
    dec = cb._compile_text(dem_text, device=local)
    compile_s = time.time() - t0
    if args.group:
        dec.set_tuning(group_threads=args.group)
    smp = cb.DeviceDemSampler(dem_text, device=local)
    S = args.shots
    nb, ob = dec.num_detector_bytes, dec.num_observable_bytes
    d_dets = torch.empty((S, nb), dtype=torch.uint8, device=dev)
    d_actual = torch.empty(S, dtype=torch.int64, device=dev)
    d_pred = torch.empty((S, ob), dtype=torch.uint8, device=dev)
    d_wint = torch.empty(S, dtype=torch.int64, device=dev)
    smp.sample_device(S, d_dets.data_ptr(), nb, d_actual.data_ptr(), seed=1000 + rank)
    torch.cuda.synchronize()

    def collectives():
        """What the path exchanges across GPUs: mistake counts are summed, predictions gathered (both tiny)."""
        pred64 = d_pred[:, 0].to(torch.int64)
        for k in range(1, ob):
            pred64 |= d_pred[:, k].to(torch.int64) << (8 * k)
        mistakes = (pred64 != d_actual).sum()
        if world > 1:
            dist.all_reduce(mistakes)
            gathered = torch.empty((world * S, ob), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(gathered, d_pred)
        return mistakes

    def device_step():
        st = dec.decode_device(d_dets.data_ptr(), S, nb, d_pred.data_ptr(), weight_int_ptr=d_wint.data_ptr())
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        m = collectives()
        e1.record()
        torch.cuda.synchronize()
        return st, e0.elapsed_time(e1), int(m.item())

    for _ in range(args.warmup):
        device_step()
    barrier()
    dev_ms = 0.0
    decode_kernel_ms = 0.0
    coll_total_ms = 0.0
    launches = 0
    events = 0
    rounds = 0
    retried = 0
    mistakes = 0
    wall0 = time.perf_counter()
    with ClockSampler(local) as clocks:
        for _ in range(args.steps):
            st, coll_ms, mistakes = device_step()
            dev_ms += st.ms_total + coll_ms
            decode_kernel_ms += st.ms_decode
            coll_total_ms += coll_ms
            launches += st.kernel_launches + 4  # + the elementwise/reduction kernels of collectives() (torch)
            events += st.num_detection_events
            rounds += st.matcher_rounds
            retried += st.num_retried_shots
        barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    # per-rank split (diagnostic): library time per step and time inside the collectives, which includes waiting for
    # the slowest rank to arrive
    mine = torch.tensor([(dev_ms - coll_total_ms) / args.steps, coll_total_ms / args.steps], dtype=torch.float64, device=dev)
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
    per_rank = [[round(float(x[0]), 2), round(float(x[1]), 2)] for x in per_rank]
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    value = world * S * args.steps / (dev_ms * 1e-3)
    group_threads = st.group_threads

    # ---- e2e: public API, host input (pinned, then pageable as a plain numpy caller has it), host output ----
    E = min(args.e2e_shots, S) if args.e2e_shots > 0 else S
    h_dets_t = torch.empty((E, nb), dtype=torch.uint8).pin_memory()
    h_dets_t.copy_(d_dets[:E])
    h_dets = h_dets_t.numpy()

    def time_e2e(host_rows):
        for _ in range(max(1, min(args.warmup, 2))):
            dec.predict_obs_flips_from_dets_bit_packed(host_rows)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = dec.predict_obs_flips_from_dets_bit_packed(host_rows)
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        tt = torch.tensor([secs], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return world * E * args.steps / float(tt[0]), out

    e2e_value, pred = time_e2e(h_dets)
    e2e_ok = bool((torch.from_numpy(pred).to(dev) == d_pred[:E]).all().item())
    pageable = np.empty((E, nb), np.uint8)
    pageable[:] = h_dets
    e2e_pageable, _ = time_e2e(pageable)
    del pageable

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        a_shot = nb + ob + 8  # algorithmic bytes per shot: detector row in, prediction out, int64 weight out
        ms_launch = decode_kernel_ms / args.steps
        achieved = S * a_shot / (ms_launch * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'decode_kernel_traffic.json')) as f:
                tj = json.load(f)
            # the ncu capture is taken at a smaller launch (ncu replays the kernel ~40 times): scaled to this launch's shots
            traffic = tj['dram_bytes_per_launch'] * S / tj.get('shots_per_launch', 65536)
        except Exception:
            pass
        cores = host_cores()
        cpu = None
        parity = None
        if not args.skip_cpu:
            import __graft_entry__ as g
            g.build_oracle()
            n_cpu = cores * args.cpu_shots_per_core
            n_have = min(n_cpu, S)
            h = np.ascontiguousarray(d_dets[:n_have].cpu().numpy())
            if len(h) < n_cpu:  # many-core hosts: repeat the batch so the sample still takes ~10-30 s
                h = np.ascontiguousarray(np.tile(h, ((n_cpu + len(h) - 1) // len(h), 1))[:n_cpu])
            rate, secs, o_obs, o_w = cpu_oracle_throughput(dem_text, h, cores, want_results=True)
            cpu = {'value': rate, 'unit': 'shots/s', 'cores': cores, 'kind': 'port',
                   'sample': '%d shots (%d per core) of the same batch, %.1f s; restated CPU oracle, decode call only' % (
                       len(h), args.cpu_shots_per_core, secs),
                   'reference_published_per_core': 305.0,
                   'reference_published_source': 'perf goal 128 shots / 420 ms, src/chromobius/decode/decoder.perf.cc:178,199'}
            # the CPU answers are not thrown away: they check the GPU's on the very same shots
            g_pred = d_pred[:n_have].cpu().numpy()
            g_mask = np.zeros(n_have, np.uint64)
            for k in range(ob):
                g_mask |= g_pred[:, k].astype(np.uint64) << np.uint64(8 * k)
            g_w = d_wint[:n_have].cpu().numpy()
            # predictions: against the oracle configured exactly like the kernels (growth_extra = 1) on a smaller sample;
            # the timed configuration above may pick another of several equal-weight matchings (counted as ties)
            import oracle
            n_exact = min(n_have, 8192)
            exact = oracle.OracleDecoder(dem_text)
            exact.set_matcher_mode(3)
            e_obs, e_w = exact.decode_batch(h[:n_exact])
            parity = {'shots': int(n_have), 'weight_diffs': int((g_w != o_w[:n_have]).sum()),
                      'tie_break_pred_diffs_vs_timed_cpu_config': int((g_mask != o_obs[:n_have]).sum()),
                      'exact_config_shots': int(n_exact), 'exact_config_weight_diffs': int((g_w[:n_exact] != e_w).sum()),
                      'pred_diffs': int((g_mask[:n_exact] != e_obs).sum()),
                      'against': 'CPU oracle (certified-local matcher, mode 3) on the first shots of the bench batch: weights '
                                 'on all `shots`, predictions on `exact_config_shots` with the kernels\' growth setting'}
        # graph-side roofline: G_shot from oracle counters (profiles/r02_g_shot.json, tools/g_shot.py) against an
        # L2-resident copy measured right here
        l2_peak = measure_l2_copy_peak(torch, dev)
        g_shot = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'r02_g_shot.json')) as f:
                g_shot = json.load(f)[args.workload]['g_shot_bytes_ell_rows']
        except Exception:
            pass
        issue_active = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'decode_kernel_traffic.json')) as f:
                issue_active = json.load(f).get('issue_active_pct')
        except Exception:
            pass
        line = {
            'metric': METRIC, 'value': value, 'unit': 'shots/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': dev_ms / args.steps, 'us_per_shot': dev_ms * 1e3 / (args.steps * S), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int64', 'data': 'synthetic',
            'config': {'workload': args.workload, 'shots_per_gpu_per_step': S, 'input_bytes_per_gpu_per_step': S * nb,
                       'l2_policy': 'inputs larger than L2 (%d MB per step vs 126 MB L2)' % (S * nb // 1000000),
                       'traffic_source': 'ncu --set full capture at 65536 shots per launch, scaled by shots (profiles/decode_kernel_traffic.json)',
                       'detection_events_per_shot': events / (args.steps * S), 'matcher_rounds_per_shot': rounds / (args.steps * S),
                       'retried_shots_per_step': retried / args.steps, 'group_threads': group_threads,
                       'mistakes_last_step_all_gpus': mistakes, 'decoder_compile_s': compile_s,
                       'decoder_table_bytes': dec.table_bytes, 'wall_ms_per_step': wall_ms / args.steps,
                       'per_rank_ms_per_step_decode_and_collectives': per_rank,
                       'timing': 'CUDA events on the library stream (per call) + torch events for the collectives, max over ranks'},
            'e2e': {'value': e2e_value, 'unit': 'shots/s', 'h2d_bytes_per_step': E * nb, 'd2h_bytes_per_step': E * ob,
                    'shots_per_gpu_per_step': E, 'matches_device_path': e2e_ok, 'host_memory': 'pinned',
                    'pageable_host_memory_value': e2e_pageable},
            'gpu_launches': int(launches),
            'clocks': clocks.summary(),
            'roofline': {'bound': 'hbm', 'kernel': 'k_decode<%d>' % group_threads, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': S * a_shot, 'avg_launch_ms': ms_launch, 'limiter': 'latency',
                         'issue_active_pct': issue_active,
                         'note': 'the contract asks for the HBM figure; the kernel is latency bound (irregular graph search), '
                                 'neither HBM nor L2 bandwidth limits it -- see roofline_l2 and DESIGN.md'},
            'roofline_l2': None if g_shot is None else {
                'bound': 'l2', 'achieved': S * g_shot / (ms_launch * 1e-3) / 1e9, 'peak': l2_peak, 'unit': 'GB/s',
                'frac': S * g_shot / (ms_launch * 1e-3) / 1e9 / l2_peak if l2_peak else None, 'g_shot_bytes': g_shot,
                'peak_source': 'L2-resident copy (2 x 24 MB buffers, read + write bytes, best of 20) measured in this run',
                'definition': 'G_shot = event-node adjacency rows + 8 B per matched path edge (profiles/r02_g_shot.json)'},
            'parity': parity,
            'cpu_baseline': cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default=WORKLOAD)
    ap.add_argument('--shots', type=int, default=1250000, help='shots per GPU per step')
    ap.add_argument('--e2e-shots', type=int, default=0, help='shots per GPU per e2e step (0 = --shots)')
    ap.add_argument('--cpu-shots-per-core', type=int, default=6144,
                    help='CPU legs: shots per host core per sample (~15 s at ~400 shots/s/core)')
    ap.add_argument('--group', type=int, default=0)
    ap.add_argument('--skip-cpu', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
