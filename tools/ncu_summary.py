#!/usr/bin/env python3
"""Turns one `ncu --set full` capture of the decode kernel into the summary kept under profiles/.

usage: ncu_summary.py report.ncu-rep out.txt traffic.json "free-text description of the capture" [shots per launch]
  out.txt      : selected raw metrics + stall samples / instructions by source function and line (tools/ncu_funcs.py)
  traffic.json : {"dram_bytes_per_launch": dram__bytes_read.sum + dram__bytes_write.sum, ...} read by bench.py
"""
import csv, io, json, os, subprocess, sys

rep, out_txt, out_json, desc = sys.argv[1:5]
shots_per_launch = int(sys.argv[5]) if len(sys.argv) > 5 else 65536
HERE = os.path.dirname(os.path.abspath(__file__))
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
WANT = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__time_duration.sum', 'l1tex__t_sector_hit_rate.pct', 'launch__block_size', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__sass_average_branch_targets_threads_uniform.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio']
lines = []
metric = {}
for i, h in enumerate(hdr):
    if h in WANT or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')):
        lines.append('%-82s %s %s' % (h, vals[i], units[i]))
    metric[h] = (vals[i], units[i])
kernel = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'


def to_bytes(name):
    v, u = metric[name]
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[u]
    return float(v) * scale


funcs = subprocess.run([sys.executable, os.path.join(HERE, 'ncu_funcs.py'), rep,
                        os.path.join(HERE, '..', 'chromobius_b200', 'csrc', 'decode_core.h'), '30'],
                       capture_output=True, text=True).stdout
with open(out_txt, 'w') as f:
    f.write('# %s\n# kernel: %s\n' % (desc, kernel))
    f.write('\n'.join(sorted(lines)) + '\n\n# stall samples / executed warp instructions by source function and line (tools/ncu_funcs.py)\n')
    f.write(funcs)
dram = to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum')
with open(out_json, 'w') as f:
    def num(name):
        try:
            return float(metric[name][0])
        except Exception:
            return None
    json.dump({'dram_bytes_per_launch': dram, 'shots_per_launch': shots_per_launch, 'kernel': kernel, 'capture': desc,
               'issue_active_pct': num('smsp__issue_active.avg.pct_of_peak_sustained_active'),
               'warps_active_pct': num('sm__warps_active.avg.pct_of_peak_sustained_active'),
               'l2_hit_pct': num('lts__t_sector_hit_rate.pct'), 'l1_hit_pct': num('l1tex__t_sector_hit_rate.pct'),
               'threads_per_warp_instruction': num('smsp__thread_inst_executed_per_inst_executed.ratio'),
               'gpu_time_ms_under_ncu': float(metric['gpu__time_duration.sum'][0]) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3}.get(metric['gpu__time_duration.sum'][1], 1)}, f, indent=1)
print(open(out_txt).read()[:3000])
