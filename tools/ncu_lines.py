#!/usr/bin/env python3
"""Aggregates ncu's per-SASS-instruction stall samples to source lines (ncu source page CSV + nvdisasm line info).

usage: ncu_lines.py report.ncu-rep lib.so kernel_substring [top_n]
"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubins = [f for f in os.listdir(tmp) if f.endswith('.cubin')]
line_of_offset = {}
for cb in cubins:
    out = subprocess.run(['nvdisasm', '--print-line-info', os.path.join(tmp, cb)], capture_output=True, text=True).stdout
    cur_fn = None
    cur_line = None
    for l in out.split('\n'):
        m = re.match(r'^\.text\.(\S+):', l)
        if m:
            cur_fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;', l)
        if m and cur_fn and kern in cur_fn:
            line_of_offset[int(m.group(1), 16)] = (cur_line, m.group(2))
csv_text = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(csv_text.split('\n')))
hdr = None
base = None
per_line = collections.Counter()
per_line_inst = collections.Counter()
stall_cols = {}
per_line_stalls = collections.defaultdict(collections.Counter)
total = 0
for r in rows:
    if not r:
        continue
    if r[0] == 'Address':
        hdr = r
        stall_cols = {i: h for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
        continue
    if hdr is None or not r[0].startswith('0x'):
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    off = addr - base
    samples = int(r[hdr.index('# Samples')] or 0)
    inst = int(r[hdr.index('Instructions Executed')] or 0)
    info = line_of_offset.get(off)
    key = info[0] if info else None
    per_line[key] += samples
    per_line_inst[key] += inst
    for i, h in stall_cols.items():
        v = int(r[i] or 0)
        if v:
            per_line_stalls[key][h] += v
    total += samples
print('total samples', total)
for key, v in per_line.most_common(topn):
    st = ', '.join('%s %d' % (h.replace('stall_', ''), c) for h, c in per_line_stalls[key].most_common(3))
    print('%6.2f%%  %-28s inst %-10d %s' % (100.0 * v / max(total, 1), '%s:%s' % key if key else '?', per_line_inst[key], st))
