#!/usr/bin/env python3
"""Local-memory traffic (and any other per-instruction ncu column) of a kernel by source line.
usage: ncu_local.py report.ncu-rep lib.so kernel_substring [column] [top_n]"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
col = sys.argv[4] if len(sys.argv) > 4 else 'L2 Theoretical Sectors Local'
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
line_of_offset = {}
for cb in [f for f in os.listdir(tmp) if f.endswith('.cubin')]:
    out = subprocess.run(['nvdisasm', '--print-line-info', os.path.join(tmp, cb)], capture_output=True, text=True).stdout
    cur_fn = cur_line = None
    for l in out.split('\n'):
        m = re.match(r'^\.text\.(\S+):', l)
        if m:
            cur_fn = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;', l)
        if m and cur_fn and kern in cur_fn:
            line_of_offset[int(m.group(1), 16)] = cur_line
text = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
hdr = base = None
per = collections.Counter(); inst = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for r in csv.reader(text.split('\n')):
    if not r: continue
    if r[0] == 'Address':
        hdr = r; continue
    if hdr is None or not r[0].startswith('0x'): continue
    addr = int(r[0], 16)
    if base is None: base = addr
    key = line_of_offset.get(addr - base)
    v = float(r[hdr.index(col)] or 0)
    if v:
        per[key] += v
        inst[key] += int(r[hdr.index('Instructions Executed')] or 0)
        ops[key][r[1].split()[0] if not r[1].startswith('@') else r[1].split()[1]] += 1
tot = sum(per.values())
print('%s total %.4g' % (col, tot))
for k, v in per.most_common(topn):
    print('%5.1f%%  %-28s inst %12d  %s' % (100 * v / tot, k, inst[k], dict(ops[k])))
