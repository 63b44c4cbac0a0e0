#!/usr/bin/env python3
"""Per-function / per-line summary of an ncu report using ncu's own CUDA<->SASS correlation (needs -lineinfo).

usage: ncu_funcs.py report.ncu-rep [source_file_for_function_names] [top_lines]
Prints warp-stall samples and executed warp instructions aggregated by source function (crude: the nearest preceding
line that looks like a function header) and the top source lines.
"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
src_path = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
text = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
cur_file = None
hdr = None
line_no = None
per = collections.defaultdict(lambda: collections.Counter())
for row in csv.reader(io.StringIO(text)):
    if not row:
        continue
    if row[0] == 'File Path':
        cur_file = row[1].split('/')[-1]
        continue
    if row[0] == 'Line No':
        hdr = row
        col = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or row[0] == 'Function Name':
        continue
    if row[0].strip().isdigit():
        line_no = int(row[0])
    if len(row) <= col['# Samples']:
        continue
    addr = row[col['Address']] if len(row) > col['Address'] else ''
    if not addr:
        continue  # the CUDA line row itself carries the sums of its SASS rows: count SASS rows only
    def num(name):
        i = col.get(name)
        try:
            return int(row[i]) if i is not None and i < len(row) and row[i] else 0
        except ValueError:
            return 0
    c = per[(cur_file, line_no)]
    c['samples'] += num('# Samples')
    c['inst'] += num('Instructions Executed')
    for h in hdr:
        if h.startswith('stall_') and 'Not Issued' not in h:
            c[h] += num(h)
total = sum(c['samples'] for c in per.values())
tinst = sum(c['inst'] for c in per.values())
print('total samples %d, warp instructions %d' % (total, tinst))
fn_of = {}
if src_path:
    cur = '?'
    base = src_path.split('/')[-1]
    for i, l in enumerate(open(src_path).read().split('\n'), 1):
        m = re.match(r'^\s*(?:static\s+)?(?:CB_HDN?|__device__|__global__)[\w\s:<>\*&,]*?\b(\w+)\s*\(', l)
        if m and not l.strip().startswith('//') and m.group(1) not in ('if', 'for', 'while', 'return', 'switch'):
            cur = m.group(1)
        fn_of[(base, i)] = cur
    agg = collections.defaultdict(lambda: collections.Counter())
    for k, c in per.items():
        f = fn_of.get(k, k[0])
        agg[f].update(c)
    print('\nby function (samples%, barrier-free samples%, inst%):')
    nb_total = sum(c['samples'] - c['stall_barrier'] for c in per.values())
    for f, c in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:45]:
        print('  %-28s %6.2f%%  %6.2f%%  %6.2f%%' % (f, 100.0 * c['samples'] / max(total, 1),
              100.0 * (c['samples'] - c['stall_barrier']) / max(nb_total, 1), 100.0 * c['inst'] / max(tinst, 1)))
print('\ntop lines:')
for k, c in sorted(per.items(), key=lambda kv: -kv[1]['samples'])[:top]:
    stalls = ', '.join('%s %d' % (h[6:], v) for h, v in c.most_common(6) if h.startswith('stall_') and v)
    print('  %5.2f%%  %s:%s  inst %d  %s' % (100.0 * c['samples'] / max(total, 1), k[0], k[1], c['inst'], stalls))
