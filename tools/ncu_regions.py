#!/usr/bin/env python3
"""Stall-reason mix of an ncu report per code region of decode_core.h (line ranges below; needs -lineinfo).

usage: ncu_regions.py report.ncu-rep [decode_core.h]
"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
text = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
cur_file=None; hdr=None; line_no=None
per = collections.defaultdict(collections.Counter)
for row in csv.reader(io.StringIO(text)):
    if not row: continue
    if row[0]=='File Path': cur_file=row[1].split('/')[-1]; continue
    if row[0]=='Line No': hdr=row; col={h:i for i,h in enumerate(hdr)}; continue
    if hdr is None or row[0]=='Function Name': continue
    if row[0].strip().isdigit(): line_no=int(row[0])
    if len(row)<=col['# Samples']: continue
    addr=row[col['Address']] if len(row)>col['Address'] else ''
    if not addr: continue
    def num(name):
        i=col.get(name)
        try: return int(row[i]) if i is not None and i<len(row) and row[i] else 0
        except ValueError: return 0
    c=per[(cur_file,line_no)]
    c['samples']+=num('# Samples'); c['inst']+=num('Instructions Executed')
    for h in hdr:
        if h.startswith('stall_') and 'Not Issued' not in h: c[h]+=num(h)
# Region boundaries are found from the source itself (function names), so they follow the code as it moves.
import re
src_path = sys.argv[2] if len(sys.argv) > 2 else 'chromobius_b200/csrc/decode_core.h'
src_lines = open(src_path).read().split('\n')
def line_of(pattern):
    for k, ln in enumerate(src_lines):
        if re.search(pattern, ln):
            return k + 1
    raise SystemExit('pattern not found: ' + pattern)
marks = [('ingest / explore / ball', r'void ingest_shot\('), ('blossom struct', r'^struct Blossom \{'), ('union-find / solve_component_* / fill / writeback', r'uint32_t uf_find\('),
         ('grow_some', r'void grow_some\('), ('match_shot (rounds, waves)', r'uint32_t match_shot\('), ('paths', r'bool near_path_emit\('), ('lift + discharge', r'^CB_HD uint32_t find_half_edge')]
bounds = [(name, line_of(pat)) for name, pat in marks] + [('end', len(src_lines) + 1)]
ranges = {'%s (%d-%d)' % (bounds[k][0], bounds[k][1], bounds[k + 1][1]): (bounds[k][1], bounds[k + 1][1]) for k in range(len(marks))}
tot=sum(c['samples'] for c in per.values())
for name,(a,b) in ranges.items():
    agg=collections.Counter()
    for (f,l),c in per.items():
        if f=='decode_core.h' and l is not None and a<=l<b: agg.update(c)
    s=agg['samples']
    print('%-50s samples %5.1f%% inst %d | %s' % (name, 100.0*s/tot, agg['inst'], ', '.join('%s %.0f%%'%(h[6:],100.0*v/max(s,1)) for h,v in agg.most_common(9) if h.startswith('stall_'))))
