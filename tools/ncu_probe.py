"""Minimal decode loop for ncu captures: python tools/ncu_probe.py <fixture> <shots> <group[:groups_per_sm]> [calls]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import chromobius_b200 as cb
from dem_util import read_dem
from synth import FastDemSampler

name, shots, spec = sys.argv[1], int(sys.argv[2]), sys.argv[3].split(':')
calls = int(sys.argv[4]) if len(sys.argv) > 4 else 2
text = read_dem(name)
dec = cb._compile_text(text, device=0)
dets, _ = FastDemSampler(text).sample(shots, 0)
dec.set_tuning(group_threads=int(spec[0]), groups_per_sm=int(spec[1]) if len(spec) > 1 else 0, ht_log2=0)
for _ in range(calls):
    dec.decode_with_integer_weights(dets)
    st = dec.last_stats
    print('decode %.2f ms groups %d' % (st.ms_decode, st.num_groups), flush=True)
