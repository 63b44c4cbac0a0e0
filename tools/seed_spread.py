"""Decode-kernel time of independent 65 536-shot batches (different sampler seeds) on ONE GPU: separates data-dependent
spread (a late heavy shot holding the persistent kernel's tail) from GPU-to-GPU differences seen in multi-GPU runs."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import chromobius_b200 as cb
from dem_util import read_dem

name = sys.argv[1] if len(sys.argv) > 1 else 'midout_color_code_d25_r100_p1000'
S = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
seeds = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else list(range(1000, 1008))
dev = torch.device('cuda', 0)
text = read_dem(name)
dec = cb._compile_text(text, device=0)
smp = cb.DeviceDemSampler(text, device=0)
nb, ob = dec.num_detector_bytes, dec.num_observable_bytes
d_dets = torch.empty((S, nb), dtype=torch.uint8, device=dev)
d_actual = torch.empty(S, dtype=torch.int64, device=dev)
d_pred = torch.empty((S, ob), dtype=torch.uint8, device=dev)
d_wint = torch.empty(S, dtype=torch.int64, device=dev)
for seed in seeds:
    smp.sample_device(S, d_dets.data_ptr(), nb, d_actual.data_ptr(), seed=seed)
    torch.cuda.synchronize()
    ms = []
    for rep in range(3):
        st = dec.decode_device(d_dets.data_ptr(), S, nb, d_pred.data_ptr(), weight_int_ptr=d_wint.data_ptr())
        ms.append(st.ms_decode)
    print('seed %d: decode ms %s  events/shot %.2f rounds/shot %.3f max weight %d' % (
        seed, ' '.join('%.1f' % m for m in ms), st.num_detection_events / S, st.matcher_rounds / S, int(d_wint.max().item())), flush=True)
