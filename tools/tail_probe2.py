"""Per-shot cost diagnostic (library built with -DCB_TAIL_PROF): cycles vs event count, largest component, rounds."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import chromobius_b200 as cb
from dem_util import read_dem
name = 'midout_color_code_d25_r100_p1000'; S = 65536
dev = torch.device('cuda', 0)
text = read_dem(name)
dec = cb._compile_text(text, device=0)
smp = cb.DeviceDemSampler(text, device=0)
nb, ob = dec.num_detector_bytes, dec.num_observable_bytes
d_dets = torch.empty((S, nb), dtype=torch.uint8, device=dev)
d_actual = torch.empty(S, dtype=torch.int64, device=dev)
d_pred = torch.empty((S, ob), dtype=torch.uint8, device=dev)
d_w = torch.empty(S, dtype=torch.int64, device=dev)
smp.sample_device(S, d_dets.data_ptr(), nb, d_actual.data_ptr(), seed=1000)
torch.cuda.synchronize()
dec.decode_device(d_dets.data_ptr(), S, nb, d_pred.data_ptr(), weight_int_ptr=d_w.data_ptr())
dec.profile_phases(True)
dec.decode_device(d_dets.data_ptr(), S, nb, d_pred.data_ptr(), weight_int_ptr=d_w.data_ptr())
dec.profile_phases(False)
w = d_w.cpu().numpy().astype(np.uint64)
cyc = (w & np.uint64(0xFFFFFFFF)).astype(np.float64) * 1024
comp = ((w >> np.uint64(32)) & np.uint64(0xFFFFFF)).astype(np.int64)
rounds = (w >> np.uint64(56)).astype(np.int64)
dets = d_dets.cpu().numpy()
cnt = np.unpackbits(dets, axis=1).sum(axis=1)
print('total cycles %.3e; mean %.2f M' % (cyc.sum(), cyc.mean() / 1e6))
order = np.argsort(-cyc)
print('top 15 shots: cycles(M) events maxcomp rounds')
for i in order[:15]:
    print('  %8.1f %5d %5d %3d' % (cyc[i] / 1e6, cnt[i], comp[i], rounds[i]))
print('corr(cycles, events) %.3f  corr(cycles, maxcomp) %.3f corr(cycles, maxcomp^3) %.3f corr(cycles, rounds) %.3f' % (
    np.corrcoef(cyc, cnt)[0, 1], np.corrcoef(cyc, comp)[0, 1], np.corrcoef(cyc, comp.astype(np.float64) ** 3)[0, 1], np.corrcoef(cyc, rounds)[0, 1]))
for lo, hi in [(0, 16), (16, 32), (32, 48), (48, 64), (64, 96), (96, 128), (128, 192), (192, 256), (256, 100000)]:
    m = (comp >= lo) & (comp < hi)
    if m.any():
        print('maxcomp [%d,%d): %6d shots, mean %.2f M cycles, max %.1f M, share of all cycles %.1f%%, mean events %.0f' % (
            lo, hi, m.sum(), cyc[m].mean() / 1e6, cyc[m].max() / 1e6, 100 * cyc[m].sum() / cyc.sum(), cnt[m].mean()))
qs = np.quantile(cyc, [0.5, 0.9, 0.99, 0.999, 0.9999])
print('cycle quantiles (M): p50 %.2f p90 %.2f p99 %.2f p99.9 %.2f p99.99 %.2f' % tuple(q / 1e6 for q in qs))
