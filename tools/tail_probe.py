"""Tail diagnostic (library built with -DCB_TAIL_PROF): when do the persistent groups finish, how long is the longest shot."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
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
for seed in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else '1000,1005,1007').split(',')]:
    smp.sample_device(S, d_dets.data_ptr(), nb, d_actual.data_ptr(), seed=seed)
    torch.cuda.synchronize()
    dec.decode_device(d_dets.data_ptr(), S, nb, d_pred.data_ptr())
    dec.profile_phases(True)
    st = dec.decode_device(d_dets.data_ptr(), S, nb, d_pred.data_ptr())
    ph = dec.profile_phases(False)
    groups = st.num_groups
    print('seed %d: kernel %.1f ms; groups %d: mean finish %.1f ms, last finish %.1f ms; longest shot %.2f ms (%.1f M cycles), shots > 8 M cycles: %d; mean shot %.2f M cycles' % (
        seed, st.ms_decode, groups, ph['ball_cycles'] / groups / 1e6, ph['ball_nodes'] / 1e6, ph['ball_pops'] / 1.965e6, ph['ball_pops'] / 1e6, ph['near_queries'],
        sum(ph[k] for k in ['ingest', 'explore_first', 'components', 'solve', 'explore_regrow', 'paths', 'lift']) / ph['shots'] / 1e6), flush=True)
