"""Where one device-resident decode call spends its time (library CUDA events): count, first decode launch, retries, finish."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import chromobius_b200 as cb
from dem_util import read_dem
name = 'midout_color_code_d25_r100_p1000'
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1250000
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
for rep in range(int(os.environ.get("REPS", 3))):
    st = dec.decode_device(d_dets.data_ptr(), S, nb, d_pred.data_ptr(), weight_int_ptr=d_w.data_ptr())
    print('rep %d: total %.1f ms = h2d %.1f + count %.1f + decode(first launch) %.1f + d2h/finish %.1f + rest (retries, host) %.1f; retried %d, max events/shot %d, launches %d' % (
        rep, st.ms_total, st.ms_h2d, st.ms_count, st.ms_decode, st.ms_d2h,
        st.ms_total - st.ms_h2d - st.ms_count - st.ms_decode - st.ms_d2h, st.num_retried_shots, st.max_events_per_shot, st.kernel_launches), flush=True)
