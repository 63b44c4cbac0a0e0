"""A/B probe for library variants (CHROMOBIUS_B200_LIB=chromobius_b200/variants/x.so): d25 r100, one group shape, decode
time of the best of three calls, oracle check on a few shots, then the 24 raw profile slots per shot."""
import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import numpy as np
import chromobius_b200 as cb
from dem_util import read_dem
from synth import FastDemSampler
label = sys.argv[1]
name = sys.argv[2] if len(sys.argv) > 2 else 'midout_color_code_d25_r100_p1000'
shots = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
check = int(sys.argv[4]) if len(sys.argv) > 4 else 0
spec = sys.argv[5] if len(sys.argv) > 5 else '256:3'
text = read_dem(name)
dec = cb._compile_text(text, device=0)
dets, obs = FastDemSampler(text).sample(shots, 0)
parts = spec.split(':')
dec.set_tuning(group_threads=int(parts[0]), groups_per_sm=int(parts[1]) if len(parts) > 1 else 0, ht_log2=0)
best = 1e30
for rep in range(3):
    pred, wint = dec.decode_with_integer_weights(dets)
    best = min(best, dec.last_stats.ms_decode)
st = dec.last_stats
print('%s %s: decode %.2f ms (%.0f shots/s) rounds/shot %.2f retried %d wsum %d' % (
    label, name[:24], best, shots / (best * 1e-3), st.matcher_rounds / shots, st.num_retried_shots, int(wint.sum())), flush=True)
if os.environ.get('PROBE_SAVE'):
    np.savez(os.environ['PROBE_SAVE'], pred=pred, wint=wint)
if check:
    import oracle
    m = np.zeros(shots, np.uint64)
    for k in range(pred.shape[1]):
        m |= pred[:, k].astype(np.uint64) << np.uint64(8 * k)
    oo, ow = oracle.OracleDecoder(text).decode_batch(dets[:check])
    print('  oracle check on %d shots: weight diffs %d obs diffs %d' % (check, (ow != wint[:check]).sum(), (oo != m[:check]).sum()), flush=True)
dec.profile_phases(True)
dec.decode_with_integer_weights(dets)
ph = dec.profile_phases(False)
n = max(ph['shots'], 1)
print('  slots/shot: ' + ' '.join('%s=%.0f' % (k, v / n) for k, v in ph.items()), flush=True)
