"""Quick GPU probe: decode sampled shots of one DEM with each group width; print timing + parity vs oracle (optional)."""
import sys, os, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import chromobius_b200 as cb
from dem_util import read_dem
from synth import FastDemSampler

name = sys.argv[1]; shots = int(sys.argv[2]); groups = [x for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else ['32', '128']
check = len(sys.argv) > 4 and sys.argv[4] == 'check'
text = read_dem(name)
t0 = time.time(); dec = cb._compile_text(text, device=0); t1 = time.time()
smp = FastDemSampler(text)
dets, obs = smp.sample(shots, 0)
t2 = time.time()
print('%s: compile %.2fs sample %.2fs D=%d table_bytes=%d' % (name, t1 - t0, t2 - t1, dec.num_detectors, dec.table_bytes), flush=True)
ref = None
def tune(spec):
    parts = str(spec).split(':')
    dec.set_tuning(group_threads=int(parts[0]), groups_per_sm=int(parts[1]) if len(parts) > 1 else 0,
                   ht_log2=int(parts[2]) if len(parts) > 2 else 0)
    return int(parts[0])
for gspec in groups:
    g = tune(gspec)
    for rep in range(2):
        t0 = time.time()
        pred, wint = dec.decode_with_integer_weights(dets)
        dt = time.time() - t0
        st = dec.last_stats
        print('  group=%7s rep%d wall %.3fs  decode %.2f ms (%.2f us/shot, %.0f shots/s) count %.3f ms h2d %.2f ms events/shot %.1f rounds/shot %.2f retried %d groups %d' % (
            gspec, rep, dt, st.ms_decode, st.ms_decode * 1e3 / shots, shots / (st.ms_decode * 1e-3), st.ms_count, st.ms_h2d,
            st.num_detection_events / shots, st.matcher_rounds / shots, st.num_retried_shots, st.num_groups), flush=True)
    if ref is None:
        ref = (pred.copy(), wint.copy())
    else:
        print('  vs first group: weight diffs %d obs diffs %d' % ((wint != ref[1]).sum(), (pred != ref[0]).any(axis=1).sum()))
m = np.zeros(shots, np.uint64)
for k in range(ref[0].shape[1]):
    m |= ref[0][:, k].astype(np.uint64) << np.uint64(8 * k)
print('  mistakes %d / %d' % ((m != obs).sum(), shots))
if check:
    import oracle
    n = min(shots, int(sys.argv[5]) if len(sys.argv) > 5 else 500)
    oo, ow = oracle.OracleDecoder(text).decode_batch(dets[:n])
    print('  oracle check on %d shots: weight diffs %d obs diffs %d' % (n, (ow != ref[1][:n]).sum(), (oo != m[:n]).sum()))
# phase profile of the first group width
tune(groups[0])
dec.profile_phases(True)
dec.decode_with_integer_weights(dets)
ph = dec.profile_phases(False)
main = ['ingest', 'explore_first', 'components', 'solve', 'explore_regrow', 'paths', 'lift']
tot = sum(ph[k] for k in main)
print('  phases (cycles/shot of thread 0, group=%s): ' % groups[0] + ', '.join('%s %.0f (%.0f%%)' % (k, v / max(ph['shots'], 1), 100.0 * v / max(tot, 1)) for k, v in ph.items() if k in main))
print('  solve split (cycles/shot): closed-form %.0f, per-warp blossoms %.0f, block-wide blossoms %.0f' % tuple(ph[k] / max(ph['shots'], 1) for k in ['solve_closed_form', 'solve_warp_blossoms', 'solve_block_blossoms']))
print('  near-table queries/shot %.0f' % (ph['near_queries'] / max(ph['shots'], 1)))
print('  balls/shot %.0f, pops/ball %.2f, nodes/ball %.1f, cycles/ball %.0f, cycles/pop %.0f' % (ph['balls'] / max(ph['shots'], 1), ph['ball_pops'] / max(ph['balls'], 1), ph['ball_nodes'] / max(ph['balls'], 1), ph['ball_cycles'] / max(ph['balls'], 1), ph['ball_cycles'] / max(ph['ball_pops'], 1)))
print('  blossoms per shot / cycles each: ' + ', '.join('%s: %.1f x %.0f' % (b, ph['n_' + b] / max(ph['shots'], 1), ph['blossom_' + b] / max(ph['n_' + b], 1)) for b in ['le8', 'le16', 'le32', 'gt32']))
