"""Small decode for compute-sanitizer runs (memcheck / racecheck): a few hundred d9 r36 shots with each group width,
checked against the CPU oracle. Usage: compute-sanitizer --tool memcheck python tools/sanitize_probe.py [shots]"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402
import chromobius_b200 as cb  # noqa: E402
import oracle  # noqa: E402
from dem_util import read_dem, NumpyDemSampler  # noqa: E402

shots = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for name in ['midout_color_code_d9_r36_p1000', 'color2surface_d7_phenom_r7_p100']:
    text = read_dem(name)
    dets, _ = NumpyDemSampler(text).sample(shots, seed=3)
    od = oracle.OracleDecoder(text)
    od.set_matcher_mode(3)
    o_obs, o_w = od.decode_batch(dets)
    dec = cb._compile_text(text, device=0)
    for group in (256, 128, 32):
        dec.set_tuning(group_threads=group)
        pred, w = dec.decode_with_integer_weights(dets)
        m = np.zeros(shots, np.uint64)
        for k in range(pred.shape[1]):
            m |= pred[:, k].astype(np.uint64) << np.uint64(8 * k)
        print('%s group %d: weight diffs %d, prediction diffs %d' % (name, group, int((w != o_w).sum()), int((m != o_obs).sum())), flush=True)
