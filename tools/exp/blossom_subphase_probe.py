import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tools')
import chromobius_b200 as cb
from dem_util import read_dem
from synth import FastDemSampler
name = 'midout_color_code_d25_r100_p1000'
text = read_dem(name)
dec = cb._compile_text(text, device=0)
dets, _ = FastDemSampler(text).sample(32768, 0)
dec.decode_with_integer_weights(dets)
dec.profile_phases(True)
dec.decode_with_integer_weights(dets)
ph = dec.profile_phases(False)
b = int(os.environ['BUCKET'])
cls = ['le8', 'le16', 'le32', 'gt32'][b]
n = ph['n_' + cls]
ph22 = ph['ball_nodes']
phases = ph22 & 0xFFFFFFFF; greedy = ph22 >> 32
print('%s: blossoms %d, avg n %.1f, total %.0f cyc each: fill %.0f prepare %.0f solve %.0f; greedy pairs %.2f, matching() calls %.2f' % (
    cls, n, ph['near_queries'] / n, ph['blossom_' + cls] / n, ph['ball_pops'] / n, ph['ball_cycles'] / n, ph['balls'] / n, greedy / n, phases / n))
