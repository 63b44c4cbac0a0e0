import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/tools')
import numpy as np, chromobius_b200 as cb
from dem_util import read_dem
from synth import FastDemSampler
text=read_dem('midout_color_code_d25_r100_p1000')
dets,_=FastDemSampler(text).sample(32768,0)
dec=cb._compile_text(text,device=0)
for g in [128,256,512,32]:
    dec.set_tuning(group_threads=g)
    for sl in [slice(29535,29536), slice(29000,30000), slice(0,32768)]:
        try:
            p,w=dec.decode_with_integer_weights(dets[sl]); print(g,sl,'ok',int(w.sum()))
        except Exception as e: print(g,sl,'ERR',str(e)[:80])
