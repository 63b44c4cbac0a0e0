cd /root/repo
V=chromobius_b200/variants
L=gpurun_out/r2p_big.log
: > $L
PROBE_SAVE=/tmp/a.npz CHROMOBIUS_B200_LIB=$V/v15.so timeout 120 python tools/exp/variant_probe.py v15 midout_color_code_d25_r100_p1000 65536 0 >> $L 2>&1
PROBE_SAVE=/tmp/b.npz CHROMOBIUS_B200_LIB=$V/big.so timeout 120 python tools/exp/variant_probe.py big midout_color_code_d25_r100_p1000 65536 300 >> $L 2>&1
python -c "
import numpy as np
a=np.load('/tmp/a.npz'); b=np.load('/tmp/b.npz')
print('v15 vs big on 65536 d25 shots: weight diffs %d, prediction diffs %d' % ((a['wint']!=b['wint']).sum(), (a['pred']!=b['pred']).any(axis=1).sum()))
" >> $L 2>&1
CHROMOBIUS_B200_LIB=$V/big.so timeout 120 python tools/exp/variant_probe.py big midout_color_code_d9_r36_p1000 262144 2000 128 >> $L 2>&1
cat $L
