cd /root/repo
V=chromobius_b200/variants
L=gpurun_out/r2k_warp_min.log
: > $L
for v in fc warp17 warp9 fc warp17; do CHROMOBIUS_B200_LIB=$V/$v.so timeout 120 python tools/exp/variant_probe.py $v midout_color_code_d25_r100_p1000 65536 200 >> $L 2>&1; done
cat $L
