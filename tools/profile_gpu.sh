#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench run, then the ncu launch list and one full capture of the decode kernel.
set -u
mkdir -p gpurun_out
SHOTS=${1:-65536}
CMD="python bench.py --steps 2 --warmup 1 --shots $SHOTS --e2e-shots 1024 --skip-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_decode -c 1 -o gpurun_out/prof_decode $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/plain.log | cut -c1-400
