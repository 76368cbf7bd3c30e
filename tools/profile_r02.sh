#!/bin/bash
# ncu evidence of round 2 (run on the GPU box through gpurun; one ncu session per call).  Outputs under gpurun_out/:
#   launches_r02.csv     every launch of ONE pass over the c2 batch with its device time (cold-cache, serialised: compare SHARES)
#   prof_full_raw.csv    `ncu --set full` raw metrics of one instance of every distinct kernel configuration of the pass
#   prof_conv6.ncu-rep   the dominant kernel (conv6, column-fused tcgen05 GEMM) with source correlation
set -u
mkdir -p gpurun_out
CMD="python tools/profile_step.py 256 3"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --profile-from-start off -c 100 -o /tmp/prof_full $CMD > gpurun_out/ncu2.log 2>&1
echo "full set rc=$?"
ncu -i /tmp/prof_full.ncu-rep --page raw --csv > gpurun_out/prof_full_raw.csv 2> gpurun_out/ncu2b.log
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -s 4 -c 1 -o gpurun_out/prof_conv6 $CMD > gpurun_out/ncu3.log 2>&1
echo "conv6 rc=$?"
ls -la gpurun_out | tail -12
