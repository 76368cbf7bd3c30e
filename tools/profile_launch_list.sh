#!/bin/bash
# ncu launch list (device time per launch) of ONE pass over the c2 batch with the final kernels -> gpurun_out/launches_r02.csv
set -u
mkdir -p gpurun_out
CMD="python tools/profile_step.py 256 3"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
