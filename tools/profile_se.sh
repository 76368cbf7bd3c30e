#!/bin/bash
# ncu --set full with source correlation of the three se_excite_kernel launches of one pass (run through gpurun).
set -u
mkdir -p gpurun_out
CMD="python tools/profile_step.py 256 3 $*"
$CMD > gpurun_out/plain_se.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_se.log; exit 1; }
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:se_excite -c 3 -o gpurun_out/prof_se $CMD > gpurun_out/ncu_se.log 2>&1
echo "se rc=$?"
ls -la gpurun_out/prof_se.ncu-rep
