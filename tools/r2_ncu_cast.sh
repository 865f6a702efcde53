#!/bin/bash
# one ncu --set full capture of a large wf_cast_rl_kernel launch (and of the fused shading kernel) on the fixed profiling
# workload (tools/wf_profile_run.py 3840x2160x4), each after the same command exited 0 without ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2a}
python tools/wf_profile_run.py 3840x2160x4 > gpurun_out/${T}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wf_cast_rl -s 2 -c 1 -o gpurun_out/prof_${T}_wf_cast_rl -f python tools/wf_profile_run.py 3840x2160x4 > gpurun_out/${T}_ncu_cast.log 2>&1
tail -3 gpurun_out/${T}_ncu_cast.log
