#!/bin/bash
# env-knob A/B on one B200: L2 fetch granularity and work-list interleaving (cast / logic ms per 16-epoch 4K batch)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-knobs}
: > gpurun_out/r2_${T}.log
for rep in 1 2; do
for g in 32 64 128; do for il in 0 10; do
  echo -n "L2_FETCH=$g INTERLEAVE=$il : " >> gpurun_out/r2_${T}.log
  B200RT_L2_FETCH=$g B200RT_WF_INTERLEAVE=$il timeout 300 python tools/wf_split_time.py 3840x2160x16 >> gpurun_out/r2_${T}.log 2>&1
done; done; done
cat gpurun_out/r2_${T}.log | cut -c1-150
