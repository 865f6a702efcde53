#!/bin/bash
# Round-2 A/B of the cast kernel on one B200: cast / logic time per 16-epoch 4K batch for the in-tree build and every
# build in tools/_variants (B200RT_LIB).  usage: tools/r2_cast_ab.sh TAG [tests]   -> gpurun_out/r2_TAG_*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-ab}
if [ "$2" = tests ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_${T}_gputests.log 2>&1; echo "gpu tests rc=$?" | tee -a gpurun_out/r2_${T}_gputests.log
  tail -3 gpurun_out/r2_${T}_gputests.log
fi
: > gpurun_out/r2_${T}_split.log
for rep in 1 2; do
  timeout 300 python tools/wf_split_time.py 3840x2160x16 >> gpurun_out/r2_${T}_split.log 2>&1
  for so in tools/_variants/*.so; do
    B200RT_LIB=$PWD/$so timeout 300 python tools/wf_split_time.py 3840x2160x16 >> gpurun_out/r2_${T}_split.log 2>&1
  done
done
cat gpurun_out/r2_${T}_split.log
