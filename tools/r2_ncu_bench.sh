#!/bin/bash
# launch list of the bench command (device time of every launch; cold-cache, serialised: compare SHARES) and one --set full
# capture of the fused shading kernel, each after the same command exited 0 without ncu.
cd "$(dirname "$0")/.."
export B200RT_WF_GRAPH=0   # ncu does not profile kernel nodes of conditional graphs: the rounds are enqueued from the host (same kernels)
mkdir -p gpurun_out
T=${1:-r2}
CMD="python bench.py --steps 2 --warmup 1 --epochs 16 --no-cpu-baseline"
$CMD > gpurun_out/${T}_bench_e16_plain.json 2> gpurun_out/${T}_bench_e16_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_bench_launches.csv $CMD > gpurun_out/${T}_bench_e16_under_ncu.log 2>&1
tail -c 300 gpurun_out/${T}_bench_e16_plain.json; echo
python tools/launch_summary.py gpurun_out/${T}_bench_launches.csv 2>/dev/null | head -20
python tools/wf_profile_run.py 3840x2160x4 > gpurun_out/${T}_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wf_logic_kernel -s 14 -c 1 -o gpurun_out/prof_${T}_wf_logic -f python tools/wf_profile_run.py 3840x2160x4 > gpurun_out/${T}_ncu_logic.log 2>&1
tail -2 gpurun_out/${T}_ncu_logic.log
