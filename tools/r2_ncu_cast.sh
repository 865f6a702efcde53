#!/bin/bash
# ncu captures of the wavefront cast kernel on the fixed profiling workload, each after the same command exited 0 without ncu:
#   1. --set full of one large wf_cast_rl_kernel launch (tools/wf_profile_run.py 3840x2160x4, third cast launch)
#   2. DRAM / time / pipe metrics of every cast launch of one 16-epoch 4K batch (the bench's batch size)
cd "$(dirname "$0")/.."
export B200RT_WF_GRAPH=0   # ncu does not profile kernel nodes of conditional graphs: the rounds are enqueued from the host (same kernels)
mkdir -p gpurun_out
T=${1:-r2a}
python tools/wf_profile_run.py 3840x2160x4 > gpurun_out/${T}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wf_cast_rl -s 2 -c 1 -o gpurun_out/prof_${T}_wf_cast_rl -f python tools/wf_profile_run.py 3840x2160x4 > gpurun_out/${T}_ncu_cast.log 2>&1
tail -2 gpurun_out/${T}_ncu_cast.log
python tools/wf_profile_run.py 3840x2160x16 > gpurun_out/${T}_plain16.log 2>&1 &&
ncu --clock-control none -k regex:wf_cast_rl -c 24 --csv --log-file gpurun_out/${T}_cast_metrics.csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum \
    python tools/wf_profile_run.py 3840x2160x16 > gpurun_out/${T}_ncu_metrics.log 2>&1
tail -2 gpurun_out/${T}_ncu_metrics.log
