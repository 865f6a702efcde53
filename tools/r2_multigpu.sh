#!/bin/bash
# multi-GPU checks on one box (gpurun --gpus N): the single-process device group test, bench.py under torchrun (C4; the
# reference arm; C5 through the acceleration structure)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_group.py -x -q -s 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_c4_n$N.json 2> gpurun_out/r2_bench_c4_n$N.err
tail -c 1200 gpurun_out/r2_bench_c4_n$N.json; tail -3 gpurun_out/r2_bench_c4_n$N.err
timeout 600 $TR --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/r2_bench_ref_n$N.json 2> gpurun_out/r2_bench_ref_n$N.err
tail -c 600 gpurun_out/r2_bench_ref_n$N.json
timeout 900 $TR --master-port 29519 bench.py --workload c5 --cast bvh --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_c5_bvh_n$N.json 2> gpurun_out/r2_bench_c5_bvh_n$N.err
tail -c 900 gpurun_out/r2_bench_c5_bvh_n$N.json; tail -3 gpurun_out/r2_bench_c5_bvh_n$N.err
