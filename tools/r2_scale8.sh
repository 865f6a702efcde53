#!/bin/bash
# 8-GPU evidence on one box (gpurun --gpus 8): C4 (epochs sharded, ncclReduce) and C5 (rows sharded, gather) through the
# device groups of the C ABI, bench.py under torchrun, one rank per GPU.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_c4_n$N.json 2> gpurun_out/r2_bench_c4_n$N.err
tail -c 1800 gpurun_out/r2_bench_c4_n$N.json; tail -3 gpurun_out/r2_bench_c4_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --workload c5 --gpus $N --steps 3 --warmup 2 > gpurun_out/r2_bench_c5_n$N.json 2> gpurun_out/r2_bench_c5_n$N.err
tail -c 1800 gpurun_out/r2_bench_c5_n$N.json; tail -3 gpurun_out/r2_bench_c5_n$N.err
timeout 600 python -m pytest tests/test_gpu_group.py -x -q 2>&1 | tail -3
