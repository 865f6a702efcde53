#!/bin/bash
# Builds the tracer for 2/3/4 resident CTAs per SM on the GPU box and benches each (dev tool).
for mb in ${1:-2 3 4}; do
  echo "=== B200RT_TRACE_MIN_BLOCKS=$mb"
  B200RT_TRACE_MIN_BLOCKS=$mb python homework-18-graphics-raytracer_b200/build.py --force --verbose 2>&1 | grep -E "trace_kernelILi[01]ELi0|Used (1|2)[0-9][0-9] reg|spill" | grep -v "Function prop" | head -8
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c1_whitted or distributed_samples" 2>&1 | tail -1
  python bench.py --epochs 8 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4x8ep ms/step', round(d['ms_per_step'],1), 'Mrays/s', round(d['value'],1), 'roof', round(d['roofline']['frac'],4))"
  python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 ms/step', round(d['ms_per_step'],2), 'Mrays/s', round(d['value'],1), 'roof', round(d['roofline']['frac'],4))"
done
