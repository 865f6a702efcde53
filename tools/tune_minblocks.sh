#!/bin/bash
# Builds the tracer for 2/3/4 resident CTAs per SM on the GPU box, reports text size, instruction-cache hit rate and time.
for mb in ${1:-2 3 4}; do
  echo "=== B200RT_TRACE_MIN_BLOCKS=$mb"
  B200RT_TRACE_MIN_BLOCKS=$mb python homework-18-graphics-raytracer_b200/build.py --force --verbose 2>&1 | grep -E "Used (1|2)[0-9][0-9] reg" | head -4 | tr '\n' ' '; echo
  cuobjdump -elf homework-18-graphics-raytracer_b200/_lib/obj/rt_kernels.o 2>/dev/null | grep -E " \.text\._ZN6b200rt12trace_kernelILi[01]ELi0" | awk '{printf "text bytes 0x%s ", $3}'; echo
  python bench.py --epochs 8 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4x8ep ms/step', round(d['ms_per_step'],1), 'Mrays/s', round(d['value'],1), 'roof', round(d['roofline']['frac'],4))"
  python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 ms/step', round(d['ms_per_step'],2), 'Mrays/s', round(d['value'],1), 'roof', round(d['roofline']['frac'],4))"
  ncu --metrics sm__icc_request_hit_rate.pct,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:trace_kernel -c 1 python tools/profile_run.py distributed 2>&1 | grep -E "icc_request|time_duration|issue_active"
done
