#!/bin/bash
for t in ${1:-128 256 512}; do
  echo "=== B200RT_DIST_THREADS=$t"
  B200RT_DIST_THREADS=$t python homework-18-graphics-raytracer_b200/build.py --force --verbose 2>&1 | grep -E "Used (1|2)[0-9][0-9] reg" | head -2 | tr '\n' ' '; echo
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "distributed" 2>&1 | tail -1
  python bench.py --epochs 8 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4x8ep ms/step', round(d['ms_per_step'],1), 'Mrays/s', round(d['value'],1), 'roof', round(d['roofline']['frac'],4))"
  ncu --metrics sm__icc_request_hit_rate.pct,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:trace_kernel -c 1 python tools/profile_run.py distributed 2>&1 | grep -E "icc_request|time_duration|issue_active"
done
