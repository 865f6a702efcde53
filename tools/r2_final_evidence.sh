#!/bin/bash
# Round-2 evidence on one B200, every ncu pass after the same command exited 0 without ncu (one ncu tool per call).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2f}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_gputests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/${T}_gputests.log; tail -2 gpurun_out/${T}_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -c 300 gpurun_out/${T}_smoke.log; echo
python bench.py > gpurun_out/${T}_bench_c4_n1.json 2> gpurun_out/${T}_bench_c4_n1.err; tail -c 400 gpurun_out/${T}_bench_c4_n1.json; echo
python bench.py --impl reference > gpurun_out/${T}_bench_c4_reference.json 2> gpurun_out/${T}_bench_c4_reference.err; tail -c 300 gpurun_out/${T}_bench_c4_reference.json; echo
python bench.py --workload k2 > gpurun_out/${T}_bench_k2_n1.json 2> gpurun_out/${T}_bench_k2_n1.err; tail -c 700 gpurun_out/${T}_bench_k2_n1.json; echo
for w in c1 c2 c3; do python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/${T}_bench_${w}_n1.json 2> gpurun_out/${T}_bench_${w}_n1.err; python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/${T}_bench_${w}_n1.json') if l.startswith('{')][0]); print('$w', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"; done
python bench.py --workload c5 --steps 2 --warmup 3 > gpurun_out/${T}_bench_c5_n1.json 2> gpurun_out/${T}_bench_c5_n1.err; tail -c 500 gpurun_out/${T}_bench_c5_n1.json; echo
timeout 300 python tools/c5_bvh_bench.py > gpurun_out/${T}_c5_bvh.txt 2>&1; cat gpurun_out/${T}_c5_bvh.txt
timeout 120 python tools/intersect_bench_mesh.py > gpurun_out/${T}_k2_mesh.txt 2>&1; tail -4 gpurun_out/${T}_k2_mesh.txt
bash tools/r2_ncu_bench.sh ${T}
bash tools/r2_ncu_cast.sh ${T}
