#!/bin/bash
# round-robin class interleave of the K 17..20 stream: tests, K=20 bench, bank conflicts
T=${1:-r3j}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_engine.py -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/${T}_tests.log
for K in 20 17; do
timeout 120 python bench.py --no-cpu --no-e2e --no-strong --factors $K --steps 30 --warmup 3 > gpurun_out/${T}_K$K.json 2> gpurun_out/${T}_K$K.err
python -c "
import json
d=json.loads(open('gpurun_out/${T}_K$K.json').read().strip().splitlines()[-1])
print('K=$K ms/step %.4f pair %.4f frac %.4f parity %s layout_s %.4f'%(d['ms_per_step'],d['roofline']['sweep_pair_ms'],d['roofline']['frac'],d['parity']['max_rel_vs_golden'],d['config']['layout_build_s']))"
done
timeout 200 ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum \
   --clock-control none -k regex:lane_sweep_kernel -s 9 -c 2 --csv --log-file gpurun_out/${T}_conflicts.csv \
   python bench.py --no-cpu --no-e2e --no-strong --no-parity --steps 3 --warmup 3 > gpurun_out/${T}_ncu.log 2>&1
grep -E "bank_conflicts|wavefronts|time_duration" gpurun_out/${T}_conflicts.csv | cut -d, -f5,13,14,15 | head -8
