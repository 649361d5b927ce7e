#!/bin/bash
# Round 2: ingest tests + timing, K=30 with two steps per block, whole suite.
T=${1:-r2j}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_ingest.py -m gpu -q > gpurun_out/${T}_tests_ingest.log 2>&1
echo "ingest tests rc=$?"; tail -30 gpurun_out/${T}_tests_ingest.log
timeout 600 python tools/time_ingest.py 20000 20000 1000 > gpurun_out/${T}_ingest.json 2> gpurun_out/${T}_ingest.err
echo "ingest timing rc=$?"; cat gpurun_out/${T}_ingest.json; tail -3 gpurun_out/${T}_ingest.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/${T}_tests.log
for K in 30 32; do
  timeout 120 python bench.py --no-cpu --no-e2e --no-strong --factors $K --steps 20 --warmup 3 > gpurun_out/${T}_lanes_K$K.json 2> gpurun_out/${T}_lanes_K$K.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/${T}_lanes_K$K.json').read().strip().splitlines()[-1])
print('K=$K ms/step %.3f pair %.3f frac %.4f parity %s'%(d['ms_per_step'],d['roofline']['sweep_pair_ms'],d['roofline']['frac'],d['parity']['max_rel_vs_golden']))"
done
