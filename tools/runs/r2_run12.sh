#!/bin/bash
T=${1:-r2l}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 600 python tools/time_ingest.py 20000 20000 1000 > gpurun_out/${T}_ingest.json 2> gpurun_out/${T}_ingest.err
echo "ingest timing rc=$?"; cat gpurun_out/${T}_ingest.json; tail -3 gpurun_out/${T}_ingest.err
SCHPF_TRACE=1 SCHPF_BENCH_VERBOSE=1 timeout 300 python bench.py --no-cpu --no-strong --no-parity > gpurun_out/${T}_trace.json 2> gpurun_out/${T}_trace.err
echo "trace rc=$?"; grep -v "^\[schpf trace\]   " gpurun_out/${T}_trace.err | tail -80
for K in 7 10; do
  timeout 120 python bench.py --no-cpu --no-e2e --no-strong --factors $K --steps 20 --warmup 3 > gpurun_out/${T}_lanes_K$K.json 2> gpurun_out/${T}_lanes_K$K.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/${T}_lanes_K$K.json').read().strip().splitlines()[-1])
print('K=$K ms/step %.3f pair %.3f frac %.4f parity %s'%(d['ms_per_step'],d['roofline']['sweep_pair_ms'],d['roofline']['frac'],d['parity']['max_rel_vs_golden']))"
done
