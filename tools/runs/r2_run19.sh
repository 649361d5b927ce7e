#!/bin/bash
# Round 2, final check of the tree as committed: whole GPU suite, smoke, the default bench line, the fp32 line.
T=${1:-r2t}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/${T}_smoke.log
SCHPF_BENCH_VERBOSE=1 timeout 900 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
echo "bench rc=$?"; cut -c1-700 gpurun_out/${T}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
echo "reference arm rc=$?"; cut -c1-400 gpurun_out/${T}_bench_reference.json
timeout 600 python bench.py --precision 32 --no-strong > gpurun_out/${T}_bench_f32_n1.json 2> gpurun_out/${T}_bench_f32_n1.err
echo "f32 bench rc=$?"; cut -c1-300 gpurun_out/${T}_bench_f32_n1.json
for K in 7 10 15 16 30 50; do
  timeout 120 python bench.py --no-cpu --no-e2e --no-strong --factors $K --steps 20 --warmup 3 > gpurun_out/${T}_bench_K$K.json 2> gpurun_out/${T}_bench_K$K.err
  python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench_K$K.json').read().strip().splitlines()[-1])
print('K=$K ms/step %.3f pair %.3f frac %.4f parity %s'%(d['ms_per_step'],d['roofline']['sweep_pair_ms'],d['roofline']['frac'],d['parity']['ok']))"
done
