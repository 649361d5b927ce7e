#!/bin/bash
T=${1:-r2x}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 200 python bench.py --no-cpu --no-strong --no-parity --no-e2e --cells 20000 --draws 1000 --steps 100 --warmup 5 > gpurun_out/${T}_mid.json 2> gpurun_out/${T}_mid.err
python -c "
import json
d=json.loads(open('gpurun_out/${T}_mid.json').read().strip().splitlines()[-1])
print('mid: ms/step %.4f pair %.4f share %.3f launches %d'%(d['ms_per_step'],d['roofline']['sweep_pair_ms'],d['roofline']['sweep_share_of_step'],d['gpu_launches']))"
timeout 200 python bench.py --no-cpu --no-strong --no-e2e > gpurun_out/${T}_cfg3.json 2> gpurun_out/${T}_cfg3.err
python -c "
import json
d=json.loads(open('gpurun_out/${T}_cfg3.json').read().strip().splitlines()[-1])
print('cfg3: ms/step %.4f pair %.4f share %.3f launches %d parity %s'%(d['ms_per_step'],d['roofline']['sweep_pair_ms'],d['roofline']['sweep_share_of_step'],d['gpu_launches'],d['parity']['max_rel_vs_golden']))"
