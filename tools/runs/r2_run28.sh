#!/bin/bash
# layout sort on (list, class) only: whole suite, traced e2e, default K=20 line without the CPU leg
T=${1:-r3m}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/${T}_tests.log
timeout 300 python bench.py --no-cpu --no-strong > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('ms/step %.4f pair %.4f frac %.4f e2e %.4g (%.4fs) layout %.4fs parity %s'%(d['ms_per_step'],d['roofline']['sweep_pair_ms'],d['roofline']['frac'],d['e2e']['value'],d['e2e']['seconds'],d['config']['layout_build_s'],d['parity']['max_rel_vs_golden']))"
SCHPF_TRACE=1 timeout 200 python bench.py --no-cpu --no-strong --no-parity --no-e2e > /dev/null 2> gpurun_out/${T}_trace.err; grep "schpf trace" gpurun_out/${T}_trace.err | tail -14
