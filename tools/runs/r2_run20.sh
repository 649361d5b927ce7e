#!/bin/bash
# Round 2: the default bench line at N ranks (library version 210).
T=${1:-r2u}; N=${2:-8}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
SCHPF_BENCH_VERBOSE=1 timeout 600 $TR bench.py --gpus $N > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
echo "bench N=$N rc=$?"; grep "bench r0" gpurun_out/${T}_bench_n$N.err | tail -4
python - "$T" "$N" <<'P'
import json, sys
d = json.loads(open("gpurun_out/%s_bench_n%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
print("value %.4g  ms/step %.3f  e2e %.4g (%.3fs) cold %.3fs" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["seconds"], d["e2e"]["cold"]["seconds"]))
print("parity", d["parity"]["ok"], d["parity"]["max_rel_vs_golden"], d["parity"]["beta_replicas_bit_identical"])
for c in d["strong"]["cases"]:
    print({k: c.get(k) for k in ("workload", "ms_per_step", "value", "speedup_vs_1gpu", "sweep_pair_ms_rank0", "error")})
P
