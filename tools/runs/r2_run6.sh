#!/bin/bash
# Round 2, N-GPU call: the NCCL parity tests (need >= 2 GPUs) and the default bench line at N ranks.
T=${1:-r2f}; N=${2:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_minibatch.py -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
SCHPF_BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
echo "bench N=$N rc=$?"; cut -c1-400 gpurun_out/${T}_bench_n$N.json; grep "bench r0" gpurun_out/${T}_bench_n$N.err | tail -12
python - "$T" "$N" <<'P'
import json, sys
d = json.loads(open("gpurun_out/%s_bench_n%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
print("value %.4g  ms/step %.3f  e2e %.4g  cold %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["cold"]["value"]))
print("parity", d["parity"])
for c in d["strong"]["cases"]:
    print({k: c.get(k) for k in ("workload", "ms_per_step", "value", "speedup_vs_1gpu", "sweep_pair_ms_rank0", "error")})
P
