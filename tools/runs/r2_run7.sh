#!/bin/bash
# Round 2: the default bench line at N ranks, then the K sweep (cfg-5) at N ranks; N-GPU box.
T=${1:-r2g}; N=${2:-8}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
SCHPF_BENCH_VERBOSE=1 timeout 600 $TR bench.py --gpus $N > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
echo "bench N=$N rc=$?"; grep "bench r0" gpurun_out/${T}_bench_n$N.err | tail -12
python - "$T" "$N" <<'P'
import json, sys
d = json.loads(open("gpurun_out/%s_bench_n%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
print("value %.4g  ms/step %.3f  e2e %.4g  cold %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["cold"]["value"]))
print("parity", d["parity"])
for c in d["strong"]["cases"]:
    print({k: c.get(k) for k in ("workload", "ms_per_step", "value", "speedup_vs_1gpu", "sweep_pair_ms_rank0", "error")})
P
for K in 7 10 15 30 50; do
  timeout 200 $TR bench.py --gpus $N --no-cpu --no-e2e --no-strong --factors $K --steps 20 --warmup 3 \
      > gpurun_out/${T}_n${N}_K$K.json 2> gpurun_out/${T}_n${N}_K$K.err
  python -c "
import json
d=json.loads(open('gpurun_out/${T}_n${N}_K$K.json').read().strip().splitlines()[-1])
print('N=$N K=$K ms/step %.3f value %.4g pair %.3f parity %s'%(d['ms_per_step'],d['value'],d['roofline']['sweep_pair_ms'],d['parity']['ok']))" || tail -3 gpurun_out/${T}_n${N}_K$K.err
done
