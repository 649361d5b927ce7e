#!/bin/bash
# 2-GPU call: the sharded NCCL tests (three exchange modes) and the N=2 bench with and without overlap
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 400 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r1h_tests_sharded.log 2>&1
echo "sharded tests rc=$?"; tail -3 gpurun_out/r1h_tests_sharded.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 --no-e2e > gpurun_out/r1h_bench_n2.json 2> gpurun_out/r1h_bench_n2.err
echo "n2 overlap rc=$?"
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 --no-e2e --no-overlap > gpurun_out/r1h_bench_n2_inorder.json 2> gpurun_out/r1h_bench_n2_inorder.err
echo "n2 in-order rc=$?"
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r1h_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f pair %.3f value %.3e" % (d["ms_per_step"], d["roofline"]["sweep_pair_ms"], d["value"]), d["config"]["parallelism"])
    except Exception as e:
        print(f, "unreadable", e)
P
