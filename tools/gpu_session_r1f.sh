#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for tag in dmma rcp4 both; do
  export SCHPF_B200_LIB=$PWD/schpf_b200/_C_$tag/libschpf_b200.so
  timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_minibatch.py -x -q > gpurun_out/r1f_tests_$tag.log 2>&1
  echo "$tag tests rc=$?"; tail -2 gpurun_out/r1f_tests_$tag.log
  timeout 200 python bench.py --no-cpu --no-e2e --steps 30 --warmup 3 > gpurun_out/r1f_bench_$tag.json 2> gpurun_out/r1f_bench_$tag.err
  echo "$tag bench rc=$?"
  timeout 200 python bench.py --no-cpu --no-e2e --factors 30 --steps 20 --warmup 3 > gpurun_out/r1f_bench_${tag}_K30.json 2> gpurun_out/r1f_bench_${tag}_K30.err
done
unset SCHPF_B200_LIB
timeout 200 python bench.py --no-cpu --no-e2e --steps 30 --warmup 3 > gpurun_out/r1f_bench_default.json 2> gpurun_out/r1f_bench_default.err
timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -k float32 > gpurun_out/r1f_tests_fp32.log 2>&1; echo "fp32 test rc=$?"
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r1f_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f pair %.3f frac %.4f" % (d["ms_per_step"], d["roofline"]["sweep_pair_ms"], d["roofline"]["frac"]))
    except Exception as e:
        print(f, "unreadable", e)
P
