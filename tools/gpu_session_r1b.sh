#!/bin/bash
# One gpurun call: new GPU tests first, then the whole GPU suite, the default bench line,
# the sweep build variants (tools: SCHPF_BUILD_TAG / SCHPF_NVCC_FLAGS, built beforehand), the
# minibatch timing.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1b_smi.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_minibatch.py -x -q > gpurun_out/r1b_tests_minibatch.log 2>&1
echo "minibatch tests rc=$?" | tee -a gpurun_out/r1b_summary.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1b_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/r1b_summary.txt
timeout 400 python bench.py --steps 30 --warmup 3 > gpurun_out/r1b_bench_default.json 2> gpurun_out/r1b_bench_default.err
echo "bench default rc=$?" | tee -a gpurun_out/r1b_summary.txt
for tag in u2 stag u2stag; do
  SCHPF_B200_LIB=$PWD/schpf_b200/_C_$tag/libschpf_b200.so timeout 200 python bench.py --no-cpu --no-e2e --steps 30 --warmup 3 \
      > gpurun_out/r1b_bench_$tag.json 2> gpurun_out/r1b_bench_$tag.err
  echo "bench $tag rc=$?" | tee -a gpurun_out/r1b_summary.txt
done
timeout 300 python tools/time_minibatch.py > gpurun_out/r1b_minibatch_timing.json 2> gpurun_out/r1b_minibatch_timing.err
echo "minibatch timing rc=$?" | tee -a gpurun_out/r1b_summary.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_tests_gpu.log 2>&1
echo "gpu suite rc=$?" | tee -a gpurun_out/r1b_summary.txt
tail -3 gpurun_out/r1b_tests_minibatch.log gpurun_out/r1b_tests_gpu.log
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r1b_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], "pair %.3f" % d["roofline"]["sweep_pair_ms"], "frac %.4f" % d["roofline"]["frac"],
              "e2e", d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
P
cat gpurun_out/r1b_minibatch_timing.json
