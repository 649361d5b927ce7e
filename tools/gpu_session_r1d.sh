#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
SCHPF_TRACE=1 timeout 300 python tools/diag_layout.py > gpurun_out/r1d_diag_layout.txt 2> gpurun_out/r1d_diag_layout.trace
echo "diag rc=$?"; cat gpurun_out/r1d_diag_layout.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1d_tests_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -3 gpurun_out/r1d_tests_gpu.log
timeout 400 python bench.py --steps 30 --warmup 3 > gpurun_out/r1d_bench_default.json 2> gpurun_out/r1d_bench_default.err
echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r1d_bench_default.json").read().strip().splitlines()[-1])
print("ms/step %.3f pair %.3f frac %.4f e2e %.3e (%.3f s) value %.3e layout %.3f" % (d["ms_per_step"], d["roofline"]["sweep_pair_ms"],
      d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["seconds"], d["value"], d["config"]["layout_build_s"]), d["clocks"])
P
