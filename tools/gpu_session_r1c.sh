#!/bin/bash
# gpurun call: layout-build diagnosis, the default bench line (new default sweep), one ncu
# --set full capture of the sweep kernels with source, and the launch list.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
SCHPF_TRACE=1 timeout 300 python tools/diag_layout.py > gpurun_out/r1c_diag_layout.txt 2> gpurun_out/r1c_diag_layout.trace
echo "diag rc=$?"
timeout 400 python bench.py --steps 30 --warmup 3 > gpurun_out/r1c_bench_default.json 2> gpurun_out/r1c_bench_default.err
echo "bench rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 9 -c 3 -f -o gpurun_out/r1c_sweep \
    python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > gpurun_out/r1c_ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r1c_launches.csv \
    python bench.py --no-cpu --no-e2e --steps 4 --warmup 1 > gpurun_out/r1c_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
cat gpurun_out/r1c_diag_layout.txt
python - <<'P'
import json
d = json.loads(open("gpurun_out/r1c_bench_default.json").read().strip().splitlines()[-1])
print("ms/step %.3f pair %.3f frac %.4f e2e %.3e (%.3f s) value %.3e" % (d["ms_per_step"], d["roofline"]["sweep_pair_ms"],
      d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["seconds"], d["value"]), d["clocks"])
P
ls -la gpurun_out/ | tail -12
