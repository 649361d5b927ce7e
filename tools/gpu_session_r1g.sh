#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B="python bench.py --no-cpu --no-e2e --steps 20 --warmup 3"
for K in 24 28 30; do
  timeout 200 $B --factors $K > gpurun_out/r1g_bench_default_K$K.json 2> gpurun_out/r1g_bench_default_K$K.err
  SCHPF_B200_LIB=$PWD/schpf_b200/_C_mid16/libschpf_b200.so timeout 200 $B --factors $K > gpurun_out/r1g_bench_mid16_K$K.json 2> gpurun_out/r1g_bench_mid16_K$K.err
done
SCHPF_B200_LIB=$PWD/schpf_b200/_C_cta2/libschpf_b200.so timeout 200 $B > gpurun_out/r1g_bench_cta2_K20.json 2> gpurun_out/r1g_bench_cta2_K20.err
timeout 200 $B > gpurun_out/r1g_bench_default_K20.json 2> gpurun_out/r1g_bench_default_K20.err
timeout 200 $B --target-ctas 2368 > gpurun_out/r1g_bench_t2368_K20.json 2> gpurun_out/r1g_bench_t2368_K20.err
timeout 200 $B --target-ctas 9472 > gpurun_out/r1g_bench_t9472_K20.json 2> gpurun_out/r1g_bench_t9472_K20.err
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r1g_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f pair %.3f frac %.4f" % (d["ms_per_step"], d["roofline"]["sweep_pair_ms"], d["roofline"]["frac"]), d["config"]["layout"]["panel_rows"], d["config"]["layout"]["grid_cells"])
    except Exception as e:
        print(f, "unreadable", e)
P
