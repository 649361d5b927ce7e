#!/bin/bash
# What a gpurun call of this repo looks like (one GPU):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/gpu_check.sh'
# GPU test suite, smoke, the default bench line, one `ncu --set full` capture of the sweep
# kernels and the launch list of the engine's kernels.  Everything lands in gpurun_out/.
#
# Build variants: SCHPF_BUILD_TAG=<tag> SCHPF_NVCC_FLAGS="-DSWEEP_...=..." python -m schpf_b200.build
# (here, before the call) gives schpf_b200/_C_<tag>/libschpf_b200.so; on the box
# SCHPF_B200_LIB=$PWD/schpf_b200/_C_<tag>/libschpf_b200.so python bench.py --no-cpu --no-e2e
# measures it (~10 s each) and the same variable makes pytest exercise it.
T=${1:-check}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -3 gpurun_out/${T}_tests_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 400 python bench.py --steps 30 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
echo "bench rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 9 -c 2 -f -o gpurun_out/${T}_sweep \
    python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > gpurun_out/${T}_ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k 'regex:sweep_kernel|finalize_kernel|fold_kernel|ex_table_kernel|sum_partials_kernel|pack_loss' -c 120 \
    --csv --log-file gpurun_out/${T}_launches.csv python bench.py --no-cpu --no-e2e --steps 8 --warmup 1 > gpurun_out/${T}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python - "$T" <<'P'
import json, sys
d = json.loads(open("gpurun_out/%s_bench_n1.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.3f  sweep pair %.3f ms  roofline frac %.4f  value %.3e  e2e %.3e  clocks %s" % (
    d["ms_per_step"], d["roofline"]["sweep_pair_ms"], d["roofline"]["frac"], d["value"], d["e2e"]["value"], d["clocks"]))
P
