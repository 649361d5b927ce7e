#!/bin/bash
# Round 2, first GPU call (one B200): the whole GPU suite with the device-state minibatch
# tests un-gated, the default bench line, `ncu --set full` captures of the sweep the driver
# times (K=20, 12 warps) and of K=30, the launch list, and the build-variant sweep.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r2_run1.sh'
T=${1:-r2a}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt 2>&1
SCHPF_TEST_UNVALIDATED=1 timeout 700 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -5 gpurun_out/${T}_tests_gpu.log
timeout 300 python bench.py --steps 30 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/${T}_bench_n1.json
for K in 20 30; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 9 -c 2 -f \
      -o gpurun_out/${T}_sweep_K$K python bench.py --no-cpu --no-e2e --factors $K --steps 3 --warmup 3 \
      > gpurun_out/${T}_ncu_full_K$K.log 2>&1
  echo "ncu full K=$K rc=$?"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k 'regex:sweep_kernel|finalize_kernel|fold_kernel|ex_table_kernel|sum_partials_kernel|pack_loss' -c 120 \
    --csv --log-file gpurun_out/${T}_launches.csv python bench.py --no-cpu --no-e2e --steps 8 --warmup 1 \
    > gpurun_out/${T}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 500 bash tools/sweep_variants.sh run
