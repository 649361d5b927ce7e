#!/bin/bash
# Round 2, final captures: ncu --set full of the sweep kernels the bench times (library version 210),
# finalize kernel, launch list of the default step, cfg-2 line.
T=${1:-r2s}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
cap() {   # name regex extra bench flags...
  local name=$1 rx=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s 9 -c 2 -f \
      -o gpurun_out/${T}_$name python bench.py --no-cpu --no-e2e --no-strong --no-parity --steps 3 --warmup 3 "$@" \
      > gpurun_out/${T}_ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
}
cap lanes_K20 lane_sweep_kernel
cap lanes_K30 lane_sweep_kernel --factors 30
cap lanes_K16 lane_sweep_kernel --factors 16
cap f32_K20 lane_sweep_f32_kernel --precision 32
timeout 300 ncu --set full --clock-control none --import-source on -k regex:finalize_kernel -s 8 -c 2 -f \
    -o gpurun_out/${T}_finalize python bench.py --no-cpu --no-e2e --no-strong --no-parity --steps 3 --warmup 3 \
    > gpurun_out/${T}_ncu_finalize.log 2>&1
echo "ncu finalize rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sweep|finalize|fold|fixup|partials|pack_loss|ex_table|prep' -c 200 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --no-cpu --no-e2e --no-strong --no-parity --steps 10 --warmup 3 > gpurun_out/${T}_launches.log 2>&1
echo "launch list rc=$?"
timeout 200 python bench.py --no-cpu --no-strong --cells 50000 --draws 1500 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
echo "cfg-2 rc=$?"; cut -c1-300 gpurun_out/${T}_bench_cfg2.json
