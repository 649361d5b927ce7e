#!/bin/bash
# last captures of the round: ncu --set full of the K=20 sweeps with the final layout, launch list
T=${1:-r3k}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lane_sweep_kernel -s 9 -c 2 -f \
    -o gpurun_out/${T}_lanes_K20 python bench.py --no-cpu --no-e2e --no-strong --no-parity --steps 3 --warmup 3 \
    > gpurun_out/${T}_ncu_K20.log 2>&1
echo "ncu K=20 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sweep|finalize|fold|fixup|partials|pack_loss|ex_table|prep' -c 200 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --no-cpu --no-e2e --no-strong --no-parity --steps 10 --warmup 3 > gpurun_out/${T}_launches.log 2>&1
echo "launch list rc=$?"
