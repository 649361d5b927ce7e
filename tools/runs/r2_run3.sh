#!/bin/bash
# Round 2, third GPU call: where does the one-lane sweep's time go?  (row loads alone, arithmetic
# alone, both), the schedule-free K=20 stream, 16 warps, and ncu captures of the lane kernels.
T=${1:-r2c}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_engine.py -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/${T}_tests.log
run() {   # tag lib K extra...
  local tag=$1 lib=$2 K=$3; shift 3
  SCHPF_B200_LIB=$lib timeout 120 python bench.py --no-cpu --no-e2e --no-strong --no-parity --factors $K --steps 20 --warmup 3 "$@" \
      > gpurun_out/${T}_${tag}_K$K.json 2> gpurun_out/${T}_${tag}_K$K.err
}
D=$PWD/schpf_b200/_C/libschpf_b200.so
for K in 16 20 30; do run lanes $D $K; done
run free $D 20 --free-schedule 1
run free_norank $D 20 --free-schedule 1 --rank-per-range 0
run pairs $D 20 --lanes 0
for tag in exp1 exp2; do for K in 16 20 30; do run $tag $PWD/schpf_b200/_C_$tag/libschpf_b200.so $K; done; done
run w20_16 $PWD/schpf_b200/_C_w20_16/libschpf_b200.so 20
for K in 20 16; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:lane_sweep_kernel -s 9 -c 2 -f \
      -o gpurun_out/${T}_lanes_K$K python bench.py --no-cpu --no-e2e --no-strong --no-parity --factors $K --steps 3 --warmup 3 \
      > gpurun_out/${T}_ncu_K$K.log 2>&1
  echo "ncu K=$K rc=$?"
done
python - "$T" <<'P'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/%s_*_K*.json" % sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        lay = d["config"]["layout"]
        print("%-36s ms/step %.3f  pair %.3f  frac %.4f  pad %.3f/%.3f  loss %s" % (
            f[11:-5], d["ms_per_step"], d["roofline"]["sweep_pair_ms"], d["roofline"]["frac"],
            lay["padded_nnz_cells"] / d["config"]["nnz_per_gpu"] - 1, lay["padded_nnz_genes"] / d["config"]["nnz_per_gpu"] - 1,
            d["loss_first_last"]))
    except Exception as e:
        print(f, "unreadable", e, open(f[:-5] + ".err").read()[-300:])
P
