#!/bin/bash
# Build-variant sweep of the sweep kernel (next round's first GPU call).
#
#   here (no GPU):   bash tools/sweep_variants.sh build
#   on the box:      /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/sweep_variants.sh run'
#
# Each variant is a tagged library schpf_b200/_C_<tag>/libschpf_b200.so (git-ignored, shipped by
# gpurun).  `run` benches every (variant, K) pair of the table with --no-cpu --no-e2e (~10 s each)
# and, for variants that beat the default, the GPU parity tests should then be run with
# SCHPF_B200_LIB pointing at the variant before it becomes a default in csrc/sweep.cu.
#
# Candidates (registers / spills from `ptxas -v`, end of round 1):
#   s4small16 : four steps per block for KP <= 12 at 16 warps      (KP=8: 122 regs, KP=12: 128, no spills)
#   s4small12 : the same for KP <= 16 at 12 warps                  (KP=16: 164 regs, no spills)
#   s4mid8    : four steps per block for 20 < KP <= 32 at 8 warps  (224 / 240 / 252 regs, no spills;
#               the default 12 x 168 spills 100 bytes at KP=32)
cd "$(dirname "$0")/.."
declare -A FLAGS=(
  [s4small16]="-DSWEEP_S4_LO=4 -DSWEEP_S4_HI=20"
  [s4small12]="-DSWEEP_S4_LO=4 -DSWEEP_S4_HI=20 -DSWEEP_WSMALL=12"
  [s4mid8]="-DSWEEP_S4_LO=20 -DSWEEP_S4_HI=32 -DSWEEP_WMID=8"
)
declare -A KS=( [s4small16]="7 10" [s4small12]="7 10 15" [s4mid8]="24 28 30" [default]="7 10 15 20 24 28 30 50" )
case "$1" in
  build)
    for tag in "${!FLAGS[@]}"; do
      SCHPF_BUILD_TAG=$tag SCHPF_NVCC_FLAGS="${FLAGS[$tag]}" python -m schpf_b200.build | tail -1
    done ;;
  run)
    mkdir -p gpurun_out
    for tag in default "${!FLAGS[@]}"; do
      lib=$PWD/schpf_b200/_C${tag:+_$tag}/libschpf_b200.so
      [ "$tag" = default ] && lib=$PWD/schpf_b200/_C/libschpf_b200.so
      for K in ${KS[$tag]}; do
        SCHPF_B200_LIB=$lib timeout 120 python bench.py --no-cpu --no-e2e --factors $K --steps 20 --warmup 3 \
            > gpurun_out/var_${tag}_K$K.json 2> gpurun_out/var_${tag}_K$K.err
      done
    done
    python - <<'P'
import glob, json
for f in sorted(glob.glob("gpurun_out/var_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-40s ms/step %.3f  pair %.3f  frac %.4f  loss %s" % (f[11:-5], d["ms_per_step"], d["roofline"]["sweep_pair_ms"],
              d["roofline"]["frac"], d["loss_first_last"]))
    except Exception as e:
        print(f, "unreadable", e)
P
    ;;
  *) echo "usage: $0 build|run" ;;
esac
