#!/bin/bash
# Round 2, second GPU call: the one-lane-per-owner sweep -- parity tests, then lanes on / off and
# the build variants at K = 15, 20, 30.
T=${1:-r2b}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_engine.py -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/${T}_tests.log
run() {   # tag lib K extra...
  local tag=$1 lib=$2 K=$3; shift 3
  SCHPF_B200_LIB=$lib timeout 120 python bench.py --no-cpu --no-e2e --factors $K --steps 20 --warmup 3 "$@" \
      > gpurun_out/${T}_${tag}_K$K.json 2> gpurun_out/${T}_${tag}_K$K.err
}
D=$PWD/schpf_b200/_C/libschpf_b200.so
for K in 15 16 20 30 32; do run lanes $D $K; run pairs $D $K --lanes 0; done
run lanes_rank1 $D 20 --rank-per-range 1
run lanes_rank0 $D 30 --rank-per-range 0
run lanes_rank0 $D 16 --rank-per-range 0
for tag in w20_8 w16_12 nobulk ns20_1 w32_ns2; do
  L=$PWD/schpf_b200/_C_$tag/libschpf_b200.so
  [ -f $L ] || continue
  case $tag in
    w20_8|ns20_1) Ks="20" ;;
    w16_12) Ks="15 16" ;;
    nobulk) Ks="16 20 30" ;;
    w32_ns2) Ks="30" ;;
  esac
  for K in $Ks; do run $tag $L $K; done
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
