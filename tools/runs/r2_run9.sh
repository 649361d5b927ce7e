#!/bin/bash
# Round 2: whole GPU suite (fp32 tests included), then build variants of the lane sweep at K = 16 / 20 / 30.
T=${1:-r2i}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/${T}_tests.log
run() {   # tag lib K extra...
  local tag=$1 lib=$2 K=$3; shift 3
  SCHPF_B200_LIB=$lib timeout 120 python bench.py --no-cpu --no-e2e --no-strong --no-parity --factors $K --steps 20 --warmup 3 "$@" \
      > gpurun_out/${T}_${tag}_K$K.json 2> gpurun_out/${T}_${tag}_K$K.err
}
D=$PWD/schpf_b200/_C/libschpf_b200.so
for K in 16 20 30; do run base $D $K; done
for tag in ctas2 pf0 l2a16 ns32_2; do
  L=$PWD/schpf_b200/_C_$tag/libschpf_b200.so
  [ -f $L ] || continue
  case $tag in
    ns32_2) Ks="30" ;;
    *) Ks="16 20 30" ;;
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
