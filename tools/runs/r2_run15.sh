#!/bin/bash
T=${1:-r2p}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
run() {   # tag lib K extra...
  local tag=$1 lib=$2 K=$3; shift 3
  SCHPF_B200_LIB=$lib timeout 120 python bench.py --no-cpu --no-e2e --no-strong --no-parity --factors $K --steps 20 --warmup 3 "$@" \
      > gpurun_out/${T}_${tag}_K$K.json 2> gpurun_out/${T}_${tag}_K$K.err
}
D=$PWD/schpf_b200/_C/libschpf_b200.so
for K in 7 16; do run base $D $K; run w16_12 $PWD/schpf_b200/_C_w16_12/libschpf_b200.so $K; done
python - "$T" <<'P'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/%s_*_K*.json" % sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        lay = d["config"]["layout"]
        print("%-30s ms/step %.3f  pair %.3f  frac %.4f  layout GB %.2f warps %s" % (
            f[11:-5], d["ms_per_step"], d["roofline"]["sweep_pair_ms"], d["roofline"]["frac"], lay["layout_bytes"] / 1e9, lay["warps_per_cta"]))
    except Exception as e:
        print(f, "unreadable", e, open(f[:-5] + ".err").read()[-300:])
P
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lane_sweep_kernel -s 9 -c 2 -f \
    -o gpurun_out/${T}_lanes_K16 python bench.py --no-cpu --no-e2e --no-strong --no-parity --factors 16 --steps 3 --warmup 3 \
    > gpurun_out/${T}_ncu_K16.log 2>&1
echo "ncu K=16 rc=$?"
