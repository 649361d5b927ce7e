#!/bin/bash
# chunked panel fills: tests, then 1 / 4 / 8 / 16 chunks at K = 16 / 20 / 30 (cfg-3) and K = 20 on the mid-size matrix
T=${1:-r2z}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_engine.py tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
run() {   # tag lib K extra...
  local tag=$1 lib=$2 K=$3; shift 3
  SCHPF_B200_LIB=$lib timeout 120 python bench.py --no-cpu --no-e2e --no-strong --no-parity --factors $K --steps 20 --warmup 3 "$@" \
      > gpurun_out/${T}_${tag}_K$K.json 2> gpurun_out/${T}_${tag}_K$K.err
}
for tag in ch1 ch4 ch8 ch16; do
  L=$PWD/schpf_b200/_C_$tag/libschpf_b200.so; [ $tag = ch8 ] && L=$PWD/schpf_b200/_C/libschpf_b200.so
  for K in 16 20 30; do run $tag $L $K; done
  run ${tag}mid $L 20 --cells 20000 --draws 1000 --steps 100
done
python - "$T" <<'P'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/%s_*_K*.json" % sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-24s ms/step %.4f  pair %.4f  frac %.4f loss %s" % (f[11:-5], d["ms_per_step"], d["roofline"]["sweep_pair_ms"], d["roofline"]["frac"], d["loss_first_last"][-1]))
    except Exception as e:
        print(f, "unreadable", e, open(f[:-5] + ".err").read()[-300:])
P
