#!/bin/bash
T=${1:-r2n}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_engine.py -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
run() {   # tag K extra...
  local tag=$1 K=$2; shift 2
  timeout 120 python bench.py --no-cpu --no-e2e --no-strong --factors $K --steps 20 --warmup 3 "$@" \
      > gpurun_out/${T}_${tag}_K$K.json 2> gpurun_out/${T}_${tag}_K$K.err
}
for K in 16 20 30; do run packed $K; run wide $K --packed 0; done
python - "$T" <<'P'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/%s_*_K*.json" % sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        lay = d["config"]["layout"]
        print("%-30s ms/step %.3f  pair %.3f  frac %.4f  layout GB %.2f  parity %s" % (
            f[11:-5], d["ms_per_step"], d["roofline"]["sweep_pair_ms"], d["roofline"]["frac"], lay["layout_bytes"] / 1e9,
            (d.get("parity") or {}).get("max_rel_vs_golden")))
    except Exception as e:
        print(f, "unreadable", e, open(f[:-5] + ".err").read()[-300:])
P
