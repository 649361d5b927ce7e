#!/bin/bash
# Round 2: packed 4-byte entries for the one-lane streams: tests, then packed vs wide at K = 16 / 20 / 30 and fp32.
T=${1:-r2m}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/${T}_tests.log
run() {   # tag K extra...
  local tag=$1 K=$2; shift 2
  timeout 120 python bench.py --no-cpu --no-e2e --no-strong --factors $K --steps 20 --warmup 3 "$@" \
      > gpurun_out/${T}_${tag}_K$K.json 2> gpurun_out/${T}_${tag}_K$K.err
}
for K in 10 16 20 30; do run packed $K; run wide $K --packed 0; done
run f32packed 20 --precision 32; run f32wide 20 --precision 32 --packed 0
run f32packed 30 --precision 32; run f32wide 30 --precision 32 --packed 0
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
