#!/bin/bash
# concurrent shape sweeps (cells-own on a second stream): whole suite, cfg-3 / mid-size / K sweep with and without
T=${1:-r3a}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
run() {   # tag K extra...
  local tag=$1 K=$2; shift 2
  timeout 120 python bench.py --no-cpu --no-e2e --no-strong --factors $K --steps 30 --warmup 3 "$@" \
      > gpurun_out/${T}_${tag}_K$K.json 2> gpurun_out/${T}_${tag}_K$K.err
}
for K in 16 20 30; do run ov $K; SCHPF_NO_OVERLAP_SWEEPS=1 run noov $K; done
run ovmid 20 --cells 20000 --draws 1000 --steps 100; SCHPF_NO_OVERLAP_SWEEPS=1 run noovmid 20 --cells 20000 --draws 1000 --steps 100
run ovsmall 20 --cells 12500 --steps 100; SCHPF_NO_OVERLAP_SWEEPS=1 run noovsmall 20 --cells 12500 --steps 100
python - "$T" <<'P'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/%s_*_K*.json" % sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-24s ms/step %.4f  pair %.4f  frac %.4f parity %s" % (f[11:-5], d["ms_per_step"], d["roofline"]["sweep_pair_ms"], d["roofline"]["frac"], d["parity"]["max_rel_vs_golden"]))
    except Exception as e:
        print(f, "unreadable", e, open(f[:-5] + ".err").read()[-300:])
P
