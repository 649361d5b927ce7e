#!/bin/bash
# Round 2, fifth GPU call (first of the re-created container): whole GPU suite, lane vs pair sweep
# across K, ncu captures of the lane kernels the bench times, launch list, full default bench line.
T=${1:-r2e}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/${T}_tests.log
run() {   # tag K extra...
  local tag=$1 K=$2; shift 2
  timeout 120 python bench.py --no-cpu --no-e2e --no-strong --no-parity --factors $K --steps 20 --warmup 3 "$@" \
      > gpurun_out/${T}_${tag}_K$K.json 2> gpurun_out/${T}_${tag}_K$K.err
}
for K in 7 10 15 16 20 24 30 32 50; do run lanes $K; done
for K in 15 20 30; do run pairs $K --lanes 0; done
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
for K in 20 30; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:lane_sweep_kernel -s 9 -c 2 -f \
      -o gpurun_out/${T}_lanes_K$K python bench.py --no-cpu --no-e2e --no-strong --no-parity --factors $K --steps 3 --warmup 3 \
      > gpurun_out/${T}_ncu_K$K.log 2>&1
  echo "ncu K=$K rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --no-cpu --no-e2e --no-strong --no-parity --steps 5 --warmup 3 > gpurun_out/${T}_launches.log 2>&1
echo "launch list rc=$?"
SCHPF_BENCH_VERBOSE=1 timeout 900 python bench.py > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err
echo "full bench rc=$?"; cut -c1-1500 gpurun_out/${T}_bench_full.json; tail -5 gpurun_out/${T}_bench_full.err
