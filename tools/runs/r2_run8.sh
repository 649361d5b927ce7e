#!/bin/bash
# Round 2: fp32 sweep -- tests, bench lines (separate from the headline), ncu capture; launch list of the fp64 step.
T=${1:-r2h}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_f32.py -m gpu -q > gpurun_out/${T}_tests_f32.log 2>&1
echo "f32 tests rc=$?"; tail -25 gpurun_out/${T}_tests_f32.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/${T}_tests.log
run() {   # tag K extra...
  local tag=$1 K=$2; shift 2
  timeout 120 python bench.py --no-cpu --no-e2e --no-strong --factors $K --steps 20 --warmup 3 "$@" \
      > gpurun_out/${T}_${tag}_K$K.json 2> gpurun_out/${T}_${tag}_K$K.err
}
for K in 10 16 20 30 32 50; do run f32 $K --precision 32; done
python - "$T" <<'P'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/%s_*_K*.json" % sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        lay = d["config"]["layout"]
        print("%-36s ms/step %.3f  pair %.3f  frac %.4f  pad %.3f/%.3f  loss %s parity %s" % (
            f[11:-5], d["ms_per_step"], d["roofline"]["sweep_pair_ms"], d["roofline"]["frac"],
            lay["padded_nnz_cells"] / d["config"]["nnz_per_gpu"] - 1, lay["padded_nnz_genes"] / d["config"]["nnz_per_gpu"] - 1,
            d["loss_first_last"], (d.get("parity") or {}).get("max_rel_vs_golden")))
    except Exception as e:
        print(f, "unreadable", e, open(f[:-5] + ".err").read()[-300:])
P
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lane_sweep_f32_kernel -s 9 -c 2 -f \
    -o gpurun_out/${T}_f32_K20 python bench.py --no-cpu --no-e2e --no-strong --no-parity --precision 32 --steps 3 --warmup 3 \
    > gpurun_out/${T}_ncu_f32_K20.log 2>&1
echo "ncu f32 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sweep|finalize|fold|fixup|partials|pack_loss|ex_table|prep' -c 200 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --no-cpu --no-e2e --no-strong --no-parity --steps 10 --warmup 3 > gpurun_out/${T}_launches.log 2>&1
echo "launch list rc=$?"
SCHPF_BENCH_VERBOSE=1 timeout 600 python bench.py --precision 32 --no-strong > gpurun_out/${T}_bench_f32_full.json 2> gpurun_out/${T}_bench_f32_full.err
echo "f32 full bench rc=$?"; cut -c1-600 gpurun_out/${T}_bench_f32_full.json; tail -3 gpurun_out/${T}_bench_f32_full.err
