#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1i_tests_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -3 gpurun_out/r1i_tests_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1i_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/r1i_smoke.log
timeout 400 python bench.py --steps 30 --warmup 3 > gpurun_out/r1i_bench_n1.json 2> gpurun_out/r1i_bench_n1.err
echo "bench rc=$?"
timeout 200 python bench.py --no-cpu --no-e2e --skewed --steps 20 --warmup 3 > gpurun_out/r1i_bench_skewed.json 2> gpurun_out/r1i_bench_skewed.err
echo "skewed rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 9 -c 2 -f -o gpurun_out/r1i_sweep \
    python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > gpurun_out/r1i_ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k 'regex:sweep_kernel|finalize_kernel|fold_kernel|ex_table_kernel|sum_partials_kernel|pack_loss' -c 120 \
    --csv --log-file gpurun_out/r1i_launches.csv python bench.py --no-cpu --no-e2e --steps 8 --warmup 1 > gpurun_out/r1i_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r1i_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f pair %.3f frac %.4f value %.3e" % (d["ms_per_step"], d["roofline"]["sweep_pair_ms"], d["roofline"]["frac"], d["value"]),
              "e2e %.3e" % d["e2e"]["value"] if "e2e" in d else "", d["config"]["nnz_total"], d["config"]["layout"]["padded_nnz_cells"] / d["config"]["nnz_per_gpu"], d["config"]["layout"]["padded_nnz_genes"] / d["config"]["nnz_per_gpu"])
    except Exception as e:
        print(f, "unreadable", e)
P
