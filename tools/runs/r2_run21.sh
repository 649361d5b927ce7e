#!/bin/bash
# mid-size problem (20k x 20k, 2e7 nnz): where does a step go when the sweeps are short?
T=${1:-r2w}
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 200 python bench.py --no-cpu --no-strong --no-parity --cells 20000 --draws 1000 --steps 100 --warmup 5 > gpurun_out/${T}_mid.json 2> gpurun_out/${T}_mid.err
python -c "
import json
d=json.loads(open('gpurun_out/${T}_mid.json').read().strip().splitlines()[-1])
print('mid: ms/step %.4f pair %.4f share %.3f e2e %.4g launches %d nnz %d'%(d['ms_per_step'],d['roofline']['sweep_pair_ms'],d['roofline']['sweep_share_of_step'],d['e2e']['value'],d['gpu_launches'],d['config']['nnz_total']))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sweep|finalize|fold|fixup|partials|pack_loss|ex_table|prep' -c 150 --csv \
    --log-file gpurun_out/${T}_mid_launches.csv python bench.py --no-cpu --no-e2e --no-strong --no-parity --cells 20000 --draws 1000 --steps 10 --warmup 3 > gpurun_out/${T}_mid_launches.log 2>&1
python - "$T" <<'P'
import csv, collections, sys
rows=[r for r in csv.reader(open('gpurun_out/%s_mid_launches.csv' % sys.argv[1])) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows[40:]:
    a=agg.setdefault(r[4].split('(')[0][-44:],[0,0.0]); a[0]+=1; a[1]+=float(r[-1])/1000.0
for k,(n,t) in agg.items(): print("%-46s n=%3d avg %8.1f us"%(k,n,t/n))
P
