#!/usr/bin/env python
"""Turn an `ncu --set full --import-source on` report of the sweep kernels into the JSON summary
kept under profiles/ (the raw .ncu-rep stays in gpurun_out/).

    python tools/ncu_summary.py gpurun_out/check_sweep.ncu-rep profiles/rN_sweep_ncu_full.json "note"
"""
import csv
import io
import json
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__cycles_elapsed.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed']
UNIT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, dst, note):
    rows = page(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('per_issue_active.ratio')]
    out = {"note": note, "kernels": []}
    for r in data:
        k = {"kernel": r[col['Kernel Name']], "grid": r[col['Grid Size']], "block": r[col['Block Size']]}
        for w in KEEP:
            if w in col:
                k[w] = r[col[w]] + ' ' + units[col[w]]
        k["stalls_per_issue"] = {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):
                                 round(float(r[col[h]]), 3) for h in stall if float(r[col[h]]) >= 0.01}
        out["kernels"].append(k)
    src = page(rep, "source", ("--kernel-id", ":::1"))
    hi = [i for i, r in enumerate(src) if r and r[0] == 'Address'][0]
    h2 = src[hi]
    d2 = [r for r in src[hi + 1:] if len(r) == len(h2) and r[0] != 'Address']
    ix = {h: i for i, h in enumerate(h2)}
    val = lambda r, h: int(r[ix[h]] or 0)
    tot = sum(val(r, '# Samples') for r in d2)
    st = [h for h in h2 if h.startswith('stall_') and 'Not Issued' not in h]
    out["source_page_first_launch"] = {
        "total_samples": tot,
        "by_reason_pct": {h[6:]: round(100 * sum(val(r, h) for r in d2) / tot, 1) for h in st
                          if sum(val(r, h) for r in d2) / tot > 0.005},
        "top_instructions": [{"sass": r[ix['Source']].strip(), "pct": round(100 * val(r, '# Samples') / tot, 2),
                              "main": max((val(r, h), h[6:]) for h in st)[1]}
                             for r in sorted(d2, key=lambda r: -val(r, '# Samples'))[:16]]}
    json.dump(out, open(dst, "w"), indent=1)
    r = data[0]
    per = sum(float(r[col[m]]) * UNIT[units[col[m]]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    print("first launch: %s, DRAM bytes %.4g (x2 per iteration -> profiles/sweep_traffic.json)" % (
        r[col['gpu__time_duration.sum']] + units[col['gpu__time_duration.sum']], per))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
