#!/usr/bin/env python
"""profiles/sweep_traffic.json from the ncu summaries of the sweep kernels (tools/ncu_summary.py
output): one entry per (K, kernel family, library version), read by bench.py for `roofline.traffic`
and `roofline.ncu` -- only an entry of the SAME library version is ever quoted.

    python tools/make_sweep_traffic.py <library_version> K:lanes:profiles/<summary>.json [...]
"""
import json
import os
import sys

UNIT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}


def num(s):
    v, _, u = s.strip().partition(' ')
    return float(v) * UNIT.get(u.strip(), 1.0)


def main(argv):
    version = int(argv[0])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "sweep_traffic.json")
    try:
        doc = json.load(open(path))
    except Exception:
        doc = {}
    caps = [c for c in doc.get("captures", [])]
    for spec in argv[1:]:
        K, lanes, src = spec.split(":")
        ks = json.load(open(src))["kernels"][:2]            # cells-own + genes-own SHAPE sweeps of one iteration
        dram = sum(num(k['dram__bytes_read.sum']) + num(k['dram__bytes_write.sum']) for k in ks)
        avg = lambda m: sum(num(k[m]) for k in ks) / len(ks)
        entry = {"K": int(K), "lanes": bool(int(lanes)), "library_version": version, "capture": src,
                 "dram_bytes_per_iteration": dram,
                 "dram_pct": avg('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                 "lsu_pct": avg('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
                 "fp64_pct": avg('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
                 "issue_pct": avg('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                 "ms_per_sweep_under_ncu": avg('gpu__time_duration.sum'),
                 "shared_wavefronts_per_sweep": avg('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'),
                 "registers": int(num(ks[0]['launch__registers_per_thread']))}
        caps = [c for c in caps if not (c["K"] == entry["K"] and c["lanes"] == entry["lanes"]
                                        and c["library_version"] == version)]
        caps.append(entry)
    json.dump({"note": "DRAM bytes and pipe utilisations of the two SHAPE sweeps of one iteration, from `ncu --set full` "
                       "captures (tools/ncu_summary.py); bench.py quotes an entry only for the same library version",
               "captures": caps}, open(path, "w"), indent=1)
    print(json.dumps(caps, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
