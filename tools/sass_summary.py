#!/usr/bin/env python
"""SASS evidence for the sweep kernels of the built library (no GPU needed).

    python tools/sass_summary.py [lib.so] > profiles/r2_sass_summary.json

For every `sweep_kernel<KP,MODE,PACKED>` (lane pairs), `lane_sweep_kernel<NA,REM,MODE,ENC>` (one lane per
owner) and `lane_sweep_f32_kernel<NP,MODE,PACKED>` (fp32) instantiation: registers, spill bytes and shared memory from
`cuobjdump --dump-resource-usage`, instruction counts of the whole kernel, and the same counts
restricted to the HOT LOOP (the innermost backward branch whose body holds LDS.128 row loads) --
which is what answers "are the STL/LDL inside the loop".  Mnemonics of interest: UBLKCP (1-D bulk
async copy on the TMA engine), UBLKRED (TMA bulk reduction, add.f64), SYNCS (mbarrier), DMMA,
LDS.128, DFMA, MUFU.RCP64H, STL / LDL (local memory = spills), REDG / ATOMG.
"""
import json
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ("UBLKCP", "UBLKRED", "SYNCS", "DMMA", "LDS", "DFMA", "DMUL", "DADD", "DSETP", "FFMA", "FMUL", "FADD", "MUFU", "I2F", "I2FP", "STL", "LDL", "REDG",
         "ATOMG", "LDG", "BAR", "LOP3", "IMAD", "ISETP")


def demangle_name(sym):
    m = re.search(r"lane_sweep_f32_kernelILi(\d+)ELi(\d+)ELb(\d)", sym)
    if m:
        planes, mode, packed = (int(x) for x in m.groups())
        return "lane_sweep_f32_kernel<floats=%d,%s,%s>" % (32 * planes, "LLH" if mode else "SHAPE", "packed" if packed else "wide")
    m = re.search(r"lane_sweep_kernelILi(\d+)ELi(\d+)ELi(\d+)ELi(\d)", sym)
    if m:
        na, rem, mode, enc = (int(x) for x in m.groups())
        return "lane_sweep_kernel<KP=%d,%s,%s>" % (16 * na + 4 * rem, "LLH" if mode else "SHAPE", ("wide-int", "wide-yhi", "packed")[enc])
    m = re.search(r"sweep_kernelILi(\d+)ELi(\d+)ELb(\d)", sym)
    if m:
        kp, mode, packed = (int(x) for x in m.groups())
        return "sweep_kernel<KP=%d,%s,%s>" % (kp, "LLH" if mode else "SHAPE", "packed" if packed else "wide")
    return None


def main(lib):
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], stdout=subprocess.PIPE, text=True).stdout
    usage, cur = {}, None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = {k.lower(): int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line)}
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
    out, name, body = {}, None, []

    def flush():
        if name is None or demangle_name(name) is None:
            return
        ins = []
        for l in body:
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        op = lambda t: [w for w in t.split() if not w.startswith("@")][0].split(".")[0]
        total = Counter(op(t) for _, t in ins)
        lds128 = sum("LDS.128" in t for _, t in ins)
        # innermost loop holding LDS.128
        best = None
        for a, t in ins:
            m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s+)?0x([0-9a-f]+)", t)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt >= a:
                continue
            loop = [x for b, x in ins if tgt <= b <= a]
            if any("LDS.128" in x for x in loop) and (best is None or len(loop) < len(best)):
                best = loop
        loop_c = Counter(op(t) for t in (best or []))
        out[demangle_name(name)] = {
            "registers": usage.get(name, {}).get("reg"), "stack_bytes": usage.get(name, {}).get("stack"),
            "whole_kernel": {k: total[k] for k in WATCH if total[k]}, "lds128_whole_kernel": lds128,
            "hot_loop_instructions": len(best or []),
            "hot_loop": {k: loop_c[k] for k in WATCH if loop_c[k]},
            "spill_instructions_in_hot_loop": loop_c["STL"] + loop_c["LDL"],
        }

    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            name, body = m.group(1), []
        else:
            body.append(line)
    flush()
    json.dump({"library": os.path.relpath(lib, ROOT), "kernels": dict(sorted(out.items()))}, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "schpf_b200", "_C", "libschpf_b200.so"))
