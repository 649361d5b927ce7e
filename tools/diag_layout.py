#!/usr/bin/env python
"""Where does set_coo spend its time?  cfg-3 matrix, three set_coo calls on one engine from
device arrays and from pinned host arrays, with SCHPF_TRACE marks (stderr) and wall clocks."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                           # noqa: E402
from schpf_b200.engine import CaviEngine               # noqa: E402
from schpf_b200.synth import synth_coo_torch           # noqa: E402

C, G, K = 100000, 20000, 20
row, col, val = synth_coo_torch(C, G, 2000, K, seed=0, device="cuda:0")
torch.cuda.synchronize()
pin = lambda t: torch.empty(t.numel(), dtype=torch.int32, pin_memory=True).copy_(t)
hrow, hcol, hval = pin(row), pin(col), pin(val)
torch.cuda.synchronize()
for stream_name, stream in (("torch stream", torch.cuda.current_stream().cuda_stream), ("default stream", None)):
    e = CaviEngine(C, G, K, device=0, stream=stream)
    for src, args in (("device", (row, col, val)), ("device", (row, col, val)), ("pinned host", (hrow, hcol, hval)),
                      ("pinned host", (hrow, hcol, hval))):
        sys.stderr.write("---- %s, %s\n" % (stream_name, src))
        t0 = time.perf_counter()
        e.set_coo(*args)
        torch.cuda.synchronize()
        print("%-14s %-12s set_coo %.1f ms" % (stream_name, src, 1e3 * (time.perf_counter() - t0)), flush=True)
    e.close()
