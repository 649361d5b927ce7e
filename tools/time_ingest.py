#!/usr/bin/env python
"""Time the device-side ingest against the reference's loaders on one synthetic MatrixMarket file
(written with pandas' C writer; the layout is what `scHPF prep` writes: integer field, 1-based).

    python tools/time_ingest.py [cells] [genes] [draws/cell] > profiles/<round>_ingest.json
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import pandas as pd
    import torch
    from scipy.io import mmread
    from schpf_b200 import io as sio
    from schpf_b200.synth import synth_coo
    C, G, D = (int(v) for v in (sys.argv[1:4] + [20000, 20000, 1000][len(sys.argv) - 1:]))
    X = synth_coo(C, G, D, 20, seed=0)
    d = tempfile.mkdtemp()
    path = os.path.join(d, "x.mtx")
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate integer general\n%%\n%d %d %d\n" % (C, G, X.nnz))
    pd.DataFrame({"r": X.row + 1, "c": X.col + 1, "v": X.data}).to_csv(path, sep=" ", header=False, index=False, mode="a")
    size = os.path.getsize(path)
    sio.load_mtx(path)                                   # warm: module load, pools
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev = sio.load_mtx(path)
    torch.cuda.synchronize()
    t_dev = time.perf_counter() - t0
    raw = np.fromfile(path, dtype=np.uint8)
    t0 = time.perf_counter()
    text = torch.from_numpy(raw).cuda()
    torch.cuda.synchronize()
    t_h2d = time.perf_counter() - t0
    head, begin = raw[:200].tobytes(), 0
    for _ in range(3):                                   # banner, comment, size line
        begin = head.index(b"\n", begin) + 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sio._parse(text, begin, 3, 1, X.nnz, 0)
    e1.record()
    torch.cuda.synchronize()
    t_parse = e0.elapsed_time(e1) * 1e-3
    t0 = time.perf_counter()
    ref = mmread(path)
    t_ref = time.perf_counter() - t0
    same = bool(np.array_equal(dev.row.cpu().numpy(), ref.row) and np.array_equal(dev.col.cpu().numpy(), ref.col)
                and np.array_equal(dev.data.cpu().numpy(), ref.data))
    os.remove(path)
    # the reference's own tsv loader (preprocessing.py:27-28) on a 2e6-line prefix (np.loadtxt is slow)
    n_tsv = min(int(X.nnz), 2_000_000)
    tsv = os.path.join(d, "x.tsv")
    pd.DataFrame({"r": X.row[:n_tsv], "c": X.col[:n_tsv], "v": X.data[:n_tsv]}).to_csv(tsv, sep="\t", header=False, index=False)
    sio.load_coo(tsv)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev_t = sio.load_coo(tsv)
    torch.cuda.synchronize()
    t_tsv = time.perf_counter() - t0
    t0 = time.perf_counter()
    raw_t = np.loadtxt(tsv, delimiter="\t", dtype=int)
    t_loadtxt = time.perf_counter() - t0
    same_t = bool(np.array_equal(dev_t.row.cpu().numpy(), raw_t[:, 0]) and np.array_equal(dev_t.col.cpu().numpy(), raw_t[:, 1])
                  and np.array_equal(dev_t.data.cpu().numpy(), raw_t[:, 2]))
    tsv_bytes = os.path.getsize(tsv)
    os.remove(tsv)
    print(json.dumps({"file_bytes": size, "nnz": int(X.nnz), "shape": [C, G],
                      "load_mtx_s": t_dev, "of_which_read_and_h2d_s": t_h2d, "of_which_device_parse_s": t_parse,
                      "device_parse_GBps": size / t_parse / 1e9, "scipy_mmread_s": t_ref, "speedup_vs_mmread": t_ref / t_dev,
                      "identical_to_mmread": same,
                      "tsv": {"lines": n_tsv, "file_bytes": tsv_bytes, "load_coo_s": t_tsv, "np_loadtxt_s": t_loadtxt,
                              "speedup_vs_loadtxt": t_loadtxt / t_tsv, "identical_to_loadtxt": same_t,
                              "what": "schpf_b200.io.load_coo vs the reference's load_coo = np.loadtxt(delimiter='\\t', dtype=int) (preprocessing.py:27-28)"},
                      "what": "schpf_b200.io.load_mtx (file -> pageable host -> HBM -> parse on device, file order) vs "
                              "scipy.io.mmread of the same file on the host (bin/scHPF:373-374)"}))


if __name__ == "__main__":
    main()
