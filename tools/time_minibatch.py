#!/usr/bin/env python
"""Wall-clock of scHPF.fit(batchsize=...) on one GPU: per-window engines (batchsize divides the
cell count into few windows) against one engine re-laid out every iteration (coprime sizes),
and the full-batch fit on the same matrix for scale.  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from schpf_b200 import scHPF                      # noqa: E402
from schpf_b200.synth import synth_coo            # noqa: E402


def timed_fit(X, K, n_iter, **kw):
    np.random.seed(0)
    m = scHPF(K, verbose=False)
    t0 = time.perf_counter()
    m.fit(X, min_iter=n_iter, max_iter=n_iter, check_freq=10, **kw)
    return time.perf_counter() - t0, m.loss


def main():
    C, G, per_cell, K = 20000, 20000, 1000, 20
    X = synth_coo(C, G, per_cell, K, seed=0)
    out = {"cells": C, "genes": G, "nnz": int(X.nnz), "K": K}
    timed_fit(X, K, 2)                                        # warm the pool / context
    for name, kw, n_iter in (("full_batch", {}, 40), ("windows_of_2000_cached", {"batchsize": 2000}, 40),
                             ("windows_of_1999_relaid", {"batchsize": 1999}, 40)):
        a, _ = timed_fit(X, K, n_iter // 2, **kw)
        b, loss = timed_fit(X, K, n_iter + n_iter // 2, **kw)
        out[name] = {"ms_per_iteration": 1e3 * (b - a) / n_iter, "fit_seconds_%d_iters" % (n_iter + n_iter // 2): b,
                     "loss_last": float(loss[-1])}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
