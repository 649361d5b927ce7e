"""Synthetic sparse UMI-count matrices for tests and benchmarks.

Planted HPF: theta* ~ Gamma(0.3, 1) (cells x K*), beta* ~ Gamma(0.3, 1)
(genes x K*); every cell draws ``nnz_per_cell`` genes (uniformly, or with
Gamma(0.5, 1) gene weights when ``skewed``); the value at (c, g) is
max(1, Poisson(theta*_c . beta*_g)); duplicate draws are merged, so the
realised nnz is a few percent below cells * nnz_per_cell and rows are sorted.

``synth_coo`` is the host/numpy generator (small and medium sizes, exactly
reproducible everywhere); ``synth_coo_torch`` generates on a torch device
(used by bench.py for the 10^8-nnz configurations, where host generation
would take minutes).
"""
import numpy as np


def synth_coo(ncells, ngenes, nnz_per_cell, nfactors, seed=0, skewed=False):
    """Returns a canonical (row-sorted, duplicate-free) scipy coo_matrix of int32."""
    from scipy.sparse import coo_matrix
    rng = np.random.default_rng(seed)
    theta = rng.gamma(0.3, 1.0, size=(ncells, nfactors))
    beta = rng.gamma(0.3, 1.0, size=(ngenes, nfactors))
    rows = np.repeat(np.arange(ncells, dtype=np.int32), nnz_per_cell)
    if skewed:
        w = rng.gamma(0.5, 1.0, size=ngenes)
        cols = rng.choice(ngenes, size=rows.shape[0], p=w / w.sum()).astype(np.int32)
    else:
        cols = rng.integers(0, ngenes, size=rows.shape[0], dtype=np.int32)
    # merge duplicate draws BEFORE sampling so that values stay max(1, Poisson)
    key = np.unique(rows.astype(np.int64) * ngenes + cols)
    rows = (key // ngenes).astype(np.int32)
    cols = (key % ngenes).astype(np.int32)
    rate = np.einsum('ik,ik->i', theta[rows], beta[cols])
    vals = np.maximum(1, rng.poisson(rate)).astype(np.int32)
    X = coo_matrix((vals, (rows, cols)), shape=(ncells, ngenes), dtype=np.int32)
    return X


def synth_coo_torch(ncells, ngenes, nnz_per_cell, nfactors, seed=0, device="cuda",
                    row_offset=0, chunk_cells=8192, skewed=False):
    """Device-side generator; returns (row, col, val) int32 torch tensors on
    ``device``, canonical order (sorted by row then col, no duplicates).
    ``skewed``: genes drawn with Gamma(0.5, 1) weights (a few genes in most cells, a long
    tail in few), like ``synth_coo(skewed=True)``.

    ``row_offset`` shifts the RNG stream so that shards generated on different
    ranks are independent (rows are returned shard-local, starting at 0).
    """
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed) * 1000003 + int(row_offset))
    # torch._standard_gamma takes no generator: seed the device's global stream
    # for the two Gamma draws and restore it afterwards.
    state = torch.random.get_rng_state() if device == "cpu" else torch.cuda.get_rng_state(device)
    try:
        if device == "cpu":
            torch.manual_seed(int(seed) * 1000003 + 7919)
        else:
            torch.cuda.manual_seed(int(seed) * 1000003 + 7919)
        beta = torch._standard_gamma(torch.full((ngenes, nfactors), 0.3, device=device,
                                                dtype=torch.float32))
        if device == "cpu":
            torch.manual_seed(int(seed) * 1000003 + 104729 + int(row_offset))
        else:
            torch.cuda.manual_seed(int(seed) * 1000003 + 104729 + int(row_offset))
        gene_cdf = None
        if skewed:
            w = torch._standard_gamma(torch.full((ngenes,), 0.5, device=device, dtype=torch.float64))
            gene_cdf = torch.cumsum(w / w.sum(), 0)
        rows_out, cols_out, vals_out = [], [], []
        for c0 in range(0, ncells, chunk_cells):
            nc = min(chunk_cells, ncells - c0)
            theta = torch._standard_gamma(torch.full((nc, nfactors), 0.3, device=device,
                                                     dtype=torch.float32))
            if gene_cdf is None:
                cols = torch.randint(0, ngenes, (nc, nnz_per_cell), device=device, generator=g)
            else:       # inverse-CDF sampling with replacement; duplicates are merged below
                u = torch.rand((nc, nnz_per_cell), device=device, dtype=torch.float64, generator=g)
                cols = torch.searchsorted(gene_cdf, u).clamp_(max=ngenes - 1)
            cols, _ = torch.sort(cols, dim=1)
            keep = torch.ones_like(cols, dtype=torch.bool)
            keep[:, 1:] = cols[:, 1:] != cols[:, :-1]
            r_local = torch.arange(nc, device=device).unsqueeze(1).expand_as(cols)
            r_sel = r_local[keep]
            c_sel = cols[keep]
            rate = (theta[r_sel] * beta[c_sel]).sum(dim=1)
            vals = torch.poisson(rate, generator=g).clamp_(min=1.0)
            rows_out.append((r_sel + c0).to(torch.int32))
            cols_out.append(c_sel.to(torch.int32))
            vals_out.append(vals.to(torch.int32))
        return torch.cat(rows_out), torch.cat(cols_out), torch.cat(vals_out)
    finally:
        if device == "cpu":
            torch.random.set_rng_state(state)
        else:
            torch.cuda.set_rng_state(state, device)
