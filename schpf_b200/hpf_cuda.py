"""Drop-in replacements for the functions of the reference's schpf/hpf_numba.py,
same names and argument meaning, computed by sm_100a CUDA through the C ABI
(include/schpf_b200.h, "function level").  Inputs and outputs are host ndarrays;
every call uploads, runs, downloads -- use `schpf_b200.engine.CaviEngine` (what
scHPF.fit uses) when state should stay on the GPU.

All arithmetic is fp64; float32 inputs are promoted and the result is cast back
to the input dtype.
"""
import numpy as np

from . import _lib
from ._lib import as_f64, as_i32, dptr, iptr, c_int, c_i64, c_dbl

_DEVICE = 0


def set_device(device):
    global _DEVICE
    _DEVICE = int(device)


def _vec(fn_name, x):
    scalar = np.ndim(x) == 0
    xa = as_f64(np.atleast_1d(x))
    out = np.empty_like(xa)
    lib = _lib.load()
    _lib.check(getattr(lib, fn_name)(c_int(_DEVICE), c_i64(xa.size), dptr(xa), dptr(out)))
    out = out.reshape(np.shape(x))
    return float(out) if scalar else out


def psi(x):
    """hpf_numba.py:16-18 (digamma); accepts scalars or arrays."""
    return _vec("schpf_psi", x)


def cgammaln(x):
    """hpf_numba.py:20-22 (log-gamma)."""
    return _vec("schpf_gammaln", x)


def compute_Xphi_data(X_data, X_row, X_col, theta_vi_shape, theta_vi_rate,
                      beta_vi_shape, beta_vi_rate):
    """hpf_numba.py:55-114 -> (nnz, nfactors) array of X * phi."""
    dtype = np.asarray(theta_vi_shape).dtype
    X_data, X_row, X_col = as_i32(X_data, "count"), as_i32(X_row), as_i32(X_col)
    ts, tr = as_f64(theta_vi_shape), as_f64(theta_vi_rate)
    bs, br = as_f64(beta_vi_shape), as_f64(beta_vi_rate)
    nnz, (ncells, K), ngenes = X_data.shape[0], ts.shape, bs.shape[0]
    assert tr.shape == ts.shape and br.shape == bs.shape and bs.shape[1] == K
    assert X_row.shape[0] == nnz and X_col.shape[0] == nnz
    out = np.empty((nnz, K), dtype=np.float64)
    lib = _lib.load()
    _lib.check(lib.schpf_compute_Xphi_data(
        c_int(_DEVICE), c_i64(nnz), c_i64(ncells), c_i64(ngenes), c_int(K),
        iptr(X_data), iptr(X_row), iptr(X_col), dptr(ts), dptr(tr), dptr(bs), dptr(br), dptr(out)))
    return out.astype(dtype, copy=False)


def compute_loading_shape_update(Xphi_data, X_keep, nkeep, shape_prior):
    """hpf_numba.py:129-156 -> (nkeep, nfactors)."""
    dtype = np.asarray(Xphi_data).dtype
    xphi, keep = as_f64(Xphi_data), as_i32(X_keep)
    nnz, K = xphi.shape
    assert keep.shape[0] == nnz
    out = np.empty((int(nkeep), K), dtype=np.float64)
    lib = _lib.load()
    _lib.check(lib.schpf_compute_loading_shape_update(
        c_int(_DEVICE), c_i64(nnz), c_int(K), dptr(xphi), iptr(keep), c_i64(int(nkeep)),
        c_dbl(float(shape_prior)), dptr(out)))
    return out.astype(dtype, copy=False)


def compute_loading_rate_update(prior_vi_shape, prior_vi_rate,
                                other_loading_vi_shape, other_loading_vi_rate):
    """hpf_numba.py:160-177 -> (len(prior), nfactors)."""
    dtype = np.asarray(prior_vi_shape).dtype
    ps, pr = as_f64(prior_vi_shape), as_f64(prior_vi_rate)
    os_, or_ = as_f64(other_loading_vi_shape), as_f64(other_loading_vi_rate)
    n, (m, K) = ps.shape[0], os_.shape
    out = np.empty((n, K), dtype=np.float64)
    lib = _lib.load()
    _lib.check(lib.schpf_compute_loading_rate_update(
        c_int(_DEVICE), c_i64(n), c_i64(m), c_int(K), dptr(ps), dptr(pr), dptr(os_), dptr(or_), dptr(out)))
    return out.astype(dtype, copy=False)


def compute_capacity_rate_update(loading_vi_shape, loading_vi_rate, prior_rate):
    """hpf_numba.py:181-188 -> (n,)."""
    dtype = np.asarray(loading_vi_shape).dtype
    ls, lr = as_f64(loading_vi_shape), as_f64(loading_vi_rate)
    n, K = ls.shape
    out = np.empty((n,), dtype=np.float64)
    lib = _lib.load()
    _lib.check(lib.schpf_compute_capacity_rate_update(
        c_int(_DEVICE), c_i64(n), c_int(K), dptr(ls), dptr(lr), c_dbl(float(prior_rate)), dptr(out)))
    return out.astype(dtype, copy=False)


def compute_pois_llh(X_data, X_row, X_col, theta_vi_shape, theta_vi_rate,
                     beta_vi_shape, beta_vi_rate):
    """hpf_numba.py:25-51 -> (nnz,) pointwise Poisson log-likelihood."""
    dtype = np.asarray(theta_vi_shape).dtype
    X_data, X_row, X_col = as_i32(X_data, "count"), as_i32(X_row), as_i32(X_col)
    ts, tr = as_f64(theta_vi_shape), as_f64(theta_vi_rate)
    bs, br = as_f64(beta_vi_shape), as_f64(beta_vi_rate)
    nnz, (ncells, K), ngenes = X_data.shape[0], ts.shape, bs.shape[0]
    out = np.empty((nnz,), dtype=np.float64)
    lib = _lib.load()
    _lib.check(lib.schpf_compute_pois_llh(
        c_int(_DEVICE), c_i64(nnz), c_i64(ncells), c_i64(ngenes), c_int(K),
        iptr(X_data), iptr(X_row), iptr(X_col), dptr(ts), dptr(tr), dptr(bs), dptr(br), dptr(out)))
    return out.astype(dtype, copy=False)


def compute_Xphi_data_numpy(X, theta, beta, theta_ix=None):
    """hpf_numba.py:117-125 keeps a second, single-threaded implementation for
    `single_process=True`; here it is the same device kernel."""
    ts, tr = theta.vi_shape, theta.vi_rate
    if theta_ix is not None:
        ts, tr = ts[theta_ix], tr[theta_ix]
    return compute_Xphi_data(X.data, X.row, X.col, ts, tr, beta.vi_shape, beta.vi_rate)
