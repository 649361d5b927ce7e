"""ctypes binding of the C oracle (oracle/hpf_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

from . import build as _build

_c_i64 = ctypes.c_int64
_c_int = ctypes.c_int
_c_dbl = ctypes.c_double
_pd = ctypes.POINTER(ctypes.c_double)
_pi = ctypes.POINTER(ctypes.c_int32)

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.OUT
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            path = _build.build()
        _lib = ctypes.CDLL(path)
        _lib.oracle_psi.restype = _c_dbl
        _lib.oracle_psi.argtypes = [_c_dbl]
        _lib.oracle_gammaln.restype = _c_dbl
        _lib.oracle_gammaln.argtypes = [_c_dbl]
        _lib.oracle_num_threads.restype = _c_int
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(_pd if a.dtype == np.float64 else _pi)


def psi(x):
    x = _d(np.atleast_1d(x))
    out = np.empty_like(x)
    lib().oracle_psi_v(_c_i64(x.size), _p(x), _p(out))
    return out


def cgammaln(x):
    x = _d(np.atleast_1d(x))
    out = np.empty_like(x)
    lib().oracle_gammaln_v(_c_i64(x.size), _p(x), _p(out))
    return out


def compute_Xphi_data(X_data, X_row, X_col, ts, tr, bs, br):
    X_data, X_row, X_col = _i(X_data), _i(X_row), _i(X_col)
    ts, tr, bs, br = _d(ts), _d(tr), _d(bs), _d(br)
    nnz, K = X_data.shape[0], ts.shape[1]
    out = np.empty((nnz, K), dtype=np.float64)
    rc = lib().oracle_compute_Xphi_data(_c_i64(nnz), _c_i64(ts.shape[0]), _c_i64(bs.shape[0]),
                                        _c_int(K), _p(X_data), _p(X_row), _p(X_col),
                                        _p(ts), _p(tr), _p(bs), _p(br), _p(out))
    assert rc == 0
    return out


def compute_loading_shape_update(Xphi, X_keep, nkeep, prior):
    Xphi, X_keep = _d(Xphi), _i(X_keep)
    nnz, K = Xphi.shape
    out = np.empty((nkeep, K), dtype=np.float64)
    lib().oracle_compute_loading_shape_update(_c_i64(nnz), _c_int(K), _p(Xphi), _p(X_keep),
                                              _c_i64(nkeep), _c_dbl(prior), _p(out))
    return out


def compute_loading_rate_update(ps, pr, os_, or_):
    ps, pr, os_, or_ = _d(ps), _d(pr), _d(os_), _d(or_)
    n, (m, K) = ps.shape[0], os_.shape
    out = np.empty((n, K), dtype=np.float64)
    lib().oracle_compute_loading_rate_update(_c_i64(n), _c_i64(m), _c_int(K), _p(ps), _p(pr),
                                             _p(os_), _p(or_), _p(out))
    return out


def compute_capacity_rate_update(shp, rte, prior_rate):
    shp, rte = _d(shp), _d(rte)
    n, K = shp.shape
    out = np.empty((n,), dtype=np.float64)
    lib().oracle_compute_capacity_rate_update(_c_i64(n), _c_int(K), _p(shp), _p(rte),
                                              _c_dbl(prior_rate), _p(out))
    return out


def compute_pois_llh(X_data, X_row, X_col, ts, tr, bs, br):
    X_data, X_row, X_col = _i(X_data), _i(X_row), _i(X_col)
    ts, tr, bs, br = _d(ts), _d(tr), _d(bs), _d(br)
    nnz, K = X_data.shape[0], ts.shape[1]
    out = np.empty((nnz,), dtype=np.float64)
    rc = lib().oracle_compute_pois_llh(_c_i64(nnz), _c_i64(ts.shape[0]), _c_i64(bs.shape[0]),
                                       _c_int(K), _p(X_data), _p(X_row), _p(X_col),
                                       _p(ts), _p(tr), _p(bs), _p(br), _p(out))
    assert rc == 0
    return out


def cavi_run(X_data, X_row, X_col, st, a, ap, bp, c, cp, dp, n_iter,
             freeze_genes=False, check_freq=0, nthreads=None):
    """In-place n_iter iterations on an oracle.hpf_numpy.State; returns loss list."""
    X_data, X_row, X_col = _i(X_data), _i(X_row), _i(X_col)
    for name in ("theta_shp", "theta_rte", "beta_shp", "beta_rte",
                 "xi_shp", "xi_rte", "eta_shp", "eta_rte"):
        setattr(st, name, _d(getattr(st, name)))
    K = st.theta_shp.shape[1]
    if nthreads:
        lib().oracle_set_num_threads(_c_int(int(nthreads)))
    cf = int(check_freq or 0)
    cap = (n_iter // cf + 1) if cf > 0 else 1
    loss = np.zeros(cap, dtype=np.float64)
    nloss = _c_int(0)
    rc = lib().oracle_cavi_run(
        _c_i64(X_data.shape[0]), _c_i64(st.theta_shp.shape[0]), _c_i64(st.beta_shp.shape[0]),
        _c_int(K), _p(X_data), _p(X_row), _p(X_col),
        _p(st.theta_shp), _p(st.theta_rte), _p(st.beta_shp), _p(st.beta_rte),
        _p(st.xi_shp), _p(st.xi_rte), _p(st.eta_shp), _p(st.eta_rte),
        _c_dbl(a), _c_dbl(ap), _c_dbl(bp), _c_dbl(c), _c_dbl(cp), _c_dbl(dp),
        _c_int(n_iter), _c_int(1 if freeze_genes else 0), _c_int(cf), _p(loss),
        ctypes.byref(nloss))
    assert rc == 0
    return list(loss[:nloss.value])


def num_threads():
    return int(lib().oracle_num_threads())
