"""CPU oracle (numpy) for the scHPF CAVI hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy/scipy, the algorithm of the reference's
numba kernels and of the body of its CAVI loop.  It exists so that the CUDA
path can be checked on machines where the reference itself is not present.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
leg may import it; nothing under ``schpf_b200/`` does.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference (``/root/reference/schpf``) in the build container, runs it on
seeded inputs and stores its outputs under ``tests/golden/*.npz``;
``tests/test_oracle.py`` checks every function below against those files.

Reference lines followed (all paths relative to the reference checkout):
  schpf/hpf_numba.py:16-22    psi / gammaln (SciPy cython_special symbols)
  schpf/hpf_numba.py:25-51    compute_pois_llh
  schpf/hpf_numba.py:55-114   compute_Xphi_data
  schpf/hpf_numba.py:129-156  compute_loading_shape_update
  schpf/hpf_numba.py:160-177  compute_loading_rate_update
  schpf/hpf_numba.py:181-188  compute_capacity_rate_update
  schpf/loss.py:107-168       pois_llh_pointwise / mean_negative_pois_llh
  schpf/scHPF_.py:605-780     _fit iteration order (full batch, simultaneous, minibatch)
  schpf/scHPF_.py:847-879     _get_empirical_hypers
"""
import numpy as np
from scipy.special import digamma, gammaln


def psi(x):
    """hpf_numba.py:16-18 -- SciPy's real digamma."""
    return digamma(np.asarray(x, dtype=np.float64))


def cgammaln(x):
    """hpf_numba.py:20-22 -- SciPy's gammaln."""
    return gammaln(np.asarray(x, dtype=np.float64))


def e_logx(vi_shape, vi_rate):
    """scHPF_.py:108-111 / hpf_numba.py:83-94."""
    return digamma(vi_shape) - np.log(vi_rate)


def compute_Xphi_data(X_data, X_row, X_col, theta_vi_shape, theta_vi_rate,
                      beta_vi_shape, beta_vi_rate):
    """hpf_numba.py:55-114.  Evaluation order of the last line is (y*rho)/sum
    as in the reference (:111-112)."""
    theta_e_logx = e_logx(theta_vi_shape, theta_vi_rate)
    beta_e_logx = e_logx(beta_vi_shape, beta_vi_rate)
    logrho = theta_e_logx[X_row, :] + beta_e_logx[X_col, :]
    largest_in = logrho.max(axis=1)
    rho_shift = np.exp(logrho - largest_in[:, None])
    # sequential left-to-right sum over k, like the reference's scalar loop
    normalizer = np.zeros(rho_shift.shape[0], dtype=rho_shift.dtype)
    for k in range(rho_shift.shape[1]):
        normalizer += rho_shift[:, k]
    return (X_data[:, None] * rho_shift) / normalizer[:, None]


def compute_loading_shape_update(Xphi_data, X_keep, nkeep, shape_prior):
    """hpf_numba.py:129-156.  Sequential accumulation in nnz order starting
    from the prior (np.add.at is unbuffered and in index order)."""
    nnz, nfactors = Xphi_data.shape
    result = shape_prior * np.ones((nkeep, nfactors), dtype=Xphi_data.dtype)
    np.add.at(result, X_keep, Xphi_data)
    return result


def compute_loading_rate_update(prior_vi_shape, prior_vi_rate,
                                other_loading_vi_shape, other_loading_vi_rate):
    """hpf_numba.py:160-177."""
    olvs, olvr = other_loading_vi_shape, other_loading_vi_rate
    other_sum = np.zeros(olvs.shape[1], dtype=prior_vi_shape.dtype)
    e_x = olvs / olvr
    for i in range(olvs.shape[0]):          # sequential over rows like :168-170
        other_sum += e_x[i]
    prior_e_x = prior_vi_shape / prior_vi_rate
    return prior_e_x[:, None] + other_sum[None, :]


def compute_capacity_rate_update(loading_vi_shape, loading_vi_rate, prior_rate):
    """hpf_numba.py:181-188 (k-outer accumulation order)."""
    result = prior_rate * np.ones((loading_vi_shape.shape[0],),
                                  dtype=loading_vi_shape.dtype)
    for k in range(loading_vi_shape.shape[1]):
        result += loading_vi_shape[:, k] / loading_vi_rate[:, k]
    return result


def compute_pois_llh(X_data, X_row, X_col, theta_vi_shape, theta_vi_rate,
                     beta_vi_shape, beta_vi_rate):
    """hpf_numba.py:25-51."""
    theta_e_x = theta_vi_shape / theta_vi_rate
    beta_e_x = beta_vi_shape / beta_vi_rate
    e_rate = np.zeros(X_data.shape[0], dtype=theta_e_x.dtype)
    for k in range(theta_e_x.shape[1]):
        e_rate += theta_e_x[X_row, k] * beta_e_x[X_col, k]
    return X_data * np.log(e_rate) - e_rate - gammaln(X_data + 1.0)


def mean_negative_pois_llh(X_data, X_row, X_col, theta_vi_shape, theta_vi_rate,
                           beta_vi_shape, beta_vi_rate):
    """loss.py:142-168."""
    return np.mean(-compute_pois_llh(X_data, X_row, X_col, theta_vi_shape,
                                     theta_vi_rate, beta_vi_shape, beta_vi_rate))


def empirical_hypers(X, ap, cp, bp=None, dp=None, freeze_genes=False, clip=True):
    """scHPF_.py:847-879 (population variance, optional dp clip)."""
    def mean_var_ratio(X, axis):
        axis_sum = X.sum(axis=axis)
        return np.mean(axis_sum) / np.var(axis_sum)
    if bp is None:
        bp = ap * mean_var_ratio(X, axis=1)
    if dp is None:
        if freeze_genes:
            raise ValueError('dp is None and cannot be set when freeze_genes is True.')
        dp = cp * mean_var_ratio(X, axis=0)
        if clip and bp > 1000 * dp:
            dp = bp / 1000
    return bp, dp


class State(object):
    """The eight variational arrays + six scalars of scHPF_.py:_fit."""

    def __init__(self, theta_shp, theta_rte, beta_shp, beta_rte,
                 xi_shp, xi_rte, eta_shp, eta_rte):
        f = lambda a: np.array(a, dtype=np.float64, copy=True)
        self.theta_shp, self.theta_rte = f(theta_shp), f(theta_rte)
        self.beta_shp, self.beta_rte = f(beta_shp), f(beta_rte)
        self.xi_shp, self.xi_rte = f(xi_shp), f(xi_rte)
        self.eta_shp, self.eta_rte = f(eta_shp), f(eta_rte)

    def copy(self):
        return State(self.theta_shp, self.theta_rte, self.beta_shp, self.beta_rte,
                     self.xi_shp, self.xi_rte, self.eta_shp, self.eta_rte)

    def arrays(self):
        return (self.theta_shp, self.theta_rte, self.beta_shp, self.beta_rte,
                self.xi_shp, self.xi_rte, self.eta_shp, self.eta_rte)


def cavi_prepare(st, a, ap, c, cp, nfactors, freeze_genes=False):
    """scHPF_.py:614-618 -- constant capacity shapes."""
    st.xi_shp[:] = ap + nfactors * a
    if not freeze_genes:
        st.eta_shp[:] = cp + nfactors * c


def cavi_iteration(X_data, X_row, X_col, st, a, bp, c, dp, freeze_genes=False,
                   Xphi=None, beta_theta_simultaneous=False, batched=False):
    """One pass of the loop body, scHPF_.py:661-714, over the cells in ``st``
    (all cells, or -- ``batched`` -- the rows of one minibatch re-based to 0..).

    ``Xphi`` may be supplied (the t==0 random-phi branch, :652-655).
    Default ordering: beta -> eta -> theta -> xi (the ``not batched`` branch);
    with ``beta_theta_simultaneous`` the gene updates are computed from the
    OLD theta but assigned after the cell updates, which use the OLD beta
    (:666-684)."""
    ncells, ngenes = st.theta_shp.shape[0], st.beta_shp.shape[0]
    if Xphi is None:
        Xphi = compute_Xphi_data(X_data, X_row, X_col, st.theta_shp, st.theta_rte,
                                 st.beta_shp, st.beta_rte)
    if beta_theta_simultaneous:
        if not freeze_genes:
            bvs = compute_loading_shape_update(Xphi, X_col, ngenes, c)
            bvr = compute_loading_rate_update(st.eta_shp, st.eta_rte,
                                              st.theta_shp, st.theta_rte)
        st.theta_shp = compute_loading_shape_update(Xphi, X_row, ncells, a)
        st.theta_rte = compute_loading_rate_update(st.xi_shp, st.xi_rte,
                                                   st.beta_shp, st.beta_rte)
        st.xi_rte = bp + (st.theta_shp / st.theta_rte).sum(1)
        if not freeze_genes:
            st.beta_shp, st.beta_rte = bvs, bvr
            st.eta_rte = dp + (st.beta_shp / st.beta_rte).sum(1)
        return st
    if batched:
        # scHPF_.py:686-694: "cell updates, must do first for batching" -- beta's rate below
        # then sees the NEW theta of the batch (:697-704)
        st.theta_shp = compute_loading_shape_update(Xphi, X_row, ncells, a)
        st.theta_rte = compute_loading_rate_update(st.xi_shp, st.xi_rte,
                                                   st.beta_shp, st.beta_rte)
        st.xi_rte = bp + (st.theta_shp / st.theta_rte).sum(1)
    if not freeze_genes:
        st.beta_shp = compute_loading_shape_update(Xphi, X_col, ngenes, c)
        st.beta_rte = compute_loading_rate_update(st.eta_shp, st.eta_rte,
                                                  st.theta_shp, st.theta_rte)
        st.eta_rte = dp + (st.beta_shp / st.beta_rte).sum(1)
    if batched:
        return st
    st.theta_shp = compute_loading_shape_update(Xphi, X_row, ncells, a)
    st.theta_rte = compute_loading_rate_update(st.xi_shp, st.xi_rte,
                                               st.beta_shp, st.beta_rte)
    st.xi_rte = bp + (st.theta_shp / st.theta_rte).sum(1)
    return st


def cavi_run(X_data, X_row, X_col, st, a, ap, bp, c, cp, dp, n_iter,
             freeze_genes=False, check_freq=None, Xphi0=None):
    """n_iter iterations from ``st`` (modified in place); returns the list of
    mean-negative-llh values taken at t % check_freq == 0 (scHPF_.py:718-730,
    loss_smoothing=1)."""
    K = st.theta_shp.shape[1]
    cavi_prepare(st, a, ap, c, cp, K, freeze_genes)
    loss = []
    for t in range(n_iter):
        cavi_iteration(X_data, X_row, X_col, st, a, bp, c, dp, freeze_genes,
                       Xphi=Xphi0 if t == 0 else None)
        if check_freq is not None and t % check_freq == 0:
            loss.append(mean_negative_pois_llh(
                X_data, X_row, X_col, st.theta_shp, st.theta_rte,
                st.beta_shp, st.beta_rte))
    return loss
