"""Loss functions with the reference's names and calling convention
(schpf/loss.py): data positional, everything else keyword-only, unused keywords
accepted.  The Poisson log-likelihood itself runs on the GPU
(hpf_cuda.compute_pois_llh; during `fit` the engine's resident copy is used
instead and nothing is transferred)."""
import functools

import numpy as np

from .hpf_cuda import compute_pois_llh


def loss_function_for_data(loss_function, X):
    """loss.py:17-34 -- bind the data argument `X`."""
    return functools.partial(loss_function, X=X)


def pois_llh_pointwise(X, *, theta, beta, single_process=False, **kwargs):
    """loss.py:107-139 -- Poisson log-likelihood of every stored entry of X."""
    return compute_pois_llh(X.data, X.row, X.col, theta.vi_shape, theta.vi_rate,
                            beta.vi_shape, beta.vi_rate)


def mean_negative_pois_llh(X, *, theta, beta, single_process=False, **kwargs):
    """loss.py:142-168."""
    return np.mean(-pois_llh_pointwise(X=X, theta=theta, beta=beta))


def projection_loss_function(loss_function, X, nfactors, model_kwargs={}, proj_kwargs={}):
    """loss.py:37-102 -- loss of held-out cells X after projecting them onto the
    model being trained (genes frozen).  Defaults as in the reference:
    reinit=False, max_iter=min_iter=10, no loss checks inside the projection."""
    from .scHPF_ import scHPF
    pmodel = scHPF(nfactors=nfactors, **model_kwargs)
    proj_kwargs = dict(proj_kwargs)
    proj_kwargs.setdefault('reinit', False)
    proj_kwargs.setdefault('max_iter', 10)
    proj_kwargs.setdefault('min_iter', 10)
    proj_kwargs.setdefault('check_freq', proj_kwargs['max_iter'] + 1)

    def _projection_loss_function(*, a, ap, bp, c, cp, dp, eta, beta, **kwargs):
        assert eta.dims[0] == beta.dims[0]
        assert beta.dims[1] == nfactors
        pmodel.a, pmodel.ap, pmodel.bp = a, ap, bp
        pmodel.c, pmodel.cp, pmodel.dp = c, cp, dp
        pmodel.eta, pmodel.beta = eta, beta
        pmodel.project(X, replace=True, **proj_kwargs)
        return loss_function(X, a=pmodel.a, ap=pmodel.ap, bp=pmodel.bp, c=pmodel.c,
                             cp=pmodel.cp, dp=pmodel.dp, xi=pmodel.xi, eta=pmodel.eta,
                             theta=pmodel.theta, beta=pmodel.beta)

    return _projection_loss_function
