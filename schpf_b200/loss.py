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


class ProjectionLoss(object):
    """loss.py:37-102 -- loss of held-out cells X after projecting them onto the model being
    trained (genes frozen).  Defaults as in the reference: reinit=False, max_iter=min_iter=10,
    no loss checks inside the projection.

    Called like the reference's closure (keyword-only host arrays) it does what the reference
    does: a projection through `scHPF.project`.  `scHPF._fit` instead calls `device_call` with its
    loop: the validation cells then keep ONE engine and device layout for the whole fit, beta / eta
    arrive device to device from the training engine (`schpf_copy_gene_state`), xi / theta of the
    validation cells stay resident between checks (the reference's `replace=True`), and the loss
    is taken on the resident matrix -- nothing but the loss scalar crosses PCIe per check."""
    accepts_device_loop = True

    def __init__(self, loss_function, X, nfactors, model_kwargs={}, proj_kwargs={}):
        from .scHPF_ import scHPF
        self.loss_function, self.X, self.nfactors = loss_function, X, nfactors
        self.pmodel = scHPF(nfactors=nfactors, **model_kwargs)
        self.proj_kwargs = dict(proj_kwargs)
        self.proj_kwargs.setdefault('reinit', False)
        self.proj_kwargs.setdefault('max_iter', 10)
        self.proj_kwargs.setdefault('min_iter', 10)
        self.proj_kwargs.setdefault('check_freq', self.proj_kwargs['max_iter'] + 1)
        self._engine = None
        self._fallbacks = 0

    # -- the reference's calling convention ------------------------------------------------
    def __call__(self, *, a, ap, bp, c, cp, dp, eta, beta, **kwargs):
        pmodel, nfactors = self.pmodel, self.nfactors
        assert eta.dims[0] == beta.dims[0]
        assert beta.dims[1] == nfactors
        pmodel.a, pmodel.ap, pmodel.bp = a, ap, bp
        pmodel.c, pmodel.cp, pmodel.dp = c, cp, dp
        pmodel.eta, pmodel.beta = eta, beta
        pmodel.project(self.X, replace=True, **self.proj_kwargs)
        return self.loss_function(self.X, a=pmodel.a, ap=pmodel.ap, bp=pmodel.bp, c=pmodel.c,
                                  cp=pmodel.cp, dp=pmodel.dp, xi=pmodel.xi, eta=pmodel.eta,
                                  theta=pmodel.theta, beta=pmodel.beta)

    # -- the device path -------------------------------------------------------------------
    def _device_path_applies(self, train_engine):
        kw, pm = self.proj_kwargs, self.pmodel
        return (train_engine is not None and hasattr(train_engine, 'copy_gene_state_from')
                and self.loss_function is mean_negative_pois_llh
                and set(kw) <= {'reinit', 'max_iter', 'min_iter', 'check_freq', 'verbose'}
                and not kw['reinit'] and kw['check_freq'] > kw['max_iter']
                and np.dtype(pm.dtype) == np.float64
                and (self._engine is not None or (pm.xi is None and pm.theta is None)))

    def device_call(self, loop, *, a, ap, bp, c, cp, dp):
        train = loop.gene_state_engine() if hasattr(loop, 'gene_state_engine') else None
        if not self._device_path_applies(train):
            from .scHPF_ import HPF_Gamma
            self._fallbacks += 1
            st = loop.host_state()
            wrap = lambda pair: HPF_Gamma(np.asarray(pair[0]), np.asarray(pair[1]))
            return self(a=a, ap=ap, bp=bp, c=c, cp=cp, dp=dp, eta=wrap(st['eta']), beta=wrap(st['beta']))
        from .scHPF_ import HPF_Gamma
        X, K, pm = self.X, self.nfactors, self.pmodel
        ncells, ngenes = X.shape
        if self._engine is None:
            # an engine of the training engine's kind (CaviEngine; the tests' oracle-backed double) on its device
            eng = type(train)(ncells, ngenes, K, device=train.device)
            try:
                eng.set_coo(X.row, X.col, X.data)
                # the draws scHPF._setup makes on the first projection, in its order (scHPF_.py:821-826)
                xi = HPF_Gamma.random_gamma_factory((ncells,), ap, bp, dtype=np.float64, rng=pm._rng)
                theta = HPF_Gamma.random_gamma_factory((ncells, K), a, bp, dtype=np.float64, rng=pm._rng)
                xi.vi_shape[:] = ap + K * a                                      # scHPF_.py:616
                eng.set_state(theta=(theta.vi_shape, theta.vi_rate), xi=(xi.vi_shape, xi.vi_rate))
            except Exception:
                eng.close()
                raise
            self._engine = eng
        eng = self._engine
        eng.set_hyper(a, ap, bp, c, cp, dp)
        eng.copy_gene_state_from(train)
        n_iter = min(self.proj_kwargs['max_iter'], pm.max_iter + 1)               # scHPF_.py:776-777
        if n_iter > 0:
            eng.step(n_iter, freeze_genes=True)
        return eng.loss()

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def projection_loss_function(loss_function, X, nfactors, model_kwargs={}, proj_kwargs={}):
    """loss.py:37-102: returns the callable described at `ProjectionLoss`."""
    return ProjectionLoss(loss_function, X, nfactors, model_kwargs, proj_kwargs)
