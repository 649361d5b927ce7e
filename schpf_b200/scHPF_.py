"""sklearn-style estimator with the public surface of the reference's
schpf/scHPF_.py -- `scHPF(nfactors=K).fit(X)`, `.project()`, `.transform()`,
`.cell_score()`, `.gene_score()`, `HPF_Gamma`, `run_trials` -- whose CAVI loop
runs on a B200 through `schpf_b200.engine.CaviEngine`.

What stays on the host, deliberately: the empirical hyperparameters b', d'
and the random initialisation (so they are bit-identical to the reference for a
given numpy seed, scHPF_.py:50-70, :783-879) and the scalar convergence logic
(scHPF_.py:717-778).  Everything that touches nnz or (cells+genes) x K numbers
is on the device; the host sees one loss scalar per check.
"""
from copy import deepcopy
from warnings import warn

import numpy as np
from scipy.sparse import coo_matrix
from scipy.special import digamma, gammaln
from sklearn.base import BaseEstimator
import joblib

from . import loss as ls
from .cavi_loop import FullBatchLoop, MinibatchLoop
from .engine import CaviEngine

# tests substitute an oracle-backed engine here; product code never does
_engine_factory = None
# float32 models use the fp32 sweep kernels (False: fp64 arithmetic on fp32-rounded inputs, as in round 1)
FP32_SWEEP = True


# ---- helpers for cell-sharded fits (torch.distributed; any backend) ----------------
def _dist_device(group, device=None):
    """Where collective operands live: the model's own CUDA device on NCCL groups (not whatever
    torch's current device happens to be), the host otherwise."""
    import torch
    import torch.distributed as dist
    if dist.get_backend(group) != "nccl":
        return torch.device("cpu")
    return torch.device("cuda", torch.cuda.current_device() if device is None else int(device))


def _global_row_offset(group, ncells, device=None):
    """Index of this rank's first cell in the rank-ordered concatenation of all shards."""
    import torch
    import torch.distributed as dist
    dev, world = _dist_device(group, device), dist.get_world_size(group)
    n = torch.tensor([int(ncells)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    return int(sum(int(c.item()) for c in counts[:dist.get_rank(group)]))


def _total_cells(group, ncells, device=None):
    import torch
    import torch.distributed as dist
    n = torch.tensor([int(ncells)], dtype=torch.int64, device=_dist_device(group, device))
    dist.all_reduce(n, group=group)
    return int(n.item())


def _mean_over_ranks(group, value, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=_dist_device(group, device))
    dist.all_reduce(t, group=group)
    return float(t.item()) / dist.get_world_size(group)


def _global_mean_var_ratio(group, local_sums, sharded_axis, device=None):
    """mean / population variance of the totals along one axis of the sharded matrix.
    Cell totals (sharded axis) are pooled as moments, gene totals are summed over ranks."""
    import torch
    import torch.distributed as dist
    dev = _dist_device(group, device)
    if sharded_axis:
        m = torch.tensor([local_sums.sum(), (local_sums ** 2).sum(), float(local_sums.shape[0])],
                         dtype=torch.float64, device=dev)
        dist.all_reduce(m, group=group)
        s1, s2, n = m.tolist()
        mean = s1 / n
        return mean / (s2 / n - mean * mean)
    t = torch.as_tensor(local_sums, dtype=torch.float64).to(dev)
    dist.all_reduce(t, group=group)
    tot = t.cpu().numpy()
    return np.mean(tot) / np.var(tot)


def _replicate_from_rank0(group, *gammas, device=None):
    """Every rank gets rank 0's copies (the gene side must be identical everywhere)."""
    import torch
    import torch.distributed as dist
    dev, src, out = _dist_device(group, device), (dist.get_global_rank(group, 0) if group is not None else 0), []
    for g in gammas:
        pair = []
        for arr in (g.vi_shape, g.vi_rate):
            t = torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64)).to(dev)
            dist.broadcast(t, src=src, group=group)
            pair.append(t.cpu().numpy().astype(arr.dtype, copy=False))
        out.append(HPF_Gamma(*pair))
    return out


def _all_gather_rows(group, arr, device=None):
    """Rows of every rank's `arr` (same trailing shape, different row counts) in rank order."""
    import torch
    import torch.distributed as dist
    dev, world = _dist_device(group, device), dist.get_world_size(group)
    n = torch.tensor([arr.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    padded = np.zeros((max(counts),) + arr.shape[1:], dtype=np.float64)
    padded[:arr.shape[0]] = arr
    mine = torch.as_tensor(padded).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return np.concatenate([p.cpu().numpy()[:c] for p, c in zip(parts, counts)]).astype(arr.dtype, copy=False)


def _shared_seed(group, rng=np.random):
    """One random-phi seed for all ranks, drawn from rank 0's numpy stream."""
    import torch.distributed as dist
    box = [int(rng.randint(0, 2 ** 31 - 1))]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return box[0]


class HPF_Gamma(object):
    """A family of independent Gamma(vi_shape, vi_rate) variational distributions
    (reference: scHPF_.py:27-178)."""

    @staticmethod
    def random_gamma_factory(dims, shape_prior, rate_prior, dtype=np.float64, rng=np.random):
        """U(0.5 p, 1.5 p) draws, shape first then rate (scHPF_.py:50-70): the
        order matters for reproducing a seeded reference run.  `rng`: numpy's global stream
        (the reference's) or a RandomState of the caller's (trials on several threads)."""
        lo_hi = lambda p: (0.5 * p, 1.5 * p)
        vi_shape = rng.uniform(*lo_hi(shape_prior), dims).astype(dtype)
        vi_rate = rng.uniform(*lo_hi(rate_prior), dims).astype(dtype)
        return HPF_Gamma(vi_shape, vi_rate)

    def __init__(self, vi_shape, vi_rate):
        assert vi_shape.shape == vi_rate.shape
        assert vi_shape.dtype == vi_rate.dtype
        assert np.all(vi_shape > 0)
        assert np.all(vi_rate > 0)
        self.vi_shape = vi_shape
        self.vi_rate = vi_rate
        self.dtype = vi_shape.dtype

    def __eq__(self, other):
        if not isinstance(other, self.__class__):
            return False
        return (np.array_equal(self.vi_shape, other.vi_shape)
                and np.array_equal(self.vi_rate, other.vi_rate)
                and self.dtype == other.dtype)

    @property
    def dims(self):
        assert self.vi_shape.shape == self.vi_rate.shape
        return self.vi_shape.shape

    @property
    def e_x(self):
        """E[x] = shape / rate"""
        return self.vi_shape / self.vi_rate

    @property
    def e_logx(self):
        """E[log x] = digamma(shape) - log(rate)"""
        return digamma(self.vi_shape) - np.log(self.vi_rate)

    @property
    def entropy(self):
        return (self.vi_shape - np.log(self.vi_rate) + gammaln(self.vi_shape)
                + (1 - self.vi_shape) * digamma(self.vi_shape))

    def sample(self, nsamples=1):
        """(dims..., nsamples) draws from the variational distributions"""
        draws = [np.random.gamma(self.vi_shape, 1 / self.vi_rate).T for _ in range(nsamples)]
        return np.stack(draws).T

    def combine(self, other, other_ixs):
        """Interleave `other`'s rows at positions `other_ixs` of the merged family."""
        assert other.dims[0] == len(other_ixs)
        assert len(np.unique(other_ixs)) == len(other_ixs)
        assert self.dims[0] + other.dims[0] > np.max(other_ixs)
        new_dims = [self.dims[0] + other.dims[0], *self.dims[1:]]
        self_ixs = np.setdiff1d(np.arange(new_dims[0]), other_ixs)
        merged = []
        for mine, theirs in ((self.vi_shape, other.vi_shape), (self.vi_rate, other.vi_rate)):
            arr = np.empty(new_dims, dtype=self.dtype)
            arr[self_ixs] = mine
            arr[other_ixs] = theirs
            merged.append(arr)
        return HPF_Gamma(*merged)


class scHPF(BaseEstimator):
    """single-cell Hierarchical Poisson Factorization (Levitin et al., MSB 2019).

    Same constructor arguments and defaults as the reference estimator
    (scHPF_.py:225-246); `device` is the only addition: a CUDA ordinal, or a list of ordinals to
    shard the cells of one fit over several GPUs from this one process (schpf_b200/multi.py).
    """

    def __init__(self, nfactors, a=0.3, ap=1, bp=None, c=0.3, cp=1, dp=None,
                 min_iter=30, max_iter=1000, check_freq=10, epsilon=0.001,
                 better_than_n_ago=5, dtype=np.float64, xi=None, theta=None,
                 eta=None, beta=None, loss=[], verbose=True, device=0):
        from . import __version__
        self.version = __version__
        self.nfactors = nfactors
        self.a = a
        self.ap = ap
        self.bp = bp
        self.c = c
        self.cp = cp
        self.dp = dp
        self.min_iter = min_iter
        self.max_iter = max_iter
        self.check_freq = check_freq
        self.epsilon = epsilon
        self.better_than_n_ago = better_than_n_ago
        self.dtype = dtype
        self.verbose = verbose
        self.device = device
        self.xi = xi
        self.eta = eta
        self.theta = theta
        self.beta = beta
        self.loss = []

    # random draws (initialisation, t == 0 Dirichlet, minibatch shuffle): numpy's global stream like
    # the reference unless a trial runner gave this model a stream of its own (trials.run_trials_pool)
    @property
    def _rng(self):
        return self.__dict__.get("_rng_state") or np.random

    def set_random_state(self, rng):
        """`rng`: a numpy RandomState (or None for numpy's global stream, the default)."""
        self.__dict__["_rng_state"] = rng

    # a and c accept -2 meaning 1/sqrt(K) (scHPF_.py:284-319)
    def _sqrtk_or(self, val):
        if val == -2:
            if self.nfactors is None:
                raise ValueError('Can only set a as a function of nfactors when'
                                 ' nfactors is not None')
            return 1 / np.sqrt(self.nfactors)
        assert val > 0
        return val

    @property
    def a(self):
        try:
            return self._a
        except AttributeError:
            warn('Automatically using a=0.3 (model saved without it).', RuntimeWarning)
            return 0.3

    @a.setter
    def a(self, val):
        self._a = self._sqrtk_or(val)

    @property
    def c(self):
        try:
            return self._c
        except AttributeError:
            warn('Automatically using c=0.3 (model saved without it).', RuntimeWarning)
            return 0.3

    @c.setter
    def c(self, val):
        self._c = self._sqrtk_or(val)

    @property
    def ngenes(self):
        return self.eta.dims[0] if self.eta is not None else None

    @property
    def ncells(self):
        return self.xi.dims[0] if self.xi is not None else None

    # ---- scores (scHPF_.py:332-369, 506-523) -------------------------------
    def _score(self, capacity, loading):
        assert loading.dims[0] == capacity.dims[0]
        return loading.e_x * capacity.e_x[:, None]

    def cell_score(self, xi=None, theta=None):
        return self._score(self.xi if xi is None else xi,
                           self.theta if theta is None else theta)

    def gene_score(self, eta=None, beta=None):
        return self._score(self.eta if eta is None else eta,
                           self.beta if beta is None else beta)

    # ---- losses (scHPF_.py:372-422) ----------------------------------------
    def pois_llh_pointwise(self, X, theta=None, beta=None):
        return ls.pois_llh_pointwise(X=X, theta=self.theta if theta is None else theta,
                                     beta=self.beta if beta is None else beta)

    def cellmean_negative_pois_llh(self, X, theta=None, beta=None):
        """Mean negative llh of the stored entries, averaged per cell (scHPF_.py:395-411): per cell
        the sum over its triples divided by the number of DISTINCT genes among them (the reference
        takes the count from a csr copy, which merges duplicate triples).  nan for empty cells."""
        theta = self.theta if theta is None else theta
        assert theta.vi_shape.shape[0] == X.shape[0]
        llh = self.pois_llh_pointwise(X=X, theta=theta, beta=beta)
        ncells, ngenes = X.shape
        sums = np.bincount(X.row, weights=-llh, minlength=ncells)
        distinct = np.unique(np.asarray(X.row, dtype=np.int64) * ngenes + X.col)
        counts = np.bincount(distinct // ngenes, minlength=ncells)
        with np.errstate(invalid='ignore', divide='ignore'):
            return sums / counts

    def mean_negative_pois_llh(self, X, theta=None, beta=None, **kwargs):
        return ls.mean_negative_pois_llh(X=X, theta=self.theta if theta is None else theta,
                                         beta=self.beta if beta is None else beta)

    # ---- fit / project / transform -----------------------------------------
    def fit(self, X, **kwargs):
        """Fit the model to a cells x genes sparse count matrix (scHPF_.py:425-445).
        Keyword arguments are those of `_fit`."""
        (self.bp, self.dp, self.xi, self.eta, self.theta, self.beta,
         self.loss) = self._fit(X, **kwargs)
        return self

    def project(self, X, recalc_bp=False, replace=False, min_iter=2, max_iter=50,
                check_freq=2, **kwargs):
        """Infer xi/theta for new cells with beta/eta frozen (scHPF_.py:448-503).
        Returns a model (replace=False) or the loss list (replace=True)."""
        if replace and recalc_bp:
            raise ValueError('Cannot replace `bp` with recalculated value')
        model = self if replace else deepcopy(self)
        if recalc_bp:
            model.bp = None
        (bp, _, xi, _, theta, _, loss) = model._fit(
            X, min_iter=min_iter, max_iter=max_iter, check_freq=check_freq,
            freeze_genes=True, **kwargs)
        if replace:
            self.xi, self.theta = xi, theta
            return loss
        model.bp, model.xi, model.theta, model.loss = bp, xi, theta, loss
        return model

    def transform(self, X, **kwargs):
        """sklearn-convention alias: cell scores of `X` projected onto this model
        (the reference has no `transform`)."""
        return self.project(X, replace=False, **kwargs).cell_score()

    def gather_cells(self, process_group=None):
        """After a cell-sharded `fit(X_shard, process_group=...)`: a copy of this model whose
        xi / theta hold the cells of ALL ranks in rank order (collective; every rank gets it).
        Genes, hyperparameters and the loss are already the same everywhere."""
        full = deepcopy(self)
        gather = lambda arr: _all_gather_rows(process_group, arr, self.device)
        full.xi = HPF_Gamma(gather(self.xi.vi_shape), gather(self.xi.vi_rate))
        full.theta = HPF_Gamma(gather(self.theta.vi_shape), gather(self.theta.vi_rate))
        return full

    def fit_transform(self, X, y=None, **kwargs):
        return self.fit(X, **kwargs).cell_score()

    # ---- the loop ------------------------------------------------------------
    def _new_engine(self, ncells, ngenes, **options):
        factory = _engine_factory or CaviEngine
        if FP32_SWEEP and np.dtype(self.dtype) == np.float32:
            # float32 models (scHPF_.py:225-246 `dtype`): the sweeps run in fp32 like the reference's
            # float32 numba kernels; the state and the rate / digamma updates stay fp64
            options.setdefault("precision", 32)
        if isinstance(self.device, (list, tuple)):
            if len(self.device) > 1:
                # several GPUs from this one process: cells sharded over them (schpf_b200/multi.py)
                from .multi import LocalShardedEngine
                return LocalShardedEngine(ncells, ngenes, self.nfactors, devices=self.device,
                                          engine_factory=factory, **options)
            return factory(ncells, ngenes, self.nfactors, device=self.device[0], **options)
        return factory(ncells, ngenes, self.nfactors, device=self.device, **options)

    def _fit(self, X, freeze_genes=False, reinit=True, loss_function=None,
             min_iter=None, max_iter=None, epsilon=None, check_freq=None,
             single_process=False, checkstep_function=None, verbose=None,
             batchsize=None, beta_theta_simultaneous=False, loss_smoothing=1,
             process_group=None):
        """Host driver of the device CAVI loop; arguments and return value as
        scHPF_.py:526-604.  `single_process` is accepted and ignored (there is
        one device path).

        `batchsize` cells per iteration run as `cavi_loop.MinibatchLoop` (same kernels on the
        batch's rows, the reference's `batched` update order).

        `process_group` (the only addition): a torch.distributed group whose ranks each pass
        their own contiguous shard of the cells as `X` (all genes).  b', d' come from the whole
        matrix, eta/beta are rank 0's draws on every rank, the loss is the loss over all
        cells, and the returned xi/theta are this rank's rows.  A custom `loss_function` is
        evaluated by every rank on ITS cells (that is all `xi` / `theta` hold) and the ranks'
        values are averaged, so that every rank takes the same stopping decision;
        `checkstep_function` sees the local cells only."""
        assert loss_smoothing > 0
        from ._lib import MAX_FACTORS
        if self.nfactors > MAX_FACTORS:
            raise ValueError('nfactors={} is above the {} factors the CUDA sweeps are built for'
                             .format(self.nfactors, MAX_FACTORS))
        nfactors, (ncells, ngenes) = self.nfactors, X.shape
        a, ap, c, cp = self.a, self.ap, self.c, self.cp
        ncells_total = ncells if process_group is None else _total_cells(process_group, ncells, self.device)
        batched = batchsize is not None and 1 < batchsize <= ncells_total       # scHPF_.py:627 (all ranks' cells)
        multi_device = isinstance(self.device, (list, tuple)) and len(self.device) > 1
        if batched and multi_device:
            raise NotImplementedError('minibatches (batchsize) with a list of devices: shard the cells over one '
                                      'process per GPU (process_group=) instead')
        if process_group is not None and multi_device:
            raise ValueError('give either a process group (one process per GPU) or a list of devices '
                             '(one process for all of them), not both')

        bp, dp, xi, eta, theta, beta = self._setup(X, freeze_genes, reinit, process_group=process_group)
        # capacity shapes are constants of the fit (scHPF_.py:614-618)
        xi.vi_shape[:] = ap + nfactors * a
        if not freeze_genes:
            eta.vi_shape[:] = cp + nfactors * c

        min_iter = self.min_iter if min_iter is None else min_iter
        max_iter = self.max_iter if max_iter is None else max_iter
        epsilon = self.epsilon if epsilon is None else epsilon   # read but unused, as in the reference (:639 vs :752)
        check_freq = self.check_freq if check_freq is None else check_freq
        verbose = self.verbose if verbose is None else verbose

        hyper = (a, ap, bp, c, cp, dp)
        state = dict(theta=(theta.vi_shape, theta.vi_rate), beta=(beta.vi_shape, beta.vi_rate),
                     xi=(xi.vi_shape, xi.vi_rate), eta=(eta.vi_shape, eta.vi_rate))
        rng = self._rng
        if batched:
            cell_range = None
            if process_group is not None:
                cell_range = (_global_row_offset(process_group, ncells, self.device), ncells_total)
            loop = MinibatchLoop(self._new_engine, X, hyper, state, nfactors, batchsize,
                                 freeze_genes, beta_theta_simultaneous, rng=rng, process_group=process_group,
                                 shared_seed=_shared_seed, cell_range=cell_range)
        else:
            engine_options = {}
            if process_group is not None:
                # the device's t == 0 draw is keyed by the GLOBAL cell index
                engine_options["row_offset"] = _global_row_offset(process_group, ncells, self.device)
            loop = FullBatchLoop(self._new_engine, X, hyper, state, nfactors, freeze_genes,
                                 beta_theta_simultaneous, process_group, _shared_seed, rng=rng,
                                 engine_options=engine_options)
        try:
            def host_state():
                st = loop.host_state()
                wrap = lambda pair: HPF_Gamma(pair[0].astype(self.dtype, copy=False),
                                              pair[1].astype(self.dtype, copy=False))
                return (wrap(st["xi"]), eta if freeze_genes else wrap(st["eta"]),
                        wrap(st["theta"]), beta if freeze_genes else wrap(st["beta"]))

            loss, unsmoothed_loss, pct_change = [], [], []
            # the reference also leaves the loop once t reaches self.max_iter (:776-777)
            n_total = min(max_iter, self.max_iter + 1)
            t = 0
            while t < n_total:
                next_check = t if t % check_freq == 0 else (t // check_freq + 1) * check_freq
                last = min(next_check, n_total - 1)
                loop.run(t, last - t + 1, reinit)
                t = last + 1
                if last % check_freq != 0:
                    continue

                # ---- loss bookkeeping and stopping rules (scHPF_.py:717-774) ----
                tc = last
                if loss_function is None:
                    curr = loop.loss()
                else:
                    if getattr(loss_function, 'accepts_device_loop', False):
                        # loss.ProjectionLoss: takes beta / eta from the training engine device to device
                        curr = loss_function.device_call(loop, a=a, ap=ap, bp=bp, c=c, cp=cp, dp=dp)
                    else:
                        hxi, heta, htheta, hbeta = host_state()
                        curr = loss_function(a=a, ap=ap, bp=bp, c=c, cp=cp, dp=dp,
                                             xi=hxi, eta=heta, theta=htheta, beta=hbeta)
                    if process_group is not None:
                        curr = _mean_over_ranks(process_group, curr, self.device)
                unsmoothed_loss.append(curr)
                if len(unsmoothed_loss) > loss_smoothing:
                    unsmoothed_loss = unsmoothed_loss[1:]
                loss.append(np.mean(unsmoothed_loss))
                if len(loss) >= 2:
                    curr, prev = loss[-1], loss[-2]
                    pct_change.append(100 * (curr - prev) / np.abs(prev))
                else:
                    pct_change.append(100)
                if verbose:
                    print('[Iter. {0: >4}]  loss:{1:.6f}  pct:{2:.9f}'.format(tc, curr, pct_change[-1]))
                if checkstep_function is not None:
                    hxi, heta, htheta, hbeta = host_state()
                    checkstep_function(bp=bp, dp=dp, xi=hxi, eta=heta, theta=htheta, beta=hbeta, t=tc)

                if len(loss) > 3 and tc >= min_iter:
                    current_small = np.abs(pct_change[-1]) < self.epsilon
                    prev_small = np.abs(pct_change[-2]) < self.epsilon
                    not_inflection = not ((np.abs(loss[-3]) < np.abs(prev))
                                          and (np.abs(prev) > np.abs(curr)))
                    if current_small and prev_small and not_inflection:
                        if verbose:
                            print('converged')
                        break
                    if len(loss) > self.better_than_n_ago and self.better_than_n_ago:
                        nprev = loss[-self.better_than_n_ago]
                        worse_than_n_ago = np.abs(nprev) < np.abs(curr)
                        getting_worse = np.abs(prev) < np.abs(curr)
                        if worse_than_n_ago and getting_worse:
                            if verbose:
                                print('getting worse break')
                            break

            xi, eta, theta, beta = host_state()
        finally:
            loop.close()
        return (bp, dp, xi, eta, theta, beta, loss)

    # ---- setup (host, bit-identical to the reference) ------------------------
    def _setup(self, X, freeze_genes=False, reinit=True, clip=True, process_group=None):
        """Empirical b', d' and (re)initialised distributions, drawing from the
        numpy global RNG in the reference's order xi, theta, eta, beta
        (scHPF_.py:783-844)."""
        nfactors, (ncells, ngenes) = self.nfactors, X.shape
        a, ap, c, cp = self.a, self.ap, self.c, self.cp
        xi, eta, theta, beta = self.xi, self.eta, self.theta, self.beta
        bp, dp = self._get_empirical_hypers(X, freeze_genes, clip, process_group=process_group)
        make = lambda *args, **kw: HPF_Gamma.random_gamma_factory(*args, rng=self._rng, **kw)
        if reinit or xi is None:
            xi = make((ncells,), ap, bp, dtype=self.dtype)
        if reinit or theta is None:
            theta = make((ncells, nfactors), a, bp, dtype=self.dtype)
        if freeze_genes:
            if eta is None or beta is None:
                raise ValueError('To fit with frozen gene variational distributions '
                                 '(`freeze_genes`==True), eta and beta must be set to '
                                 'valid HPF_Gamma instances.')
        else:
            if reinit or eta is None:
                eta = make((ngenes,), cp, dp, dtype=self.dtype)
            if reinit or beta is None:
                beta = make((ngenes, nfactors), c, dp, dtype=self.dtype)
            if process_group is not None:
                eta, beta = _replicate_from_rank0(process_group, eta, beta, device=self.device)
        return (bp, dp, xi, eta, theta, beta)

    def _get_empirical_hypers(self, X, freeze_genes=False, clip=True, process_group=None):
        """b' = a' * mean/var of the cell totals, d' = c' * mean/var of the gene
        totals (population variance), d' clipped to b'/1000 (scHPF_.py:847-879).
        With a process group the totals are those of the whole (sharded) matrix."""
        bp, dp = self.bp, self.dp

        def mean_var_ratio(axis):
            axis_sum = X.sum(axis=axis)
            if process_group is not None:
                return _global_mean_var_ratio(process_group, np.asarray(axis_sum, dtype=np.float64).ravel(),
                                              sharded_axis=(axis == 1), device=self.device)
            return np.mean(axis_sum) / np.var(axis_sum)
        if bp is None:
            bp = self.ap * mean_var_ratio(1)
        if dp is None:
            if freeze_genes:
                raise ValueError('dp is None and cannot be set when freeze_genes is True.')
            dp = self.cp * mean_var_ratio(0)
            if clip and bp > 1000 * dp:
                old_val = dp
                dp = bp / 1000
                print('Clipping dp: was {} now {}'.format(old_val, dp))
        return bp, dp

    def _initialize(self, X, freeze_genes=False):
        """Set random distributions and empirical hyperparameters on self (scHPF_.py:882-892)."""
        (self.bp, self.dp, self.xi, self.eta, self.theta,
         self.beta) = self._setup(X, freeze_genes, reinit=True)


def load_model(file_name):
    """scHPF_.py:895-909"""
    return joblib.load(file_name)


def save_model(model, file_name):
    """scHPF_.py:912-925"""
    joblib.dump(model, file_name)


def combine_across_cells(x, y, y_ixs):
    """Model with the cells of `y` merged into those of `x` at rows `y_ixs`; both
    must share d', eta and beta.  bp becomes None when the two differ
    (scHPF_.py:928-965)."""
    assert x.dp == y.dp
    assert x.eta == y.eta
    assert x.beta == y.beta
    merged = deepcopy(x)
    if y.bp != x.bp:
        merged.bp = None
    merged.xi = x.xi.combine(y.xi, y_ixs)
    merged.theta = x.theta.combine(y.theta, y_ixs)
    return merged
