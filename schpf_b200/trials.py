"""Model selection over random restarts with the reference's signatures
(schpf/scHPF_.py:968-1332 `run_trials`, `run_trials_pool`).

Host glue only: every trial is an ordinary `scHPF.fit` on a GPU.  Two things
differ from the reference on purpose:

* when the loss is the default (mean negative Poisson log-likelihood of the
  training matrix itself) the trial asks `fit` for its built-in loss, which is
  evaluated on the copy of the matrix already resident in HBM instead of
  re-uploading it at every check; the value is the same number;
* `run_trials_pool` spreads trials over GPUs (`devices`, one worker thread per
  device -- the C ABI releases the GIL) instead of over joblib CPU processes;
  `njobs` / `max_threads` are accepted and ignored.  Each trial is the reference's
  (a full `fit` with `reinit=True`); it draws from a RandomState of its own.
"""
from functools import partial
import threading

import numpy as np

from . import loss as ls
from .scHPF_ import scHPF


def _gene_count_warning(ngenes):
    if ngenes >= 20000:
        print('WARNING: you are running scHPF with {} genes, which is more than the ~20k '
              'protein coding genes in the human genome. We suggest running scHPF on '
              'protein-coding genes only.'.format(ngenes))


def _loss_setup(X, nfactors, vcells, vX, loss_function, check_freq):
    """-> (loss function to hand to fit or None for the resident default, training-loss printer)"""
    default_loss = loss_function is None
    if default_loss:
        loss_function = ls.mean_negative_pois_llh
    if vcells is not None:
        assert X.shape[1] == vcells.shape[1]
    if vX is not None:
        assert vX.shape == X.shape
    if vcells is not None:
        # every loss check projects the validation cells onto the current genes (loss.py:37-102)
        proj_kwargs = dict(reinit=False, min_iter=1, max_iter=min(10, check_freq),
                           check_freq=check_freq + 1, verbose=False)
        fit_loss = ls.projection_loss_function(loss_function, vcells, nfactors, proj_kwargs=proj_kwargs)
        train_loss = ls.loss_function_for_data(loss_function, X)

        def checkstep(**kwargs):
            print('\ttrain:', '{0:.6f}'.format(train_loss(**kwargs)))
        return fit_loss, checkstep
    if default_loss and (vX is None or vX is X):
        return None, None                       # resident matrix, no transfer per check
    return ls.loss_function_for_data(loss_function, X if vX is None else vX), None


def run_trials(X, nfactors, ntrials=5, min_iter=30, max_iter=1000, check_freq=10, epsilon=0.001,
               better_than_n_ago=5, dtype=np.float64, verbose=True, vcells=None, vX=None,
               loss_function=None, model_kwargs={}, return_all=False, reproject=False,
               reproject_kwargs={}, batchsize=0, beta_theta_simultaneous=False, loss_smoothing=1):
    """Train `ntrials` randomly initialised models and keep the one with the lowest final
    loss (scHPF_.py:968-1148).  Returns the best model, or (best, others ordered by loss)
    when `return_all`."""
    _gene_count_warning(X.shape[1])
    best_loss, best_model, best_t = np.finfo(np.float64).max, None, None
    models, losses = [], []
    for t in range(ntrials):
        fit_loss, checkstep = _loss_setup(X, nfactors, vcells, vX, loss_function, check_freq)
        model = scHPF(nfactors=nfactors, min_iter=min_iter, max_iter=max_iter, check_freq=check_freq,
                      epsilon=epsilon, better_than_n_ago=better_than_n_ago, verbose=verbose,
                      dtype=dtype, **model_kwargs)
        model.fit(X, loss_function=fit_loss, checkstep_function=checkstep, batchsize=batchsize,
                  loss_smoothing=loss_smoothing, beta_theta_simultaneous=beta_theta_simultaneous)
        if hasattr(fit_loss, 'close'):
            fit_loss.close()                    # the validation cells' engine
        if reproject:
            print('Reprojecting data...')
            kw = dict(reproject_kwargs, replace=True, reinit=False)
            proj_loss = model.project(X, **kw)
            model.loss.append(proj_loss)
            loss = proj_loss[-1]
        else:
            loss = model.loss[-1]
        if loss < best_loss:
            best_model, best_loss, best_t = model, loss, t
            if verbose:
                print('New best!')
        if return_all:
            models.append(model)
            losses.append(loss)
        if verbose:
            print('Trial {0} loss: {1:.6f}'.format(t, loss))
            print('Best loss: {0:.6f} (trial {1})'.format(best_loss, best_t))
    if return_all:
        order = np.argsort(losses)
        ordered = [models[i] for i in order]
        assert ordered[0] is best_model
        return best_model, ordered[1:]
    return best_model


def run_trials_pool(X, nfactors, ntrials=5, njobs=0, max_threads=None, min_iter=30, max_iter=1000,
                    check_freq=10, epsilon=0.001, better_than_n_ago=5, dtype=np.float64, verbose=True,
                    vcells=None, vX=None, loss_function=None, model_kwargs={}, return_all=False,
                    reproject=False, reproject_kwargs={}, batchsize=0, beta_theta_simultaneous=False,
                    loss_smoothing=1, devices=None):
    """Trials for one or several K, spread over GPUs (scHPF_.py:1151-1332).  Returns the list
    of best models, one per K (plus the rejected ones per K when `return_all`).

    Every trial is what the reference's pool runs (scHPF_.py:1286-1297): a full `fit` with
    `reinit=True`, i.e. random initialisation AND the t == 0 random-phi iteration.  Trials run on
    one worker thread per device, so each gets a random stream of its own: one integer seed per
    trial is drawn from numpy's global stream on the calling thread, in (K, trial) order, before
    any work is dispatched, and the trial's draws (initialisation, t == 0 Dirichlet, minibatch
    shuffle, validation-cell projection) all come from `RandomState(seed)`.  A seeded call is
    therefore reproducible whatever the number of devices, and trial i equals
    `m = scHPF(K); m.set_random_state(RandomState(seed_i)); m.fit(X)`.  An exception in a worker
    is re-raised here after all workers have stopped."""
    _gene_count_warning(X.shape[1])
    ks = [nfactors] if isinstance(nfactors, (int, np.integer)) else list(nfactors)
    if devices is None:
        from . import _lib
        n = _lib.load().schpf_device_count()
        devices = list(range(max(n, 1)))
    jobs = []
    for K in ks:
        for _ in range(ntrials):
            model = scHPF(nfactors=K, min_iter=min_iter, max_iter=max_iter, check_freq=check_freq,
                          epsilon=epsilon, better_than_n_ago=better_than_n_ago, verbose=False,
                          dtype=dtype, **model_kwargs)
            model.set_random_state(np.random.RandomState(int(np.random.randint(0, 2 ** 31 - 1))))
            jobs.append(model)
    errors = []

    def work(dev):
        try:
            for i in range(dev, len(jobs), len(devices)):
                if errors:
                    return
                m = jobs[i]
                m.device = devices[dev]
                fit_loss, _ = _loss_setup(X, m.nfactors, vcells, vX, loss_function, check_freq)
                if hasattr(fit_loss, 'pmodel'):            # validation-cell projections draw from the trial's stream
                    fit_loss.pmodel.set_random_state(m._rng)
                    fit_loss.pmodel.device = m.device
                m.fit(X, loss_function=fit_loss, batchsize=batchsize, loss_smoothing=loss_smoothing,
                      beta_theta_simultaneous=beta_theta_simultaneous)
                if reproject:
                    kw = dict(reproject_kwargs, replace=True, reinit=False)
                    m.loss.append(m.project(X, **kw))
                if hasattr(fit_loss, 'close'):
                    fit_loss.close()
                m.set_random_state(None)                   # the returned model pickles like any other
        except BaseException as exc:                       # noqa: B902 -- re-raised on the calling thread
            errors.append((dev, exc))

    threads = [threading.Thread(target=work, args=(d,)) for d in range(len(devices))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    if errors:
        dev, exc = errors[0]
        raise RuntimeError('run_trials_pool: the worker of device {} failed: {!r}'.format(devices[dev], exc)) from exc

    best, rejected = [], []
    for i, K in enumerate(ks):
        cand = jobs[i * ntrials:(i + 1) * ntrials]
        final = [m.loss[-1][-1] if reproject else m.loss[-1] for m in cand]
        order = np.argsort(final)
        best.append(cand[order[0]])
        rejected.append([cand[j] for j in order[1:]])
    return (best, rejected) if return_all else best
