#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the REAL reference.

Run in the build container only (needs /root/reference; the GPU box does not
have it):

    python tests/golden/make_golden.py

It imports the unmodified reference package from /root/reference, runs its
numba hot path (schpf/hpf_numba.py, schpf/scHPF_.py:_fit, schpf/loss.py) on
seeded inputs and stores inputs + outputs as small .npz files.  The oracle
(oracle/) and the CUDA path (schpf_b200/) are both checked against these.

Fixtures
  kernels_k4.npz    the reference's own test fixture recipe (tests/conftest.py:
                    seed 42, 300 x 1000, 3 % fill, K = 4, fp64) and the output
                    of each of the six kernels on it (tests/test_inference.py).
  cavi_cfg1.npz     BASELINE cfg-1 (1k x 2k, 100 draws/cell, K = 5): state
                    after 1, 10 and 50 iterations of scHPF._fit from a seeded
                    init (reinit=False, min_iter=max_iter) + loss lists.
  cavi_k20.npz      50 iterations at K = 20 (800 x 1200) from a seeded init: the BASELINE parity
                    target (theta / beta within 1e-6 after 50 iterations) at the headline K.
  project_cfg1.npz  scHPF.project of 200 held-out cells onto the 50-iteration
                    model (genes frozen): projected xi/theta + loss.
  reinit_small.npz  fit with reinit=True (the t==0 Dirichlet branch) under a
                    fixed numpy seed: final state + loss.
  simul_small.npz   beta_theta_simultaneous=True variant of the loop.
  minibatch_small.npz  scHPF.fit(batchsize=...) (schpf/scHPF_.py:626-631, 642-650, 686-704):
                    A: reinit=True, 240 cells in windows of 64 (wrapping), 9 iterations;
                    B: reinit=False, windows of 100, beta_theta_simultaneous, loss_smoothing=2;
                    C: reinit=False, windows of 100, the default (cells first) order.
  trials_small.npz  run_trials: three seeded restarts (final losses in return order, the best model)
                    and two restarts scored on projected validation cells.
  fp32_small.npz    the same loop with dtype=np.float32 (mixed precision in the reference).
  fp32_k20.npz      dtype=np.float32 at K = 20 (600 x 900), 20 iterations, next to the reference's
                    float64 run from the same init (the float32 noise of the reference itself).
"""
import os
import sys

import numpy as np

REF = os.environ.get("SCHPF_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import schpf                                    # noqa: E402  (the reference)
from schpf import scHPF, hpf_numba              # noqa: E402
from scipy.sparse import coo_matrix             # noqa: E402

from schpf_b200.synth import synth_coo          # noqa: E402

assert os.path.realpath(schpf.__file__).startswith(os.path.realpath(REF)), schpf.__file__


def state_dict(model, prefix=""):
    return {
        prefix + "theta_shp": model.theta.vi_shape.copy(), prefix + "theta_rte": model.theta.vi_rate.copy(),
        prefix + "beta_shp": model.beta.vi_shape.copy(), prefix + "beta_rte": model.beta.vi_rate.copy(),
        prefix + "xi_shp": model.xi.vi_shape.copy(), prefix + "xi_rte": model.xi.vi_rate.copy(),
        prefix + "eta_shp": model.eta.vi_shape.copy(), prefix + "eta_rte": model.eta.vi_rate.copy(),
    }


def kernels_k4():
    # tests/conftest.py:8-38 recipe
    np.random.seed(42)
    N_CELLS, N_GENES, NZ_FRAC, N_FACTORS = (300, 1000, 0.03, 4)
    NNZ = int(N_CELLS * N_GENES * NZ_FRAC)
    X_data = np.random.negative_binomial(2, 0.5, NNZ)
    X_data[X_data == 0] = 1
    cell_ix = np.random.randint(0, N_CELLS, NNZ, dtype=np.int32)
    gene_ix = np.random.randint(0, N_GENES, NNZ, dtype=np.int32)
    X = coo_matrix((X_data, (cell_ix, gene_ix)), (N_CELLS, N_GENES), dtype=np.int32)
    X.sum_duplicates()
    model = scHPF(N_FACTORS, dtype=np.float64)
    model._initialize(X)
    random_phi = np.random.dirichlet(np.ones(N_FACTORS), X.data.shape[0])
    Xphi_rand = X.data[:, None] * random_phi

    out = dict(row=X.row.astype(np.int32), col=X.col.astype(np.int32),
               data=X.data.astype(np.int32), shape=np.array(X.shape),
               a=model.a, ap=model.ap, bp=model.bp, c=model.c, cp=model.cp, dp=model.dp,
               Xphi_rand=Xphi_rand)
    out.update(state_dict(model))
    th, be, xi, eta = model.theta, model.beta, model.xi, model.eta
    out["Xphi"] = hpf_numba.compute_Xphi_data(X.data, X.row, X.col, th.vi_shape, th.vi_rate,
                                              be.vi_shape, be.vi_rate)
    out["theta_shape_upd"] = hpf_numba.compute_loading_shape_update(Xphi_rand, X.row, N_CELLS, model.a)
    out["beta_shape_upd"] = hpf_numba.compute_loading_shape_update(Xphi_rand, X.col, N_GENES, model.c)
    out["theta_rate_upd"] = hpf_numba.compute_loading_rate_update(xi.vi_shape, xi.vi_rate,
                                                                  be.vi_shape, be.vi_rate)
    out["beta_rate_upd"] = hpf_numba.compute_loading_rate_update(eta.vi_shape, eta.vi_rate,
                                                                 th.vi_shape, th.vi_rate)
    out["eta_rate_upd"] = hpf_numba.compute_capacity_rate_update(be.vi_shape, be.vi_rate, model.dp)
    out["xi_rate_upd"] = hpf_numba.compute_capacity_rate_update(th.vi_shape, th.vi_rate, model.bp)
    out["llh"] = hpf_numba.compute_pois_llh(X.data, X.row, X.col, th.vi_shape, th.vi_rate,
                                            be.vi_shape, be.vi_rate)
    out["mean_neg_llh"] = model.mean_negative_pois_llh(X)
    out["cell_score"] = model.cell_score()
    out["gene_score"] = model.gene_score()
    xs = np.array([1e-4, 1e-3, 1e-2, 0.1, 1, 10, 100, 1000], dtype=np.float64)   # tests/test_inference.py:24
    out["psi_x"] = xs
    out["psi_y"] = np.array([hpf_numba.psi(x) for x in xs])
    out["gammaln_y"] = np.array([hpf_numba.cgammaln(x) for x in xs])
    np.savez_compressed(os.path.join(HERE, "kernels_k4.npz"), **out)
    print("kernels_k4: nnz", X.nnz)


def cavi_cfg1():
    X = synth_coo(1000, 2000, 100, 5, seed=0)
    out = dict(row=X.row.astype(np.int32), col=X.col.astype(np.int32),
               data=X.data.astype(np.int32), shape=np.array(X.shape))
    np.random.seed(1)
    base = scHPF(5, verbose=False)
    base._initialize(X)
    out.update(dict(a=base.a, ap=base.ap, bp=base.bp, c=base.c, cp=base.cp, dp=base.dp))
    out.update(state_dict(base, "init_"))
    final = None
    for n, cf in ((1, 1), (10, 3), (50, 10)):
        from copy import deepcopy
        m = deepcopy(base)
        m.fit(X, reinit=False, min_iter=n, max_iter=n, check_freq=cf, verbose=False)
        out.update(state_dict(m, "it%d_" % n))
        out["it%d_loss" % n] = np.array(m.loss)
        out["it%d_check_freq" % n] = cf
        final = m
    np.savez_compressed(os.path.join(HERE, "cavi_cfg1.npz"), **out)
    print("cavi_cfg1: nnz", X.nnz, "loss50", final.loss)

    # ---- projection of held-out cells onto the 50-iteration model ----------
    Xn = synth_coo(200, 2000, 100, 5, seed=7)
    np.random.seed(3)
    proj = final.project(Xn, min_iter=10, max_iter=10, check_freq=2, verbose=False)
    pout = dict(row=Xn.row.astype(np.int32), col=Xn.col.astype(np.int32),
                data=Xn.data.astype(np.int32), shape=np.array(Xn.shape),
                bp=proj.bp, loss=np.array(proj.loss), seed=3,
                theta_shp=proj.theta.vi_shape, theta_rte=proj.theta.vi_rate,
                xi_shp=proj.xi.vi_shape, xi_rte=proj.xi.vi_rate,
                cell_score=proj.cell_score())
    # the exact init the projection started from (np.random.seed(3) then _setup order)
    np.random.seed(3)
    from schpf import HPF_Gamma
    xi0 = HPF_Gamma.random_gamma_factory((200,), final.ap, final.bp)
    th0 = HPF_Gamma.random_gamma_factory((200, 5), final.a, final.bp)
    pout.update(init_xi_shp=xi0.vi_shape, init_xi_rte=xi0.vi_rate,
                init_theta_shp=th0.vi_shape, init_theta_rte=th0.vi_rate)
    np.savez_compressed(os.path.join(HERE, "project_cfg1.npz"), **pout)
    print("project_cfg1: loss", proj.loss)


def reinit_small():
    X = synth_coo(200, 300, 40, 3, seed=5)
    np.random.seed(11)
    m = scHPF(3, verbose=False)
    m.fit(X, min_iter=6, max_iter=6, check_freq=2, verbose=False)    # reinit=True default
    out = dict(row=X.row.astype(np.int32), col=X.col.astype(np.int32),
               data=X.data.astype(np.int32), shape=np.array(X.shape), seed=11,
               bp=m.bp, dp=m.dp, loss=np.array(m.loss))
    out.update(state_dict(m))
    np.savez_compressed(os.path.join(HERE, "reinit_small.npz"), **out)
    print("reinit_small: loss", m.loss)


def simul_small():
    X = synth_coo(200, 300, 40, 3, seed=5)
    np.random.seed(12)
    base = scHPF(3, verbose=False)
    base._initialize(X)
    out = dict(row=X.row.astype(np.int32), col=X.col.astype(np.int32),
               data=X.data.astype(np.int32), shape=np.array(X.shape),
               a=base.a, ap=base.ap, bp=base.bp, c=base.c, cp=base.cp, dp=base.dp)
    out.update(state_dict(base, "init_"))
    base.fit(X, reinit=False, min_iter=7, max_iter=7, check_freq=2, verbose=False,
             beta_theta_simultaneous=True)
    out.update(state_dict(base, "fin_"))
    out["loss"] = np.array(base.loss)
    np.savez_compressed(os.path.join(HERE, "simul_small.npz"), **out)
    print("simul_small: loss", base.loss)


def minibatch_small():
    X = synth_coo(240, 300, 40, 3, seed=5)
    out = dict(row=X.row.astype(np.int32), col=X.col.astype(np.int32),
               data=X.data.astype(np.int32), shape=np.array(X.shape))
    # A: everything from the numpy stream (init, batch shuffle, t == 0 Dirichlet of the batch)
    np.random.seed(21)
    m = scHPF(3, verbose=False)
    m.fit(X, batchsize=64, min_iter=9, max_iter=9, check_freq=2, verbose=False)
    out.update(A_seed=21, A_batchsize=64, A_iters=9, A_check_freq=2, A_bp=m.bp, A_dp=m.dp,
               A_loss=np.array(m.loss))
    out.update(state_dict(m, "A_"))
    # B: seeded init, then only the shuffle comes from the stream
    np.random.seed(22)
    base = scHPF(3, verbose=False)
    base._initialize(X)
    out.update(dict(a=base.a, ap=base.ap, bp=base.bp, c=base.c, cp=base.cp, dp=base.dp))
    out.update(state_dict(base, "B_init_"))
    np.random.seed(23)
    base.fit(X, reinit=False, batchsize=100, min_iter=8, max_iter=8, check_freq=2, verbose=False,
             beta_theta_simultaneous=True, loss_smoothing=2)
    out.update(B_seed=23, B_batchsize=100, B_iters=8, B_check_freq=2, B_loss=np.array(base.loss))
    out.update(state_dict(base, "B_"))
    # C: the same seeded init, the default (`batched`) order: theta/xi of the batch first, then beta from the
    #    same Xphi with the new theta in its rate (scHPF_.py:686-704); the case the sharded minibatch is held to
    gam = lambda n: schpf.HPF_Gamma(out["B_init_%s_shp" % n].copy(), out["B_init_%s_rte" % n].copy())
    m3 = scHPF(3, verbose=False, bp=float(out["bp"]), dp=float(out["dp"]), xi=gam("xi"), theta=gam("theta"),
               eta=gam("eta"), beta=gam("beta"))
    np.random.seed(24)
    m3.fit(X, reinit=False, batchsize=100, min_iter=8, max_iter=8, check_freq=2, verbose=False)
    out.update(C_seed=24, C_batchsize=100, C_iters=8, C_check_freq=2, C_loss=np.array(m3.loss))
    out.update(state_dict(m3, "C_"))
    np.savez_compressed(os.path.join(HERE, "minibatch_small.npz"), **out)
    print("minibatch_small: loss A", m.loss, "loss B", base.loss)


def cavi_k20():
    """The BASELINE parity target at the headline K: 50 iterations at K = 20 from a seeded init
    (800 x 1200, 100 draws per cell)."""
    X = synth_coo(800, 1200, 100, 20, seed=2)
    np.random.seed(41)
    m = scHPF(20, verbose=False)
    m._initialize(X)
    out = dict(row=X.row.astype(np.int32), col=X.col.astype(np.int32),
               data=X.data.astype(np.int32), shape=np.array(X.shape), seed=41,
               a=m.a, ap=m.ap, bp=m.bp, c=m.c, cp=m.cp, dp=m.dp)
    out.update(state_dict(m, "init_"))
    m.fit(X, reinit=False, min_iter=50, max_iter=50, check_freq=10, verbose=False)
    out.update(state_dict(m, "it50_"))
    out["it50_loss"] = np.array(m.loss)
    out["cell_score"] = m.cell_score()
    out["gene_score"] = m.gene_score()
    np.savez_compressed(os.path.join(HERE, "cavi_k20.npz"), **out)
    print("cavi_k20: nnz", X.nnz, "loss", m.loss)


def trials_small():
    """run_trials (scHPF_.py:968-1148): three seeded restarts, and two restarts scored on
    projected validation cells (loss.py:37-102)."""
    import contextlib
    import io
    from schpf import run_trials
    X = synth_coo(200, 300, 40, 3, seed=5)
    V = synth_coo(40, 300, 40, 3, seed=6)
    out = dict(row=X.row.astype(np.int32), col=X.col.astype(np.int32), data=X.data.astype(np.int32),
               shape=np.array(X.shape), vrow=V.row.astype(np.int32), vcol=V.col.astype(np.int32),
               vdata=V.data.astype(np.int32), vshape=np.array(V.shape))
    np.random.seed(51)
    with contextlib.redirect_stdout(io.StringIO()):
        best, others = run_trials(X, 3, ntrials=3, min_iter=4, max_iter=4, check_freq=2, verbose=False,
                                  return_all=True)
    out.update(A_seed=51, A_final_losses=np.array([best.loss[-1]] + [m.loss[-1] for m in others]),
               A_best_loss=np.array(best.loss), A_best_bp=best.bp, A_best_dp=best.dp)
    out.update(state_dict(best, "A_best_"))
    out["A_cellmean"] = best.cellmean_negative_pois_llh(X)                   # scHPF_.py:395-411
    # the same with duplicate triples: the reference counts DISTINCT genes per cell (its csr sums them)
    dup = np.concatenate([np.arange(X.nnz), np.arange(0, X.nnz, 7)])
    Xd = coo_matrix((X.data[dup], (X.row[dup], X.col[dup])), shape=X.shape)
    out["dup_index"] = dup
    out["A_cellmean_dup"] = best.cellmean_negative_pois_llh(Xd)
    np.random.seed(52)
    with contextlib.redirect_stdout(io.StringIO()):
        vbest = run_trials(X, 3, ntrials=2, min_iter=4, max_iter=4, check_freq=2, verbose=False, vcells=V)
    out.update(B_seed=52, B_best_loss=np.array(vbest.loss))
    out.update(state_dict(vbest, "B_best_"))
    np.savez_compressed(os.path.join(HERE, "trials_small.npz"), **out)
    print("trials_small: A", out["A_final_losses"], "B", vbest.loss)


def fp32_small():
    """dtype=np.float32 (tests/conftest.py:29 parametrises the reference's tests over it): seeded
    fp32 init, 10 iterations.  The reference's result is mixed precision."""
    X = synth_coo(200, 300, 40, 3, seed=5)
    np.random.seed(31)
    m = scHPF(3, verbose=False, dtype=np.float32)
    m._initialize(X)
    out = dict(row=X.row.astype(np.int32), col=X.col.astype(np.int32),
               data=X.data.astype(np.int32), shape=np.array(X.shape), seed=31,
               a=m.a, ap=m.ap, bp=m.bp, c=m.c, cp=m.cp, dp=m.dp)
    out.update(state_dict(m, "init_"))
    m.fit(X, reinit=False, min_iter=10, max_iter=10, check_freq=5, verbose=False)
    out.update(state_dict(m, "fin_"))
    out["loss"] = np.array(m.loss)
    out["dtypes"] = np.array([str(getattr(m, n).vi_shape.dtype) + "/" + str(getattr(m, n).vi_rate.dtype)
                              for n in ("theta", "beta", "xi", "eta")])
    np.savez_compressed(os.path.join(HERE, "fp32_small.npz"), **out)
    print("fp32_small: loss", m.loss, out["dtypes"])


def fp32_k20():
    """dtype=np.float32 at the headline K: 20 iterations of the reference's float32 run from a seeded
    fp32 init (`fin_`), and the reference's float64 run from the SAME (fp32-valued) init (`f64_`): the
    distance between the two is the float32 noise of the reference itself, which bounds what a
    second float32 implementation can be asked to reproduce."""
    X = synth_coo(600, 900, 80, 20, seed=9)
    np.random.seed(77)
    m = scHPF(20, verbose=False, dtype=np.float32)
    m._initialize(X)
    out = dict(row=X.row.astype(np.int32), col=X.col.astype(np.int32),
               data=X.data.astype(np.int32), shape=np.array(X.shape), seed=77,
               a=m.a, ap=m.ap, bp=m.bp, c=m.c, cp=m.cp, dp=m.dp)
    out.update(state_dict(m, "init_"))
    gam = lambda d: schpf.HPF_Gamma(d.vi_shape.astype(np.float64), d.vi_rate.astype(np.float64))
    m64 = scHPF(20, verbose=False, bp=float(m.bp), dp=float(m.dp), xi=gam(m.xi), theta=gam(m.theta),
                eta=gam(m.eta), beta=gam(m.beta))
    m.fit(X, reinit=False, min_iter=20, max_iter=20, check_freq=5, verbose=False)
    out.update(state_dict(m, "fin_"))
    out["loss"] = np.array(m.loss)
    m64.fit(X, reinit=False, min_iter=20, max_iter=20, check_freq=5, verbose=False)
    out.update(state_dict(m64, "f64_"))
    out["f64_loss"] = np.array(m64.loss)
    rel = lambda a, b: float(np.max(np.abs(a.astype(np.float64) - b) / np.abs(b)))
    noise = {n: rel(out["fin_" + n], out["f64_" + n]) for n in ("theta_shp", "theta_rte", "beta_shp", "beta_rte", "xi_rte", "eta_rte")}
    out["f32_vs_f64_max_rel"] = np.array([noise[k] for k in sorted(noise)])
    np.savez_compressed(os.path.join(HERE, "fp32_k20.npz"), **out)
    print("fp32_k20: loss", m.loss, "f64 loss", m64.loss, "\n  reference float32 vs float64, max rel:", noise)


if __name__ == "__main__":
    only = sys.argv[1:]
    for fn in (kernels_k4, cavi_cfg1, cavi_k20, reinit_small, simul_small, minibatch_small, trials_small, fp32_small,
               fp32_k20):
        if not only or fn.__name__ in only:
            fn()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
