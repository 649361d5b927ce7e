"""CPU-only tests: the C ABI exports what include/schpf_b200.h declares, the
product path refuses to run without a GPU, and the host logic of the estimator
(setup, RNG order, loop schedule, convergence rules, error behaviour) reproduces
golden runs of the reference when its engine is swapped for the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_equal
from scipy.sparse import coo_matrix

import schpf_b200
from schpf_b200 import _lib, scHPF, HPF_Gamma, combine_across_cells
from schpf_b200 import scHPF_ as shell
from schpf_b200.engine import shard_bounds_by_nnz
from conftest import ROOT, has_cuda, max_rel
from oracle_engine import OracleEngine

STATE = ("theta_shp", "theta_rte", "beta_shp", "beta_rte", "xi_shp", "xi_rte", "eta_shp", "eta_rte")


def _X(g):
    return coo_matrix((g["data"], (g["row"], g["col"])), shape=tuple(int(v) for v in g["shape"]))


def _gam(g, name, prefix=""):
    return HPF_Gamma(g[prefix + name + "_shp"].copy(), g[prefix + name + "_rte"].copy())


@pytest.fixture()
def oracle_backend(monkeypatch):
    """Host logic under test, arithmetic from the oracle (tests only)."""
    from oracle import hpf_numpy
    import schpf_b200.loss
    monkeypatch.setattr(shell, "_engine_factory", OracleEngine)
    monkeypatch.setattr(schpf_b200.loss, "compute_pois_llh", hpf_numpy.compute_pois_llh)


# ---------------------------------------------------------------- C ABI ------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "schpf_b200.h")).read()
    declared = set(re.findall(r"\b(schpf_[a-zA-Z_]+)\s*\(", header))
    declared.discard("schpf_engine")
    assert len(declared) >= 28
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export: " + name
    assert set(_lib.SIGNATURES) == declared
    assert _lib.load().schpf_version() >= 100


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_gpu():
    from schpf_b200 import hpf_cuda
    with pytest.raises(_lib.SchpfError):
        hpf_cuda.psi(np.array([1.0, 2.0]))
    X = coo_matrix((np.array([1, 2], dtype=np.int32), (np.array([0, 1]), np.array([1, 0]))), shape=(2, 2))
    with pytest.raises(_lib.SchpfError):
        scHPF(2, verbose=False).fit(X, max_iter=1)


# ------------------------------------------------------- estimator shell -----
def test_setup_matches_reference_rng_and_hypers(g_cavi):
    g, X = g_cavi, _X(g_cavi)
    np.random.seed(1)
    m = scHPF(5, verbose=False)
    m._initialize(X)
    assert m.bp == float(g["bp"]) and m.dp == float(g["dp"])       # reference asserts equality too
    for name in ("theta", "beta", "xi", "eta"):
        assert_equal(getattr(m, name).vi_shape, g["init_%s_shp" % name])
        assert_equal(getattr(m, name).vi_rate, g["init_%s_rte" % name])
    assert m.ncells == 1000 and m.ngenes == 2000


def test_hyperparameter_rules():
    m = scHPF(9, a=-2, c=-2)
    assert m.a == 1 / 3 and m.c == 1 / 3                            # -2 -> 1/sqrt(K)
    with pytest.raises(ValueError):
        scHPF(None, a=-2)
    with pytest.raises(ValueError):
        scHPF(3).project(None, replace=True, recalc_bp=True)
    X = coo_matrix((np.ones(3, dtype=np.int32), (np.arange(3), np.arange(3))), shape=(3, 3))
    with pytest.raises(ValueError):
        scHPF(2, bp=1.0)._setup(X, freeze_genes=True)                # frozen genes without dp
    with pytest.raises(ValueError):
        scHPF(2, bp=1.0, dp=1.0)._setup(X, freeze_genes=True)        # ... without eta / beta


@pytest.mark.parametrize("n", [1, 10, 50])
def test_fit_loop_reproduces_reference(oracle_backend, g_cavi, n):
    g, X = g_cavi, _X(g_cavi)
    m = scHPF(5, verbose=False, bp=float(g["bp"]), dp=float(g["dp"]),
              xi=_gam(g, "xi", "init_"), theta=_gam(g, "theta", "init_"),
              eta=_gam(g, "eta", "init_"), beta=_gam(g, "beta", "init_"))
    m.fit(X, reinit=False, min_iter=n, max_iter=n, check_freq=int(g["it%d_check_freq" % n]))
    for name in ("theta", "beta", "xi", "eta"):
        assert max_rel(getattr(m, name).vi_shape, g["it%d_%s_shp" % (n, name)]) < 1e-10
        assert max_rel(getattr(m, name).vi_rate, g["it%d_%s_rte" % (n, name)]) < 1e-10
    assert_allclose(m.loss, g["it%d_loss" % n], rtol=1e-12)


def test_fifty_iterations_at_k20_reproduce_reference(oracle_backend, g_k20):
    """BASELINE.json's parity target (theta / beta within 1e-6 of the reference after 50 iterations)
    at the headline K, for the host loop over the oracle: seeded init drawn here, then 50 iterations."""
    g, X = g_k20, _X(g_k20)
    np.random.seed(int(g["seed"]))
    m = scHPF(20, verbose=False)
    m._initialize(X)
    assert m.bp == float(g["bp"]) and m.dp == float(g["dp"])
    assert_equal(m.theta.vi_shape, g["init_theta_shp"])
    m.fit(X, reinit=False, min_iter=50, max_iter=50, check_freq=10)
    for name in ("theta", "beta", "xi", "eta"):
        assert max_rel(getattr(m, name).vi_shape, g["it50_%s_shp" % name]) < 1e-9
        assert max_rel(getattr(m, name).vi_rate, g["it50_%s_rte" % name]) < 1e-9
    assert_allclose(m.loss, g["it50_loss"], rtol=1e-11)
    assert_allclose(m.cell_score(), g["cell_score"], rtol=1e-9)
    assert_allclose(m.gene_score(), g["gene_score"], rtol=1e-9)


def test_fit_with_reinit_reproduces_seeded_reference(oracle_backend, g_reinit):
    g, X = g_reinit, _X(g_reinit)
    np.random.seed(int(g["seed"]))
    m = scHPF(3, verbose=False).fit(X, min_iter=6, max_iter=6, check_freq=2)
    assert m.bp == float(g["bp"]) and m.dp == float(g["dp"])
    for name in ("theta", "beta", "xi", "eta"):
        assert max_rel(getattr(m, name).vi_shape, g[name + "_shp"]) < 1e-11
        assert max_rel(getattr(m, name).vi_rate, g[name + "_rte"]) < 1e-11
    assert_allclose(m.loss, g["loss"], rtol=1e-12)


def test_simultaneous_updates(oracle_backend, g_simul):
    g, X = g_simul, _X(g_simul)
    m = scHPF(3, verbose=False, bp=float(g["bp"]), dp=float(g["dp"]),
              xi=_gam(g, "xi", "init_"), theta=_gam(g, "theta", "init_"),
              eta=_gam(g, "eta", "init_"), beta=_gam(g, "beta", "init_"))
    m.fit(X, reinit=False, min_iter=7, max_iter=7, check_freq=2, beta_theta_simultaneous=True)
    for name in ("theta", "beta", "xi", "eta"):
        assert max_rel(getattr(m, name).vi_shape, g["fin_%s_shp" % name]) < 1e-11
    assert_allclose(m.loss, g["loss"], rtol=1e-12)


def test_project_reproduces_reference(oracle_backend, g_cavi, g_project):
    g, p = g_cavi, g_project
    trained = scHPF(5, verbose=False, bp=float(g["bp"]), dp=float(g["dp"]),
                    xi=_gam(g, "xi", "it50_"), theta=_gam(g, "theta", "it50_"),
                    eta=_gam(g, "eta", "it50_"), beta=_gam(g, "beta", "it50_"))
    Xn = _X(p)
    np.random.seed(int(p["seed"]))
    proj = trained.project(Xn, min_iter=10, max_iter=10, check_freq=2)
    assert proj is not trained and proj.ncells == 200
    assert proj.eta == trained.eta and proj.beta == trained.beta     # bit-exact, as the reference tests
    assert proj.bp == trained.bp == float(p["bp"])
    assert max_rel(proj.theta.vi_shape, p["theta_shp"]) < 1e-11
    assert max_rel(proj.theta.vi_rate, p["theta_rte"]) < 1e-11
    assert max_rel(proj.xi.vi_rate, p["xi_rte"]) < 1e-11
    assert_allclose(proj.loss, p["loss"], rtol=1e-12)
    assert_allclose(proj.cell_score(), p["cell_score"], rtol=1e-10)
    # transform == cell scores of a fresh projection
    np.random.seed(int(p["seed"]))
    assert_allclose(trained.transform(Xn, min_iter=10, max_iter=10, check_freq=2), p["cell_score"], rtol=1e-10)
    # replace=True writes xi/theta into the model and returns the loss list
    np.random.seed(int(p["seed"]))
    loss = trained.project(Xn, replace=True, min_iter=10, max_iter=10, check_freq=2)
    assert_allclose(loss, p["loss"], rtol=1e-12)
    assert trained.ncells == 200
    # recalc_bp recomputes b' from the new data
    recalc = trained.project(Xn, recalc_bp=True, min_iter=2, max_iter=2)
    rs = np.asarray(Xn.sum(axis=1))
    assert recalc.bp == trained.ap * np.mean(rs) / np.var(rs)


def test_scores_against_reference(g_kernels):
    g = g_kernels
    m = scHPF(4, xi=_gam(g, "xi"), theta=_gam(g, "theta"), eta=_gam(g, "eta"), beta=_gam(g, "beta"))
    assert_equal(m.cell_score(), g["cell_score"])
    assert_equal(m.gene_score(), g["gene_score"])


def test_convergence_rules(oracle_backend, g_cavi):
    """Scripted losses drive the stopping rules of scHPF_.py:750-774.  The expected
    lengths were produced by the reference itself on the same scripts."""
    g, X = g_cavi, _X(g_cavi)

    def run(seq, **kw):
        it = iter(seq)
        m = scHPF(5, verbose=False, epsilon=0.001, better_than_n_ago=5)
        m.fit(X, loss_function=lambda **k: next(it), min_iter=0, max_iter=200, check_freq=1, **kw)
        return m.loss
    # converges at the 2nd consecutive small change once more than 3 checks exist
    flat = [10, 9, 8, 7.99999, 7.999985, 7.99998, 1, 1, 1]
    assert len(run(flat)) == 5
    # an inflection (prev is a local max) blocks convergence for that check
    infl = [10.0, 8.0, 8.00001, 8.000005, 8.0000049, 3, 3, 3, 3]
    assert len(run(infl)) == 5
    # getting worse than better_than_n_ago checks ago, and rising -> stop
    worse = [5, 4, 3, 2, 1, 1.5, 2.5, 9, 9, 9]
    assert len(run(worse)) == 8
    # loss_smoothing averages the last n raw values
    m_loss = run([4.0, 2.0, 6.0, 8.0] + [8.0] * 200, loss_smoothing=2)
    assert m_loss[:4] == [4.0, 3.0, 4.0, 7.0]
    # check schedule: losses at t = 0, cf, 2cf, ... ; max_iter=10, check_freq=10 -> one entry
    m = scHPF(5, verbose=False).fit(X, max_iter=10, min_iter=10, check_freq=10)
    assert len(m.loss) == 1
    m = scHPF(5, verbose=False).fit(X, max_iter=11, min_iter=11, check_freq=5)
    assert len(m.loss) == 3


def test_checkstep_function_and_custom_loss_get_host_state(oracle_backend, g_cavi):
    X = _X(g_cavi)
    seen = []

    def checkstep(bp, dp, xi, eta, theta, beta, t):
        assert isinstance(theta, HPF_Gamma) and theta.dims == (1000, 5) and beta.dims == (2000, 5)
        seen.append(t)
    np.random.seed(0)
    m = scHPF(5, verbose=False).fit(X, max_iter=7, min_iter=7, check_freq=3, checkstep_function=checkstep)
    assert seen == [0, 3, 6] and len(m.loss) == 3


# ------------------------------------------------------------- minibatches ---
def test_minibatch_windows_follow_the_reference_schedule():
    """util.py:218-231: one shuffle, then consecutive wrapping windows."""
    from schpf_b200.cavi_loop import minibatch_windows
    np.random.seed(4)
    order = np.arange(10)
    np.random.shuffle(order)
    from schpf_b200.cavi_loop import MinibatchSchedule
    sched = MinibatchSchedule(10, 4)
    assert sched.pieces(0) == [(0, 4)] and sched.pieces(8) == [(8, 2), (0, 2)] and sched.pieces(6) == [(6, 4)]
    np.random.seed(4)
    gen = minibatch_windows(10, 4)
    got = [next(gen) for _ in range(6)]
    assert [s for s, _ in got] == [0, 4, 8, 2, 6, 0]
    assert_equal(got[0][1], order[0:4])
    assert_equal(got[2][1], np.concatenate([order[8:], order[:2]]))       # wraps
    assert_equal(got[5][1], order[0:4])                                  # the cycle repeats
    np.random.seed(4)
    gen = minibatch_windows(10, 10)                                      # equality is allowed (:219)
    assert_equal(next(gen)[1], order)
    assert_equal(next(gen)[1], order)


@pytest.mark.parametrize("device_state", [False, True])
@pytest.mark.parametrize("cache", [64, 0])
def test_minibatch_fit_reproduces_seeded_reference(oracle_backend, g_minibatch, monkeypatch, cache, device_state):
    """Case A of minibatch_small.npz: init, batch shuffle and the batch's t == 0 Dirichlet all
    come from numpy's stream in the reference's order; with and without per-window engines,
    with theta / xi kept on the host or (permuted once) in a master engine."""
    from schpf_b200 import cavi_loop
    monkeypatch.setattr(cavi_loop, "MINIBATCH_ENGINE_CACHE", cache)
    monkeypatch.setattr(cavi_loop, "MINIBATCH_DEVICE_STATE", device_state)
    g, X = g_minibatch, _X(g_minibatch)
    np.random.seed(int(g["A_seed"]))
    m = scHPF(3, verbose=False).fit(X, batchsize=int(g["A_batchsize"]), min_iter=int(g["A_iters"]),
                                    max_iter=int(g["A_iters"]), check_freq=int(g["A_check_freq"]))
    assert m.bp == float(g["A_bp"]) and m.dp == float(g["A_dp"])
    for name in ("theta", "beta", "xi", "eta"):
        assert max_rel(getattr(m, name).vi_shape, g["A_%s_shp" % name]) < 1e-11
        assert max_rel(getattr(m, name).vi_rate, g["A_%s_rte" % name]) < 1e-11
    assert_allclose(m.loss, g["A_loss"], rtol=1e-12)


@pytest.mark.parametrize("device_state", [False, True])
def test_minibatch_simultaneous_and_smoothing(oracle_backend, g_minibatch, monkeypatch, device_state):
    """Case B: reinit=False, beta_theta_simultaneous, loss_smoothing=2."""
    from schpf_b200 import cavi_loop
    monkeypatch.setattr(cavi_loop, "MINIBATCH_DEVICE_STATE", device_state)
    g, X = g_minibatch, _X(g_minibatch)
    m = scHPF(3, verbose=False, bp=float(g["bp"]), dp=float(g["dp"]),
              xi=_gam(g, "xi", "B_init_"), theta=_gam(g, "theta", "B_init_"),
              eta=_gam(g, "eta", "B_init_"), beta=_gam(g, "beta", "B_init_"))
    np.random.seed(int(g["B_seed"]))
    m.fit(X, reinit=False, batchsize=int(g["B_batchsize"]), min_iter=int(g["B_iters"]),
          max_iter=int(g["B_iters"]), check_freq=int(g["B_check_freq"]),
          beta_theta_simultaneous=True, loss_smoothing=2)
    for name in ("theta", "beta", "xi", "eta"):
        assert max_rel(getattr(m, name).vi_shape, g["B_%s_shp" % name]) < 1e-11
        assert max_rel(getattr(m, name).vi_rate, g["B_%s_rte" % name]) < 1e-11
    assert_allclose(m.loss, g["B_loss"], rtol=1e-12)


@pytest.mark.parametrize("device_state", [False, True])
def test_minibatch_default_order_from_a_seeded_init(oracle_backend, g_minibatch, monkeypatch, device_state):
    """Case C: reinit=False, the default `batched` order (cells first, scHPF_.py:686-704)."""
    from schpf_b200 import cavi_loop
    monkeypatch.setattr(cavi_loop, "MINIBATCH_DEVICE_STATE", device_state)
    g, X = g_minibatch, _X(g_minibatch)
    m = scHPF(3, verbose=False, bp=float(g["bp"]), dp=float(g["dp"]),
              xi=_gam(g, "xi", "B_init_"), theta=_gam(g, "theta", "B_init_"),
              eta=_gam(g, "eta", "B_init_"), beta=_gam(g, "beta", "B_init_"))
    np.random.seed(int(g["C_seed"]))
    m.fit(X, reinit=False, batchsize=int(g["C_batchsize"]), min_iter=int(g["C_iters"]),
          max_iter=int(g["C_iters"]), check_freq=int(g["C_check_freq"]))
    for name in ("theta", "beta", "xi", "eta"):
        assert max_rel(getattr(m, name).vi_shape, g["C_%s_shp" % name]) < 1e-11
        assert max_rel(getattr(m, name).vi_rate, g["C_%s_rte" % name]) < 1e-11
    assert_allclose(m.loss, g["C_loss"], rtol=1e-12)


@pytest.mark.parametrize("device_state", [False, True])
def test_minibatch_edge_cases(oracle_backend, g_minibatch, monkeypatch, device_state):
    from schpf_b200 import cavi_loop
    monkeypatch.setattr(cavi_loop, "MINIBATCH_DEVICE_STATE", device_state)
    g, X = g_minibatch, _X(g_minibatch)
    # batchsize 0 / 1 / None / > ncells mean "all cells" (scHPF_.py:627)
    np.random.seed(3)
    full = scHPF(3, verbose=False).fit(X, min_iter=3, max_iter=3, check_freq=1)
    for bs in (0, 1, None, X.shape[0] + 1):
        np.random.seed(3)
        m = scHPF(3, verbose=False).fit(X, batchsize=bs, min_iter=3, max_iter=3, check_freq=1)
        assert_equal(m.theta.vi_shape, full.theta.vi_shape)
    # custom loss / checkstep functions see full-size host arrays; frozen genes stay untouched
    seen = []
    np.random.seed(3)
    m = scHPF(3, verbose=False).fit(X, batchsize=50, min_iter=4, max_iter=4, check_freq=2,
                                    checkstep_function=lambda **k: seen.append((k["t"], k["theta"].dims, k["beta"].dims)))
    assert seen == [(0, (240, 3), (300, 3)), (2, (240, 3), (300, 3))]
    np.random.seed(3)
    p = m.project(X, batchsize=60, min_iter=3, max_iter=3, check_freq=1)
    assert p.beta == m.beta and p.eta == m.eta and len(p.loss) == 3


def test_float32_models_keep_their_dtype_and_track_the_reference(oracle_backend, g_fp32):
    """dtype=np.float32: the reference's own result is mixed precision (beta.vi_shape and
    eta.vi_rate come back fp64); here the arrays are fp32 in and out and the
    arithmetic is fp64, so the two agree to fp32 rounding (the reference's tests use
    rtol 1e-5 .. 1e-6 for fp32, tests/test_inference.py:46-121)."""
    g = g_fp32
    X = coo_matrix((g["data"], (g["row"], g["col"])), shape=tuple(int(v) for v in g["shape"]))
    np.random.seed(int(g["seed"]))
    m = scHPF(3, verbose=False, dtype=np.float32)
    m._initialize(X)
    for n in ("theta", "beta", "xi", "eta"):
        assert getattr(m, n).vi_shape.dtype == np.float32
        assert np.array_equal(getattr(m, n).vi_shape, g["init_%s_shp" % n])     # same fp32 draws
    m.fit(X, reinit=False, min_iter=10, max_iter=10, check_freq=5)
    for n in ("theta", "beta", "xi", "eta"):
        d = getattr(m, n)
        assert d.vi_shape.dtype == np.float32 and d.vi_rate.dtype == np.float32
        assert_allclose(d.vi_shape, g["fin_%s_shp" % n], rtol=2e-5)
        assert_allclose(d.vi_rate, g["fin_%s_rte" % n], rtol=2e-5)
    assert_allclose(m.loss, g["loss"], rtol=2e-5)


def test_combine_across_cells(g_kernels):
    g = g_kernels
    x = scHPF(4, bp=1.0, dp=2.0, xi=_gam(g, "xi"), theta=_gam(g, "theta"), eta=_gam(g, "eta"), beta=_gam(g, "beta"))
    y = scHPF(4, bp=3.0, dp=2.0, xi=HPF_Gamma(g["xi_shp"][:5] + 1, g["xi_rte"][:5] + 1),
              theta=HPF_Gamma(g["theta_shp"][:5] + 1, g["theta_rte"][:5] + 1), eta=x.eta, beta=x.beta)
    ixs = np.array([0, 2, 4, 6, 304])
    xy = combine_across_cells(x, y, ixs)
    assert xy.bp is None and xy.ncells == 305
    assert_equal(xy.theta.vi_shape[ixs], y.theta.vi_shape)
    assert_equal(xy.theta.vi_shape[np.setdiff1d(np.arange(305), ixs)], x.theta.vi_shape)


def test_model_roundtrips_through_joblib(tmp_path, g_kernels):
    g = g_kernels
    m = scHPF(4, bp=1.0, dp=2.0, xi=_gam(g, "xi"), theta=_gam(g, "theta"), eta=_gam(g, "eta"), beta=_gam(g, "beta"))
    f = str(tmp_path / "m.joblib")
    schpf_b200.save_model(m, f)
    m2 = schpf_b200.load_model(f)
    assert m2.theta == m.theta and m2.beta == m.beta and m2.a == m.a


# --------------------------------------------------------------- sharding ----
def test_shard_bounds_balance_nnz():
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 500, size=10007)
    for ws in (1, 2, 3, 8):
        b = shard_bounds_by_nnz(counts, ws)
        assert b[0] == 0 and b[-1] == counts.shape[0] and np.all(np.diff(b) >= 0) and len(b) == ws + 1
        per = np.array([counts[b[r]:b[r + 1]].sum() for r in range(ws)])
        assert per.sum() == counts.sum()
        assert per.max() - per.min() <= 2 * counts.max()
    # degenerate: all nonzeros in one cell, empty matrix
    assert list(shard_bounds_by_nnz([0, 0, 9, 0], 2)) in ([0, 2, 4], [0, 3, 4])
    assert list(shard_bounds_by_nnz([0, 0, 0], 2))[0] == 0


# ------------------------------------------------------------ run_trials -----
def test_run_trials_picks_lowest_loss_like_the_reference(oracle_backend, g_reinit, capsys):
    from schpf_b200 import run_trials
    X = _X(g_reinit)
    np.random.seed(21)
    best, others = run_trials(X, 3, ntrials=3, min_iter=4, max_iter=4, check_freq=2, verbose=False,
                              return_all=True)
    losses = [best.loss[-1]] + [m.loss[-1] for m in others]
    assert losses == sorted(losses) and len(others) == 2
    # the first trial consumes the RNG exactly like a plain seeded fit
    np.random.seed(21)
    first = scHPF(3, verbose=False, min_iter=4, max_iter=4, check_freq=2).fit(X)
    assert any(np.array_equal(first.theta.vi_shape, m.theta.vi_shape) for m in [best] + others)
    # validation cells: every check projects them (loss.py:37-102) and the training loss is printed
    np.random.seed(22)
    vcells = coo_matrix((g_reinit["data"][:400], (g_reinit["row"][:400] % 20, g_reinit["col"][:400])),
                        shape=(20, X.shape[1]))
    vcells.sum_duplicates()
    m = run_trials(X, 3, ntrials=1, min_iter=3, max_iter=3, check_freq=1, verbose=False, vcells=vcells,
                   reproject=True)
    assert "train:" in capsys.readouterr().out
    assert isinstance(m.loss[-1], list) and len(m.loss) == 4          # 3 checks + the reprojection's list


def test_run_trials_reproduces_the_reference(oracle_backend, g_trials, capsys):
    """run_trials under a fixed numpy seed against the reference's own run_trials: the final losses
    of all restarts in return order, the selected model, and the variant scored on projected
    validation cells (scHPF_.py:968-1148, loss.py:37-102)."""
    from schpf_b200 import run_trials
    g = g_trials
    X = coo_matrix((g["data"], (g["row"], g["col"])), shape=tuple(int(v) for v in g["shape"]))
    V = coo_matrix((g["vdata"], (g["vrow"], g["vcol"])), shape=tuple(int(v) for v in g["vshape"]))
    np.random.seed(int(g["A_seed"]))
    best, others = run_trials(X, 3, ntrials=3, min_iter=4, max_iter=4, check_freq=2, verbose=False, return_all=True)
    assert_allclose([best.loss[-1]] + [m.loss[-1] for m in others], g["A_final_losses"], rtol=1e-11)
    assert_allclose(best.loss, g["A_best_loss"], rtol=1e-11)
    assert best.bp == float(g["A_best_bp"]) and best.dp == float(g["A_best_dp"])
    for n in ("theta", "beta", "xi", "eta"):
        assert max_rel(getattr(best, n).vi_shape, g["A_best_%s_shp" % n]) < 1e-10
        assert max_rel(getattr(best, n).vi_rate, g["A_best_%s_rte" % n]) < 1e-10
    # per-cell mean of the pointwise llh (scHPF_.py:395-411), also with duplicate triples
    assert_allclose(best.cellmean_negative_pois_llh(X), g["A_cellmean"], rtol=1e-11)
    dup = g["dup_index"]
    Xd = coo_matrix((X.data[dup], (X.row[dup], X.col[dup])), shape=X.shape)
    assert_allclose(best.cellmean_negative_pois_llh(Xd), g["A_cellmean_dup"], rtol=1e-11)
    np.random.seed(int(g["B_seed"]))
    vbest = run_trials(X, 3, ntrials=2, min_iter=4, max_iter=4, check_freq=2, verbose=False, vcells=V)
    assert_allclose(vbest.loss, g["B_best_loss"], rtol=1e-11)
    for n in ("theta", "beta"):
        assert max_rel(getattr(vbest, n).vi_shape, g["B_best_%s_shp" % n]) < 1e-10
    assert "train:" in capsys.readouterr().out


def test_run_trials_pool_over_devices(oracle_backend, g_reinit):
    from schpf_b200 import run_trials_pool
    X = _X(g_reinit)
    np.random.seed(5)
    best, rejected = run_trials_pool(X, [2, 3], ntrials=2, min_iter=3, max_iter=3, check_freq=1,
                                     return_all=True, devices=[0, 0])
    assert [m.nfactors for m in best] == [2, 3] and [len(r) for r in rejected] == [1, 1]
    assert all(b.loss[-1] <= r[0].loss[-1] for b, r in zip(best, rejected))
    np.random.seed(5)
    again = run_trials_pool(X, [2, 3], ntrials=2, min_iter=3, max_iter=3, check_freq=1, devices=[0])
    assert all(np.array_equal(a.beta.vi_shape, b.beta.vi_shape) for a, b in zip(again, best))


def test_run_trials_pool_trials_are_full_fits_with_their_own_stream(oracle_backend, g_reinit):
    """a pool trial is what the reference's pool runs -- a fit with reinit=True, i.e.
    random initialisation AND the t == 0 random-phi iteration -- from a RandomState of its own,
    seeded from numpy's global stream on the calling thread in (K, trial) order."""
    from schpf_b200 import run_trials_pool, scHPF
    X = _X(g_reinit)
    np.random.seed(11)
    best, rejected = run_trials_pool(X, 3, ntrials=3, min_iter=4, max_iter=4, check_freq=2, return_all=True,
                                     devices=[0, 0])
    models = [best[0]] + rejected[0]
    np.random.seed(11)
    seeds = [int(np.random.randint(0, 2 ** 31 - 1)) for _ in range(3)]
    direct = []
    for s in seeds:
        m = scHPF(3, min_iter=4, max_iter=4, check_freq=2, verbose=False)
        m.set_random_state(np.random.RandomState(s))
        m.fit(X)                                     # reinit=True by default
        direct.append(m)
    by_loss = sorted(direct, key=lambda m: m.loss[-1])
    for a, b in zip(models, by_loss):
        assert np.array_equal(a.beta.vi_shape, b.beta.vi_shape) and a.loss == b.loss
    # and it is NOT the reinit=False trajectory from the same initial draws (the old pool)
    m0 = scHPF(3, min_iter=4, max_iter=4, check_freq=2, verbose=False)
    m0.set_random_state(np.random.RandomState(seeds[0]))
    m0._initialize(X)
    m0.fit(X, reinit=False)
    assert not np.allclose(m0.beta.vi_shape, direct[0].beta.vi_shape)


def test_run_trials_pool_reraises_worker_errors(oracle_backend, g_reinit, monkeypatch):
    from schpf_b200 import run_trials_pool
    from schpf_b200 import scHPF_ as shell
    X = _X(g_reinit)

    class Boom(shell._engine_factory):
        def set_coo(self, *a):
            raise MemoryError("device out of memory (simulated)")
    monkeypatch.setattr(shell, "_engine_factory", Boom)
    with pytest.raises(RuntimeError, match="worker of device 0 failed.*out of memory"):
        run_trials_pool(X, 3, ntrials=2, min_iter=2, max_iter=2, check_freq=1, devices=[0])


def test_more_factors_than_the_kernels_support_is_a_python_error():
    from schpf_b200 import scHPF
    X = coo_matrix((np.ones(4, dtype=np.int32), (np.arange(4), np.arange(4))), shape=(4, 4))
    with pytest.raises(ValueError, match="above the 64 factors"):
        scHPF(65, verbose=False).fit(X, max_iter=1, min_iter=1)


def test_fit_on_a_list_of_devices_shards_the_cells_in_one_process(oracle_backend, g_cavi):
    """scHPF(K, device=[0, 1, 2]).fit(X): one process, cells cut by nnz over the devices, the
    exchange buffers summed every iteration (here through the oracle-backed engines): the
    reference's unsharded golden run, incl. the t == 0 branch being well defined."""
    g, X = g_cavi, _X(g_cavi)
    gam = lambda n: HPF_Gamma(g["init_" + n + "_shp"].copy(), g["init_" + n + "_rte"].copy())
    m = scHPF(5, verbose=False, bp=float(g["bp"]), dp=float(g["dp"]), device=[0, 1, 2],
              xi=gam("xi"), theta=gam("theta"), eta=gam("eta"), beta=gam("beta"))
    m.fit(X, reinit=False, min_iter=10, max_iter=10, check_freq=3)
    for n in ("theta", "beta", "xi", "eta"):
        assert max_rel(getattr(m, n).vi_shape, g["it10_" + n + "_shp"]) < 1e-11, n
        assert max_rel(getattr(m, n).vi_rate, g["it10_" + n + "_rte"]) < 1e-11, n
    assert_allclose(m.loss, g["it10_loss"], rtol=1e-12)
    # projection (frozen genes: no exchange) and a reinit=True fit run too
    proj = m.project(X, min_iter=2, max_iter=2, check_freq=1, verbose=False)
    assert proj.theta.vi_shape.shape == m.theta.vi_shape.shape and np.isfinite(proj.loss[-1])
    np.random.seed(0)
    m2 = scHPF(5, verbose=False, device=[0, 1]).fit(X, min_iter=2, max_iter=2, check_freq=1)
    assert abs((m2.beta.vi_shape - m2.c).sum() / X.data.sum() - 1) < 1e-9
    with pytest.raises(NotImplementedError):
        scHPF(5, verbose=False, device=[0, 1]).fit(X, batchsize=100, max_iter=1, min_iter=1)


# ------------------------------------------------------------------ ingest ---
def test_mtx_header_parsing_matches_scipy(tmp_path):
    """host half of schpf_b200.io.load_mtx (the data lines are parsed on the device): size line and data
    offset agree with what scipy.io.mminfo / mmread see for files written by the reference's writer"""
    from scipy.io import mminfo, mmwrite
    from schpf_b200.io import parse_mtx_header
    X = coo_matrix((np.array([3, 1, 2], dtype=np.int32), (np.array([0, 2, 2]), np.array([1, 0, 3]))), shape=(3, 5))
    for field, comment in (("integer", ""), ("real", "two\nlines"), ("pattern", "x")):
        path = str(tmp_path / ("h_%s.mtx" % field))
        mmwrite(path, X.astype(np.float64) if field == "real" else X, field=field, comment=comment)
        raw = open(path, "rb").read()
        nrows, ncols, nnz, fld, begin = parse_mtx_header(raw)
        assert (nrows, ncols, nnz, fld) == (mminfo(path)[0], mminfo(path)[1], mminfo(path)[2], field)
        data = [l for l in raw[begin:].decode().splitlines() if l.strip()]
        assert len(data) == nnz and data[0].split()[:2] == ["1", "2"]
    ok = b"%%MatrixMarket matrix coordinate integer general\n% c\n\n2 2 1\n1 1 5"
    assert parse_mtx_header(ok) == (2, 2, 1, "integer", ok.index(b"1 1 5"))
    for bad, msg in ((b"1 1 1\n", "banner"), (b"%%MatrixMarket matrix array real general\n1 1\n", "coordinate"),
                     (b"%%MatrixMarket matrix coordinate complex general\n1 1 1\n", "field"),
                     (b"%%MatrixMarket matrix coordinate integer symmetric\n1 1 1\n", "symmetry"),
                     (b"%%MatrixMarket matrix coordinate integer general\n% only comments\n", "size line"),
                     (b"%%MatrixMarket matrix coordinate integer general\n1 1\n", "size line"),
                     (b"%%MatrixMarket matrix coordinate integer general\n1 x 1\n", "size line")):
        with pytest.raises(ValueError, match=msg):
            parse_mtx_header(bad)
