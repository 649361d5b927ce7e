"""GPU parity of the fp32 sweep (engine option precision=32, schpf_b200/csrc/sweep_f32.cu), which
float32 models use (reference: scHPF(dtype=np.float32), scHPF_.py:225-246; its numba kernels
then run in float32, hpf_numba.py:25-51,55-114,129-156).

Tolerances.  Two float32 implementations that sum in different orders cannot agree better than
each of them agrees with the exact result.  tests/golden/fp32_k20.npz therefore holds the
reference's float32 run AND its float64 run from the same initial state: the reference's own
float32 noise after 20 iterations at K=20 is max 2.5e-5 (median 3e-7).  Asserted here: this
path is within 1e-4 of the reference's float32 result (the reference's tests use rtol 1e-5 for
single float32 kernels, tests/test_inference.py:46-121) and NO FURTHER from the float64 result
than twice the reference's float32 run is.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose
from scipy.sparse import coo_matrix

from schpf_b200 import scHPF
from schpf_b200.engine import CaviEngine
from conftest import load_golden, max_rel
from oracle import hpf_c as oc
from test_gpu_engine import _engine_from, _prep_capacity_shapes, _random_problem

pytestmark = pytest.mark.gpu

NAMES = ("theta", "beta", "xi", "eta")
F32_RTOL = 1e-4


@pytest.fixture(scope="module")
def g20():
    return load_golden("fp32_k20.npz")


def test_engine_fp32_tracks_reference_float32_run(g20):
    g = _prep_capacity_shapes({k: (v.astype(np.float64) if v.dtype == np.float32 else v) for k, v in g20.items()},
                              "init_", 20)
    with _engine_from(g, "init_", 20, precision=32) as e:
        assert e.counter("precision") == 32
        losses = []
        for t in range(20):
            e.step(1)
            if t % 5 == 0:
                losses.append(e.loss())
        st = e.get_state()
        assert e.counter("slow_path_hits") == 0
    noise = float(np.max(g["f32_vs_f64_max_rel"]))            # the reference's own float32 noise
    worst64 = 0.0
    for n in NAMES:
        for i, s in ((0, "shp"), (1, "rte")):
            assert_allclose(st[n][i], g["fin_%s_%s" % (n, s)].astype(np.float64), rtol=F32_RTOL, err_msg=n + s)
            worst64 = max(worst64, max_rel(st[n][i], g["f64_%s_%s" % (n, s)]))
    assert worst64 < 2 * noise, (worst64, noise)
    assert_allclose(losses, g["loss"], rtol=2e-5)
    assert_allclose(losses, g["f64_loss"], rtol=2e-5)


def test_estimator_routes_float32_models_to_the_fp32_sweep(g20, monkeypatch):
    g = g20
    seen = []
    import schpf_b200.scHPF_ as mod

    class Spy(CaviEngine):
        def __init__(self, *a, **kw):
            super().__init__(*a, **kw)
            seen.append(self)

        def close(self):
            if self._h is not None:
                seen.append(self.counter("precision"))
            super().close()
    monkeypatch.setattr(mod, "CaviEngine", Spy)
    X = coo_matrix((g["data"], (g["row"], g["col"])), shape=tuple(int(v) for v in g["shape"]))
    np.random.seed(int(g["seed"]))
    m = scHPF(20, verbose=False, dtype=np.float32)
    m._initialize(X)
    for n in NAMES:
        assert np.array_equal(getattr(m, n).vi_shape, g["init_%s_shp" % n])     # same fp32 draws
    m.fit(X, reinit=False, min_iter=20, max_iter=20, check_freq=5)
    assert 32.0 in seen
    for n in NAMES:
        d = getattr(m, n)
        assert d.vi_shape.dtype == np.float32 and d.vi_rate.dtype == np.float32
        assert_allclose(d.vi_shape, g["fin_%s_shp" % n], rtol=F32_RTOL)
        assert_allclose(d.vi_rate, g["fin_%s_rte" % n], rtol=F32_RTOL)
    assert_allclose(m.loss, g["loss"], rtol=2e-5)
    # float64 models never take it
    seen.clear()
    m64 = scHPF(20, verbose=False)
    m64.fit(X, min_iter=2, max_iter=2, check_freq=2)
    assert 32.0 not in seen and 64.0 in seen


@pytest.mark.parametrize("K", [1, 3, 7, 16, 20, 30, 32, 33, 50, 64])
def test_fp32_sweep_against_oracle_every_plane_count(K):
    """one plane (K <= 32), two planes (K <= 64), odd K (atomic epilogue), ragged input with
    duplicates, empty cells / genes and explicit zeros; fp64 oracle as the exact answer"""
    C, G, nnz, n_iter = 301, 173, 4000, 5
    row, col, data, st = _random_problem(C, G, K, nnz, K, zeros=True)
    hyp = (0.3, 1.0, 0.7, 0.3, 1.0, 1.3)
    with CaviEngine(C, G, K, precision=32, panel_rows=64) as e:
        e.set_coo(row, col, data)
        e.set_hyper(*hyp)
        e.set_state(theta=(st.theta_shp, st.theta_rte), beta=(st.beta_shp, st.beta_rte),
                    xi=(st.xi_shp, st.xi_rte), eta=(st.eta_shp, st.eta_rte))
        e.step(n_iter)
        got, loss = e.get_state(), e.loss()
        assert e.counter("slow_path_hits") == 0
    want_loss = oc.cavi_run(data, row, col, st, *hyp, n_iter, check_freq=0)
    assert max_rel(got["theta"][0], st.theta_shp) < 2e-5 and max_rel(got["theta"][1], st.theta_rte) < 2e-5
    assert max_rel(got["beta"][0], st.beta_shp) < 2e-5 and max_rel(got["beta"][1], st.beta_rte) < 2e-5
    assert max_rel(got["xi"][1], st.xi_rte) < 2e-5 and max_rel(got["eta"][1], st.eta_rte) < 2e-5
    want_llh = oc.compute_pois_llh(data, row, col, st.theta_shp, st.theta_rte, st.beta_shp, st.beta_rte)
    assert_allclose(loss, np.mean(-want_llh), rtol=1e-5)


def test_fp32_mass_conservation_at_scale():
    """sum_k (theta_shape - a) over a cell == the cell's total count, whatever the rounding of phi:
    the factored tables the finalisation multiplies by are the values the sweep read."""
    C, G, K = 20000, 6000, 20
    rng = np.random.default_rng(3)
    nnz = 3_000_000
    row = rng.integers(0, C, nnz).astype(np.int32)
    col = rng.integers(0, G, nnz).astype(np.int32)
    data = rng.integers(1, 9, nnz).astype(np.int32)
    _, _, _, st = _random_problem(C, G, K, 1, 5)
    with CaviEngine(C, G, K, precision=32) as e:
        e.set_coo(row, col, data)
        e.set_hyper(0.3, 1.0, 0.7, 0.3, 1.0, 1.3)
        e.set_state(theta=(st.theta_shp, st.theta_rte), beta=(st.beta_shp, st.beta_rte),
                    xi=(st.xi_shp, st.xi_rte), eta=(st.eta_shp, st.eta_rte))
        e.step(3)
        got = e.get_state()
    rows = np.bincount(row, weights=data, minlength=C)
    cols = np.bincount(col, weights=data, minlength=G)
    assert_allclose((got["theta"][0] - 0.3).sum(1), rows, rtol=2e-6, atol=1e-6)
    assert_allclose((got["beta"][0] - 0.3).sum(1), cols, rtol=2e-6, atol=1e-6)


def test_fp32_normaliser_underflow_is_redone_in_log_space():
    """cells and genes concentrated on different factors: the fp32 factored normaliser underflows
    (e^-100 per table entry, fp32 flushes below ~1e-38) where fp64 does not; the nonzero is queued
    and redone in log space in fp64 like the reference does every nonzero (hpf_numba.py:98-112)"""
    C, G, K = 64, 48, 4
    rng = np.random.default_rng(0)
    nnz = 600
    row = rng.integers(0, C, nnz).astype(np.int32)
    col = rng.integers(0, G, nnz).astype(np.int32)
    data = rng.integers(1, 5, nnz).astype(np.int32)
    ts = np.full((C, K), 0.3)
    tr = np.full((C, K), 1e44)
    tr[:, 0] = 1.0          # cells live on factor 0 ...
    bs = np.full((G, K), 0.3)
    br = np.full((G, K), 1e44)
    br[:, 1] = 1.0          # ... genes on factor 1: every product of table entries is ~1e-44
    from oracle import hpf_numpy as onp
    st = onp.State(ts, tr, bs, br, np.full(C, 2.2), np.ones(C), np.full(G, 2.2), np.ones(G))
    hyp = (0.3, 1.0, 0.7, 0.3, 1.0, 1.3)
    with CaviEngine(C, G, K, precision=32) as e:
        e.set_coo(row, col, data)
        e.set_hyper(*hyp)
        e.set_state(theta=(ts, tr), beta=(bs, br), xi=(st.xi_shp, st.xi_rte), eta=(st.eta_shp, st.eta_rte))
        e.step(1)
        got = e.get_state()
        assert e.counter("slow_path_hits") > 0
    oc.cavi_run(data, row, col, st, *hyp, 1, check_freq=0)
    assert max_rel(got["theta"][0], st.theta_shp) < 1e-6
    assert max_rel(got["beta"][0], st.beta_shp) < 1e-6


def test_float32_minibatch_and_projection_take_the_fp32_sweep(g20):
    """every engine a float32 model creates (minibatch windows, the master engine, projections) runs the
    fp32 sweep; results stay within float32 noise of the float64 model's on the same inputs"""
    g = g20
    X = coo_matrix((g["data"], (g["row"], g["col"])), shape=tuple(int(v) for v in g["shape"]))
    out = {}
    for dt in (np.float32, np.float64):
        np.random.seed(11)
        m = scHPF(20, verbose=False, dtype=dt)
        m.fit(X, batchsize=150, min_iter=12, max_iter=12, check_freq=4)
        np.random.seed(12)
        p = m.project(X, min_iter=5, max_iter=5, check_freq=5)
        assert m.theta.vi_shape.dtype == dt and p.theta.vi_rate.dtype == dt
        out[dt] = (m, p)
    (m32, p32), (m64, p64) = out[np.float32], out[np.float64]
    # the two models start from the same draws rounded to float32 / float64: agreement ~ float32 eps, amplified
    # by 12 minibatch iterations
    med = lambda a, b: float(np.median(np.abs(a.astype(np.float64) - b) / np.abs(b)))
    assert med(m32.beta.vi_shape, m64.beta.vi_shape) < 1e-4
    assert med(m32.theta.vi_shape, m64.theta.vi_shape) < 1e-4
    assert_allclose(m32.loss, m64.loss, rtol=1e-3)
    assert_allclose(p32.loss, p64.loss, rtol=1e-3)
    assert np.all(np.isfinite(p32.theta.vi_shape)) and np.all(p32.theta.vi_shape > 0)
