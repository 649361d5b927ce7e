"""The oracle (oracle/hpf_numpy.py, oracle/hpf_oracle.c) against golden outputs
of the real reference (tests/golden/*.npz, made by tests/golden/make_golden.py).
CPU only."""
import numpy as np
import pytest
from numpy.testing import assert_allclose
from scipy.special import digamma, gammaln

from oracle import hpf_numpy as onp
from oracle import hpf_c as oc
from conftest import max_rel

IMPLS = [pytest.param(onp, id="numpy"), pytest.param(oc, id="c")]
STATE = ("theta_shp", "theta_rte", "beta_shp", "beta_rte", "xi_shp", "xi_rte", "eta_shp", "eta_rte")


def _coo(g):
    return g["data"], g["row"], g["col"]


@pytest.mark.parametrize("impl", IMPLS)
def test_psi_gammaln_reference_points(impl, g_kernels):
    # reference tests/test_inference.py:24-37 (rtol 1e-7 there; far tighter here)
    assert_allclose(impl.psi(g_kernels["psi_x"]), g_kernels["psi_y"], rtol=1e-14)
    assert_allclose(impl.cgammaln(g_kernels["psi_x"]), g_kernels["gammaln_y"], rtol=1e-14)


def test_c_psi_dense_grid_against_scipy():
    x = np.exp(np.linspace(np.log(1e-4), np.log(1e6), 20001))
    got, want = oc.psi(x), digamma(x)
    # absolute error scaled by max(1,|psi|): psi crosses zero at 1.4616
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    assert err.max() < 4e-15
    # Gauss special values (scipy/special/tests/test_digamma.py)
    assert_allclose(oc.psi([1.0])[0], -np.euler_gamma, rtol=1e-14)
    assert_allclose(oc.psi([0.5])[0], -2 * np.log(2) - np.euler_gamma, rtol=1e-14)
    assert_allclose(oc.psi([1 / 3.])[0], -np.pi / (2 * np.sqrt(3)) - 1.5 * np.log(3) - np.euler_gamma, rtol=1e-14)
    assert_allclose(oc.psi([0.25])[0], -np.pi / 2 - 3 * np.log(2) - np.euler_gamma, rtol=1e-14)


def test_c_gammaln_grid():
    x = np.concatenate([np.arange(1, 200, dtype=np.float64), np.exp(np.linspace(-9, 13, 500))])
    assert_allclose(oc.cgammaln(x), gammaln(x), rtol=1e-13, atol=1e-14)


@pytest.mark.parametrize("impl", IMPLS)
def test_compute_Xphi_data(impl, g_kernels):
    g = g_kernels
    got = impl.compute_Xphi_data(*_coo(g), g["theta_shp"], g["theta_rte"], g["beta_shp"], g["beta_rte"])
    assert_allclose(got, g["Xphi"], rtol=1e-12, atol=0)
    assert_allclose(got.sum(1), g["data"], rtol=1e-13)


@pytest.mark.parametrize("impl", IMPLS)
def test_shape_updates(impl, g_kernels):
    g = g_kernels
    nc, ng = g["shape"]
    assert_allclose(impl.compute_loading_shape_update(g["Xphi_rand"], g["row"], int(nc), float(g["a"])),
                    g["theta_shape_upd"], rtol=1e-13)
    assert_allclose(impl.compute_loading_shape_update(g["Xphi_rand"], g["col"], int(ng), float(g["c"])),
                    g["beta_shape_upd"], rtol=1e-13)


@pytest.mark.parametrize("impl", IMPLS)
def test_rate_updates(impl, g_kernels):
    g = g_kernels
    assert_allclose(impl.compute_loading_rate_update(g["xi_shp"], g["xi_rte"], g["beta_shp"], g["beta_rte"]),
                    g["theta_rate_upd"], rtol=1e-13)
    assert_allclose(impl.compute_loading_rate_update(g["eta_shp"], g["eta_rte"], g["theta_shp"], g["theta_rte"]),
                    g["beta_rate_upd"], rtol=1e-13)
    assert_allclose(impl.compute_capacity_rate_update(g["beta_shp"], g["beta_rte"], float(g["dp"])),
                    g["eta_rate_upd"], rtol=1e-13)
    assert_allclose(impl.compute_capacity_rate_update(g["theta_shp"], g["theta_rte"], float(g["bp"])),
                    g["xi_rate_upd"], rtol=1e-13)


@pytest.mark.parametrize("impl", IMPLS)
def test_pois_llh(impl, g_kernels):
    g = g_kernels
    got = impl.compute_pois_llh(*_coo(g), g["theta_shp"], g["theta_rte"], g["beta_shp"], g["beta_rte"])
    assert_allclose(got, g["llh"], rtol=1e-12)
    assert_allclose(np.mean(-got), g["mean_neg_llh"], rtol=1e-13)


def test_empirical_hypers(g_cavi):
    from scipy.sparse import coo_matrix
    g = g_cavi
    X = coo_matrix((g["data"], (g["row"], g["col"])), shape=tuple(g["shape"]))
    bp, dp = onp.empirical_hypers(X, float(g["ap"]), float(g["cp"]))
    assert bp == float(g["bp"]) and dp == float(g["dp"])        # bit-exact (reference asserts equality)


def _init_state(g, prefix="init_"):
    return onp.State(*[g[prefix + n] for n in STATE])


@pytest.mark.parametrize("n", [1, 10, 50])
@pytest.mark.parametrize("impl", IMPLS)
def test_cavi_loop_against_reference(impl, n, g_cavi):
    g = g_cavi
    st = _init_state(g)
    hyp = [float(g[k]) for k in ("a", "ap", "bp", "c", "cp", "dp")]
    cf = int(g["it%d_check_freq" % n])
    loss = impl.cavi_run(*_coo(g), st, *hyp, n, check_freq=cf)
    for name in STATE:
        assert max_rel(getattr(st, name), g["it%d_%s" % (n, name)]) < 1e-10, name
    assert_allclose(loss, g["it%d_loss" % n], rtol=1e-12)


def test_cavi_simultaneous_against_reference(g_simul):
    g = g_simul
    st = _init_state(g)
    K = st.theta_shp.shape[1]
    onp.cavi_prepare(st, float(g["a"]), float(g["ap"]), float(g["c"]), float(g["cp"]), K)
    loss = []
    for t in range(7):
        onp.cavi_iteration(*_coo(g), st, float(g["a"]), float(g["bp"]), float(g["c"]), float(g["dp"]),
                           beta_theta_simultaneous=True)
        if t % 2 == 0:
            loss.append(onp.mean_negative_pois_llh(*_coo(g), st.theta_shp, st.theta_rte,
                                                   st.beta_shp, st.beta_rte))
    for name in STATE:
        assert max_rel(getattr(st, name), g["fin_" + name]) < 1e-11, name
    assert_allclose(loss, g["loss"], rtol=1e-12)
