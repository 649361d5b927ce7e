"""GPU parity, function level: every replacement of a reference numba kernel
(schpf_b200/hpf_cuda.py -> C ABI -> CUDA) against the golden outputs of the
reference (tests/golden/kernels_k4.npz) and against the oracle on seeded
inputs.  Mirrors the reference's tests/test_inference.py, with its tolerances
(rtol 1e-7, atol 0) as the outer bound and the tolerances actually asserted
written at each check."""
import numpy as np
import pytest
from numpy.testing import assert_allclose
from scipy.special import digamma, gammaln

from schpf_b200 import hpf_cuda
from oracle import hpf_c as oc

pytestmark = pytest.mark.gpu


def _coo(g):
    return g["data"], g["row"], g["col"]


@pytest.mark.parametrize("x", [0.0001, 0.001, 0.01, 0.1, 1, 10, 100, 1000])
def test_digamma_gammaln_reference_points(x, g_kernels):
    # reference tests/test_inference.py:24-37 (assert_allclose default rtol 1e-7)
    assert_allclose(hpf_cuda.psi(np.float64(x)), digamma(x), rtol=1e-13)
    assert_allclose(hpf_cuda.cgammaln(np.float64(x)), gammaln(x), rtol=1e-13)


def test_digamma_dense_grid():
    x = np.exp(np.linspace(np.log(1e-4), np.log(1e6), 200001))
    got, want = hpf_cuda.psi(x), digamma(x)
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    assert err.max() < 5e-15
    # device and C oracle: same series, the device sums its recurrence as one fraction
    assert (np.abs(got - oc.psi(x)) / np.maximum(1.0, np.abs(want))).max() < 1e-14
    # Gauss special values (scipy/special/tests/test_digamma.py)
    eg = np.euler_gamma
    vals = hpf_cuda.psi(np.array([1.0, 0.5, 1 / 3., 0.25]))
    want = [-eg, -2 * np.log(2) - eg, -np.pi / (2 * np.sqrt(3)) - 1.5 * np.log(3) - eg,
            -np.pi / 2 - 3 * np.log(2) - eg]
    assert_allclose(vals, want, rtol=1e-14)
    assert_allclose(hpf_cuda.psi(np.arange(1, 12, dtype=np.float64)), digamma(np.arange(1, 12)), rtol=1e-14)


def test_gammaln_grid():
    x = np.concatenate([np.arange(1, 300, dtype=np.float64), np.exp(np.linspace(-9, 13, 5000))])
    assert_allclose(hpf_cuda.cgammaln(x), gammaln(x), rtol=2e-13, atol=2e-14)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_compute_Xphi_data(g_kernels, dtype):
    # reference tests/test_inference.py:40-57: rtol 1e-7 (fp64) / 1e-5 (fp32), atol 0
    g = g_kernels
    args = [g[k].astype(dtype) for k in ("theta_shp", "theta_rte", "beta_shp", "beta_rte")]
    got = hpf_cuda.compute_Xphi_data(*_coo(g), *args)
    assert got.dtype == dtype
    assert_allclose(got, g["Xphi"], rtol=1e-12 if dtype == np.float64 else 1e-5, atol=0)
    if dtype == np.float64:
        assert_allclose(got.sum(1), g["data"], rtol=1e-14)
    # the reference's single-process twin (hpf_numba.py:117-125) is the same device kernel here
    from scipy.sparse import coo_matrix
    from schpf_b200 import HPF_Gamma
    data, row, col = _coo(g)
    X = coo_matrix((data, (row, col)), shape=tuple(int(v) for v in g["shape"]))
    twin = hpf_cuda.compute_Xphi_data_numpy(X, HPF_Gamma(args[0], args[1]), HPF_Gamma(args[2], args[3]))
    assert twin.dtype == dtype and np.array_equal(twin, got)


def test_shape_updates(g_kernels):
    # reference tests/test_inference.py:60-88, rtol 1e-7
    g = g_kernels
    nc, ng = (int(v) for v in g["shape"])
    assert_allclose(hpf_cuda.compute_loading_shape_update(g["Xphi_rand"], g["row"], nc, float(g["a"])),
                    g["theta_shape_upd"], rtol=1e-13)
    assert_allclose(hpf_cuda.compute_loading_shape_update(g["Xphi_rand"], g["col"], ng, float(g["c"])),
                    g["beta_shape_upd"], rtol=1e-13)


def test_rate_updates(g_kernels):
    # reference tests/test_inference.py:91-108
    g = g_kernels
    assert_allclose(hpf_cuda.compute_loading_rate_update(g["xi_shp"], g["xi_rte"], g["beta_shp"], g["beta_rte"]),
                    g["theta_rate_upd"], rtol=1e-13)
    assert_allclose(hpf_cuda.compute_loading_rate_update(g["eta_shp"], g["eta_rte"], g["theta_shp"], g["theta_rte"]),
                    g["beta_rate_upd"], rtol=1e-13)
    assert_allclose(hpf_cuda.compute_capacity_rate_update(g["beta_shp"], g["beta_rte"], float(g["dp"])),
                    g["eta_rate_upd"], rtol=1e-13)
    assert_allclose(hpf_cuda.compute_capacity_rate_update(g["theta_shp"], g["theta_rte"], float(g["bp"])),
                    g["xi_rate_upd"], rtol=1e-13)


def test_pois_llh(g_kernels):
    # reference tests/test_inference.py:111-121, rtol 1e-7
    g = g_kernels
    got = hpf_cuda.compute_pois_llh(*_coo(g), g["theta_shp"], g["theta_rte"], g["beta_shp"], g["beta_rte"])
    assert_allclose(got, g["llh"], rtol=1e-12)


@pytest.mark.parametrize("K", [1, 3, 7, 20, 50, 64])
def test_kernels_against_oracle_random(K):
    rng = np.random.default_rng(K)
    C, G, nnz = 257, 131, 5000
    row = rng.integers(0, C, nnz).astype(np.int32)
    col = rng.integers(0, G, nnz).astype(np.int32)
    data = rng.integers(0, 40, nnz).astype(np.int32)            # explicit zeros and duplicates included
    ts, tr = rng.gamma(2.0, 1.0, (C, K)) + 1e-3, rng.gamma(2.0, 1.0, (C, K)) + 1e-3
    bs, br = rng.gamma(0.5, 1.0, (G, K)) + 1e-3, rng.gamma(2.0, 1.0, (G, K)) + 1e-3
    xphi = hpf_cuda.compute_Xphi_data(data, row, col, ts, tr, bs, br)
    # atol: entries that underflow to subnormals differ in their last bits
    assert_allclose(xphi, oc.compute_Xphi_data(data, row, col, ts, tr, bs, br), rtol=1e-11, atol=1e-300)
    assert_allclose(hpf_cuda.compute_loading_shape_update(xphi, col, G, 0.3),
                    oc.compute_loading_shape_update(xphi, col, G, 0.3), rtol=1e-13)
    assert_allclose(hpf_cuda.compute_pois_llh(data, row, col, ts, tr, bs, br),
                    oc.compute_pois_llh(data, row, col, ts, tr, bs, br), rtol=1e-11, atol=1e-12)


def test_empty_inputs_and_bad_arguments():
    from schpf_b200._lib import SchpfError
    K = 4
    e_i = np.zeros(0, dtype=np.int32)
    ts = np.ones((3, K))
    out = hpf_cuda.compute_Xphi_data(e_i, e_i, e_i, ts, ts, ts, ts)
    assert out.shape == (0, K)
    assert hpf_cuda.compute_pois_llh(e_i, e_i, e_i, ts, ts, ts, ts).shape == (0,)
    assert_allclose(hpf_cuda.compute_loading_shape_update(np.zeros((0, K)), e_i, 3, 0.25), 0.25)
    with pytest.raises(SchpfError):                              # row index out of range
        hpf_cuda.compute_Xphi_data(np.ones(1, np.int32), np.array([3], np.int32), np.zeros(1, np.int32),
                                   ts, ts, ts, ts)
    with pytest.raises(SchpfError):                              # K beyond the instantiated kernels
        big = np.ones((3, 65))
        hpf_cuda.compute_Xphi_data(e_i, e_i, e_i, big, big, big, big)
