"""GPU parity of the one-lane-per-owner sweep (csrc/sweep_lanes.cu; K <= 20 and 29..32) and the
exact integer round trip of every device layout: the stream decoded back to
triples is, as a multiset, exactly the COO input -- for the lane-pair stream, the scheduled
one-lane stream (K=20), the schedule-free one (K=16, 32) and per-panel owner ranking.

Tolerance 1e-9 relative against the oracle (measured ~1e-13: summation order only).
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_equal

from schpf_b200.engine import CaviEngine
from conftest import max_rel
from oracle import hpf_c as oc
from oracle import hpf_numpy as onp

pytestmark = pytest.mark.gpu
TOL = 1e-9
HYP = (0.3, 1.0, 0.7, 0.3, 1.0, 1.3)


def _problem(C, G, K, nnz, seed, zeros=True, big=False):
    rng = np.random.default_rng(seed)
    row = rng.integers(0, C, nnz).astype(np.int32)
    col = rng.integers(0, G, nnz).astype(np.int32)
    data = rng.integers(0 if zeros else 1, 30, nnz).astype(np.int32)
    if big:                                  # counts whose double has a non-zero low word
        data[::97] = (1 << 21) + 12345
    st = onp.State(rng.uniform(0.15, 0.45, (C, K)), rng.uniform(0.5, 1.5, (C, K)),
                   rng.uniform(0.15, 0.45, (G, K)), rng.uniform(0.5, 1.5, (G, K)),
                   np.full(C, 1.0 + K * 0.3), rng.uniform(0.5, 1.5, C),
                   np.full(G, 1.0 + K * 0.3), rng.uniform(0.5, 1.5, G))
    return row, col, data, st


def _load(e, row, col, data, st):
    e.set_coo(row, col, data)
    e.set_hyper(*HYP)
    e.set_state(theta=(st.theta_shp, st.theta_rte), beta=(st.beta_shp, st.beta_rte),
                xi=(st.xi_shp, st.xi_rte), eta=(st.eta_shp, st.eta_rte))


@pytest.mark.parametrize("K", [1, 2, 5, 7, 10, 12, 13, 15, 16, 17, 19, 20, 29, 30, 31, 32])
@pytest.mark.parametrize("opts", [dict(), dict(panel_rows=64), dict(panel_rows=64, rank_per_range=1),
                                  dict(panel_rows=128, rank_per_range=0, warps_per_cta=2),
                                  dict(panel_rows=64, free_schedule=1), dict(free_schedule=0, rank_per_range=0),
                                  dict(packed_entries=0), dict(panel_rows=64, packed_entries=0, rank_per_range=0)])
def test_lane_sweep_against_oracle_and_lane_pairs(K, opts):
    C, G, nnz, n_iter = 517, 389, 24000, 4
    row, col, data, st = _problem(C, G, K, nnz, K)
    out = {}
    for lanes in (1, 0):
        with CaviEngine(C, G, K, lanes=lanes, **opts) as e:
            _load(e, row, col, data, st)
            assert e.counter("lanes") == lanes
            e.step(n_iter)
            out[lanes] = (e.get_state(), e.loss(), e.counter("slow_path_hits"))
    want_loss = oc.cavi_run(data, row, col, st, *HYP, n_iter, check_freq=0)
    for lanes in (1, 0):
        got, loss, hits = out[lanes]
        assert max_rel(got["theta"][0], st.theta_shp) < TOL and max_rel(got["theta"][1], st.theta_rte) < TOL
        assert max_rel(got["beta"][0], st.beta_shp) < TOL and max_rel(got["beta"][1], st.beta_rte) < TOL
        assert max_rel(got["xi"][1], st.xi_rte) < TOL and max_rel(got["eta"][1], st.eta_rte) < TOL
        want_llh = oc.compute_pois_llh(data, row, col, st.theta_shp, st.theta_rte, st.beta_shp, st.beta_rte)
        assert_allclose(loss, np.mean(-want_llh), rtol=1e-11)
        assert hits == 0
    assert max_rel(out[1][0]["theta"][0], out[0][0]["theta"][0]) < 1e-11


@pytest.mark.parametrize("K", [16, 20, 32])
def test_lane_sweep_counts_beyond_the_high_word_encoding(K):
    """counts >= 2^21 do not fit the high word of a double: the stream then carries plain integers"""
    C, G, nnz = 300, 200, 9000
    row, col, data, st = _problem(C, G, K, nnz, 5, big=True)
    with CaviEngine(C, G, K) as e:
        _load(e, row, col, data, st)
        e.step(2)
        got, loss = e.get_state(), e.loss()
    oc.cavi_run(data, row, col, st, *HYP, 2, check_freq=0)
    assert max_rel(got["theta"][0], st.theta_shp) < TOL and max_rel(got["beta"][0], st.beta_shp) < TOL
    want_llh = oc.compute_pois_llh(data, row, col, st.theta_shp, st.theta_rte, st.beta_shp, st.beta_rte)
    assert_allclose(loss, np.mean(-want_llh), rtol=1e-11)


@pytest.mark.parametrize("K", [4, 16, 20, 30])
def test_underflowed_nonzeros_are_queued_and_redone_in_log_space(K):
    """Rows whose Elog maxima sit on different factors by more than ~640 nats underflow the
    factored softmax; the sweep queues those nonzeros and slow_fixup_kernel redoes them like the
    reference does every nonzero (hpf_numba.py:98-112)."""
    C, G = 64, 48
    rng = np.random.default_rng(2)
    row = np.repeat(np.arange(C, dtype=np.int32), 6)
    col = rng.integers(0, G, row.shape[0]).astype(np.int32)
    data = rng.integers(1, 9, row.shape[0]).astype(np.int32)
    ts = np.full((C, K), 1e-3); ts[np.arange(C), np.arange(C) % K] = 50.0     # psi(1e-3) ~ -1000
    bs = np.full((G, K), 1e-3); bs[np.arange(G), (np.arange(G) + 1) % K] = 50.0
    st = onp.State(ts, np.ones((C, K)), bs, np.ones((G, K)), np.full(C, 1.0), np.ones(C), np.full(G, 1.0), np.ones(G))
    hyp = (1e-3, 1.0, 0.7, 1e-3, 1.0, 1.3)
    with CaviEngine(C, G, K) as e:
        e.set_coo(row, col, data)
        e.set_hyper(*hyp)
        e.set_state(theta=(st.theta_shp, st.theta_rte), beta=(st.beta_shp, st.beta_rte),
                    xi=(st.xi_shp, st.xi_rte), eta=(st.eta_shp, st.eta_rte))
        e.step(1)
        got = e.get_state()
        hits = e.counter("slow_path_hits")
    assert hits > 0
    onp.cavi_iteration(data, row, col, st, hyp[0], hyp[2], hyp[3], hyp[5])
    assert max_rel(got["theta"][0], st.theta_shp) < 1e-9
    assert max_rel(got["beta"][0], st.beta_shp) < 1e-9
    assert np.all(np.isfinite(got["theta"][0])) and np.all(np.isfinite(got["beta"][1]))


def _multiset(a, b, c):
    order = np.lexsort((c, b, a))
    return np.stack([a[order], b[order], c[order]])


@pytest.mark.parametrize("K,opts", [
    (5, dict(panel_rows=64)), (5, dict(panel_rows=64, packed_entries=1)), (20, dict(lanes=0, panel_rows=128)),
    (20, dict()), (20, dict(panel_rows=64, rank_per_range=1)), (16, dict(panel_rows=64)), (16, dict(rank_per_range=0)),
    (19, dict(panel_rows=64, free_schedule=1)), (20, dict(free_schedule=1, rank_per_range=0)),
    (30, dict(panel_rows=32, warps_per_cta=1)), (50, dict(panel_rows=16)),
    (20, dict(packed_entries=0)), (16, dict(panel_rows=64, packed_entries=0)), (30, dict(panel_rows=32, packed_entries=0)),
    (7, dict()),
])
@pytest.mark.parametrize("big", [False, True])
def test_layout_round_trip_is_exact(K, opts, big):
    """multiset{(row, col, y)} of the device streams == the COO input, np.array_equal (duplicates,
    explicit zeros, empty rows / columns included), for both sweep directions"""
    C, G, nnz = 700, 450, 30000
    row, col, data, st = _problem(C, G, K, nnz, 17, zeros=True, big=big)
    want = _multiset(row, col, data)
    with CaviEngine(C, G, K, **opts) as e:
        e.set_coo(row, col, data)
        for side in (0, 1):
            own, oth, cnt = e.layout_dump(side)
            assert own.shape[0] == int(e.counter("padded_nnz_cells" if side == 0 else "padded_nnz_genes"))
            real = own >= 0
            assert real.sum() == nnz                                   # pads carry no triple
            assert np.all(oth[~real] == -1) and np.all(cnt[~real] == 0)
            r, c = (own[real], oth[real]) if side == 0 else (oth[real], own[real])
            assert np.array_equal(_multiset(r, c, cnt[real]), want)


@pytest.mark.parametrize("K", [7, 16, 20, 32])
def test_stream_encoding_follows_the_counts(K):
    """schedule-free one-lane streams can use 4-byte entries (pad<<31 | count<<12 | row) when every
    count is below 2^19, wide entries otherwise; both give the same result"""
    C, G, nnz = 400, 300, 12000
    out = {}
    for big, opts, want in ((False, dict(packed_entries=1), 1), (False, dict(packed_entries=0), 0),
                            (True, dict(packed_entries=1), 0)):
        row, col, data, st = _problem(C, G, K, nnz, 3, big=big)
        with CaviEngine(C, G, K, **opts) as e:
            _load(e, row, col, data, st)
            assert e.counter("packed_entries") == want
            e.step(3)
            out[(big, want)] = (e.get_state(), e.counter("layout_bytes"))
    assert max_rel(out[(False, 1)][0]["theta"][0], out[(False, 0)][0]["theta"][0]) < 1e-12
    assert max_rel(out[(False, 1)][0]["beta"][0], out[(False, 0)][0]["beta"][0]) < 1e-12
    assert out[(False, 1)][1] < out[(False, 0)][1]
    # defaults: packed where it is not slower (K <= 16, fp32), wide where the count conversion costs fp64 issue slots
    row, col, data, st = _problem(C, G, K, nnz, 3)
    for opts, want in ((dict(), 1 if K <= 16 else 0), (dict(precision=32), 1)):
        with CaviEngine(C, G, K, **opts) as e:
            e.set_coo(row, col, data)
            assert e.counter("packed_entries") == want


def test_per_panel_ranking_removes_the_padding_of_schedule_free_streams():
    """K=16 rows all start at bank group 0 (no bank schedule): with the owners ranked panel by
    panel the only pads left are the count differences of neighbours in the sorted order."""
    C, G, K = 6000, 3000, 16
    rng = np.random.default_rng(0)
    row = np.repeat(np.arange(C, dtype=np.int32), 300)
    col = rng.integers(0, G, row.shape[0]).astype(np.int32)
    data = np.ones_like(row)
    pads = {}
    for rank in (0, 1):
        with CaviEngine(C, G, K, rank_per_range=rank, panel_rows=512) as e:
            e.set_coo(row, col, data)
            pads[rank] = e.counter("padded_nnz_cells") / row.shape[0] - 1.0
    assert pads[1] < 0.05 < pads[0]
