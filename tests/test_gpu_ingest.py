"""Device-side ingest (schpf_b200/io.py, csrc/ingest.cu) against what the reference's loaders return
for the same file: `scipy.io.mmread` (bin/scHPF:373-374) and `load_coo` = np.loadtxt
(schpf/preprocessing.py:11-29).  Integer work: everything is np.array_equal, in file order."""
import numpy as np
import pytest
from numpy.testing import assert_equal
from scipy.io import mmread, mmwrite
from scipy.sparse import coo_matrix

from schpf_b200 import scHPF, HPF_Gamma
from schpf_b200 import io as sio
from conftest import load_golden, max_rel

pytestmark = pytest.mark.gpu


def _random_coo(C, G, nnz, seed, big=False):
    rng = np.random.default_rng(seed)
    row = rng.integers(0, C, nnz).astype(np.int32)
    col = rng.integers(0, G, nnz).astype(np.int32)
    data = (rng.integers(1, 2 ** 31 - 1, nnz) if big else rng.integers(1, 5000, nnz)).astype(np.int32)
    return coo_matrix((data, (row, col)), shape=(C, G))           # duplicates kept, arbitrary order


def _same(dev, ref):
    ref = coo_matrix(ref)
    assert dev.shape == ref.shape
    assert_equal(dev.row.cpu().numpy(), ref.row.astype(np.int32))
    assert_equal(dev.col.cpu().numpy(), ref.col.astype(np.int32))
    assert_equal(dev.data.cpu().numpy().astype(np.int64), np.asarray(ref.data).astype(np.int64))


@pytest.mark.parametrize("field", ["integer", "real", "pattern"])
def test_mtx_as_written_by_the_reference_pipeline(tmp_path, field):
    X = _random_coo(3000, 1700, 250_000, 1)
    path = str(tmp_path / "m.mtx")
    if field == "pattern":
        X = coo_matrix((np.ones(X.nnz, dtype=np.int32), (X.row, X.col)), shape=X.shape)
    mmwrite(path, X.astype(np.float64) if field == "real" else X, field=field,
            comment="written by scHPF prep\nsecond comment line")
    dev, ref = sio.load_mtx(path), mmread(path)
    _same(dev, ref)
    assert dev.nnz == ref.nnz
    # totals the empirical hyperparameters are made of: same integers as scipy's
    assert_equal(np.asarray(dev.sum(axis=1)), np.asarray(X.sum(axis=1)))
    assert_equal(np.asarray(dev.sum(axis=0)), np.asarray(X.sum(axis=0)))
    assert dev.sum() == int(X.sum())


def test_tsv_like_the_reference_load_coo(tmp_path):
    X = _random_coo(900, 40000, 180_000, 2, big=True)
    path = str(tmp_path / "m.tsv")
    np.savetxt(path, np.stack([X.row, X.col, X.data], axis=1), fmt="%d", delimiter="\t")
    raw = np.loadtxt(path, delimiter="\t", dtype=int)             # preprocessing.py:27-28
    _same(sio.load_coo(path), coo_matrix((raw[:, 2], (raw[:, 0], raw[:, 1]))))
    _same(sio.load(path), coo_matrix((raw[:, 2], (raw[:, 0], raw[:, 1]))))


def test_ragged_text(tmp_path):
    """CR LF, blank lines, comments between data lines, no trailing newline, leading zeros, '+',
    blanks for tabs, trailing blanks, reals with exponents, lines longer than a thread's 32 bytes"""
    body = ("%%MatrixMarket matrix coordinate real general\r\n% c\r\n\r\n 5 7 6 \r\n"
            "1 1 3\r\n\r\n% mid comment\r\n0005\t007   2.000e+00  \r\n+2 +3 +40e-1\r\n"
            "3   2   1234567.0000000000000000000000000000000000000\r\n5 7 0\r\n4 4 2147483647")
    path = str(tmp_path / "r.mtx")
    open(path, "wb").write(body.encode())
    dev = sio.load_mtx(path)
    assert dev.shape == (5, 7)
    assert_equal(dev.row.cpu().numpy(), [0, 4, 1, 2, 4, 3])
    assert_equal(dev.col.cpu().numpy(), [0, 6, 2, 1, 6, 3])
    assert_equal(dev.data.cpu().numpy(), [3, 2, 4, 1234567, 0, 2147483647])


@pytest.mark.parametrize("bad,why", [
    ("2 2 -1\n", "negative count"), ("2 2 2.5\n", "fractional count"), ("2 x 1\n", "not a number"),
    ("2 2\n", "too few fields"), ("0 2 1\n", "index below the base"), ("2 2 1 9\n", "extra field"),
    ("2 2 2147483648\n", "count overflows int32"), ("99999999999 2 1\n", "index overflows int32")])
def test_malformed_lines_are_rejected_with_their_offset(tmp_path, bad, why):
    head = "%%MatrixMarket matrix coordinate integer general\n3 3 3\n"
    good = "1 1 1\n"
    path = str(tmp_path / "bad.mtx")
    open(path, "w").write(head + good + bad + good)
    with pytest.raises(ValueError, match="byte %d" % len(head + good)):
        sio.load_mtx(path)


def test_unsupported_headers_raise(tmp_path):
    path = str(tmp_path / "s.mtx")
    open(path, "w").write("%%MatrixMarket matrix coordinate integer symmetric\n2 2 1\n1 1 1\n")
    with pytest.raises(ValueError, match="symmetry"):
        sio.load_mtx(path)
    open(path, "w").write("%%MatrixMarket matrix array real general\n1 1\n1.0\n")
    with pytest.raises(ValueError, match="coordinate"):
        sio.load_mtx(path)
    open(path, "w").write("%%MatrixMarket matrix coordinate integer general\n2 2 2\n1 1 1\n")
    with pytest.raises(ValueError, match="announces 2"):
        sio.load_mtx(path)
    open(path, "w").write("%%MatrixMarket matrix coordinate integer general\n2 2 1\n1 3 1\n")
    with pytest.raises(ValueError, match="outside"):
        sio.load_mtx(path)


def test_fit_from_a_device_matrix_equals_fit_from_scipy(tmp_path):
    """the golden 10-iteration run of the real reference, fed through file -> GPU parse -> fit without
    the matrix ever being a host array; empirical b', d' bit-identical (tests/test_scHPF_model.py:21-22)"""
    g = load_golden("cavi_cfg1.npz")
    X = coo_matrix((g["data"], (g["row"], g["col"])), shape=tuple(int(v) for v in g["shape"]))
    path = str(tmp_path / "cfg1.mtx")
    mmwrite(path, X, field="integer")
    Xd = sio.load_mtx(path)
    m = scHPF(5, verbose=False)
    m._initialize(Xd)
    assert m.bp == float(g["bp"]) and m.dp == float(g["dp"])
    for n in ("theta", "beta", "xi", "eta"):
        setattr(m, n, HPF_Gamma(g["init_%s_shp" % n].copy(), g["init_%s_rte" % n].copy()))
    m.fit(Xd, reinit=False, min_iter=10, max_iter=10, check_freq=5)
    for n in ("theta", "beta", "xi", "eta"):
        assert max_rel(getattr(m, n).vi_shape, g["it10_%s_shp" % n]) < 1e-9
        assert max_rel(getattr(m, n).vi_rate, g["it10_%s_rte" % n]) < 1e-9
    # project (frozen genes) takes the device matrix too
    np.random.seed(3)
    p = m.project(Xd, min_iter=3, max_iter=3, check_freq=3)
    np.random.seed(3)
    q = m.project(X, min_iter=3, max_iter=3, check_freq=3)
    # same kernels on the same triples; the order in which CTAs add their panels' sums is not fixed
    assert max_rel(p.theta.vi_shape, q.theta.vi_shape) < 1e-12
