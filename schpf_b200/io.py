"""Device-side ingest of the files the reference loads before a fit.

    X = schpf_b200.io.load_mtx("filtered.mtx", device=0)      # scipy.io.mmread,  bin/scHPF:373-374
    X = schpf_b200.io.load_coo("matrix.tsv", device=0)        # preprocessing.load_coo, preprocessing.py:11-29
    model = scHPF(K).fit(X)

The file's bytes go to the GPU unparsed and the triples are parsed there (csrc/ingest.cu); the
result is a `DeviceCOO` whose row / col / data stay in device memory, in file order, and which
`scHPF.fit` / `project` / `CaviEngine.set_coo` take directly (no host copy of the matrix is ever
made).  `DeviceCOO.tocoo()` returns the scipy matrix `mmread` / `load_coo` would have returned.

Only what the reference's own pipeline produces is accepted: MatrixMarket `matrix coordinate`
with field integer / real / pattern and symmetry `general`; anything else raises ValueError.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import c_i64, c_int, c_vp


class DeviceCOO(object):
    """COO triples resident on one GPU: `.row`, `.col`, `.data` are int32 CUDA tensors."""

    def __init__(self, row, col, data, shape):
        self.row, self.col, self.data = row, col, data
        self.shape = (int(shape[0]), int(shape[1]))
        self.nnz = int(row.numel())
        self.dtype = np.dtype(np.int32)

    @property
    def device(self):
        return self.row.device

    def sum(self, axis=None):
        """Like scipy's sparse `.sum`: axis=1 -> (ncells, 1) cell totals, axis=0 -> (1, ngenes) gene totals
        (what the empirical hyperparameters need, scHPF_.py:847-879); exact (int64 accumulation)."""
        import torch
        if axis is None:
            return int(self.data.sum(dtype=torch.int64).item())
        if axis not in (0, 1, -1, -2):
            raise ValueError("axis out of range")
        by_row = axis in (1, -1)
        n = self.shape[0] if by_row else self.shape[1]
        idx = (self.row if by_row else self.col).long()
        out = torch.zeros(n, dtype=torch.int64, device=self.row.device).index_add_(0, idx, self.data.long())
        out = out.cpu().numpy()
        return np.asmatrix(out[:, None] if by_row else out[None, :])

    def tocoo(self):
        from scipy.sparse import coo_matrix
        return coo_matrix((self.data.cpu().numpy(), (self.row.cpu().numpy(), self.col.cpu().numpy())), shape=self.shape)

    def tocsr(self):
        return self.tocoo().tocsr()


def _device_index(device):
    return device.index if hasattr(device, "index") and device.index is not None else int(device)


def _upload(path, device):
    import torch
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size == 0:
        raise ValueError("%s is empty" % path)
    return torch.from_numpy(raw).to(torch.device("cuda", _device_index(device))), raw


def _parse(text, begin, nfields, index_base, capacity, device):
    import torch
    lib = _lib.load()
    dev = _device_index(device)
    tdev = torch.device("cuda", dev)
    stream = torch.cuda.current_stream(tdev).cuda_stream
    nbytes = int(text.numel())
    if capacity is None:
        n = c_i64(0)
        _lib.check(lib.schpf_count_lines(c_int(dev), c_vp(stream), c_vp(text.data_ptr()), c_i64(nbytes), c_i64(begin),
                                         ctypes.byref(n)))
        capacity = int(n.value)
    if capacity == 0:
        n = c_i64(0)
        _lib.check(lib.schpf_count_lines(c_int(dev), c_vp(stream), c_vp(text.data_ptr()), c_i64(nbytes), c_i64(begin),
                                         ctypes.byref(n)))
        if n.value:
            raise ValueError("schpf_b200.io: %d data lines where none were announced" % n.value)
        empty = torch.empty(0, dtype=torch.int32, device=tdev)
        return empty, empty.clone(), empty.clone()
    row = torch.empty(capacity, dtype=torch.int32, device=tdev)
    col = torch.empty(capacity, dtype=torch.int32, device=tdev)
    val = torch.empty(capacity, dtype=torch.int32, device=tdev)
    n_out, err = c_i64(0), c_i64(-1)
    rc = lib.schpf_parse_triples(c_int(dev), c_vp(stream), c_vp(text.data_ptr()), c_i64(nbytes), c_i64(begin),
                                 c_int(nfields), c_int(index_base), c_vp(row.data_ptr()), c_vp(col.data_ptr()),
                                 c_vp(val.data_ptr()), c_i64(capacity), ctypes.byref(n_out), ctypes.byref(err))
    if rc != 0:
        msg = lib.schpf_last_error()
        raise ValueError("schpf_b200.io: %s" % (msg.decode() if msg else "parse error %d" % rc))
    n = int(n_out.value)
    return row[:n], col[:n], val[:n]


def parse_mtx_header(head, total_bytes=None, path="<mtx>"):
    """The few lines in front of the data of a MatrixMarket file: banner, comments, size line.
    `head` = the first bytes of the file.  Returns (nrows, ncols, nnz, field, offset of the first data
    byte).  Raises ValueError for anything that is not a `matrix coordinate` file with field integer /
    real / pattern and symmetry `general` (what the reference's `prep` writes, bin/scHPF:327,361)."""
    total_bytes = len(head) if total_bytes is None else total_bytes
    pos, banner, size = 0, None, None
    while size is None:
        nl = head.find(b"\n", pos)
        if nl < 0:
            if len(head) < total_bytes or pos >= len(head):
                raise ValueError("%s: no size line in the first %d bytes" % (path, len(head)))
            nl = len(head)                                   # last line without a newline
        line = head[pos:nl].decode("ascii", "replace").strip()
        pos = nl + 1
        if banner is None:
            banner = line
            if not line.lower().startswith("%%matrixmarket"):
                raise ValueError("%s: not a MatrixMarket file (no %%%%MatrixMarket banner)" % path)
        elif line and not line.startswith("%"):
            size = line.split()
    tok = banner.lower().split()
    if len(tok) < 5 or tok[1] != "matrix" or tok[2] != "coordinate":
        raise ValueError("%s: only `matrix coordinate` MatrixMarket files hold a sparse count matrix" % path)
    field, symmetry = tok[3], tok[4]
    if field not in ("integer", "real", "pattern") or symmetry != "general":
        raise ValueError("%s: MatrixMarket field %r / symmetry %r is not a count matrix the reference writes "
                         "(integer | real | pattern, general)" % (path, field, symmetry))
    if len(size) != 3:
        raise ValueError("%s: bad size line %r" % (path, " ".join(size)))
    try:
        nrows, ncols, nnz = (int(v) for v in size)
    except ValueError:
        raise ValueError("%s: bad size line %r" % (path, " ".join(size)))
    if nrows < 0 or ncols < 0 or nnz < 0:
        raise ValueError("%s: bad size line %r" % (path, " ".join(size)))
    return nrows, ncols, nnz, field, min(pos, total_bytes)


def load_mtx(path, device=0):
    """MatrixMarket coordinate file -> DeviceCOO (cells x genes as stored; the reference's `prep`
    writes cells as rows)."""
    text, raw = _upload(path, device)
    nrows, ncols, nnz, field, begin = parse_mtx_header(bytes(raw[:1 << 22]), raw.size, path)
    row, col, val = _parse(text, begin, 2 if field == "pattern" else 3, 1, nnz, device)
    if int(row.numel()) != nnz:
        raise ValueError("%s: size line announces %d entries, file holds %d" % (path, nnz, int(row.numel())))
    if nnz and (int(row.max()) >= nrows or int(col.max()) >= ncols):
        raise ValueError("%s: index outside the %d x %d matrix of the size line" % (path, nrows, ncols))
    return DeviceCOO(row, col, val, (nrows, ncols))


def load_coo(path, device=0):
    """The reference's tab-separated triples (preprocessing.py:11-29: 0-indexed "cell<TAB>gene<TAB>count"
    lines) -> DeviceCOO; like `coo_matrix((data, (row, col)))` the shape is (max row + 1, max col + 1)."""
    text, _ = _upload(path, device)
    row, col, val = _parse(text, 0, 3, 0, None, device)
    if int(row.numel()) == 0:
        raise ValueError("%s holds no triples" % path)
    return DeviceCOO(row, col, val, (int(row.max()) + 1, int(col.max()) + 1))


def load(path, device=0):
    """What bin/scHPF:373 does: `.mtx` -> mmread, anything else -> load_coo."""
    return load_mtx(path, device) if str(path).endswith(".mtx") else load_coo(path, device)
