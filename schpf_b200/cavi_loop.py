"""Device-side drivers of the iteration body of the reference's `_fit`
(schpf/scHPF_.py:642-715).  `scHPF._fit` keeps the scalar bookkeeping (loss list,
percent change, stopping rules) and asks one of these for iterations, the loss and
the host copy of the state:

  FullBatchLoop   every iteration sees all cells (the default, `batched == False`)
  MinibatchLoop   `batchsize` cells per iteration (scHPF_.py:626-631, 642-650, 686-704)

Both talk to the GPU only through `engine.CaviEngine` (the C ABI).
"""
from math import gcd

import numpy as np

# nnz * K above which the t == 0 random-phi draw moves from numpy (bit-compatible
# with a seeded reference run, scHPF_.py:652-655) to the device generator
HOST_DIRICHLET_LIMIT = 1 << 24

# a minibatch window gets an engine (= its own device layout) of its own while the number of
# distinct windows is at most this; beyond, one engine is re-laid out every iteration
MINIBATCH_ENGINE_CACHE = 64

# theta / xi of all cells stay on the device between minibatch iterations (see MinibatchLoop);
# False keeps them in host arrays (a batch then moves 4*batchsize*K doubles over PCIe per iteration).
# Both modes reproduce the reference's seeded minibatch runs (tests/test_gpu_minibatch.py).
MINIBATCH_DEVICE_STATE = True

# windows that are re-laid out every iteration (more distinct windows than MINIBATCH_ENGINE_CACHE) are cut
# from a device-resident copy of the permuted matrix by row-pointer slicing (needs torch for the slices)
DEVICE_WINDOW_SLICES = True


class MinibatchSchedule(object):
    """The reference's batch schedule (schpf/util.py:218-231): ONE `np.random.shuffle` of the
    cell order at the first draw, then consecutive windows of `batchsize` cells that wrap
    around the end.  `next()` returns (window start, cell indices); `order` is the shuffled
    cell order (None until the first draw, like the reference's lazy generator)."""

    def __init__(self, ncells, batchsize, rng=np.random):
        assert ncells >= batchsize
        self.ncells, self.batchsize = int(ncells), int(batchsize)
        self.order, self._start, self.rng = None, 0, rng
        self._offsets = np.arange(self.batchsize)

    def draw_order(self):
        if self.order is None:
            self.order = np.arange(self.ncells)
            self.rng.shuffle(self.order)
        return self.order

    def next(self):
        order, start = self.draw_order(), self._start
        self._start = (start + self.batchsize) % self.ncells
        return start, order[(start + self._offsets) % self.ncells]

    def pieces(self, start):
        """The window starting at `start` as contiguous ranges of the shuffled order:
        [(first position, length)], two entries when it wraps."""
        n1 = min(self.batchsize, self.ncells - start)
        return [(start, n1)] + ([(0, self.batchsize - n1)] if n1 < self.batchsize else [])


def minibatch_windows(ncells, batchsize):
    """Generator form of MinibatchSchedule: yields (window start, cell indices)."""
    sched = MinibatchSchedule(ncells, batchsize)
    while True:
        yield sched.next()


def _random_phi_step(engine, data, nfactors, rng=np.random, **flags):
    """First iteration from y * Dirichlet(1_K) instead of the E-step (scHPF_.py:652-655):
    numpy's stream while the draw is small enough to make on the host, the device's
    counter-based generator (seeded from numpy's stream) beyond."""
    nnz = data.shape[0]
    if nnz * nfactors <= HOST_DIRICHLET_LIMIT:
        random_phi = rng.dirichlet(np.ones(nfactors), nnz)
        if hasattr(data, "is_cuda"):            # io.DeviceCOO: the counts live on the device
            data = data.cpu().numpy()
        engine.step_with_xphi(data[:, None] * random_phi, **flags)
    else:
        engine.step_random_phi(int(rng.randint(0, 2 ** 31 - 1)), **flags)


class FullBatchLoop(object):
    """All cells every iteration; with `process_group`, this rank's shard of the cells."""

    def __init__(self, new_engine, X, hyper, state, nfactors, freeze_genes, simultaneous,
                 process_group=None, shared_seed=None, rng=np.random, engine_options=None):
        self.X, self.nfactors, self.rng = X, nfactors, rng
        self.flags = dict(freeze_genes=freeze_genes, simultaneous=simultaneous)
        self.freeze_genes = freeze_genes
        self.process_group, self.shared_seed = process_group, shared_seed
        self.engine = new_engine(*X.shape, **(engine_options or {}))
        if process_group is not None:
            from .engine import ShardedEngine
            self.engine = ShardedEngine(self.engine, process_group)
        try:
            self.engine.set_coo(X.row, X.col, X.data)
            self.engine.set_hyper(*hyper)
            self.engine.set_state(**state)
        except Exception:
            self.close()
            raise

    def run(self, t, n, reinit):
        """iterations t .. t+n-1"""
        if n > 0 and t == 0 and reinit:
            if self.process_group is not None:
                self.engine.step(1, random_phi_seed=self.shared_seed(self.process_group, self.rng), **self.flags)
            else:
                _random_phi_step(self.engine, self.X.data, self.nfactors, self.rng, **self.flags)
            n -= 1
        if n > 0:
            self.engine.step(n, **self.flags)

    def loss(self):
        return self.engine.loss()

    def gene_state_engine(self):
        """The device engine holding the newest beta / eta (None when the engine is not a CaviEngine)."""
        eng = getattr(self.engine, "local", self.engine)
        return eng if hasattr(eng, "copy_gene_state_from") else None

    def host_state(self):
        """{name: (vi_shape, vi_rate)}; the gene side is absent when it is frozen"""
        return self.engine.get_state(("theta", "xi") if self.freeze_genes else ("theta", "beta", "xi", "eta"))

    def close(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None


class MinibatchLoop(object):
    """`batchsize` cells per iteration, on the device.

    The reference slices `X.tocsr()[batch_ix, :].tocoo()` on the host every iteration and runs
    its kernels on that submatrix (scHPF_.py:642-650).  Here a batch is an engine of its own
    over the batch's rows (re-based to 0..batchsize-1, all genes): its sweeps and
    finalisations are the full-batch kernels, in the `batched` order of scHPF_.py:686-704
    (theta/xi first, then beta from the same Xphi with the NEW theta in its rate;
    SCHPF_CELLS_FIRST).  beta/eta stay on the device and are handed from one batch's engine
    to the next device-to-device.  The batch schedule is a fixed cyclic sequence of windows
    over one shuffled cell order, so while there are few distinct windows each keeps its
    engine and device layout; otherwise one engine is re-laid out every iteration.

    Where theta/xi of all cells live between iterations:
      host (default)   numpy arrays; a batch moves 4*batchsize*K doubles over PCIe and the loop
                       synchronises twice per iteration;
      device (`device_state=True`, opt-in until it has run on hardware) the cells are permuted
                       ONCE by the shuffle, so every window is a contiguous row range (two when it
                       wraps) of a full-size "master" engine, and a batch moves with
                       `schpf_copy_cell_state` (D2D, no host synchronisation in the loop).
    The default loss is taken by a full-matrix engine (the master in device mode) that gets
    its layout only if that loss is used.

    With `process_group` (cells sharded over ranks, `X` = this rank's rows): the schedule runs over
    ALL cells -- rank 0 draws the one shuffle from its numpy stream and broadcasts it, so a seeded run
    walks the same windows as the unsharded loop -- and every rank takes the cells of a window that
    it owns (in window order).  theta/xi of those cells are updated locally; beta then needs the
    sums over the whole batch, so the one exchange of the iteration happens AFTER the cell update
    (`ShardedEngine.step(cells_first=True)`).  theta/xi stay in host arrays (a rank's part of a
    window is not a contiguous range of its cells); the loss is the loss over all ranks' cells.
    """

    def __init__(self, new_engine, X, hyper, state, nfactors, batchsize, freeze_genes, simultaneous,
                 device_state=None, rng=np.random, process_group=None, shared_seed=None, cell_range=None):
        self.rng = rng
        self.process_group, self.shared_seed = process_group, shared_seed
        self.new_engine, self.hyper, self.nfactors = new_engine, hyper, nfactors
        self.ncells, self.ngenes = X.shape
        self.batchsize = int(batchsize)
        self.freeze_genes = freeze_genes
        self.flags = dict(freeze_genes=freeze_genes, simultaneous=simultaneous, cells_first=True)
        self.device_state = MINIBATCH_DEVICE_STATE if device_state is None else bool(device_state)
        self.X = X
        # like the reference's X.tocsr(): duplicates summed, columns sorted within a row
        self.Xcsr = X.tocsr()
        f = lambda pair: [np.array(pair[0], dtype=np.float64, copy=True), np.array(pair[1], dtype=np.float64, copy=True)]
        self.theta, self.xi = f(state["theta"]), f(state["xi"])      # host mode: the live copy
        self._gene_init = dict(beta=state["beta"], eta=state["eta"])
        # cell_range = (first global cell of this rank, cells of all ranks); unsharded: (0, ncells)
        self.lo, self.ncells_total = (0, self.ncells) if cell_range is None else (int(cell_range[0]), int(cell_range[1]))
        self.schedule = MinibatchSchedule(self.ncells_total, self.batchsize, rng)
        if process_group is not None:
            self.device_state = False
            self.schedule.order = self._broadcast_order(process_group, self.ncells_total, rng)
        n_windows = self.ncells_total // gcd(self.ncells_total, self.batchsize)
        self.cache_engines = n_windows <= MINIBATCH_ENGINE_CACHE
        self.engines = {}          # window start -> engine (only one entry when not caching)
        self.current = None        # the engine holding the newest beta / eta
        self.full = None           # full-matrix engine: loss; in device mode also all cells' theta / xi
        self.full_has_coo = False
        self.Xp_csr = None         # device mode: rows in shuffled order
        self._dev_rows = None      # ... and their COO arrays on the device (windows that are re-laid out)

    @staticmethod
    def _broadcast_order(group, ncells_total, rng):
        """the reference's one shuffle (util.py:220-221), drawn on rank 0, the same on every rank"""
        import torch.distributed as dist
        src = dist.get_global_rank(group, 0) if group is not None else 0
        box = [None]
        if dist.get_rank(group) == 0:
            order = np.arange(ncells_total)
            rng.shuffle(order)
            box = [order]
        dist.broadcast_object_list(box, src=src, group=group)
        return np.asarray(box[0])

    def _wrap(self, engine):
        if self.process_group is None:
            return engine
        from .engine import ShardedEngine
        return ShardedEngine(engine, self.process_group)

    # -- engines -------------------------------------------------------------
    def _full_engine(self, with_coo):
        if self.full is None:
            self.full = self._wrap(self.new_engine(self.ncells, self.ngenes))
            self.full.set_hyper(*self.hyper)
            if self.device_state:
                order = self.schedule.draw_order()
                self.Xp_csr = self.Xcsr[order, :]
                self.full.set_state(theta=(self.theta[0][order], self.theta[1][order]),
                                    xi=(self.xi[0][order], self.xi[1][order]), **self._gene_init)
        if with_coo and not self.full_has_coo:
            Xf = self.Xp_csr.tocoo() if self.device_state else self.X
            self.full.set_coo(Xf.row, Xf.col, Xf.data)
            self.full_has_coo = True
        return self.full

    def _device_rows(self):
        """The permuted matrix as row-sorted COO arrays ON THE DEVICE plus the host row pointer: a window is
        then one or two contiguous slices of them (CSR row-pointer slicing) instead of a scipy
        `tocsr()[ix].tocoo()` and an upload per iteration (the reference's scHPF_.py:645-650)."""
        if self._dev_rows is None:
            import torch
            Xp = self.Xp_csr.tocoo()           # rows ascending, columns ascending within a row
            dev = torch.device("cuda", int(getattr(self.full, "device", 0)))
            up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
            self._dev_rows = (up(Xp.row), up(Xp.col), up(Xp.data), np.asarray(self.Xp_csr.indptr, dtype=np.int64))
        return self._dev_rows

    def _batch_coo_device(self, start):
        import torch
        row, col, data, indptr = self._device_rows()
        rows, cols, vals, at = [], [], [], 0
        for a, n in self.schedule.pieces(start):
            lo, hi = int(indptr[a]), int(indptr[a + n])
            rows.append(row[lo:hi] - (a - at))
            cols.append(col[lo:hi])
            vals.append(data[lo:hi])
            at += n
        if len(rows) == 1:
            return rows[0], cols[0], vals[0]
        return torch.cat(rows), torch.cat(cols), torch.cat(vals)

    def _batch_coo(self, start, batch_ix):
        if self.device_state:       # the same rows in the same order, sliced from the permuted matrix
            from scipy.sparse import vstack
            parts = [self.Xp_csr[a:a + n, :] for a, n in self.schedule.pieces(start)]
            return (parts[0] if len(parts) == 1 else vstack(parts, format="csr")).tocoo()
        return self.Xcsr[batch_ix, :].tocoo()

    def _batch_engine(self, start, batch_ix):
        key = start if self.cache_engines else 0
        eng, Xb = self.engines.get(key), None
        on_device = (self.device_state and self.process_group is None and not self.cache_engines
                     and DEVICE_WINDOW_SLICES and hasattr(self.full, "_h"))
        if on_device and eng is not None:
            # one engine re-laid out every iteration: the window's triples never leave the device
            eng.set_coo(*self._batch_coo_device(start))
        elif eng is None or not self.cache_engines:
            Xb = self._batch_coo(start, batch_ix)
            if eng is not None and eng.ncells != len(batch_ix):     # sharded, not cached: this rank's part varies
                eng.close()
                eng = None
            if eng is None:
                # sharded: the device's t == 0 draw is keyed by (row + offset, gene); shift every rank's rows apart
                opts = {"row_offset": self.lo} if self.process_group is not None else {}
                eng = self._wrap(self.new_engine(len(batch_ix), self.ngenes, **opts))
                eng.set_hyper(*self.hyper)
                self.engines[key] = eng
            eng.set_coo(Xb.row, Xb.col, Xb.data)
        if self.current is None:
            eng.set_state(**self._gene_init)
        elif eng is not self.current:
            eng.copy_gene_state_from(self.current)
        return eng, Xb

    def run(self, t, n, reinit):
        for tt in range(t, t + n):
            start, batch_ix = self.schedule.next()
            if self.process_group is not None:
                # this rank's cells of the window, in window order, as local row indices
                batch_ix = batch_ix[(batch_ix >= self.lo) & (batch_ix < self.lo + self.ncells)] - self.lo
                if batch_ix.shape[0] == 0:
                    raise ValueError("minibatch window %d holds none of this rank's %d cells: batchsize=%d is too "
                                     "small for this sharding" % (start, self.ncells, self.batchsize))
            if self.device_state:
                master = self._full_engine(with_coo=False)
            eng, Xb = self._batch_engine(start, batch_ix)
            if self.device_state:
                at = 0
                for a, m in self.schedule.pieces(start):
                    eng.copy_cell_state_from(master, at, a, m)
                    at += m
            else:
                eng.set_state(theta=(self.theta[0][batch_ix], self.theta[1][batch_ix]),
                              xi=(self.xi[0][batch_ix], self.xi[1][batch_ix]))
            if tt == 0 and reinit and self.process_group is not None:
                eng.step(1, random_phi_seed=self.shared_seed(self.process_group, self.rng), **self.flags)
            elif tt == 0 and reinit:
                if Xb is None:
                    Xb = self._batch_coo(start, batch_ix)
                _random_phi_step(eng, Xb.data, self.nfactors, self.rng, **self.flags)
            else:
                eng.step(1, **self.flags)
            if self.device_state:
                at = 0
                for a, m in self.schedule.pieces(start):
                    master.copy_cell_state_from(eng, a, at, m)
                    at += m
            else:
                st = eng.get_state(("theta", "xi"))
                for k in (0, 1):
                    self.theta[k][batch_ix] = st["theta"][k]
                    self.xi[k][batch_ix] = st["xi"][k]
            self.current = eng

    # -- read-outs -------------------------------------------------------------
    def loss(self):
        """mean negative Poisson llh of ALL of X (the reference's default loss is bound to the
        whole training matrix, scHPF_.py:622-624) under the current state."""
        full = self._full_engine(with_coo=True)
        if not self.device_state:
            full.set_state(theta=tuple(self.theta), xi=tuple(self.xi))
        if self.current is None:
            full.set_state(**self._gene_init)
        else:
            full.copy_gene_state_from(self.current)
        return full.loss()

    def gene_state_engine(self):
        eng = self.current
        return eng if eng is not None and hasattr(eng, "copy_gene_state_from") else None

    def host_state(self):
        if self.device_state and self.full is not None:
            st, order = self.full.get_state(("theta", "xi")), self.schedule.order
            out = {}
            for name in ("theta", "xi"):        # back from the shuffled to the caller's cell order
                pair = []
                for arr in st[name]:
                    back = np.empty_like(arr)
                    back[order] = arr
                    pair.append(back)
                out[name] = tuple(pair)
        else:
            out = {"theta": (self.theta[0].copy(), self.theta[1].copy()),
                   "xi": (self.xi[0].copy(), self.xi[1].copy())}
        if not self.freeze_genes:
            if self.current is None:
                out.update({k: (np.array(v[0], dtype=np.float64), np.array(v[1], dtype=np.float64))
                            for k, v in self._gene_init.items()})
            else:
                out.update(self.current.get_state(("beta", "eta")))
        return out

    def close(self):
        for eng in list(self.engines.values()) + ([self.full] if self.full is not None else []):
            eng.close()
        self.engines, self.full, self.current = {}, None, None
