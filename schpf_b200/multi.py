"""Several GPUs from ONE process: `scHPF(K, device=[0, 1, 2, 3]).fit(X)` needs no launcher
("the multi-GPU handle is single-process, so `fit` stays an ordinary blocking
Python call").

`LocalShardedEngine` has the interface `cavi_loop.FullBatchLoop` expects from an engine.  The
cells are cut into contiguous, nnz-balanced ranges (`engine.shard_bounds_by_nnz`), one
`CaviEngine` per device, beta / eta replicated; every device is driven by a host thread of its
own (the C ABI drops the GIL), which is also what NCCL asks of a process that owns several
devices.  The one exchange step of an iteration -- the all-reduce of
[G*K beta-shape partial sums | K column sums of theta.e_x] -- is

  nccl  (default on CUDA engines) issued by the engines themselves: the communicators are opened
        once per device set from one ncclUniqueId (`schpf_comm_create` on every thread) and
        `schpf_step` then runs the all-reduce under the cells-own sweep, exactly as in the
        one-process-per-GPU path;
  sum   the split-phase calls with the buffers summed through torch tensors (any engine exposing
        `exchange_tensor`; what the CPU test drives with the oracle-backed engine).
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .engine import CaviEngine, shard_bounds_by_nnz

_COMMS = {}      # tuple(devices) -> [communicator per device]; opened once per process


class LocalShardedEngine(object):
    def __init__(self, ncells, ngenes, nfactors, devices, engine_factory=None, exchange=None, **options):
        self.ncells, self.ngenes, self.nfactors = int(ncells), int(ngenes), int(nfactors)
        self.devices = [int(d) for d in devices]
        if len(self.devices) < 1 or len(set(self.devices)) != len(self.devices):
            raise ValueError("devices must be a non-empty list of distinct CUDA ordinals, got %r" % (devices,))
        self.factory = engine_factory or CaviEngine
        self.exchange = exchange or ("nccl" if self.factory is CaviEngine else "sum")
        self.options = dict(options)
        self.device = self.devices[0]
        self.pool = ThreadPoolExecutor(max_workers=len(self.devices))
        self.engines, self.bounds, self.keep, self.nnz = [], None, None, 0
        self._hyper = None

    # -- helpers -----------------------------------------------------------------
    def _each(self, fn, *per_engine_args):
        """fn(rank, engine, *args_of_that_rank) on every device's thread; results in rank order."""
        futs = [self.pool.submit(fn, r, e, *[a[r] for a in per_engine_args]) for r, e in enumerate(self.engines)]
        return [f.result() for f in futs]

    def _rows(self, r):
        return slice(int(self.bounds[r]), int(self.bounds[r + 1]))

    # -- configuration ---------------------------------------------------------------
    def set_option(self, key, value):
        self.options[key] = value
        for e in self.engines:
            e.set_option(key, value)

    def set_coo(self, row, col, data):
        from ._lib import to_host
        row, col, data = (to_host(a) for a in (row, col, data))      # shards are cut on the host
        world = len(self.devices)
        self.nnz = int(row.shape[0])
        self.bounds = shard_bounds_by_nnz(np.bincount(row, minlength=self.ncells), world)
        self.keep = [np.flatnonzero((row >= self.bounds[r]) & (row < self.bounds[r + 1])) for r in range(world)]
        self.close_engines()
        self.engines = [None] * world

        def make(r):
            lo, hi = int(self.bounds[r]), int(self.bounds[r + 1])
            e = self.factory(hi - lo, self.ngenes, self.nfactors, device=self.devices[r], row_offset=lo, **self.options)
            k = self.keep[r]
            e.set_coo(row[k] - lo, col[k], data[k])
            return e
        self.engines = [f.result() for f in [self.pool.submit(make, r) for r in range(world)]]
        if self.exchange == "nccl" and world > 1:
            key = tuple(self.devices)
            if key not in _COMMS:
                uid = CaviEngine.comm_unique_id()
                _COMMS[key] = self._each(lambda r, e: CaviEngine.comm_create(uid, r, world, self.devices[r]))
            self._each(lambda r, e: e.comm_attach(_COMMS[key][r]))
        if self._hyper is not None:
            self.set_hyper(*self._hyper)

    def set_hyper(self, *hyper):
        self._hyper = hyper
        for e in self.engines:
            e.set_hyper(*hyper)

    def set_state(self, theta=None, beta=None, xi=None, eta=None):
        def put(r, e):
            s = self._rows(r)
            cut = lambda p: None if p is None else (np.asarray(p[0])[s], np.asarray(p[1])[s])
            e.set_state(theta=cut(theta), beta=beta, xi=cut(xi), eta=eta)
        self._each(put)

    def get_state(self, which=("theta", "beta", "xi", "eta")):
        parts = self._each(lambda r, e: e.get_state(tuple(n for n in which if n in ("theta", "xi")) +
                                                     (tuple(n for n in which if n in ("beta", "eta")) if r == 0 else ())))
        out = {}
        for n in which:
            if n in ("theta", "xi"):
                out[n] = tuple(np.concatenate([p[n][i] for p in parts]) for i in (0, 1))
            else:
                out[n] = parts[0][n]
        return out

    # -- iterations ------------------------------------------------------------------
    def _sum_exchange(self):
        bufs = [e.exchange_tensor() for e in self.engines]
        total = bufs[0].clone()
        for b in bufs[1:]:
            total += b.to(total.device)
        for b in bufs:
            b.copy_(total.to(b.device))

    def _step_split(self, begin_kwargs_per_rank, flags):
        self._each(lambda r, e, kw: e.step_begin(**kw), begin_kwargs_per_rank)
        if not flags.get("freeze_genes"):
            for e in self.engines:
                e.synchronize()
            self._sum_exchange()
        self._each(lambda r, e: e.step_end(**flags))

    def step(self, n_iters=1, freeze_genes=False, simultaneous=False, cells_first=False):
        if cells_first and not freeze_genes and len(self.engines) > 1:
            raise NotImplementedError("the minibatch update order is not available on several devices")
        flags = dict(freeze_genes=freeze_genes, simultaneous=simultaneous)
        if self.exchange == "nccl" or len(self.engines) == 1 or freeze_genes:
            self._each(lambda r, e: e.step(n_iters, **flags))
            return
        for _ in range(int(n_iters)):
            self._step_split([dict(flags)] * len(self.engines), flags)

    def step_random_phi(self, seed, freeze_genes=False, simultaneous=False, cells_first=False):
        flags = dict(freeze_genes=freeze_genes, simultaneous=simultaneous)
        if self.exchange == "nccl" or len(self.engines) == 1 or freeze_genes:
            self._each(lambda r, e: e.step_random_phi(seed, **flags))
            return
        self._step_split([dict(flags, random_phi_seed=seed)] * len(self.engines), flags)

    def step_with_xphi(self, xphi, freeze_genes=False, simultaneous=False, cells_first=False):
        """t == 0 with the caller's Xphi (rows in the order the triples were given to set_coo)."""
        xphi = np.asarray(xphi, dtype=np.float64)
        flags = dict(freeze_genes=freeze_genes, simultaneous=simultaneous)
        if len(self.engines) == 1 or self.exchange == "nccl" or freeze_genes:
            self._each(lambda r, e: e.step_with_xphi(xphi[self.keep[r]], **flags))
            return
        try:        # engines whose split phase takes the caller's Xphi (the tests' oracle-backed double)
            self._step_split([dict(flags, xphi=xphi[k]) for k in self.keep], flags)
        except TypeError:
            raise NotImplementedError("step_with_xphi needs the engines' own exchange (exchange='nccl')")

    # -- read-outs -------------------------------------------------------------------
    def loss(self):
        if self.exchange == "nccl" and len(self.engines) > 1:
            return self._each(lambda r, e: e.loss())[0]          # all-reduced inside: same number everywhere
        parts = self._each(lambda r, e: e.loss_parts())
        s, n = sum(p[0] for p in parts), sum(p[1] for p in parts)
        return -s / n if n > 0 else 0.0

    def counter(self, what):
        vals = self._each(lambda r, e: e.counter(what))
        return float(sum(vals)) if what in ("nnz", "padded_nnz_cells", "padded_nnz_genes", "layout_bytes") else float(vals[0])

    def synchronize(self):
        self._each(lambda r, e: e.synchronize())

    def close_engines(self):
        for e in self.engines:
            if e is not None:
                e.close()
        self.engines = []

    def close(self):
        self.close_engines()
        self.pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
