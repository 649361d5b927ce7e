"""Host-side handle on the device-resident CAVI engine (C ABI "engine level").

`CaviEngine` owns one GPU's copy of the problem: the COO matrix re-laid out for
the two sweeps, the eight variational arrays, the lookup tables.  `ShardedEngine`
adds the one exchange step per iteration that cell sharding needs (one process
per GPU, `torch.distributed`): an in-place all-reduce of
[G*K beta-shape partial sums | K column sums of theta.e_x].
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import CELLS_FIRST, FREEZE_GENES, PHASE_CELLS, PHASE_GENES, SIMULTANEOUS, as_f64, as_i32, c_i64, c_int, c_dbl, c_u64, c_vp, dptr

STATE_NAMES = ("theta", "beta", "xi", "eta")


def _flags(freeze_genes, simultaneous, cells_first=False):
    """`cells_first` is the minibatch order of scHPF_.py:686-704; `simultaneous` wins over it,
    as in the reference (:666)."""
    return ((FREEZE_GENES if freeze_genes else 0) | (SIMULTANEOUS if simultaneous else 0)
            | (CELLS_FIRST if cells_first and not simultaneous else 0))


class CaviEngine(object):
    """One GPU's CAVI state.  Not thread-safe; one handle = one fit."""

    def __init__(self, ncells, ngenes, nfactors, device=0, stream=None, **options):
        self._h = None
        self._lib = _lib.load()
        self.ncells, self.ngenes, self.nfactors = int(ncells), int(ngenes), int(nfactors)
        self.device = int(device)
        self.nnz = 0
        h = c_vp()
        _lib.check(self._lib.schpf_create(ctypes.byref(h), c_int(self.device), c_i64(self.ncells),
                                          c_i64(self.ngenes), c_int(self.nfactors),
                                          c_vp(int(stream)) if stream else c_vp(None)))
        self._h = h
        for k, v in options.items():
            self.set_option(k, v)

    # -- lifetime ----------------------------------------------------------
    def close(self):
        if self._h is not None:
            self._lib.schpf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- configuration -----------------------------------------------------
    def set_option(self, key, value):
        _lib.check(self._lib.schpf_set_option(self._h, key.encode(), c_i64(int(value))))

    def set_coo(self, row, col, data):
        """COO triples: host arrays (numpy / anything array-like) or CUDA torch tensors."""
        if hasattr(row, "is_cuda") and row.is_cuda:
            import torch
            assert row.dtype == torch.int32 and col.dtype == torch.int32 and data.dtype == torch.int32
            row, col, data = row.contiguous(), col.contiguous(), data.contiguous()
            self.nnz = int(row.numel())
            _lib.check(self._lib.schpf_set_coo_device(self._h, c_vp(row.data_ptr()), c_vp(col.data_ptr()),
                                                      c_vp(data.data_ptr()), c_i64(self.nnz)))
            return
        if hasattr(row, "numpy"):               # CPU torch tensors (pinned or not)
            row, col, data = row.numpy(), col.numpy(), data.numpy()
        row, col, data = as_i32(row), as_i32(col), as_i32(data, "count")
        assert row.shape == col.shape == data.shape and row.ndim == 1
        self.nnz = int(row.shape[0])
        _lib.check(self._lib.schpf_set_coo(self._h, c_vp(row.ctypes.data), c_vp(col.ctypes.data),
                                           c_vp(data.ctypes.data), c_i64(self.nnz)))

    def set_hyper(self, a, ap, bp, c, cp, dp):
        _lib.check(self._lib.schpf_set_hyper(self._h, c_dbl(a), c_dbl(ap), c_dbl(bp),
                                             c_dbl(c), c_dbl(cp), c_dbl(dp)))

    def _shape_of(self, name):
        n = self.ncells if name in ("theta", "xi") else self.ngenes
        return (n, self.nfactors) if name in ("theta", "beta") else (n,)

    def set_state(self, theta=None, beta=None, xi=None, eta=None):
        """Each argument is a (vi_shape, vi_rate) pair of host arrays, or None to keep."""
        args, keep = [], []
        for name, pair in zip(STATE_NAMES, (theta, beta, xi, eta)):
            if pair is None:
                args += [c_vp(None), c_vp(None)]
                continue
            shp, rte = as_f64(pair[0]), as_f64(pair[1])
            if shp.shape != self._shape_of(name) or rte.shape != shp.shape:
                raise ValueError("%s: expected shape %s, got %s / %s"
                                 % (name, self._shape_of(name), shp.shape, rte.shape))
            keep += [shp, rte]
            args += [c_vp(shp.ctypes.data), c_vp(rte.ctypes.data)]
        _lib.check(self._lib.schpf_set_state(self._h, *args))

    def get_state(self, which=STATE_NAMES):
        """-> {name: (vi_shape, vi_rate)} of fresh host fp64 arrays."""
        out, args = {}, []
        for name in STATE_NAMES:
            if name in which:
                shp = np.empty(self._shape_of(name), dtype=np.float64)
                rte = np.empty(self._shape_of(name), dtype=np.float64)
                out[name] = (shp, rte)
                args += [c_vp(shp.ctypes.data), c_vp(rte.ctypes.data)]
            else:
                args += [c_vp(None), c_vp(None)]
        _lib.check(self._lib.schpf_get_state(self._h, *args))
        return out

    # -- iterations --------------------------------------------------------
    def step(self, n_iters=1, freeze_genes=False, simultaneous=False, cells_first=False):
        _lib.check(self._lib.schpf_step(self._h, c_int(int(n_iters)),
                                        c_int(_flags(freeze_genes, simultaneous, cells_first))))

    def step_with_xphi(self, xphi, freeze_genes=False, simultaneous=False, cells_first=False):
        xphi = as_f64(xphi)
        if xphi.shape != (self.nnz, self.nfactors):
            raise ValueError("Xphi must be (nnz, nfactors)")
        _lib.check(self._lib.schpf_step_with_xphi(self._h, dptr(xphi),
                                                  c_int(_flags(freeze_genes, simultaneous, cells_first))))

    def step_random_phi(self, seed, freeze_genes=False, simultaneous=False, cells_first=False):
        _lib.check(self._lib.schpf_step_random_phi(self._h, c_u64(int(seed) & (2 ** 64 - 1)),
                                                   c_int(_flags(freeze_genes, simultaneous, cells_first))))

    def copy_cell_state_from(self, other, dst_row0, src_row0, nrows):
        """theta / xi of `nrows` cells of `other` starting at `src_row0` -> this engine's cells
        starting at `dst_row0`, device to device (same device and nfactors)."""
        _lib.check(self._lib.schpf_copy_cell_state(self._h, c_i64(int(dst_row0)), other._h,
                                                   c_i64(int(src_row0)), c_i64(int(nrows))))

    def copy_gene_state_from(self, other):
        """beta / eta of `other` (same device, ngenes, nfactors) -> this engine, device to device."""
        _lib.check(self._lib.schpf_copy_gene_state(self._h, other._h))

    def step_begin(self, freeze_genes=False, simultaneous=False, random_phi_seed=None, cells_first=False):
        mode, seed = (0, 0) if random_phi_seed is None else (1, int(random_phi_seed) & (2 ** 64 - 1))
        _lib.check(self._lib.schpf_step_begin(self._h, c_int(_flags(freeze_genes, simultaneous, cells_first)),
                                              c_int(mode), c_u64(seed)))

    def step_end(self, freeze_genes=False, simultaneous=False, cells_first=False, phase=None):
        """`phase` splits the minibatch order (cells_first) for callers that reduce the exchange buffer
        themselves: "cells" = theta/xi + the new theta's column sums into the buffer, "genes" = beta/eta."""
        extra = {None: 0, "cells": PHASE_CELLS, "genes": PHASE_GENES}[phase]
        _lib.check(self._lib.schpf_step_end(self._h, c_int(_flags(freeze_genes, simultaneous, cells_first) | extra)))

    # -- cell sharding with the exchange inside the engine --------------------
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL unique id (call on one rank, hand the bytes to all ranks)."""
        buf = ctypes.create_string_buffer(128)
        _lib.check(_lib.load().schpf_comm_unique_id(buf))
        return buf.raw

    @staticmethod
    def comm_create(unique_id, rank, world_size, device):
        """Open this process's NCCL communicator (seconds; keep it and share it between engines)."""
        assert len(unique_id) == 128
        comm = c_vp()
        _lib.check(_lib.load().schpf_comm_create(ctypes.byref(comm), c_int(int(device)), unique_id,
                                                 c_int(int(rank)), c_int(int(world_size))))
        return comm

    def comm_attach(self, comm):
        """From now on step()/loss() include the all-reduce, enqueued in order on this engine's
        stream (collective: all ranks call the same sequence).  The engine does not own `comm`."""
        _lib.check(self._lib.schpf_comm_attach(self._h, comm))

    def exchange_buffer(self):
        """(device pointer, number of doubles) of the per-iteration exchange buffer."""
        p, n = c_vp(), c_i64()
        _lib.check(self._lib.schpf_exchange_buffer(self._h, ctypes.byref(p), ctypes.byref(n)))
        return int(p.value), int(n.value)

    def exchange_tensor(self):
        """The exchange buffer as a zero-copy torch CUDA tensor (for torch.distributed)."""
        import torch
        ptr, n = self.exchange_buffer()

        class _View(object):
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False),
                                        "version": 2, "strides": None}
        return torch.as_tensor(_View(), device=torch.device("cuda", self.device))

    # -- read-outs ---------------------------------------------------------
    def loss(self):
        """loss.py:142-168: mean negative Poisson llh over the nonzeros."""
        v = c_dbl()
        _lib.check(self._lib.schpf_loss(self._h, ctypes.byref(v)))
        return float(v.value)

    def loss_parts(self):
        s, n = c_dbl(), c_i64()
        _lib.check(self._lib.schpf_loss_parts(self._h, ctypes.byref(s), ctypes.byref(n)))
        return float(s.value), int(n.value)

    def llh_pointwise(self):
        out = np.empty((self.nnz,), dtype=np.float64)
        _lib.check(self._lib.schpf_llh_pointwise(self._h, dptr(out)))
        return out

    def xphi(self):
        out = np.empty((self.nnz, self.nfactors), dtype=np.float64)
        _lib.check(self._lib.schpf_xphi_debug(self._h, dptr(out)))
        return out

    def layout_dump(self, side):
        """The device layout of one sweep direction (0: cells own, 1: genes own) decoded back to
        triples in stream order: (owner, other, count) int32 arrays, pads as (-1, -1, 0)."""
        n = c_i64()
        _lib.check(self._lib.schpf_layout_dump(self._h, c_int(int(side)), c_i64(0), None, None, None,
                                               ctypes.byref(n)))
        own, oth, cnt = (np.empty(n.value, dtype=np.int32) for _ in range(3))
        if n.value:
            _lib.check(self._lib.schpf_layout_dump(self._h, c_int(int(side)), n, _lib.iptr(own), _lib.iptr(oth),
                                                   _lib.iptr(cnt), ctypes.byref(n)))
        return own, oth, cnt

    def synchronize(self):
        _lib.check(self._lib.schpf_synchronize(self._h))

    def counter(self, what):
        v = c_dbl()
        _lib.check(self._lib.schpf_counter(self._h, what.encode(), ctypes.byref(v)))
        return float(v.value)


def release_scratch(device=0):
    """Free the work area `set_coo` keeps on `device` between calls (~32 B per nonzero)."""
    _lib.check(_lib.load().schpf_release_scratch(c_int(int(device))))


def shard_bounds_by_nnz(row_counts, world_size):
    """Contiguous cell ranges with (nearly) equal numbers of nonzeros.

    Returns world_size+1 increasing cell indices b with b[0]=0, b[-1]=ncells; rank r owns
    cells [b[r], b[r+1]).  Cut points are where the prefix sum of per-cell counts crosses
    r * nnz / world_size."""
    row_counts = np.asarray(row_counts, dtype=np.int64)
    ncells = row_counts.shape[0]
    prefix = np.concatenate([[0], np.cumsum(row_counts)])
    total = prefix[-1]
    bounds = [0]
    for r in range(1, world_size):
        target = (total * r) // world_size
        b = int(np.searchsorted(prefix, target, side="left"))
        b = min(max(b, bounds[-1]), ncells)
        bounds.append(b)
    bounds.append(ncells)
    return np.asarray(bounds, dtype=np.int64)


def shard_coo_rows(X, rank, world_size):
    """This rank's share of a cells x genes matrix for a cell-sharded fit: the contiguous,
    nnz-balanced range of rows `shard_bounds_by_nnz` assigns to `rank`, re-based to start at 0.

    Returns (X_local: coo_matrix of shape (hi - lo, ngenes), lo, hi).  Every rank must call it
    on the same matrix (or on a matrix with the same per-cell counts)."""
    from scipy.sparse import coo_matrix
    X = X.tocoo()
    bounds = shard_bounds_by_nnz(np.bincount(X.row, minlength=X.shape[0]), world_size)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    keep = (X.row >= lo) & (X.row < hi)
    local = coo_matrix((X.data[keep], (X.row[keep] - lo, X.col[keep])), shape=(hi - lo, X.shape[1]),
                       dtype=X.dtype)
    return local, lo, hi


def bind_host_to_device(device):
    """Pin the calling process to the CPU cores next to CUDA device `device` (NVML's ideal
    affinity).  One process per GPU on a multi-socket host: pinned staging buffers allocated
    afterwards come from the GPU's own NUMA node, so eight ranks uploading their shards at once do
    not all pull from one socket's memory.  Returns the number of cores bound, 0 when NVML is not
    usable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = int(device)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            idx = int(vis.split(",")[idx])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cores = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cores &= set(os.sched_getaffinity(0))
        if not cores:
            return 0
        os.sched_setaffinity(0, cores)
        return len(cores)
    except Exception:
        return 0


class ShardedEngine(object):
    """Cells sharded over the ranks of a torch.distributed process group.

    `local` is this rank's engine (a CaviEngine on its GPU) holding only this rank's cells
    (rows re-based to start at 0) and a full replica of beta / eta.  Every iteration:
    local sweeps -> all-reduce(sum) of the exchange buffer -> local finalisation; beta and
    eta stay bit-identical on all ranks because every rank applies the same reduced
    buffer.  With frozen genes there is nothing to exchange.

    Two transports for the one exchange step:
      * native (default on NCCL groups): torch.distributed only bootstraps -- rank 0's
        ncclUniqueId is broadcast and each engine opens its own communicator; the
        all-reduce is then issued by the engine in order on its own stream, so an
        iteration is one C call with no stream hand-offs (measured: ~100 us per iteration
        saved against the torch path, tools/diag_allreduce.py);
      * torch: `torch.distributed.all_reduce` on a zero-copy view of the buffer (any
        backend; what the gloo CPU tests exercise).
    """

    def __init__(self, local, group=None, native=None):
        import torch.distributed as dist
        self.local = local
        self.group = group
        self._dist = dist
        self._buf = None
        if native is None:
            native = hasattr(local, "comm_attach") and dist.get_backend(group) == "nccl"
        self.native = bool(native)
        if self.native:
            local.comm_attach(self._communicator(local, group))

    _comms = {}     # (group, device) -> communicator: created once per process, like torch's own

    @classmethod
    def _communicator(cls, local, group):
        import torch.distributed as dist
        key = (group, local.device)
        if key not in cls._comms:
            rank, world = dist.get_rank(group), dist.get_world_size(group)
            box = [local.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0,
                                       group=group)
            cls._comms[key] = local.comm_create(box[0], rank, world, local.device)
        return cls._comms[key]

    def _exchange(self):
        if self._buf is None:
            self._buf = self.local.exchange_tensor()
        self._dist.all_reduce(self._buf, op=self._dist.ReduceOp.SUM, group=self.group)

    def step(self, n_iters=1, freeze_genes=False, simultaneous=False, random_phi_seed=None, cells_first=False):
        """`cells_first` (the minibatch order, scHPF_.py:686-704): theta/xi of this rank's cells of the
        batch first, THEN the exchange (beta's rate sums the new theta of the whole batch), then beta/eta."""
        n_iters = int(n_iters)
        if self.native:
            if random_phi_seed is not None and n_iters > 0:
                self.local.step_random_phi(random_phi_seed, freeze_genes, simultaneous, cells_first)
                n_iters -= 1
            if n_iters > 0:
                self.local.step(n_iters, freeze_genes, simultaneous, cells_first)
            return
        late = cells_first and not simultaneous and not freeze_genes
        for i in range(n_iters):
            self.local.step_begin(freeze_genes, simultaneous, random_phi_seed if i == 0 else None,
                                  cells_first=cells_first)
            if late:
                self.local.step_end(freeze_genes, simultaneous, cells_first=True, phase="cells")
                self._exchange()
                self.local.step_end(freeze_genes, simultaneous, cells_first=True, phase="genes")
                continue
            if not freeze_genes:
                self._exchange()
            self.local.step_end(freeze_genes, simultaneous, cells_first=cells_first)

    def close(self):
        self.local.close()

    def loss(self):
        if self.native:
            return self.local.loss()
        import torch
        s, n = self.local.loss_parts()
        dev = self._buf.device if self._buf is not None else self.local.exchange_tensor().device
        t = torch.tensor([s, float(n)], dtype=torch.float64, device=dev)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        s, n = t.tolist()
        return -s / n if n > 0 else 0.0

    def __getattr__(self, name):        # everything else is per-rank
        return getattr(self.local, name)
