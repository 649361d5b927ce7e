"""An engine with CaviEngine's interface whose arithmetic is the CPU oracle.

TEST INFRASTRUCTURE: lets the host-side logic of schpf_b200 (the estimator's
loop / convergence rules, the sharded exchange step) run on a machine without a
GPU, and gives the GPU tests an independent implementation to compare with.
The product never imports this.
"""
import numpy as np

from oracle import hpf_numpy as onp


class OracleEngine(object):
    def __init__(self, ncells, ngenes, nfactors, device=0, stream=None, **options):
        self.ncells, self.ngenes, self.nfactors = int(ncells), int(ngenes), int(nfactors)
        self.device = int(device)
        self.row_offset = int(options.get("row_offset", 0))
        self.nnz = 0
        self.st = None
        self._exch = None
        self._xphi = None

    def close(self):
        pass

    def set_option(self, key, value):
        pass

    def set_coo(self, row, col, data):
        self.row = np.asarray(row, dtype=np.int32)
        self.col = np.asarray(col, dtype=np.int32)
        self.data = np.asarray(data, dtype=np.int32)
        self.nnz = self.row.shape[0]

    def set_hyper(self, a, ap, bp, c, cp, dp):
        self.a, self.ap, self.bp, self.c, self.cp, self.dp = a, ap, bp, c, cp, dp

    def set_state(self, theta=None, beta=None, xi=None, eta=None):
        if self.st is None:
            K = self.nfactors
            z = lambda *s: np.ones(s)
            self.st = onp.State(z(self.ncells, K), z(self.ncells, K), z(self.ngenes, K), z(self.ngenes, K),
                                z(self.ncells), z(self.ncells), z(self.ngenes), z(self.ngenes))
        f = lambda a: np.array(a, dtype=np.float64, copy=True)
        if theta is not None:
            self.st.theta_shp, self.st.theta_rte = f(theta[0]), f(theta[1])
        if beta is not None:
            self.st.beta_shp, self.st.beta_rte = f(beta[0]), f(beta[1])
        if xi is not None:
            self.st.xi_shp, self.st.xi_rte = f(xi[0]), f(xi[1])
        if eta is not None:
            self.st.eta_shp, self.st.eta_rte = f(eta[0]), f(eta[1])

    def get_state(self, which=("theta", "beta", "xi", "eta")):
        s = self.st
        full = {"theta": (s.theta_shp, s.theta_rte), "beta": (s.beta_shp, s.beta_rte),
                "xi": (s.xi_shp, s.xi_rte), "eta": (s.eta_shp, s.eta_rte)}
        return {k: (v[0].copy(), v[1].copy()) for k, v in full.items() if k in which}

    # ---- whole iterations --------------------------------------------------
    def _iter(self, freeze_genes, simultaneous, Xphi=None, cells_first=False):
        onp.cavi_iteration(self.data, self.row, self.col, self.st, self.a, self.bp, self.c, self.dp,
                           freeze_genes=freeze_genes, Xphi=Xphi,
                           beta_theta_simultaneous=simultaneous, batched=cells_first)

    def step(self, n_iters=1, freeze_genes=False, simultaneous=False, cells_first=False):
        for _ in range(int(n_iters)):
            self._iter(freeze_genes, simultaneous, cells_first=cells_first)

    def step_with_xphi(self, xphi, freeze_genes=False, simultaneous=False, cells_first=False):
        self._iter(freeze_genes, simultaneous, Xphi=np.asarray(xphi, dtype=np.float64), cells_first=cells_first)

    def step_random_phi(self, seed, freeze_genes=False, simultaneous=False, cells_first=False):
        rng = np.random.default_rng(seed)
        phi = rng.dirichlet(np.ones(self.nfactors), self.nnz)
        self._iter(freeze_genes, simultaneous, Xphi=self.data[:, None] * phi, cells_first=cells_first)

    def copy_cell_state_from(self, other, dst_row0, src_row0, nrows):
        if self.st is None:
            self.set_state()
        d, s_, n = int(dst_row0), int(src_row0), int(nrows)
        assert 0 <= d and d + n <= self.ncells and 0 <= s_ and s_ + n <= other.ncells
        for name in ("theta_shp", "theta_rte", "xi_shp", "xi_rte"):
            getattr(self.st, name)[d:d + n] = getattr(other.st, name)[s_:s_ + n]

    def copy_gene_state_from(self, other):
        o = other.st
        self.set_state(beta=(o.beta_shp, o.beta_rte), eta=(o.eta_shp, o.eta_rte))

    # ---- split phase (cell sharding) ----------------------------------------
    def step_begin(self, freeze_genes=False, simultaneous=False, random_phi_seed=None, xphi=None, cells_first=False):
        s, K = self.st, self.nfactors
        if xphi is not None:
            self._xphi = np.asarray(xphi, dtype=np.float64)
        elif random_phi_seed is not None:
            # t == 0: y * Dirichlet(1_K), a stream of this shard's own (keyed like the device's by the shard offset)
            rng = np.random.default_rng([int(random_phi_seed), self.row_offset])
            self._xphi = self.data[:, None] * rng.dirichlet(np.ones(K), self.nnz)
        else:
            self._xphi = onp.compute_Xphi_data(self.data, self.row, self.col, s.theta_shp, s.theta_rte,
                                               s.beta_shp, s.beta_rte)
        if not freeze_genes:
            buf = self.exchange_tensor().numpy()
            part = onp.compute_loading_shape_update(self._xphi, self.col, self.ngenes, 0.0)
            buf[:self.ngenes * K] = part.ravel()
            buf[self.ngenes * K:] = (s.theta_shp / s.theta_rte).sum(0)

    def exchange_tensor(self):
        import torch
        if self._exch is None:
            self._exch = torch.zeros(self.ngenes * self.nfactors + self.nfactors, dtype=torch.float64)
        return self._exch

    def _cells_update(self):
        s = self.st
        s.theta_shp = onp.compute_loading_shape_update(self._xphi, self.row, self.ncells, self.a)
        s.theta_rte = onp.compute_loading_rate_update(s.xi_shp, s.xi_rte, s.beta_shp, s.beta_rte)
        s.xi_rte = self.bp + (s.theta_shp / s.theta_rte).sum(1)

    def _genes_update(self):
        s, K = self.st, self.nfactors
        buf = self.exchange_tensor().numpy()
        s.beta_shp = self.c + buf[:self.ngenes * K].reshape(self.ngenes, K)
        s.beta_rte = (s.eta_shp / s.eta_rte)[:, None] + buf[self.ngenes * K:][None, :]
        s.eta_rte = self.dp + (s.beta_shp / s.beta_rte).sum(1)

    def step_end(self, freeze_genes=False, simultaneous=False, cells_first=False, phase=None):
        s, K = self.st, self.nfactors
        assert not simultaneous
        if cells_first:
            # minibatch order (scHPF_.py:686-704): cells first, from the OLD beta; beta's rate sums the NEW theta
            if phase in (None, "cells"):
                self._cells_update()
                if not freeze_genes:
                    self.exchange_tensor().numpy()[self.ngenes * K:] = (s.theta_shp / s.theta_rte).sum(0)
            if phase in (None, "genes") and not freeze_genes:
                self._genes_update()
            return
        if not freeze_genes:
            buf = self.exchange_tensor().numpy()
            s.beta_shp = self.c + buf[:self.ngenes * K].reshape(self.ngenes, K)
            s.beta_rte = (s.eta_shp / s.eta_rte)[:, None] + buf[self.ngenes * K:][None, :]
            s.eta_rte = self.dp + (s.beta_shp / s.beta_rte).sum(1)
        s.theta_shp = onp.compute_loading_shape_update(self._xphi, self.row, self.ncells, self.a)
        s.theta_rte = onp.compute_loading_rate_update(s.xi_shp, s.xi_rte, s.beta_shp, s.beta_rte)
        s.xi_rte = self.bp + (s.theta_shp / s.theta_rte).sum(1)

    # ---- read-outs -----------------------------------------------------------
    def llh_pointwise(self):
        s = self.st
        return onp.compute_pois_llh(self.data, self.row, self.col, s.theta_shp, s.theta_rte,
                                    s.beta_shp, s.beta_rte)

    def loss_parts(self):
        return float(self.llh_pointwise().sum()), int(self.nnz)

    def loss(self):
        return float(np.mean(-self.llh_pointwise()))

    def xphi(self):
        s = self.st
        return onp.compute_Xphi_data(self.data, self.row, self.col, s.theta_shp, s.theta_rte,
                                     s.beta_shp, s.beta_rte)

    def synchronize(self):
        pass
