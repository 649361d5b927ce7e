"""The cell-sharded path on CPU: two gloo ranks run schpf_b200.engine.ShardedEngine
(the product's exchange logic) around oracle-backed local engines and must reproduce
the unsharded golden run of the reference: same beta/eta on both ranks, theta/xi equal
to the corresponding rows."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, GOLDEN


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir, freeze):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from schpf_b200.engine import ShardedEngine, shard_bounds_by_nnz
    from oracle_engine import OracleEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = dict(np.load(os.path.join(GOLDEN, "cavi_cfg1.npz")))
    C, G = (int(v) for v in g["shape"])
    K = g["init_theta_shp"].shape[1]
    counts = np.bincount(g["row"], minlength=C)
    b = shard_bounds_by_nnz(counts, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    keep = (g["row"] >= lo) & (g["row"] < hi)
    local = OracleEngine(hi - lo, G, K)
    local.set_coo(g["row"][keep] - lo, g["col"][keep], g["data"][keep])
    local.set_hyper(*[float(g[k]) for k in ("a", "ap", "bp", "c", "cp", "dp")])
    pre = "it50_" if freeze else "init_"
    xi_shp = np.full(hi - lo, float(g["ap"]) + K * float(g["a"]))
    eta_shp = g[pre + "eta_shp"] if freeze else np.full(G, float(g["cp"]) + K * float(g["c"]))
    local.set_state(theta=(g["init_theta_shp"][lo:hi], g["init_theta_rte"][lo:hi]),
                    beta=(g[pre + "beta_shp"], g[pre + "beta_rte"]),
                    xi=(xi_shp, g["init_xi_rte"][lo:hi]), eta=(eta_shp, g[pre + "eta_rte"]))
    eng = ShardedEngine(local, None)
    loss = []
    for t in range(10):
        eng.step(1, freeze_genes=freeze)
        if t % 3 == 0:
            loss.append(eng.loss())
    st = local.get_state()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, loss=np.array(loss),
             **{n + s: st[n][i] for n in st for i, s in ((0, "_shp"), (1, "_rte"))})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("freeze", [False, True])
def test_two_rank_gloo_matches_unsharded(tmp_path, freeze):
    import torch.multiprocessing as mp
    from oracle import hpf_numpy as onp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), freeze), nprocs=world, join=True)
    g = dict(np.load(os.path.join(GOLDEN, "cavi_cfg1.npz")))
    r = [dict(np.load(str(tmp_path / ("rank%d.npz" % k)))) for k in range(world)]
    assert int(r[0]["lo"]) == 0 and int(r[0]["hi"]) == int(r[1]["lo"]) and int(r[1]["hi"]) == 1000
    # nnz-balanced cut
    counts = np.bincount(g["row"], minlength=1000)
    assert abs(counts[:int(r[0]["hi"])].sum() - counts[int(r[0]["hi"]):].sum()) <= 2 * counts.max()
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.abs(b)))
    if not freeze:
        # replicas of the gene side are bit-identical across ranks and equal the reference's run
        for n in ("beta_shp", "beta_rte", "eta_rte"):
            assert np.array_equal(r[0][n], r[1][n])
            assert rel(r[0][n], g["it10_" + n]) < 1e-11
        for n in ("theta_shp", "theta_rte", "xi_rte"):
            assert rel(np.concatenate([r[0][n], r[1][n]]), g["it10_" + n]) < 1e-11
        assert np.allclose(r[0]["loss"], g["it10_loss"], rtol=1e-12) and np.array_equal(r[0]["loss"], r[1]["loss"])
    else:
        # projection: genes untouched, cells equal to an unsharded oracle projection
        assert np.array_equal(r[0]["beta_shp"], g["it50_beta_shp"]) and np.array_equal(r[1]["eta_rte"], g["it50_eta_rte"])
        st = onp.State(g["init_theta_shp"], g["init_theta_rte"], g["it50_beta_shp"], g["it50_beta_rte"],
                       g["init_xi_shp"], g["init_xi_rte"], g["it50_eta_shp"], g["it50_eta_rte"])
        hyp = [float(g[k]) for k in ("a", "ap", "bp", "c", "cp", "dp")]
        onp.cavi_run(g["data"], g["row"], g["col"], st, *hyp, 10, freeze_genes=True)
        assert rel(np.concatenate([r[0]["theta_shp"], r[1]["theta_shp"]]), st.theta_shp) < 1e-12
        assert rel(np.concatenate([r[0]["xi_rte"], r[1]["xi_rte"]]), st.xi_rte) < 1e-12


def _estimator_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from scipy.sparse import coo_matrix
    from schpf_b200 import scHPF, HPF_Gamma
    from schpf_b200 import scHPF_ as shell
    from schpf_b200.engine import shard_coo_rows
    from oracle_engine import OracleEngine
    shell._engine_factory = OracleEngine               # host logic under test, arithmetic from the oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = dict(np.load(os.path.join(GOLDEN, "cavi_cfg1.npz")))
    C, G = (int(v) for v in g["shape"])
    X, lo, hi = shard_coo_rows(coo_matrix((g["data"], (g["row"], g["col"])), shape=(C, G)), rank, world)
    assert X.shape == (hi - lo, G)
    gam = lambda n, sl: HPF_Gamma(g["init_" + n + "_shp"][sl].copy(), g["init_" + n + "_rte"][sl].copy())
    # rank 1 is handed garbage for the gene side: it must end up with rank 0's
    scale = 1.0 if rank == 0 else 3.0
    beta = HPF_Gamma(g["init_beta_shp"] * scale, g["init_beta_rte"] * scale)
    eta = HPF_Gamma(g["init_eta_shp"] * scale, g["init_eta_rte"] * scale)
    m = scHPF(5, verbose=False, xi=gam("xi", slice(lo, hi)), theta=gam("theta", slice(lo, hi)), eta=eta, beta=beta)
    m.fit(X, reinit=False, min_iter=10, max_iter=10, check_freq=3, process_group=dist.group.WORLD)
    full = m.gather_cells(dist.group.WORLD)
    assert full.ncells == C and m.ncells == hi - lo and full.beta == m.beta
    np.savez(os.path.join(out_dir, "est%d.npz" % rank), lo=lo, hi=hi, bp=m.bp, dp=m.dp, loss=np.array(m.loss),
             theta_shp=m.theta.vi_shape, xi_rte=m.xi.vi_rate, beta_shp=m.beta.vi_shape, eta_rte=m.eta.vi_rate,
             full_theta_shp=full.theta.vi_shape, full_xi_rte=full.xi.vi_rate)
    dist.barrier()
    dist.destroy_process_group()


def test_estimator_fit_with_process_group(tmp_path):
    """scHPF.fit(X_shard, process_group=...) on two gloo ranks: global b', d', gene side taken
    from rank 0, global loss, and the same result as the reference's unsharded run."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_estimator_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = dict(np.load(os.path.join(GOLDEN, "cavi_cfg1.npz")))
    r = [dict(np.load(str(tmp_path / ("est%d.npz" % k)))) for k in range(world)]
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.abs(b)))
    for k in range(world):
        assert abs(float(r[k]["bp"]) / float(g["bp"]) - 1) < 1e-12       # pooled moments, not np.var: ~1 ulp
        assert abs(float(r[k]["dp"]) / float(g["dp"]) - 1) < 1e-12
    assert np.array_equal(r[0]["beta_shp"], r[1]["beta_shp"]) and np.array_equal(r[0]["loss"], r[1]["loss"])
    assert rel(r[0]["beta_shp"], g["it10_beta_shp"]) < 1e-10
    assert rel(np.concatenate([r[0]["theta_shp"], r[1]["theta_shp"]]), g["it10_theta_shp"]) < 1e-10
    assert rel(np.concatenate([r[0]["xi_rte"], r[1]["xi_rte"]]), g["it10_xi_rte"]) < 1e-10
    assert np.allclose(r[0]["loss"], g["it10_loss"], rtol=1e-10)
    # gather_cells: every rank ends up with all cells, in rank order
    for k in range(world):
        assert np.array_equal(r[k]["full_theta_shp"], np.concatenate([r[0]["theta_shp"], r[1]["theta_shp"]]))
        assert np.array_equal(r[k]["full_xi_rte"], np.concatenate([r[0]["xi_rte"], r[1]["xi_rte"]]))


def _custom_loss_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from scipy.sparse import coo_matrix
    from schpf_b200 import scHPF
    from schpf_b200 import scHPF_ as shell
    from schpf_b200 import loss as ls
    from schpf_b200.engine import shard_coo_rows
    from oracle_engine import OracleEngine
    from oracle import hpf_numpy as onp
    shell._engine_factory = OracleEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = dict(np.load(os.path.join(GOLDEN, "cavi_cfg1.npz")))
    C, G = (int(v) for v in g["shape"])
    X, lo, hi = shard_coo_rows(coo_matrix((g["data"], (g["row"], g["col"])), shape=(C, G)), rank, world)
    offsets = []

    class Spy(OracleEngine):
        def __init__(self, *a, **kw):
            offsets.append(kw.get("row_offset"))
            OracleEngine.__init__(self, *a, **kw)
    shell._engine_factory = Spy

    def local_loss(*, theta, beta, **kw):        # each rank sees ITS cells only: values differ between ranks
        return float(np.mean(-onp.compute_pois_llh(X.data, X.row, X.col, theta.vi_shape, theta.vi_rate,
                                                   beta.vi_shape, beta.vi_rate))) * (1.0 + 0.5 * rank)
    np.random.seed(3 + rank)                     # different streams: the shared seed must come from rank 0
    m = scHPF(5, verbose=False, epsilon=5.0)     # (the estimator's epsilon is the one _fit reads, as in the reference)
    # reinit=True: random initialisation and the t == 0 random-phi step under a process group; a
    # rank-local custom loss with loose stopping rules: every rank must stop at the same iteration
    m.fit(X, min_iter=2, max_iter=40, check_freq=2, loss_function=local_loss, process_group=dist.group.WORLD)
    np.savez(os.path.join(out_dir, "cl%d.npz" % rank), loss=np.array(m.loss), beta_shp=m.beta.vi_shape,
             offset=np.array(offsets[0]), lo=lo)
    dist.barrier()
    dist.destroy_process_group()


def test_custom_loss_and_reinit_under_a_process_group(tmp_path):
    """a rank-local custom loss must not let ranks stop at different iterations (the
    values are averaged over ranks before the stopping rules), the engines get the global cell
    offset of their shard, and reinit=True (random-phi first iteration) works when sharded."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_custom_loss_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [dict(np.load(str(tmp_path / ("cl%d.npz" % k)))) for k in range(world)]
    assert np.array_equal(r[0]["loss"], r[1]["loss"]) and len(r[0]["loss"]) >= 2
    assert len(r[0]["loss"]) < 20                                  # stopped early, together
    assert np.array_equal(r[0]["beta_shp"], r[1]["beta_shp"])      # replicas still identical
    assert int(r[0]["offset"]) == 0 and int(r[1]["offset"]) == int(r[1]["lo"]) > 0


def _minibatch_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from scipy.sparse import coo_matrix
    from schpf_b200 import scHPF, HPF_Gamma
    from schpf_b200 import scHPF_ as shell
    from schpf_b200.engine import shard_coo_rows
    from oracle_engine import OracleEngine
    shell._engine_factory = OracleEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = dict(np.load(os.path.join(GOLDEN, "minibatch_small.npz")))
    C, G = (int(v) for v in g["shape"])
    X, lo, hi = shard_coo_rows(coo_matrix((g["data"], (g["row"], g["col"])), shape=(C, G)), rank, world)
    gam = lambda n, rows: HPF_Gamma(g["B_init_%s_shp" % n][rows].copy(), g["B_init_%s_rte" % n][rows].copy())
    cells, genes = slice(lo, hi), slice(None)
    m = scHPF(3, verbose=False, bp=float(g["bp"]), dp=float(g["dp"]), xi=gam("xi", cells), theta=gam("theta", cells),
              eta=gam("eta", genes), beta=gam("beta", genes))
    np.random.seed(int(g["C_seed"]) if rank == 0 else 999)       # the one shuffle comes from rank 0's stream
    m.fit(X, reinit=False, batchsize=int(g["C_batchsize"]), min_iter=int(g["C_iters"]), max_iter=int(g["C_iters"]),
          check_freq=int(g["C_check_freq"]), process_group=dist.group.WORLD)
    # reinit=True: random initialisation, and the batch's t == 0 random-phi step with one seed for all ranks
    np.random.seed(50 + rank)
    m2 = scHPF(3, verbose=False)
    m2.fit(X, batchsize=120, min_iter=4, max_iter=4, check_freq=1, process_group=dist.group.WORLD)
    np.savez(os.path.join(out_dir, "mb%d.npz" % rank), lo=lo, hi=hi, loss=np.array(m.loss),
             theta_shp=m.theta.vi_shape, theta_rte=m.theta.vi_rate, xi_rte=m.xi.vi_rate,
             beta_shp=m.beta.vi_shape, beta_rte=m.beta.vi_rate, eta_rte=m.eta.vi_rate,
             loss2=np.array(m2.loss), beta2=m2.beta.vi_shape, theta2=m2.theta.vi_shape)
    dist.barrier()
    dist.destroy_process_group()


def test_minibatch_fit_with_process_group_reproduces_the_reference(tmp_path):
    """scHPF.fit(X_shard, batchsize=..., process_group=...): the windows run over ALL cells (rank 0's
    shuffle), every rank updates the cells of a window it owns, and the one exchange of the iteration
    follows the cell update.  Two gloo ranks reproduce case C of minibatch_small.npz -- a seeded
    minibatch run of the REAL, single-process reference (scHPF_.py:626-631, 642-650, 686-704)."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_minibatch_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = dict(np.load(os.path.join(GOLDEN, "minibatch_small.npz")))
    r = [dict(np.load(str(tmp_path / ("mb%d.npz" % k)))) for k in range(world)]
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.abs(b)))
    assert int(r[0]["lo"]) == 0 and int(r[0]["hi"]) == int(r[1]["lo"]) and int(r[1]["hi"]) == int(g["shape"][0])
    for n in ("beta_shp", "beta_rte", "eta_rte", "loss"):
        assert np.array_equal(r[0][n], r[1][n]), n                       # replicas identical
    assert rel(r[0]["beta_shp"], g["C_beta_shp"]) < 1e-10 and rel(r[0]["beta_rte"], g["C_beta_rte"]) < 1e-10
    assert rel(r[0]["eta_rte"], g["C_eta_rte"]) < 1e-10
    for n in ("theta_shp", "theta_rte", "xi_rte"):
        assert rel(np.concatenate([r[0][n], r[1][n]]), g["C_" + n]) < 1e-10, n
    assert np.allclose(r[0]["loss"], g["C_loss"], rtol=1e-10)
    # the reinit=True fit: same gene side and loss on both ranks, everything finite and positive
    assert np.array_equal(r[0]["beta2"], r[1]["beta2"]) and np.array_equal(r[0]["loss2"], r[1]["loss2"])
    assert len(r[0]["loss2"]) == 4 and np.all(np.isfinite(r[0]["loss2"]))
    assert all(np.all(r[k]["theta2"] > 0) for k in range(world)) and np.all(r[0]["beta2"] > 0)
