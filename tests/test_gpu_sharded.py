"""Cell sharding on real GPUs (needs >= 2 devices; skipped otherwise): one process per GPU,
NCCL.  All transports of schpf_b200.engine.ShardedEngine -- the engine's own ncclAllReduce
(overlapped with the cells-own sweep, or in order on its stream) and
torch.distributed.all_reduce on the zero-copy buffer view -- must reproduce the unsharded
golden run of the reference."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, GOLDEN, has_cuda

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir, mode):
    native = mode != "torch"
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from schpf_b200.engine import CaviEngine, ShardedEngine, shard_bounds_by_nnz
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    g = dict(np.load(os.path.join(GOLDEN, "cavi_cfg1.npz")))
    C, G = (int(v) for v in g["shape"])
    K = g["init_theta_shp"].shape[1]
    b = shard_bounds_by_nnz(np.bincount(g["row"], minlength=C), world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    keep = (g["row"] >= lo) & (g["row"] < hi)
    local = CaviEngine(hi - lo, G, K, device=rank, row_offset=lo,
                       overlap_exchange=0 if mode == "native-in-stream" else 1)
    local.set_coo(g["row"][keep] - lo, g["col"][keep], g["data"][keep])
    local.set_hyper(*[float(g[k]) for k in ("a", "ap", "bp", "c", "cp", "dp")])
    local.set_state(theta=(g["init_theta_shp"][lo:hi], g["init_theta_rte"][lo:hi]),
                    beta=(g["init_beta_shp"], g["init_beta_rte"]),
                    xi=(np.full(hi - lo, float(g["ap"]) + K * float(g["a"])), g["init_xi_rte"][lo:hi]),
                    eta=(np.full(G, float(g["cp"]) + K * float(g["c"])), g["init_eta_rte"]))
    eng = ShardedEngine(local, None, native=native)
    assert eng.native == native
    loss = []
    for t in range(10):
        eng.step(1)
        if t % 3 == 0:
            loss.append(eng.loss())
    st = local.get_state()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, loss=np.array(loss),
             **{n + s: st[n][i] for n in st for i, s in ((0, "_shp"), (1, "_rte"))})
    dist.barrier()
    local.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("mode", ["native-overlapped", "native-in-stream", "torch"])
def test_two_gpu_nccl_matches_unsharded(tmp_path, mode):
    """native-overlapped: the engine's all-reduce on a second stream under the cells-own sweep
    (default); native-in-stream: the same call in order on the engine's stream; torch:
    torch.distributed.all_reduce on the zero-copy view of the exchange buffer."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), mode), nprocs=world, join=True)
    g = dict(np.load(os.path.join(GOLDEN, "cavi_cfg1.npz")))
    r = [dict(np.load(str(tmp_path / ("rank%d.npz" % k)))) for k in range(world)]
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.abs(b)))
    for n in ("beta_shp", "beta_rte", "eta_rte"):
        assert np.array_equal(r[0][n], r[1][n])                 # replicas bit-identical
        assert rel(r[0][n], g["it10_" + n]) < 1e-9
    for n in ("theta_shp", "theta_rte", "xi_rte"):
        assert rel(np.concatenate([r[0][n], r[1][n]]), g["it10_" + n]) < 1e-9
    assert np.allclose(r[0]["loss"], g["it10_loss"], rtol=1e-11)
    assert np.array_equal(r[0]["loss"], r[1]["loss"])


def _estimator_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from scipy.sparse import coo_matrix
    from schpf_b200 import scHPF, HPF_Gamma
    from schpf_b200.engine import shard_coo_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    # torch's current device is deliberately NOT this rank's GPU: the estimator must use model.device
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    g = dict(np.load(os.path.join(GOLDEN, "cavi_cfg1.npz")))
    C, G = (int(v) for v in g["shape"])
    X, lo, hi = shard_coo_rows(coo_matrix((g["data"], (g["row"], g["col"])), shape=(C, G)), rank, world)
    gam = lambda n, sl: HPF_Gamma(g["init_" + n + "_shp"][sl].copy(), g["init_" + n + "_rte"][sl].copy())
    scale = 1.0 if rank == 0 else 3.0           # rank 1 is handed garbage for the gene side
    beta = HPF_Gamma(g["init_beta_shp"] * scale, g["init_beta_rte"] * scale)
    eta = HPF_Gamma(g["init_eta_shp"] * scale, g["init_eta_rte"] * scale)
    m = scHPF(5, verbose=False, device=rank, xi=gam("xi", slice(lo, hi)), theta=gam("theta", slice(lo, hi)),
              eta=eta, beta=beta)
    m.fit(X, reinit=False, min_iter=10, max_iter=10, check_freq=3, process_group=dist.group.WORLD)
    full = m.gather_cells(dist.group.WORLD)
    # a second fit with the default reinit=True: random initialisation + the device's t == 0 draw,
    # keyed by the global cell index (row_offset) and one seed shared by the ranks
    np.random.seed(100 + rank)
    m2 = scHPF(5, verbose=False, device=rank)
    import schpf_b200.cavi_loop as cl
    cl.HOST_DIRICHLET_LIMIT = 0                 # force the device generator
    m2.fit(X, min_iter=1, max_iter=1, check_freq=1, process_group=dist.group.WORLD)
    np.savez(os.path.join(out_dir, "est%d.npz" % rank), lo=lo, hi=hi, bp=m.bp, dp=m.dp, loss=np.array(m.loss),
             theta_shp=m.theta.vi_shape, xi_rte=m.xi.vi_rate, beta_shp=m.beta.vi_shape, eta_rte=m.eta.vi_rate,
             full_theta_shp=full.theta.vi_shape, full_xi_rte=full.xi.vi_rate,
             t0_theta_shp=m2.theta.vi_shape, t0_beta_shp=m2.beta.vi_shape, t0_a=m2.a, t0_c=m2.c,
             local_mass=float(X.data.sum()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_two_gpu_estimator_fit_with_process_group(tmp_path):
    """scHPF.fit(X_shard, process_group=) over NCCL on two GPUs + gather_cells: the reference's
    unsharded golden run, bit-identical gene side on both ranks; and a reinit=True first iteration
    (device random phi) conserving the counts' mass on both sides."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_estimator_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = dict(np.load(os.path.join(GOLDEN, "cavi_cfg1.npz")))
    r = [dict(np.load(str(tmp_path / ("est%d.npz" % k)))) for k in range(world)]
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.abs(b)))
    for k in range(world):
        assert abs(float(r[k]["bp"]) / float(g["bp"]) - 1) < 1e-12
        assert abs(float(r[k]["dp"]) / float(g["dp"]) - 1) < 1e-12
    assert np.array_equal(r[0]["beta_shp"], r[1]["beta_shp"]) and np.array_equal(r[0]["loss"], r[1]["loss"])
    assert rel(r[0]["beta_shp"], g["it10_beta_shp"]) < 1e-9
    assert rel(np.concatenate([r[0]["theta_shp"], r[1]["theta_shp"]]), g["it10_theta_shp"]) < 1e-9
    assert rel(np.concatenate([r[0]["xi_rte"], r[1]["xi_rte"]]), g["it10_xi_rte"]) < 1e-9
    assert np.allclose(r[0]["loss"], g["it10_loss"], rtol=1e-10)
    for k in range(world):
        assert np.array_equal(r[k]["full_theta_shp"], np.concatenate([r[0]["theta_shp"], r[1]["theta_shp"]]))
    # t == 0 with reinit: theta shape - a sums to this shard's counts, beta shape - c to all counts
    total = float(r[0]["local_mass"] + r[1]["local_mass"])
    for k in range(world):
        assert abs((r[k]["t0_theta_shp"] - float(r[k]["t0_a"])).sum() / float(r[k]["local_mass"]) - 1) < 1e-10
        assert abs((r[k]["t0_beta_shp"] - float(r[k]["t0_c"])).sum() / total - 1) < 1e-10
    assert np.array_equal(r[0]["t0_beta_shp"], r[1]["t0_beta_shp"])


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_single_process_two_gpu_fit_matches_unsharded():
    """scHPF(K, device=[0, 1]).fit(X): both GPUs from this one process (schpf_b200/multi.py; one host
    thread and one NCCL communicator per device, the engines' own all-reduce): the reference's
    unsharded golden run."""
    from scipy.sparse import coo_matrix
    from schpf_b200 import scHPF, HPF_Gamma
    g = dict(np.load(os.path.join(GOLDEN, "cavi_cfg1.npz")))
    C, G = (int(v) for v in g["shape"])
    X = coo_matrix((g["data"], (g["row"], g["col"])), shape=(C, G))
    gam = lambda n: HPF_Gamma(g["init_" + n + "_shp"].copy(), g["init_" + n + "_rte"].copy())
    m = scHPF(5, verbose=False, bp=float(g["bp"]), dp=float(g["dp"]), device=[0, 1],
              xi=gam("xi"), theta=gam("theta"), eta=gam("eta"), beta=gam("beta"))
    m.fit(X, reinit=False, min_iter=10, max_iter=10, check_freq=3)
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.abs(b)))
    for n in ("theta", "beta", "xi", "eta"):
        assert rel(getattr(m, n).vi_shape, g["it10_" + n + "_shp"]) < 1e-9, n
        assert rel(getattr(m, n).vi_rate, g["it10_" + n + "_rte"]) < 1e-9, n
    assert np.allclose(m.loss, g["it10_loss"], rtol=1e-10)
    # the t == 0 branch (host Dirichlet, rows split by shard) and a projection on both devices
    np.random.seed(3)
    m2 = scHPF(5, verbose=False, device=[0, 1]).fit(X, min_iter=2, max_iter=2, check_freq=1)
    assert abs((m2.beta.vi_shape - m2.c).sum() / X.data.sum() - 1) < 1e-9
    p = m2.project(X, min_iter=2, max_iter=2, check_freq=1)
    assert p.beta == m2.beta and np.isfinite(p.loss[-1])


def _minibatch_worker(rank, world, port, out_dir, native):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from scipy.sparse import coo_matrix
    from schpf_b200 import scHPF, HPF_Gamma
    from schpf_b200 import engine as eng_mod
    from schpf_b200.engine import shard_coo_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    if not native:
        # the split-phase path: theta/xi, torch.distributed all-reduce of the buffer, then beta/eta
        orig = eng_mod.ShardedEngine.__init__
        eng_mod.ShardedEngine.__init__ = lambda self, local, group=None, native=None: orig(self, local, group, False)
    g = dict(np.load(os.path.join(GOLDEN, "minibatch_small.npz")))
    C, G = (int(v) for v in g["shape"])
    X, lo, hi = shard_coo_rows(coo_matrix((g["data"], (g["row"], g["col"])), shape=(C, G)), rank, world)
    gam = lambda n, rows: HPF_Gamma(g["B_init_%s_shp" % n][rows].copy(), g["B_init_%s_rte" % n][rows].copy())
    m = scHPF(3, verbose=False, device=rank, bp=float(g["bp"]), dp=float(g["dp"]), xi=gam("xi", slice(lo, hi)),
              theta=gam("theta", slice(lo, hi)), eta=gam("eta", slice(None)), beta=gam("beta", slice(None)))
    np.random.seed(int(g["C_seed"]) if rank == 0 else 999)
    m.fit(X, reinit=False, batchsize=int(g["C_batchsize"]), min_iter=int(g["C_iters"]), max_iter=int(g["C_iters"]),
          check_freq=int(g["C_check_freq"]), process_group=dist.group.WORLD)
    # reinit=True under sharding: the device's t == 0 draw with one shared seed; must run and stay finite
    np.random.seed(5)
    m2 = scHPF(3, verbose=False, device=rank)
    m2.fit(X, batchsize=120, min_iter=3, max_iter=3, check_freq=1, process_group=dist.group.WORLD)
    np.savez(os.path.join(out_dir, "mb%d.npz" % rank), lo=lo, hi=hi, loss=np.array(m.loss),
             theta_shp=m.theta.vi_shape, theta_rte=m.theta.vi_rate, xi_rte=m.xi.vi_rate,
             beta_shp=m.beta.vi_shape, beta_rte=m.beta.vi_rate, eta_rte=m.eta.vi_rate,
             loss2=np.array(m2.loss), beta2=m2.beta.vi_shape)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("native", [True, False])
def test_two_gpu_minibatch_fit_reproduces_the_reference(tmp_path, native):
    """fit(batchsize=, process_group=) over NCCL: case C of minibatch_small.npz (a seeded minibatch run of
    the real, single-process reference) with the cells of every window spread over two GPUs; the exchange
    after the cell update is issued by the engine (native) or by torch.distributed between the two
    phases of schpf_step_end."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_minibatch_worker, args=(world, _free_port(), str(tmp_path), native), nprocs=world, join=True)
    g = dict(np.load(os.path.join(GOLDEN, "minibatch_small.npz")))
    r = [dict(np.load(str(tmp_path / ("mb%d.npz" % k)))) for k in range(world)]
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.abs(b)))
    for n in ("beta_shp", "beta_rte", "eta_rte", "loss", "beta2", "loss2"):
        assert np.array_equal(r[0][n], r[1][n]), n
    for n in ("beta_shp", "beta_rte", "eta_rte"):
        assert rel(r[0][n], g["C_" + n]) < 1e-9, n
    for n in ("theta_shp", "theta_rte", "xi_rte"):
        assert rel(np.concatenate([r[0][n], r[1][n]]), g["C_" + n]) < 1e-9, n
    assert np.allclose(r[0]["loss"], g["C_loss"], rtol=1e-11)
    assert np.all(np.isfinite(r[0]["loss2"])) and np.all(np.isfinite(r[0]["beta2"]))
