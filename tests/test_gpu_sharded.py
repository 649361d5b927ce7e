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
