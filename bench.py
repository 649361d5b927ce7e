#!/usr/bin/env python
"""Benchmark of the CAVI hot path (BASELINE.json metric: CAVI iterations/s and
nnz-updates/s at K=20 on 100k x 20k sparse counts; 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one full CAVI iteration (both sweeps, beta/eta and theta/xi
finalisation, plus the Poisson log-likelihood every 10th iteration, the
reference's default check_freq) over one synthetic count matrix.  N=1 runs
BASELINE cfg-3 (100k cells x 20k genes, 2000 draws/cell, K=20); under torchrun
every rank holds a cfg-3-sized shard of cells (weak scaling) and the gene-side
statistics are all-reduced once per iteration.

Prints ONE JSON line (rank 0).  `value` is whole-job nnz-updates/s with the
matrix already resident in HBM; `e2e` is the same metric through the public
estimator (`scHPF.fit`) from pinned host buffers, transfers and layout build
included; `roofline` is the Xphi+scatter sweep pair against the measured HBM
peak; `cpu_baseline` is the reference's numba path (or the oracle port) timed
on this box's host cores on a bounded row-prefix of the same matrix.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(cells_per_gpu=100000, genes=20000, draws_per_cell=2000, nfactors=20, check_freq=10)
HYPER = dict(a=0.3, ap=1.0, c=0.3, cp=1.0)


_T0 = time.perf_counter()


def log(msg):
    """progress to stderr (SCHPF_BENCH_VERBOSE=1): where a multi-rank run is, and when"""
    if os.environ.get("SCHPF_BENCH_VERBOSE"):
        sys.stderr.write("[bench r%s +%.1fs] %s\n" % (os.environ.get("RANK", "0"), time.perf_counter() - _T0, msg))
        sys.stderr.flush()


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------ clocks ---------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, "/tmp/schpf_clocks_%d.csv" % os.getpid()

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


# ------------------------------------------------------ problem set-up -------
def make_problem(torch, device, rank, cfg):
    from schpf_b200.synth import synth_coo_torch
    C, G, K = cfg["cells_per_gpu"], cfg["genes"], cfg["nfactors"]
    row, col, val = synth_coo_torch(C, G, cfg["draws_per_cell"], K, seed=0, device=device,
                                    row_offset=rank * C, skewed=cfg.get("skewed", False))
    return row, col, val


def empirical_hypers(torch, row, col, val, C, G, group_reduce=None):
    """b', d' as scHPF_.py:847-879 (mean/var of cell and gene totals)."""
    rs = torch.zeros(C, dtype=torch.float64, device=row.device).index_add_(0, row.long(), val.double())
    cs = torch.zeros(G, dtype=torch.float64, device=row.device).index_add_(0, col.long(), val.double())
    stats = torch.stack([rs.sum(), (rs * rs).sum(), torch.tensor(float(C), device=row.device, dtype=torch.float64)])
    if group_reduce is not None:
        group_reduce(stats)
        group_reduce(cs)
    mean = stats[0] / stats[2]
    var = stats[1] / stats[2] - mean * mean
    bp = float(HYPER["ap"] * mean / var)
    dp = float(HYPER["cp"] * cs.mean() / cs.var(unbiased=False))
    if bp > 1000 * dp:
        dp = bp / 1000
    return bp, dp


def init_state(C, G, K, bp, dp, seed):
    """U(0.5p, 1.5p) like HPF_Gamma.random_gamma_factory; beta/eta identical on every rank."""
    rl = np.random.default_rng(1000 + seed)      # cell side: per rank
    rg = np.random.default_rng(7)                 # gene side: shared
    u = lambda r, p, shape: r.uniform(0.5 * p, 1.5 * p, shape)
    return dict(xi=(np.full(C, HYPER["ap"] + K * HYPER["a"]), u(rl, bp, C)),
                theta=(u(rl, HYPER["a"], (C, K)), u(rl, bp, (C, K))),
                eta=(np.full(G, HYPER["cp"] + K * HYPER["c"]), u(rg, dp, G)),
                beta=(u(rg, HYPER["c"], (G, K)), u(rg, dp, (G, K))))


class HostCOO(object):
    """What scHPF.fit needs from a scipy coo_matrix: row / col / data / shape."""

    def __init__(self, row, col, data, shape):
        self.row, self.col, self.data, self.shape = row, col, data, shape
        self.nnz = row.shape[0]


# ------------------------------------------------------ CPU baseline ---------
def cpu_reference_run(row, col, val, C_sample, G, K, bp, dp, state, n_iter, warmup=1):
    """Times the reference's own numba path (baseline/_ref, if it travelled) or
    the oracle port on the first C_sample cells.  Returns dict for `cpu_baseline`."""
    from scipy.sparse import coo_matrix
    keep = row < C_sample
    r, c, d = row[keep], col[keep], val[keep]
    nnz = int(r.shape[0])
    th = (state["theta"][0][:C_sample].copy(), state["theta"][1][:C_sample].copy())
    xi = (state["xi"][0][:C_sample].copy(), state["xi"][1][:C_sample].copy())
    kind, cores, extra = None, os.cpu_count() or 1, {}
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    t_iter = None
    if os.path.isdir(os.path.join(ref_dir, "schpf")) and not os.environ.get("SCHPF_BENCH_FORCE_PORT"):
        try:
            sys.path.insert(0, ref_dir)
            import numba
            import schpf as ref
            numba.set_num_threads(min(cores, numba.config.NUMBA_NUM_THREADS))
            X = coo_matrix((d, (r, c)), shape=(C_sample, G), dtype=np.int32)
            mk = lambda p: ref.HPF_Gamma(p[0].copy(), p[1].copy())
            model = ref.scHPF(K, bp=bp, dp=dp, verbose=False, xi=mk(xi), theta=mk(th),
                              eta=mk(state["eta"]), beta=mk(state["beta"]), **HYPER)
            model.fit(X, reinit=False, min_iter=warmup, max_iter=warmup, check_freq=warmup + 1)   # JIT
            t0 = time.perf_counter()
            model.fit(X, reinit=False, min_iter=n_iter, max_iter=n_iter, check_freq=n_iter + 1)
            t_iter = (time.perf_counter() - t0) / n_iter
            kind, cores = "reference", int(numba.get_num_threads())
            extra = {"threading_layer": numba.threading_layer(), "impl": "schpf %s numba %s" % (ref.__version__, numba.__version__)}
        except Exception as exc:                                   # reference did not travel / import failed
            extra = {"reference_unavailable": repr(exc)[:200]}
            kind = None
        finally:
            if ref_dir in sys.path:
                sys.path.remove(ref_dir)
    if kind is None:
        from oracle import hpf_c, hpf_numpy
        st = hpf_numpy.State(th[0], th[1], state["beta"][0], state["beta"][1], xi[0], xi[1],
                             state["eta"][0], state["eta"][1])
        args = (d, r, c, st, HYPER["a"], HYPER["ap"], bp, HYPER["c"], HYPER["cp"], dp)
        hpf_c.cavi_run(*args, warmup)
        t0 = time.perf_counter()
        hpf_c.cavi_run(*args, n_iter)
        t_iter = (time.perf_counter() - t0) / n_iter
        kind, cores = "port", hpf_c.num_threads()
    out = {"value": nnz / t_iter, "unit": "nnz-updates/s", "cores": cores, "kind": kind,
           "sample": "first %d cells of the same matrix (%d nnz), %d timed CAVI iterations after %d warm-up, "
                     "no loss checks" % (C_sample, nnz, n_iter, warmup),
           "iters_per_sec_on_sample": 1.0 / t_iter}
    out.update(extra)
    return out


# ------------------------------------------------------------- main ----------
_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, torchrun) write to fd 1; the contract is ONE JSON
    line there.  Route fd 1 to stderr for the run and keep the real one for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, line)
    else:
        os.write(_REAL_STDOUT, line)


def main():
    quiet_stdout()
    # a hung collective must not hang the box: dump every thread's stack and exit
    import faulthandler
    faulthandler.dump_traceback_later(env_int("SCHPF_BENCH_WATCHDOG_S", 420), exit=True, file=sys.stderr)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=CFG["cells_per_gpu"])
    ap.add_argument("--genes", type=int, default=CFG["genes"])
    ap.add_argument("--draws", type=int, default=CFG["draws_per_cell"])
    ap.add_argument("--factors", type=int, default=CFG["nfactors"])
    ap.add_argument("--cpu-cells", type=int, default=4000, help="row prefix timed on the CPU")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--skewed", action="store_true",
                    help="genes drawn with Gamma(0.5,1) weights instead of uniformly (not the BASELINE workload)")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--panel-rows", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--target-ctas", type=int, default=0)
    ap.add_argument("--no-overlap", action="store_true",
                    help="N>1: the engine's all-reduce in order on its stream instead of under the cells-own sweep")
    ap.add_argument("--torch-exchange", action="store_true",
                    help="N>1: all-reduce through torch.distributed instead of the engine's own NCCL call")
    args = ap.parse_args()
    cfg = dict(CFG, cells_per_gpu=args.cells, genes=args.genes, draws_per_cell=args.draws, nfactors=args.factors,
               skewed=args.skewed)
    C, G, K, cf = cfg["cells_per_gpu"], cfg["genes"], cfg["nfactors"], cfg["check_freq"]
    steps, warmup = args.steps, max(args.warmup, 0)

    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    workload = "%dk cells x %dk genes, %d draws/cell%s, K=%d, fp64%s" % (
        C // 1000, G // 1000, cfg["draws_per_cell"], " (skewed gene weights)" if args.skewed else "", K,
        " per GPU (cells sharded)" if world > 1 else "")

    if args.impl == "reference":
        return main_reference(args, cfg, rank, world, workload)

    import torch
    import torch.distributed as dist
    from schpf_b200.engine import CaviEngine, ShardedEngine
    from schpf_b200 import scHPF, HPF_Gamma

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = "cuda:%d" % local_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    log("process group up")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    reduce_sum = (lambda t: dist.all_reduce(t)) if world > 1 else None
    row, col, val = make_problem(torch, device, rank, cfg)
    nnz_local = int(row.numel())
    log("synthetic shard ready: nnz %d" % nnz_local)
    bp, dp = empirical_hypers(torch, row, col, val, C, G, reduce_sum)
    state = init_state(C, G, K, bp, dp, rank)
    nnz_t = torch.tensor([float(nnz_local)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(nnz_t)
    nnz_total = int(nnz_t.item())

    stream = torch.cuda.current_stream().cuda_stream
    opts = dict(timing=1, variant=args.variant, overlap_exchange=0 if args.no_overlap else 1)
    if args.panel_rows:
        opts["panel_rows"] = args.panel_rows
    if args.warps:
        opts["warps_per_cta"] = args.warps
    if args.target_ctas:
        opts["target_ctas"] = args.target_ctas
    local = CaviEngine(C, G, K, device=local_rank, stream=stream, row_offset=rank * C, **opts)
    t0 = time.perf_counter()
    local.set_coo(row, col, val)
    torch.cuda.synchronize()
    layout_s = time.perf_counter() - t0
    log("layout built in %.2fs" % layout_s)
    local.set_hyper(HYPER["a"], HYPER["ap"], bp, HYPER["c"], HYPER["cp"], dp)
    local.set_state(**state)
    engine = ShardedEngine(local, None, native=not args.torch_exchange) if world > 1 else local
    log("engine ready (exchange: %s)" % ("single GPU" if world == 1 else "native" if engine.native else "torch"))

    def run(n, t_start):
        """n CAVI iterations with the loss every cf-th, like _fit does"""
        losses = []
        for t in range(t_start, t_start + n):
            engine.step(1)
            if t % cf == 0:
                losses.append(engine.loss())
        return losses

    run(warmup, 0)
    log("warm-up done")
    barrier()
    local.counter("reset")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    losses = run(steps, warmup)
    e1.record()
    barrier()
    log("timed region done")
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else {}
    total_ms = float(ms.item())
    sweep_ms = local.counter("sweep_ms")
    n_sweeps = local.counter("sweep_launches")
    launches = local.counter("kernel_launches")
    n_llh = sum(1 for t in range(warmup, warmup + steps) if t % cf == 0)
    n_shape_sweeps = n_sweeps - n_llh
    info = {k: local.counter(k) for k in ("padded_nnz_cells", "padded_nnz_genes", "panel_rows", "grid_cells",
                                          "grid_genes", "layout_bytes", "slow_path_hits")}

    # ---- roofline of the dominant kernel: the two shape sweeps of an iteration --
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    algo_bytes_iter = 12.0 * nnz_local + 16.0 * K * (C + G)           # B_xphi
    # sweep_ms covers shape sweeps and llh sweeps; an llh sweep costs about one shape sweep
    sweep_pair_ms = 2.0 * sweep_ms / max(n_sweeps, 1)
    achieved = algo_bytes_iter / (sweep_pair_ms * 1e-3) / 1e9 if sweep_pair_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_iteration")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": "sweep_kernel<KP=%d,SHAPE> x2 per iteration (cells-own + genes-own)" % ((K + 3) // 4 * 4),
                "algorithmic_bytes_per_iteration": algo_bytes_iter,
                "sweep_pair_ms": sweep_pair_ms, "sweep_share_of_step": sweep_ms / total_ms if total_ms else None}

    value = nnz_total * steps / (total_ms * 1e-3)
    result = {
        "metric": "nnz_updates_per_sec", "value": value, "unit": "nnz-updates/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "iters_per_sec": steps / (total_ms * 1e-3),
        "config": {"workload": workload, "nnz_total": nnz_total, "nnz_per_gpu": nnz_local, "check_freq": cf,
                   "l2_policy": "inputs_exceed_L2 (entry streams %.1f GB per sweep vs 126 MB L2)" %
                                (info["padded_nnz_cells"] * 8 / 1e9),
                   "parallelism": ("cells sharded over %d GPU(s); one NCCL all-reduce of G*K+K doubles per iteration, %s"
                                   % (world, "torch.distributed" if args.torch_exchange else
                                      "issued by the engine in order on its stream" if args.no_overlap else
                                      "issued by the engine on a second stream under the cells-own sweep"))
                   if world > 1 else "single GPU",
                   "variant": "tiled two-pass sweep" if args.variant == 0 else "literal per-nnz atomics",
                   "layout": info, "layout_build_s": layout_s, "bp": bp, "dp": dp},
        "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
        "loss_first_last": [losses[0], losses[-1]] if losses else None,
    }

    # ---- end to end through the public estimator, from pinned host memory ------
    if not args.no_e2e:
        hrow = torch.empty(nnz_local, dtype=torch.int32, pin_memory=True).copy_(row)
        hcol = torch.empty(nnz_local, dtype=torch.int32, pin_memory=True).copy_(col)
        hval = torch.empty(nnz_local, dtype=torch.int32, pin_memory=True).copy_(val)
        torch.cuda.synchronize()
        log("pinned host copy of the shard ready")
        local.close()
        log("first engine closed")
        del row, col, val
        torch.cuda.empty_cache()
        e2e_iters = steps
        if world == 1:
            X = HostCOO(hrow.numpy(), hcol.numpy(), hval.numpy(), (C, G))
            gam = lambda p: HPF_Gamma(p[0].copy(), p[1].copy())
            model = scHPF(K, bp=bp, dp=dp, verbose=False, device=local_rank, xi=gam(state["xi"]),
                          theta=gam(state["theta"]), eta=gam(state["eta"]), beta=gam(state["beta"]), **HYPER)
            barrier()
            t0 = time.perf_counter()
            model.fit(X, reinit=False, min_iter=e2e_iters, max_iter=e2e_iters, check_freq=cf)
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
            n_checks = len(model.loss)
        else:
            barrier()
            t0 = time.perf_counter()
            loc = CaviEngine(C, G, K, device=local_rank, stream=stream, row_offset=rank * C,
                             overlap_exchange=0 if args.no_overlap else 1)
            loc.set_coo(hrow, hcol, hval)
            loc.set_hyper(HYPER["a"], HYPER["ap"], bp, HYPER["c"], HYPER["cp"], dp)
            loc.set_state(**state)
            eng = ShardedEngine(loc, None, native=not args.torch_exchange)
            n_checks = 0
            for t in range(e2e_iters):
                eng.step(1)
                if t % cf == 0:
                    eng.loss()
                    n_checks += 1
            loc.get_state()
            barrier()
            e2e_s = time.perf_counter() - t0
            log("e2e loop done")
            loc.close()
        ts = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e2e_s = float(ts.item())
        state_bytes = 8.0 * (2 * C * K + 2 * G * K + 2 * C + 2 * G)
        result["e2e"] = {"value": nnz_total * e2e_iters / e2e_s, "unit": "nnz-updates/s",
                         "h2d_bytes_per_step": (12.0 * nnz_local + state_bytes) / e2e_iters,
                         "d2h_bytes_per_step": (state_bytes + 8.0 * n_checks) / e2e_iters,
                         "iters": e2e_iters, "seconds": e2e_s, "iters_per_sec": e2e_iters / e2e_s,
                         "what": "scHPF(K).fit(X, reinit=False, max_iter=%d, check_freq=%d) from pinned host COO: "
                                 "upload, device layout build, iterations, loss checks, state download" % (e2e_iters, cf)
                         if world == 1 else "per-rank engine from pinned host COO shard + ShardedEngine loop + state download"}
        cpu_src = (hrow.numpy(), hcol.numpy(), hval.numpy())
    else:
        cpu_src = (row.cpu().numpy(), col.cpu().numpy(), val.cpu().numpy())

    if rank == 0 and world == 1 and not args.no_cpu:
        result["cpu_baseline"] = cpu_reference_run(*cpu_src, min(args.cpu_cells, C), G, K, bp, dp, state, n_iter=3)

    if rank == 0:
        emit(result)
    log("result emitted")
    if world > 1:
        # leave without tearing communicators down: teardown order between torch's NCCL
        # communicator and the engine's is not worth a hang at exit
        dist.barrier()
        sys.stderr.flush()
        os._exit(0)


def main_reference(args, cfg, rank, world, workload):
    """The reference's own CPU implementation of the path on this box's host cores
    (rank 0 only): each step is one CAVI iteration on a bounded row-prefix of the
    same synthetic matrix."""
    if rank != 0:
        return
    from schpf_b200.synth import synth_coo
    C, G, K = cfg["cells_per_gpu"], cfg["genes"], cfg["nfactors"]
    Cs = min(args.cpu_cells, C)
    # the same generator family on the host (the device stream cannot be reproduced without a GPU)
    X = synth_coo(Cs, G, cfg["draws_per_cell"], K, seed=0)
    rs, cs = np.asarray(X.sum(axis=1)).ravel(), np.asarray(X.sum(axis=0)).ravel()
    bp = HYPER["ap"] * rs.mean() / rs.var()
    dp = HYPER["cp"] * cs.mean() / cs.var()
    if bp > 1000 * dp:
        dp = bp / 1000
    state = init_state(Cs, G, K, bp, dp, 0)
    base = cpu_reference_run(X.row, X.col, X.data, Cs, G, K, bp, dp, state, n_iter=max(args.steps, 1),
                             warmup=max(args.warmup, 1))
    nnz = int(X.nnz)
    line = {
        "impl": "reference", "metric": "nnz_updates_per_sec", "value": base["value"], "unit": "nnz-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * nnz / base["value"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "iters_per_sec": base["iters_per_sec_on_sample"],
        "config": {"workload": workload, "sample": base["sample"], "nnz_sample": nnz},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "nnz-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


if __name__ == "__main__":
    main()
