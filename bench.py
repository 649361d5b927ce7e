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
estimator (`scHPF.fit`, at N>1 `scHPF.fit(process_group=)`) from pinned host
buffers, transfers and layout build included (`e2e.cold` = the first fit of the
process); `roofline` is the Xphi+scatter sweep pair against the measured HBM
peak; `cpu_baseline` is the reference's numba path (or the oracle port) timed
on this box's host cores on a bounded row-prefix of the same matrix.

Round 2 additions, all inside the same line:
  parity   the golden 50-iteration run of the real reference at K=20
           (tests/golden/cavi_k20.npz) repeated through the SAME engine path the
           timed loop uses, cells sharded over the N ranks: max relative error of
           theta/beta/xi/eta against the golden state, loss error, and whether the
           beta/eta replicas of all ranks are bit-identical.  At N=1 also the live
           comparison with the CPU leg: the reference and the GPU engine run the
           same 50 iterations from the same initial state on the same 10k-cell
           prefix of the benchmark matrix (`parity.vs_reference_live`).
  strong   strong scaling: BASELINE cfg-4 (500k cells x 20k genes, 2000 draws/cell,
           K=30) and cfg-3 (100k x 20k, K=20) with the TOTAL matrix fixed and its
           cells sharded over the N ranks.
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
    """SM clock and throttle reasons DURING the timed region: an in-process NVML poll every 2 ms on
    a thread (the timed loop sits in ctypes calls, which drop the GIL); `nvidia-smi -lms 20` as a
    fallback when pynvml is not importable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, "/tmp/schpf_clocks_%d.csv" % os.getpid()
        self.thread, self.samples, self.stop_flag, self.nvml = None, [], False, None

    def _poll(self):
        n = self.nvml
        h = n.nvmlDeviceGetHandleByIndex(self.idx)
        bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap}
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        while not self.stop_flag:
            try:
                r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.samples.append((n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), mx,
                                     [k for k, b in bits.items() if r & b]))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            # the visible ordinal is not the NVML index when CUDA_VISIBLE_DEVICES is set
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    self.idx = int(vis.split(",")[self.idx])
                except Exception:
                    pass
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.samples:
                out.update(sm_mhz=float(np.median([s[0] for s in self.samples])), sm_max_mhz=float(self.samples[0][1]),
                           reasons=sorted({r for s in self.samples for r in s[2]}), samples=len(self.samples),
                           source="nvml, 2 ms poll inside the timed region")
            return out
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
                       samples=len(sm), source="nvidia-smi -lms 20")
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
def cpu_reference_run(row, col, val, C_sample, G, K, bp, dp, state, n_iter, warmup=1, keep_state=False):
    """Times the reference's own numba path (baseline/_ref, if it travelled) or
    the oracle port on the first C_sample cells.  Returns the dict for `cpu_baseline`
    and, with keep_state, the state after warmup + n_iter iterations (the parity leg)."""
    from scipy.sparse import coo_matrix
    keep = row < C_sample
    r, c, d = row[keep], col[keep], val[keep]
    nnz = int(r.shape[0])
    th = (state["theta"][0][:C_sample].copy(), state["theta"][1][:C_sample].copy())
    xi = (state["xi"][0][:C_sample].copy(), state["xi"][1][:C_sample].copy())
    kind, cores, extra, final = None, os.cpu_count() or 1, {}, None
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
            if keep_state:
                final = {n: (getattr(model, n).vi_shape, getattr(model, n).vi_rate) for n in ("theta", "beta", "xi", "eta")}
        except Exception as exc:                                   # reference did not travel / import failed
            extra = {"reference_unavailable": repr(exc)[:200]}
            kind = None
        finally:
            if ref_dir in sys.path:
                sys.path.remove(ref_dir)
    if kind is None:
        from oracle import hpf_c, hpf_numpy
        st = hpf_numpy.State(th[0], th[1], state["beta"][0].copy(), state["beta"][1].copy(), xi[0], xi[1],
                             state["eta"][0].copy(), state["eta"][1].copy())
        args = (d, r, c, st, HYPER["a"], HYPER["ap"], bp, HYPER["c"], HYPER["cp"], dp)
        hpf_c.cavi_run(*args, warmup)
        t0 = time.perf_counter()
        hpf_c.cavi_run(*args, n_iter)
        t_iter = (time.perf_counter() - t0) / n_iter
        kind, cores = "port", hpf_c.num_threads()
        if keep_state:
            final = {"theta": (st.theta_shp, st.theta_rte), "beta": (st.beta_shp, st.beta_rte),
                     "xi": (st.xi_shp, st.xi_rte), "eta": (st.eta_shp, st.eta_rte)}
    out = {"value": nnz / t_iter, "unit": "nnz-updates/s", "cores": cores, "kind": kind,
           "sample": "first %d cells of the same matrix (%d nnz), %d timed CAVI iterations after %d warm-up, "
                     "no loss checks" % (C_sample, nnz, n_iter, warmup),
           "iters_per_sec_on_sample": 1.0 / t_iter}
    out.update(extra)
    return (out, final, (r, c, d)) if keep_state else out


def max_rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


# ------------------------------------------------------------ parity ---------
def parity_golden(torch, dist, rank, world, local_rank, stream, engine_opts):
    """The golden 50-iteration run of the REAL reference at K=20 (tests/golden/cavi_k20.npz,
    generator committed) through the same engine path the timed loop uses, cells sharded over
    the ranks by nnz.  Every number is the max over ranks."""
    from schpf_b200.engine import CaviEngine, ShardedEngine, shard_bounds_by_nnz
    f32 = engine_opts.get("precision") == 32
    # float32 models: the reference's float32 run (20 iterations at K=20; its own float32 noise
    # against its float64 run is 2.5e-5, tests/test_gpu_f32.py) at 1e-4
    name, n_iter, every, fin, tol = (("fp32_k20.npz", 20, 5, "fin_", 1e-4) if f32 else ("cavi_k20.npz", 50, 10, "it50_", 1e-9))
    path = os.path.join(ROOT, "tests", "golden", name)
    if not os.path.exists(path):
        return {"unavailable": "tests/golden/%s missing" % name}
    g = dict(np.load(path))
    g = {k: (v.astype(np.float64) if v.dtype == np.float32 else v) for k, v in g.items()}
    C, G = (int(v) for v in g["shape"])
    K = g["init_theta_shp"].shape[1]
    b = shard_bounds_by_nnz(np.bincount(g["row"], minlength=C), world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    keep = (g["row"] >= lo) & (g["row"] < hi)
    a, ap, c, cp = (float(g[k]) for k in ("a", "ap", "c", "cp"))
    local = CaviEngine(hi - lo, G, K, device=local_rank, stream=stream, row_offset=lo, **engine_opts)
    try:
        local.set_coo(g["row"][keep] - lo, g["col"][keep], g["data"][keep])
        local.set_hyper(a, ap, float(g["bp"]), c, cp, float(g["dp"]))
        local.set_state(theta=(g["init_theta_shp"][lo:hi], g["init_theta_rte"][lo:hi]),
                        beta=(g["init_beta_shp"], g["init_beta_rte"]),
                        xi=(np.full(hi - lo, ap + K * a), g["init_xi_rte"][lo:hi]),
                        eta=(np.full(G, cp + K * c), g["init_eta_rte"]))
        eng = ShardedEngine(local, None) if world > 1 else local
        loss = []
        for t in range(n_iter):
            eng.step(1)
            if t % every == 0:
                loss.append(eng.loss())
        st = local.get_state()
        lanes = bool(local.counter("lanes"))
    finally:
        local.close()
    err = max(max(max_rel(st["theta"][i], g[fin + "theta_" + s][lo:hi]), max_rel(st["beta"][i], g[fin + "beta_" + s]))
              for i, s in ((0, "shp"), (1, "rte")))
    err = max(err, max_rel(st["xi"][1], g[fin + "xi_rte"][lo:hi]), max_rel(st["eta"][1], g[fin + "eta_rte"]))
    loss_rel = max_rel(np.array(loss), g["loss" if f32 else "it50_loss"])
    identical = True
    if world > 1:
        dev = torch.device("cuda", local_rank)
        e = torch.tensor([err, loss_rel], dtype=torch.float64, device=dev)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        err, loss_rel = (float(v) for v in e.tolist())
        # bit patterns of beta / eta on every rank: max == min over ranks <=> identical replicas
        bits = np.concatenate([st["beta"][0].ravel(), st["beta"][1].ravel(), st["eta"][1].ravel()]).view(np.int64)
        hi_t = torch.from_numpy(bits.copy()).to(dev)
        lo_t = hi_t.clone()
        dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
        identical = bool(torch.equal(hi_t, lo_t))
    return {"what": "tests/golden/%s: %d iterations of the real reference (%d x %d, K=%d%s), repeated by "
                    "the engine path of the timed loop with cells sharded over %d rank(s)" % (
                        name, n_iter, C, G, K, ", dtype=float32" if f32 else "", world),
            "max_rel_vs_golden": err, "loss_rel": loss_rel, "beta_replicas_bit_identical": identical,
            "tolerance": tol, "ok": bool(err < tol and loss_rel < tol and identical), "lanes_kernel": lanes}


# ------------------------------------------------------ strong scaling -------
def strong_case(torch, dist, rank, world, local_rank, stream, barrier, name, cells_total, genes, draws, K, steps,
                warmup, engine_opts, cf):
    """One strong-scaling point: a matrix of `cells_total` cells in TOTAL, cells sharded evenly over the ranks."""
    from schpf_b200.engine import CaviEngine, ShardedEngine, release_scratch
    from schpf_b200.synth import synth_coo_torch
    device = "cuda:%d" % local_rank
    C = cells_total // world + (1 if rank < cells_total % world else 0)
    row_offset = rank * (cells_total // world) + min(rank, cells_total % world)
    row, col, val = synth_coo_torch(C, genes, draws, K, seed=0, device=device, row_offset=row_offset)
    nnz_local = int(row.numel())
    reduce_sum = (lambda t: dist.all_reduce(t)) if world > 1 else None
    bp, dp = empirical_hypers(torch, row, col, val, C, genes, reduce_sum)
    state = init_state(C, genes, K, bp, dp, rank)
    local = CaviEngine(C, genes, K, device=local_rank, stream=stream, row_offset=row_offset, timing=1, **engine_opts)
    try:
        t0 = time.perf_counter()
        local.set_coo(row, col, val)
        torch.cuda.synchronize()
        layout_s = time.perf_counter() - t0
        del row, col, val
        torch.cuda.empty_cache()
        local.set_hyper(HYPER["a"], HYPER["ap"], bp, HYPER["c"], HYPER["cp"], dp)
        local.set_state(**state)
        eng = ShardedEngine(local, None) if world > 1 else local

        def run(n, t_start):
            out = []
            for t in range(t_start, t_start + n):
                eng.step(1)
                if t % cf == 0:
                    out.append(eng.loss())
            return out
        run(warmup, 0)
        barrier()
        local.counter("reset")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        losses = run(steps, warmup)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        nn = torch.tensor([float(nnz_local)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(nn)
        total_ms, nnz_total = float(ms.item()), int(nn.item())
        sweep_ms = local.counter("sweep_ms_shape")
        n_shape = local.counter("shape_sweep_launches")
        pair_ms = 2.0 * sweep_ms / max(n_shape, 1)
        algo = 12.0 * nnz_local + 16.0 * K * (C + genes)
        info = {"padded_nnz_cells": local.counter("padded_nnz_cells"), "padded_nnz_genes": local.counter("padded_nnz_genes"),
                "lanes_kernel": bool(local.counter("lanes"))}
    finally:
        local.close()
        release_scratch(local_rank)
        torch.cuda.empty_cache()
    return {"workload": name, "cells_total": cells_total, "cells_per_gpu": C, "nnz_total": nnz_total, "K": K,
            "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
            "value": nnz_total * steps / (total_ms * 1e-3), "unit": "nnz-updates/s",
            "iters_per_sec": steps / (total_ms * 1e-3), "sweep_pair_ms_rank0": pair_ms,
            "roofline_frac_rank0": algo / (pair_ms * 1e-3) / 1e9 / hbm_peak()[0] if pair_ms > 0 else None,
            "layout_build_s_rank0": layout_s, "loss_first_last": [losses[0], losses[-1]] if losses else None,
            "pad_fraction_rank0": [info["padded_nnz_cells"] / nnz_local - 1.0, info["padded_nnz_genes"] / nnz_local - 1.0],
            "lanes_kernel": info["lanes_kernel"]}


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def strong_block(torch, dist, rank, world, local_rank, stream, barrier, engine_opts, cf, steps):
    out = {"what": "strong scaling: the TOTAL matrix is fixed and its cells are sharded evenly over the %d rank(s); "
                   "value = nnz-updates/s of the whole job, timed like the main line (CUDA events, max over ranks)" % world,
           "cases": []}
    cases = [("BASELINE cfg-4: 500k cells x 20k genes, 2000 draws/cell, K=30, fp64", 500000, 20000, 2000, 30),
             ("BASELINE cfg-3 strong-scaled: 100k cells x 20k genes, 2000 draws/cell, K=20, fp64", 100000, 20000, 2000, 20)]
    ref_path = os.path.join(ROOT, "profiles", "r2_strong_n1.json")
    ref = json.load(open(ref_path)) if os.path.exists(ref_path) else {}
    for name, ct, g, d, K in cases:
        log("strong case: " + name)
        try:
            res = strong_case(torch, dist, rank, world, local_rank, stream, barrier, name, ct, g, d, K, steps, 3,
                              engine_opts, cf)
            base = ref.get(name)
            if base:
                res["speedup_vs_1gpu"] = res["value"] / base["value"]
                res["speedup_reference"] = "profiles/r2_strong_n1.json (builder-run N=1 line of this same block: %.4g nnz-updates/s)" % base["value"]
            out["cases"].append(res)
        except Exception as exc:
            out["cases"].append({"workload": name, "error": repr(exc)[:300]})
            torch.cuda.empty_cache()
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


def sweep_profile_note(K, lanes, lib_version):
    """ncu numbers of the dominant kernel (static: taken from the `ncu --set full` capture of the
    SAME kernel version, profiles/sweep_traffic.json; null when the capture is of another version)."""
    tpath = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    try:
        t = json.load(open(tpath))
    except Exception:
        return None, None
    for entry in t.get("captures", []):
        if entry.get("K") == K and bool(entry.get("lanes")) == bool(lanes) and entry.get("library_version") == lib_version:
            return entry.get("dram_bytes_per_iteration"), entry
    return None, None


def main():
    quiet_stdout()
    # a hung collective must not hang the box: dump every thread's stack and exit
    import faulthandler
    faulthandler.dump_traceback_later(env_int("SCHPF_BENCH_WATCHDOG_S", 900), exit=True, file=sys.stderr)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=CFG["cells_per_gpu"])
    ap.add_argument("--genes", type=int, default=CFG["genes"])
    ap.add_argument("--draws", type=int, default=CFG["draws_per_cell"])
    ap.add_argument("--factors", type=int, default=CFG["nfactors"])
    ap.add_argument("--cpu-cells", type=int, default=10000, help="row prefix run on the CPU (timing + live parity)")
    ap.add_argument("--cpu-iters", type=int, default=50, help="CAVI iterations of the CPU leg (BASELINE.md: 50 for parity)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block (cfg-4, cfg-3)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--strong-steps", type=int, default=10)
    ap.add_argument("--skewed", action="store_true",
                    help="genes drawn with Gamma(0.5,1) weights instead of uniformly (not the BASELINE workload)")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--panel-rows", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--target-ctas", type=int, default=0)
    ap.add_argument("--precision", type=int, default=64, choices=[32, 64],
                    help="32: the fp32 sweep that float32 models use (a separate line, never the headline)")
    ap.add_argument("--packed", type=int, default=-1, help="4-byte stream entries: -1 automatic, 0 wide entries")
    ap.add_argument("--lanes", type=int, default=1, help="0: lane-pair sweep for every K (round-1 kernels)")
    ap.add_argument("--rank-per-range", type=int, default=-1, help="owners re-ranked inside every panel range (-1 auto)")
    ap.add_argument("--free-schedule", type=int, default=-1, help="K 17..20: plane B without the bank schedule (-1 auto)")
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
    workload = "%dk cells x %dk genes, %d draws/cell%s, K=%d, %s%s" % (
        C // 1000, G // 1000, cfg["draws_per_cell"], " (skewed gene weights)" if args.skewed else "", K,
        "fp32 sweep (float32 model)" if args.precision == 32 else "fp64", " per GPU (cells sharded)" if world > 1 else "")

    if args.impl == "reference":
        return main_reference(args, cfg, rank, world, workload)

    import torch
    import torch.distributed as dist
    from schpf_b200 import _lib
    from schpf_b200.engine import CaviEngine, ShardedEngine
    from schpf_b200 import scHPF, HPF_Gamma

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = "cuda:%d" % local_rank
    bound = 0
    if world > 1:
        # one process per GPU: stay on the cores (and the memory) of this GPU's NUMA node
        from schpf_b200.engine import bind_host_to_device
        if not os.environ.get("SCHPF_BENCH_NO_BIND"):
            bound = bind_host_to_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(device))
    log("process group up (bound to %d cores)" % bound)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.current_stream().cuda_stream
    opts = dict(variant=args.variant, overlap_exchange=0 if args.no_overlap else 1)
    if args.panel_rows:
        opts["panel_rows"] = args.panel_rows
    if args.warps:
        opts["warps_per_cta"] = args.warps
    if args.target_ctas:
        opts["target_ctas"] = args.target_ctas
    if not args.lanes:
        opts["lanes"] = 0
    f32 = args.precision == 32
    if f32:
        opts["precision"] = 32
    if args.packed >= 0:
        opts["packed_entries"] = args.packed
    if os.environ.get("SCHPF_NO_OVERLAP_SWEEPS"):
        opts["overlap_sweeps"] = 0          # A/B switch: the two shape sweeps one after the other on one stream
    if args.rank_per_range >= 0:
        opts["rank_per_range"] = args.rank_per_range
    if args.free_schedule >= 0:
        opts["free_schedule"] = args.free_schedule

    # ---- parity first: the numbers below mean nothing if this is red ---------------
    parity = None
    if not args.no_parity:
        parity = parity_golden(torch, dist, rank, world, local_rank, stream, opts)
        log("golden parity: %s" % parity)

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
    state_bytes = 8.0 * (2 * C * K + 2 * G * K + 2 * C + 2 * G)

    hrow = hcol = hval = None
    if not args.no_e2e or not args.no_cpu:
        hrow = torch.empty(nnz_local, dtype=torch.int32, pin_memory=True).copy_(row)
        hcol = torch.empty(nnz_local, dtype=torch.int32, pin_memory=True).copy_(col)
        hval = torch.empty(nnz_local, dtype=torch.int32, pin_memory=True).copy_(val)
        torch.cuda.synchronize()
        log("pinned host copy of the shard ready")

    def e2e_fit(iters):
        """scHPF.fit through the public estimator from pinned host COO; at N>1 every rank fits its
        shard with process_group= (one model over all cells).  Returns (seconds max over ranks, loss checks)."""
        X = HostCOO(hrow.numpy(), hcol.numpy(), hval.numpy(), (C, G))
        gam = lambda p: HPF_Gamma(p[0].copy(), p[1].copy())
        if f32:
            gam = lambda p: HPF_Gamma(p[0].astype(np.float32), p[1].astype(np.float32))
        model = scHPF(K, bp=bp, dp=dp, verbose=False, device=local_rank, xi=gam(state["xi"]),
                      theta=gam(state["theta"]), eta=gam(state["eta"]), beta=gam(state["beta"]),
                      dtype=np.float32 if f32 else np.float64, **HYPER)
        kw = dict(process_group=dist.group.WORLD) if world > 1 else {}
        barrier()
        t0 = time.perf_counter()
        model.fit(X, reinit=False, min_iter=iters, max_iter=iters, check_freq=cf, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ts = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        return float(ts.item()), len(model.loss), model

    # ---- cold end-to-end: the FIRST fit of this process (device scratch, pools and
    #      module load are all cold), before anything else touches the engine -------
    e2e_cold = None
    if not args.no_e2e:
        cold_s, _, cold_model = e2e_fit(steps)
        e2e_cold = {"value": nnz_total * steps / cold_s, "unit": "nnz-updates/s", "seconds": cold_s,
                    "what": "first scHPF.fit of the process (cold device scratch / memory pool)"}
        del cold_model
        log("cold e2e fit: %.3fs" % cold_s)

    local = CaviEngine(C, G, K, device=local_rank, stream=stream, row_offset=rank * C, timing=1, **opts)
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
    shape_ms = local.counter("sweep_ms_shape")
    llh_ms = local.counter("sweep_ms_llh")
    n_shape = local.counter("shape_sweep_launches")
    launches = local.counter("kernel_launches")
    lanes = bool(local.counter("lanes"))
    info = {k: local.counter(k) for k in ("padded_nnz_cells", "padded_nnz_genes", "panel_rows", "grid_cells",
                                          "grid_genes", "layout_bytes", "slow_path_hits", "warps_per_cta", "packed_entries")}
    info["lanes_kernel"] = lanes

    # ---- roofline of the dominant kernel: the two shape sweeps of an iteration --
    peak, peak_src = hbm_peak()
    algo_bytes_iter = 12.0 * nnz_local + 16.0 * K * (C + G)           # B_xphi
    sweep_pair_ms = 2.0 * shape_ms / max(n_shape, 1)                  # shape sweeps only (llh sweeps timed apart)
    achieved = algo_bytes_iter / (sweep_pair_ms * 1e-3) / 1e9 if sweep_pair_ms > 0 else 0.0
    lib_version = int(_lib.load().schpf_version())
    traffic, prof = sweep_profile_note(K, lanes and not f32, lib_version) if not f32 else (None, None)
    kp = (K + 3) // 4 * 4
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": ("lane_sweep_f32_kernel<K=%d,SHAPE> x2 per iteration (cells-own + genes-own; one lane per owner, fp32 tables)" % K)
                if f32 else ("lane_sweep_kernel<K=%d,SHAPE> x2 per iteration (cells-own + genes-own; one lane per owner)" % K)
                if lanes else ("sweep_kernel<KP=%d,SHAPE> x2 per iteration (cells-own + genes-own; lane pairs)" % kp),
                "algorithmic_bytes_per_iteration": algo_bytes_iter,
                "sweep_pair_ms": sweep_pair_ms, "sweep_share_of_step": (shape_ms + llh_ms) / total_ms if total_ms else None,
                "llh_sweep_ms_total": llh_ms,
                "ncu": {k: prof[k] for k in ("capture", "dram_pct", "lsu_pct", "fp64_pct", "issue_pct", "ms_per_sweep_under_ncu")
                        if k in prof} if prof else None}

    value = nnz_total * steps / (total_ms * 1e-3)
    result = {
        "metric": "nnz_updates_per_sec", "value": value, "unit": "nnz-updates/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if f32 else "f64", "data": "synthetic",
        "iters_per_sec": steps / (total_ms * 1e-3),
        "config": {"workload": workload, "nnz_total": nnz_total, "nnz_per_gpu": nnz_local, "check_freq": cf,
                   "l2_policy": "inputs_exceed_L2 (entry streams %.1f GB per sweep vs 126 MB L2)" %
                                (info["padded_nnz_cells"] * (4 if info["packed_entries"] else 8) / 1e9),
                   "parallelism": ("cells sharded over %d GPU(s); one NCCL all-reduce of G*K+K doubles per iteration, %s"
                                   % (world, "torch.distributed" if args.torch_exchange else
                                      "issued by the engine in order on its stream" if args.no_overlap else
                                      "issued by the engine on a second stream under the cells-own sweep"))
                   if world > 1 else "single GPU",
                   "variant": "tiled two-pass sweep" if args.variant == 0 else "literal per-nnz atomics",
                   "layout": info, "layout_build_s": layout_s, "bp": bp, "dp": dp, "library_version": lib_version},
        "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
        "loss_first_last": [losses[0], losses[-1]] if losses else None,
        "parity": parity,
    }
    local.close()
    del row, col, val
    torch.cuda.empty_cache()

    # ---- end to end through the public estimator, from pinned host memory ------
    if not args.no_e2e:
        e2e_s, n_checks, model = e2e_fit(steps)
        del model
        result["e2e"] = {"value": nnz_total * steps / e2e_s, "unit": "nnz-updates/s",
                         "h2d_bytes_per_step": (12.0 * nnz_local + state_bytes) / steps,
                         "d2h_bytes_per_step": (state_bytes + 8.0 * n_checks) / steps,
                         "iters": steps, "seconds": e2e_s, "iters_per_sec": steps / e2e_s,
                         "what": ("scHPF(K).fit(X, reinit=False, max_iter=%d, check_freq=%d" % (steps, cf))
                                 + (", process_group=WORLD) on every rank's shard" if world > 1 else ")")
                                 + " from pinned host COO: upload, device layout build, iterations, loss checks, state download",
                         "cold": e2e_cold}

    # ---- strong scaling (cfg-4 and cfg-3 with the total matrix fixed) ---------------
    if not args.no_strong:
        result["strong"] = strong_block(torch, dist, rank, world, local_rank, stream, barrier, opts, cf, args.strong_steps)

    # ---- CPU leg: the reference's own path on this box's cores + live parity against it -----
    if rank == 0 and world == 1 and not args.no_cpu:
        Cs = min(args.cpu_cells, C)
        base, ref_state, (r, c, d) = cpu_reference_run(hrow.numpy(), hcol.numpy(), hval.numpy(), Cs, G, K, bp, dp, state,
                                                        n_iter=args.cpu_iters - 1, warmup=1, keep_state=True)
        result["cpu_baseline"] = base
        if ref_state is not None and parity is not None:
            with CaviEngine(Cs, G, K, device=local_rank, stream=stream, **opts) as e:
                e.set_coo(r, c, d)
                e.set_hyper(HYPER["a"], HYPER["ap"], bp, HYPER["c"], HYPER["cp"], dp)
                e.set_state(theta=(state["theta"][0][:Cs], state["theta"][1][:Cs]), beta=state["beta"],
                            xi=(state["xi"][0][:Cs], state["xi"][1][:Cs]), eta=state["eta"])
                e.step(args.cpu_iters)
                got = e.get_state()
            errs = {n + "_" + s: max_rel(got[n][i], ref_state[n][i]) for n in ("theta", "beta", "xi", "eta")
                    for i, s in ((0, "shape"), (1, "rate"))}
            parity["vs_reference_live"] = {
                "what": "%d CAVI iterations from the same initial state on the first %d cells of the benchmark matrix "
                        "(%d nnz, K=%d): this engine vs the %s on the host" % (
                            args.cpu_iters, Cs, int(r.shape[0]), K,
                            "unmodified reference (baseline/_ref, numba)" if base["kind"] == "reference" else "oracle port"),
                "max_rel": max(errs.values()), "per_array": errs, "target": 1e-4 if f32 else 1e-6,
                "ok": bool(max(errs.values()) < (1e-4 if f32 else 1e-6))}

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
        "config": {"workload": workload, "sample": base["sample"], "nnz_sample": nnz,
                   "note": "the reference materialises Xphi (8*K*nnz bytes: 32 GB at the full 100k-cell matrix), so it is "
                           "timed on a row prefix; its nnz-updates/s is flat in nnz (SURVEY 8d); a small prefix keeps "
                           "its tables in cache, which flatters the reference"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "nnz-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


if __name__ == "__main__":
    main()
