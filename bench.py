#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on B200:

    rNNGP (spNNGP method="response") MCMC iterations/s at n = 1,000,000 locations, m = 15 neighbours,
    plus updateBF locations/s of the dominant kernel against the measured FP64 / HBM roofline.

A "step" is one Metropolis iteration of the reference's `rNNGP` sampler (src/rNNGP.cpp:272-378): draw
beta, evaluate the log-likelihood (B/F solve + Q) at the current and at the proposed theta, accept/reject.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle/_ref)

Prints ONE JSON line on rank 0.  N > 1 is launched with torchrun (one rank per GPU): the locations are
sharded over the ranks (strong scaling, total work fixed) and the likelihood partials (4 doubles per rank and
iteration) are exchanged through the engine's shared-memory mailbox (include/nngp_b200.h: nngp_comm_init; set
NNGP_REDUCE=nccl for the ncclAllReduce path).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

# stdout carries ONE JSON line: NCCL_DEBUG=VERSION (from the environment or from /etc/nccl.conf) makes libnccl print its
# banner there.  Two guards: the variable is pinned to WARN unless the caller asked for more, and file descriptor 1 points at
# stderr for the whole run -- the JSON line is written to the saved descriptor (emit_line).
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
_REAL_STDOUT = None


def protect_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

COV = {"exponential": 0, "spherical": 1, "matern": 2, "gaussian": 3}
# The reference's C++ is unmodified, but it runs here without R: Rmath's bessel_k (RKBESL) is stood in for by libstdc++'s
# std::cyl_bessel_k (spnngp_b200/rshim/rmath_shim.cpp), which is several times slower.  The Matern CPU figure is therefore a
# property of that stand-in; the exponential model (--cov exponential) is the CPU comparison that does not depend on it.
BESSEL_NOTE = ("Matern under the reference arm calls libstdc++ std::cyl_bessel_k (rshim stand-in for Rmath bessel_k, no R in "
               "this image): slower than R's RKBESL -- use the exponential model's ratio as the CPU comparison")


def flops_per_location(m: int, fused: bool = True) -> float:
    """Algorithmic FLOPs per location: distance 6P + correlation 3P + nugget m + Cholesky
    (m^3/3 + m^2/2 + m/6) + solves 2m^2 + F (2m+2) + log 1 [+ fused Q 2m+4], P = m(m+1)/2."""
    P = m * (m + 1) / 2
    f = 6 * P + 3 * P + m + (m ** 3 / 3 + m ** 2 / 2 + m / 6) + 2 * m * m + (2 * m + 2) + 1
    return f + (2 * m + 4 if fused else 0)


def workload_label(n: int, m: int, cov: str) -> str:
    """config.workload, identical in both arms."""
    tag = (" (BASELINE.json configs[2])" if (n, m, cov) == (1000000, 15, "matern") else
           " (BASELINE.json configs[0])" if (n, m, cov) == (5000, 15, "exponential") else "")
    return f"rNNGP response model n={n} m={m} p=2 cov={cov}{tag}"


def bytes_per_location(m: int, tri: bool = False) -> float:
    """Algorithmic bytes per location with B, F stored: coords 16 + nnIndx 4m + B 8m + F 8 + resid 8.
    tri: the distance-triangle kernel streams the 8 * m(m+1)/2 bytes of resident pair distances instead of the sites."""
    return (4 * m * (m + 1) + 16 + 12 * m) if tri else (32 + 12 * m)


def make_workload(n: int, m: int, seed: int = 1):
    """coords ~ U(0,1)^2 * sqrt(n/5000) (constant density), ordered by x; X = [1, N(0,1)];
    y = X beta + smooth field + noise.  numpy default_rng(seed); identical on every rank."""
    rng = np.random.default_rng(seed)
    coords = rng.random((n, 2)) * np.sqrt(n / 5000.0)
    coords = coords[np.argsort(coords[:, 0], kind="stable")]
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    beta = np.array([1.0, 5.0])
    w = np.sin(3.0 * coords[:, 0]) + np.cos(2.0 * coords[:, 1])
    y = X @ beta + np.sqrt(5.0) * w + rng.standard_normal(n)
    # column-major like the R matrices the reference receives (REAL(coords_r), REAL(X_r)): the engine's wrapper then
    # passes the buffers through without a layout copy
    return dict(n=n, m=m, p=2, coords=np.asfortranarray(coords), X=np.asfortranarray(X), y=y, beta=beta)


def sampler_settings(cov: int):
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0, "nu": 0.75}
    # well-mixing proposal sd; at n = 1e6 the posterior is sharp, so keep it small
    tuning = {"sigma.sq": 0.002, "tau.sq": 0.002, "phi": 0.002, "nu": 0.002}
    priors = dict(sigmaSqIG=[2.0, 5.0], tauSqIG=[2.0, 1.0], phiUnif=[3.0, 300.0], nuUnif=[0.1, 2.0] if cov == 2 else None)
    return starting, tuning, priors


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled from a
    background thread every ~2 ms (the timed region can be shorter than one nvidia-smi -lms period)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.masks = set()
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            # CUDA_VISIBLE_DEVICES remaps indices: resolve through the UUID when possible
            h = None
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.index < len(ids):
                    tok = ids[self.index]
                    h = pynvml.nvmlDeviceGetHandleByUUID(tok) if tok.startswith("GPU-") else \
                        pynvml.nvmlDeviceGetHandleByIndex(int(tok))
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self._h = h
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
            return

        def loop():
            nv = self._nvml
            while not self._stop.is_set():
                try:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    try:
                        m = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    except Exception:
                        m = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for bit in self.REASONS:
                        if m & bit:
                            self.masks.add(self.REASONS[bit])
                except Exception:
                    pass
                time.sleep(0.002)

        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": []}
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        if self.samples:
            out = {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.masks),
                   "samples": len(self.samples), "source": "NVML, 2 ms polling inside the timed region"}
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def traffic_from_profile(n, m, cov_name):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if one matches."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        for row in json.load(open(p)):
            if row["n"] == n and row["m"] == m and row["cov"] == cov_name:
                return row["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


# =====================================================================================================
# reference arm: the reference's own CPU implementation on the box's host cores
# =====================================================================================================
def run_reference(args, rank, world):
    if rank != 0:
        return 0
    from oracle import pyoracle as po
    cov = COV[args.cov]
    cores = os.cpu_count() or 1
    have_ref = po.Ref.available()
    port = po.Port()
    port.set_threads(cores)
    ref = po.Ref() if have_ref else None
    kind = "reference" if have_ref else "port"
    starting, tuning, pri = sampler_settings(cov)

    # size the per-step sample so that (W + K) iterations finish in a few minutes: probe one B/F pass
    n_full, m = args.n, args.m
    probe_n = min(n_full, 50000)
    wl = make_workload(probe_n, m)
    idx, _, lu = port.mk_nn_indx_cb(wl["coords"], m)
    theta = [5.0, 1.0, 6.0, 0.75][: 4 if cov == 2 else 3]
    t0 = time.perf_counter()
    if ref is not None:
        ref.update_bf(wl["coords"], idx, lu, probe_n, m, theta, cov, threads=cores)
    else:
        port.update_bf(wl["coords"], idx, lu, probe_n, m, theta, cov)
    t_loc = (time.perf_counter() - t0) / probe_n  # seconds per location per B/F pass
    budget = float(args.ref_budget_s)
    iters = args.steps + args.warmup
    n_s = int(min(n_full, max(2000, budget / (iters * 2.2 * t_loc))))
    if n_s < n_full:
        wl = make_workload(n_s, m)
        idx, _, lu = port.mk_nn_indx_cb(wl["coords"], m)
    elif probe_n != n_full:
        wl = make_workload(n_full, m)
        idx, _, lu = port.mk_nn_indx_cb(wl["coords"], m)

    def run(k):
        if ref is not None:
            ref.set_seed(1)
            t0 = time.perf_counter()
            ref.rnngp(wl["y"], wl["X"], wl["coords"], cov, idx, lu, m, pri["sigmaSqIG"], pri["tauSqIG"], pri["phiUnif"],
                      pri["nuUnif"], starting, tuning, k, nThreads=cores)
            return time.perf_counter() - t0
        from spnngp_b200 import _lib
        R = _lib.load_rshim(); R.rshim_set_seed(1)
        t0 = time.perf_counter()
        port.rnngp(wl["y"], wl["X"], wl["coords"], cov, idx, lu, m, pri["sigmaSqIG"], pri["tauSqIG"], pri["phiUnif"],
                   pri["nuUnif"], starting, tuning, k, lambda a, b: R.rnorm(a, b), lambda a, b: R.runif(a, b))
        return time.perf_counter() - t0

    if args.warmup > 0:
        run(args.warmup)
    dt = run(args.steps)
    iter_s_sample = args.steps / dt
    value = iter_s_sample * (n_s / n_full)  # work per iteration is linear in n (2 x updateBF + 2 x Q)
    sample = (f"{args.steps} iterations of the reference's rNNGP (2 x updateBF + 2 x Q each) on the first-built "
              f"n_s={n_s} locations of the workload, OpenMP {cores} threads; "
              + ("full size" if n_s == n_full else f"scaled to n={n_full} by n_s/n (work is linear in n)"))
    line = {
        "impl": "reference", "metric": "rNNGP_mcmc_iter_per_s", "value": value, "unit": "iter/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_label(n_full, m, args.cov), "n": n_full, "m": m, "cov_model": args.cov, "p": 2,
                   "sampler": "reference (src/rNNGP.cpp), OpenMP"},
        "cpu_baseline": {"value": value, "unit": "iter/s", "cores": cores, "kind": kind, "sample": sample,
                         "bessel_k": BESSEL_NOTE if cov == 2 else None},
        "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)
    return 0


# =====================================================================================================
# this repo's arm
# =====================================================================================================
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from spnngp_b200 import Engine, _lib
    from spnngp_b200.dist import balanced_shard_range, make_allreduce, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; spnngp_b200 has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cov = COV[args.cov]
    n, m = args.n, args.m
    wl = make_workload(n, m)
    starting, tuning, pri = sampler_settings(cov)
    R = _lib.load_rshim()

    stream = torch.cuda.Stream(dev)  # a real (non-default) stream: the engine and the timing events share it
    torch.cuda.set_stream(stream)
    eng = Engine(local_rank, stream=stream.cuda_stream)
    t0 = time.perf_counter()
    nnIndx, nnDist, lu = eng.mk_nn_indx(wl["coords"], m, "cb")  # setup, not timed: the sampler reuses neighbours
    t_knn = time.perf_counter() - t0
    lo, hi = shard_range(n, m, rank, world)
    # the engine's own NCCL communicator: likelihood partials are summed on the device right behind the
    # kernel (python-level hook `make_allreduce` is the slower, backend-agnostic alternative)
    allreduce = None
    if world > 1:
        if args.py_allreduce:
            allreduce = make_allreduce(device=dev)
        else:
            eng.comm_init_from_torch()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def chain(k, mode):
        R.rshim_set_seed(1)  # same stream on every rank -> identical accept decisions
        return eng.rnngp(cov, pri["sigmaSqIG"], pri["tauSqIG"], pri["phiUnif"], pri["nuUnif"], starting, tuning, k,
                         R.rnorm, R.runif, mode=mode, allreduce=allreduce)

    def timed_chain(k, mode, w):
        """ONE chain of w + k iterations: the first w are the warm-up, the k after them are timed (CUDA events on the engine's
        stream, barrier + synchronize on both sides).  The boundary is taken in the sampler's once-per-iteration hook, so
        the timed region holds exactly k whole iterations of a running chain -- each with its residual kernel, fused launch,
        reduction, accept step and, on accepted proposals, the cross forms -- and not the chain's one-off initial solve."""
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        seen = [0, 0]

        def hook():
            seen[0] += 1
            if seen[0] == w:
                barrier()
                seen[1] = eng.launch_count()
                e0.record(stream)
            return 0

        R.rshim_set_seed(1)
        _, ts, _ = eng.rnngp(cov, pri["sigmaSqIG"], pri["tauSqIG"], pri["phiUnif"], pri["nuUnif"], starting, tuning, w + k,
                             R.rnorm, R.runif, mode=mode, allreduce=allreduce, interrupt=hook)
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        acc = int(np.sum(np.any(ts[:, w:] != ts[:, w - 1:-1], axis=0)))  # accepted proposals inside the timed window
        return float(ms.item()), acc, eng.launch_count() - seen[1]

    # ---- HBM-resident: data uploaded once, then W warm-up and K timed iterations -----------------------
    shard_sizes = None
    if world > 1 and not args.no_balance:
        # shard boundaries equalised by measured launch time (setup, like the neighbour search): the head of the ordering
        # is more expensive per location for Matern, and the slowest rank paces every iteration (spnngp_b200/dist.py)
        theta_b = [5.0, 1.0, 6.0, 0.75][: 4 if cov == 2 else 3]

        def gather(x):
            t = torch.zeros(world, dtype=torch.float64, device=dev)
            t[rank] = x
            dist.all_reduce(t)
            return t.cpu().tolist()

        lo, hi = balanced_shard_range(
            eng, n, m, rank, world,
            lambda a, b: eng.set_data(wl["coords"], nnIndx, lu, wl["X"], wl["y"], m, loc_range=(a, b)),
            lambda: (eng.time_bf_loglik(cov, [theta_b], [wl["beta"]], iters=3), eng.time_bf_loglik(cov, [theta_b], [wl["beta"]], iters=20))[1],
            gather)
        shard_sizes = [int(v) for v in gather(float(hi - lo))]
    eng.set_data(wl["coords"], nnIndx, lu, wl["X"], wl["y"], m, loc_range=(lo, hi))
    W = max(args.warmup, 3)
    chain(W, args.mode)  # first touch of every buffer and code path
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms, acc, launches = timed_chain(args.steps, args.mode, W)
    ms_alt, _, _ = timed_chain(args.steps, 1 - args.mode, W)
    # for comparison: a whole nngp_rnngp call of K iterations, its one-off initial solve and first cross forms included
    barrier()
    ew0 = torch.cuda.Event(enable_timing=True); ew1 = torch.cuda.Event(enable_timing=True)
    ew0.record(stream)
    chain(args.steps, args.mode)
    ew1.record(stream)
    barrier()
    ms_call = torch.tensor([ew0.elapsed_time(ew1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_call, op=dist.ReduceOp.MAX)
    ms_call = float(ms_call.item())
    clk = clocks.stop() if rank == 0 else {}
    value = args.steps / (ms * 1e-3)

    # ---- end to end: host buffers in, samples out (H2D of the data set inside the timed region) ---------
    barrier()
    t0 = time.perf_counter()
    eng.set_data(wl["coords"], nnIndx, lu, wl["X"], wl["y"], m, loc_range=(lo, hi))
    t_set = time.perf_counter() - t0
    _, _, acc_e = chain(args.steps, args.mode)
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    data_bytes = wl["coords"].nbytes + nnIndx.nbytes + wl["X"].nbytes + wl["y"].nbytes
    nb = 2 if args.mode == 0 else 1
    h2d_step = data_bytes / args.steps + nb * (2 * 8 + 0)           # data set amortised + beta per launch
    d2h_step = nb * (2 * 8 + 4) + (acc_e / args.steps) * 9 * 8 + (2 + 3) * 8  # partials, S on accepts, samples

    # ---- N > 1: self-check of the sharded numbers against the unsharded engine on rank 0 -------------------
    parity = None
    theta0 = [5.0, 1.0, 6.0, 0.75][: 4 if cov == 2 else 3]
    if world > 1 and not args.no_parity:
        out_sh = eng.bf_loglik(cov, [theta0], [wl["beta"]], slots=[0])           # every rank: summed over the shards
        _, ts_sh, _ = chain(args.steps, args.mode)
        if rank == 0:
            full = Engine(local_rank, stream=stream.cuda_stream)
            full.set_data(wl["coords"], nnIndx, lu, wl["X"], wl["y"], m)
            out_1 = full.bf_loglik(cov, [theta0], [wl["beta"]], slots=[0])
            R.rshim_set_seed(1)
            _, ts_1, _ = full.rnngp(cov, pri["sigmaSqIG"], pri["tauSqIG"], pri["phiUnif"], pri["nuUnif"], starting, tuning,
                                    args.steps, R.rnorm, R.runif, mode=args.mode)
            full.close()
            parity = {"rel_logdet": float(abs(out_sh[0, 0] - out_1[0, 0]) / abs(out_1[0, 0])),
                      "rel_q": float(abs(out_sh[0, 1] - out_1[0, 1]) / abs(out_1[0, 1])),
                      "chain_max_rel_diff_to_n1": float(np.max(np.abs(ts_sh - ts_1) / np.abs(ts_1))),
                      "chain_equal_to_n1": bool(np.allclose(ts_sh, ts_1, rtol=1e-9, atol=0.0)),
                      "note": "sharded vs unsharded engine on rank 0, same seed, after the timed region"}
        barrier()

    launches_t = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(launches_t)
    line = None
    if rank == 0:
        # ---- dominant kernel: the fused B/F + log-likelihood launch over this rank's shard ----------------
        theta = [5.0, 1.0, 6.0, 0.75][: 4 if cov == 2 else 3]
        eng.time_bf_loglik(cov, [theta], [wl["beta"]], iters=3)
        k_ms = eng.time_bf_loglik(cov, [theta], [wl["beta"]], iters=20)
        nloc = hi - lo
        loc_s = nloc / (k_ms * 1e-3)
        ab = {}
        if cov == 2:  # A/B of the launch-level series choice (csrc/matern_series.cuh): same launch with the 14-term series
            ab["matern_short_series_used"] = bool(eng.get_option("matern_short_used"))
            eng.set_option("matern_short", 0)
            eng.time_bf_loglik(cov, [theta], [wl["beta"]], iters=3)
            ab["kernel_ms_full_series"] = eng.time_bf_loglik(cov, [theta], [wl["beta"]], iters=20)
            eng.set_option("matern_short", 1)
        fp64_peak = eng.fp64_peak_tflops()
        peaks, peak_src = measured_peaks()
        fl = flops_per_location(m)
        tri = bool(eng.get_option("tri_cache_used"))
        ach_tf = loc_s * fl / 1e12
        ach_gbs = loc_s * bytes_per_location(m, tri) / 1e9
        roofline = {
            "bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
            "traffic": traffic_from_profile(n, m, args.cov),
            "kernel": "bf_tri_kernel" if tri else ("bf_ws_kernel" if 6 <= m <= 24 else ("bf_tpl_kernel" if m <= 5 else "bf_kernel")),
            "kernel_ms": k_ms, "locations_per_launch": nloc, "updateBF_loc_per_s": loc_s,
            "flops_per_location": fl, "bytes_per_location": bytes_per_location(m, tri),
            "peak_source": "FP64 FMA peak measured live by nngp_fp64_peak (DFMA micro-benchmark; MEASURED_PEAKS.json has no "
                           "FP64 figure); transcendentals count as 1 flop (SURVEY §8d)",
            "hbm_view": {"achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach_gbs / peaks["hbm_gbs"],
                         "peak_source": peak_src},
        }
        # also the other common covariance at the same size (one line each, kernel only)
        other = "exponential" if args.cov != "exponential" else "matern"
        th_o = [5.0, 1.0, 6.0, 0.75][: 4 if COV[other] == 2 else 3]
        eng.time_bf_loglik(COV[other], [th_o], [wl["beta"]], iters=3)
        k_ms_o = eng.time_bf_loglik(COV[other], [th_o], [wl["beta"]], iters=10)

        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_baseline = cpu_baseline_leg(args, wl, nnIndx, lu, cov, starting, tuning, pri)
        line = {
            "metric": "rNNGP_mcmc_iter_per_s", "value": value, "unit": "iter/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_label(n, m, args.cov),
                "n": n, "m": m, "cov_model": args.cov, "p": 2,
                "sampler": "reference-faithful: 2 B/F solves + 2 Q per iteration, one batched launch" if args.mode == 0
                else "double-buffered: B, F of the current theta stay resident; one launch per iteration solves the "
                     "candidate and re-forms Q of the current theta (identical chain, tests/test_gpu_golden_and_glue.py)",
                "l2": f"inputs larger than L2: {bytes_per_location(m) * (hi - lo) / 1e6:.0f} MB touched per launch vs 126 MB L2"
                if bytes_per_location(m) * (hi - lo) > 126e6 else "working set fits L2 at this shard size (no flush; strong scaling)",
                "parallelism": f"locations (and their nnIndx / B / F / distance storage) sharded over {world} GPU(s); "
                               + ("ncclAllReduce of the partials per iteration" if os.environ.get("NNGP_REDUCE") == "nccl" else
                                  "partials stamped into a shared-memory mailbox by the kernel's last CTA, summed in rank order by every host"),
                "timed_region": f"{args.steps} consecutive iterations of one running chain, after {W} warm-up iterations of the same chain "
                                "(boundary taken in the sampler's once-per-iteration hook; barrier + synchronize on both sides); the "
                                "chain's one-off initial solve lies before it -- extra.iter_per_s_whole_call and e2e include it",
                "shard_sizes": shard_sizes,
                "acceptance_rate": acc / args.steps, "rng": "R-compatible Mersenne-Twister + inversion (librshim), seed 1",
            },
            "e2e": {"value": args.steps / e2e_s, "unit": "iter/s", "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                    "set_data_s": t_set,
                    "note": "nngp_set_data from host buffers + K iterations + samples back, wall clock, max over ranks"},
            "gpu_launches": int(launches_t.item()),
            "clocks": clk,
            "parity": parity,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "extra": {
                "iter_per_s_whole_call": args.steps / (ms_call * 1e-3),
                "alt_sampler_iter_per_s": args.steps / (ms_alt * 1e-3),
                "alt_sampler": "double-buffered" if args.mode == 0 else "reference-faithful",
                f"updateBF_loc_per_s_{other}": (hi - lo) / (k_ms_o * 1e-3),
                "knn_cb_setup_s": t_knn, **ab,
            },
        }
        emit_line(line)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cpu_baseline_leg(args, wl, nnIndx, lu, cov, starting, tuning, pri):
    """The reference's own OpenMP path on this box's host cores, bounded to ~10-30 s (rank 0, N = 1)."""
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    n, m = wl["n"], wl["m"]
    theta = [5.0, 1.0, 6.0, 0.75][: 4 if cov == 2 else 3]
    if po.Ref.available():
        ref = po.Ref()
        t0 = time.perf_counter()
        ref.update_bf(wl["coords"], nnIndx, lu, n, m, theta, cov, threads=cores)
        t_bf = time.perf_counter() - t0
        k = int(min(20, max(1, round(15.0 / (2.2 * t_bf)))))
        ref.set_seed(1)
        t0 = time.perf_counter()
        ref.rnngp(wl["y"], wl["X"], wl["coords"], cov, nnIndx, lu, m, pri["sigmaSqIG"], pri["tauSqIG"], pri["phiUnif"],
                  pri["nuUnif"], starting, tuning, k, nThreads=cores)
        dt = time.perf_counter() - t0
        return {"value": k / dt, "unit": "iter/s", "cores": cores, "kind": "reference",
                "bessel_k": BESSEL_NOTE if cov == 2 else None,
                "sample": f"{k} iterations of the reference's rNNGP at full n={n} (plus its initial updateBF), OpenMP {cores} threads",
                "updateBF_loc_per_s": n / t_bf}
    port = po.Port(); port.set_threads(cores)
    t0 = time.perf_counter()
    port.update_bf(wl["coords"], nnIndx, lu, n, m, theta, cov)
    t_bf = time.perf_counter() - t0
    return {"value": 1.0 / (2.2 * t_bf), "unit": "iter/s", "cores": cores, "kind": "port",
            "sample": f"one updateBF pass of the C oracle port at full n={n}; iteration = 2 x updateBF + 2 x Q",
            "updateBF_loc_per_s": n / t_bf}


def cpu_baseline_other_configs(args):
    """`--cpu-configs`: the reference's own CPU code (oracle/_ref; rank 0, no GPU work) on bounded samples of the other
    BASELINE configs, next to which tools/bench_configs.py's GPU timings are read -- SURVEY §8(d) "CPU baseline timing".
    Prints one JSON line."""
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    if not po.Ref.available():
        emit_line({"impl": "reference", "cpu_configs": None, "unavailable": "oracle/_ref not built"})
        return 0
    ref = po.Ref(); port = po.Port(); port.set_threads(cores)
    rng = np.random.default_rng(11)
    out = {"cores": cores, "kind": "reference", "threads": cores}

    def timed(f):
        t = time.perf_counter(); r = f(); return time.perf_counter() - t, r

    def by_x(c):
        return c[np.argsort(c[:, 0], kind="stable")]

    # configs[4]: ordered search
    c = by_x(rng.random((1000000, 2)) * np.sqrt(1000000 / 5000.0))
    t, (idx6, _, lu6) = timed(lambda: ref.mk_nn_indx_cb(c, 15, threads=cores))
    out["mkNNIndxCB_n1e6_m15_s"] = t
    t, _ = timed(lambda: ref.mk_nn_indx_cb(c, 32, threads=cores)); out["mkNNIndxCB_n1e6_m32_s"] = t
    t, _ = timed(lambda: ref.mk_nn_indx(c[:100000], 15, threads=cores)); out["mkNNIndx_brute_n1e5_m15_s"] = t
    # configs[3]: a slice of the prediction job, q = 5,000 sites x S = 1,000 samples at n = 10^6
    n, q, S = 1000000, 5000, 1000
    X = np.column_stack([np.ones(n), rng.standard_normal(n)]); beta = np.array([1.0, 5.0])
    y = X @ beta + rng.standard_normal(n)
    c0 = rng.random((q, 2)) * np.sqrt(n / 5000.0); X0 = np.column_stack([np.ones(q), rng.standard_normal(q)])
    nn0 = port.nn_query(c, c0, 15)
    bs = beta[:, None] + 0.05 * rng.standard_normal((2, S))
    th = np.tile(np.array([[5.0], [1.0], [6.0]]), (1, S))
    ref.set_seed(1)
    t, _ = timed(lambda: ref.predict(X, y, c, 15, X0, c0, nn0, bs, th, 0, nThreads=cores))
    out["rNNGPPredict_q5000_S1000_s"] = t
    out["rNNGPPredict_draws_per_s"] = q * S / t
    # configs[1]: one fold's fit + hold-out prediction for ONE parameter set at BCEF's fold size
    n2, n_ho = 150974, 37743
    c2 = by_x(rng.random((n2, 2)) * np.array([21.5, 17.0]))
    X2 = np.column_stack([np.ones(n2), rng.standard_normal(n2)]); y2 = X2 @ beta + rng.standard_normal(n2)
    t_nn, (idx2, _, lu2) = timed(lambda: ref.mk_nn_indx_cb(c2, 15, threads=cores))
    ch = rng.random((n_ho, 2)) * np.array([21.5, 17.0]); Xh = np.column_stack([np.ones(n_ho), rng.standard_normal(n_ho)])
    nnh = port.nn_query(c2, ch, 15)
    t_fit, _ = timed(lambda: ref.cnngp(y2, X2, c2, np.array([[3.0], [0.25]]), idx2, lu2, Xh, ch, nnh, 15, [2.0, 1.0], 0,
                                        nThreads=cores))
    out["cNNGP_one_set_one_fold_s"] = t_fit
    out["mkNNIndxCB_fold_n151k_s"] = t_nn
    out["spConjNNGP_kfold_5x25_extrapolated_s"] = 5 * (t_nn + 25 * t_fit)
    out["note"] = ("reference C++ (oracle/_ref) with OpenMP on the box's host cores; the prediction and k-fold figures are "
                   "bounded samples, linear in q*S and in folds*sets; the RANN query search is not included")
    emit_line({"impl": "reference", "cpu_configs": out})
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--locations", type=int, default=1000000)
    ap.add_argument("--m", type=int, default=15)
    ap.add_argument("--cov", default="matern", choices=list(COV))
    ap.add_argument("--no-balance", action="store_true", help="N > 1: keep the neighbour-count shard boundaries (no timing-based rebalancing)")
    ap.add_argument("--no-parity", action="store_true", help="skip the N > 1 self-check against the unsharded engine")
    ap.add_argument("--mode", type=int, default=1, help="1 double-buffered sampler (default), 0 reference-faithful")
    ap.add_argument("--ref-budget-s", type=float, default=150.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--py-allreduce", action="store_true", help="sum partials through torch.distributed in Python")
    ap.add_argument("--cpu-configs", action="store_true",
                    help="time the reference's CPU code on samples of BASELINE configs[1], [3], [4] (no GPU work)")
    args = ap.parse_args()
    protect_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.cpu_configs:
        return cpu_baseline_other_configs(args) if rank == 0 else 0
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
