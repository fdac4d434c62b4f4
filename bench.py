#!/usr/bin/env python3
"""Headline benchmark: 3-curve power-scan parameter-sample likelihoods per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--samples S] [--impl ours|reference]

Workload (BASELINE.json configs[1]): L=128 nodes, Length=2000 nm, Time=2000 ns,
T=80000 implicit steps, tol=7, MAX=10000, the 3 Power_scan excitations, default prior of the
reference entry script, synthetic observations (PL of the truth sample, 80001 points per curve).
One "step" = one pass of the fused forward-model+likelihood path over a batch of S samples per
GPU (S*3 simulations); the nominal 1M-sample configuration is this step repeated, so throughput
is reported on whole-wave batches (S defaults to 2x the number of simulations resident on the
GPU).  Under torchrun every rank owns its own S samples (weak scaling, no data-path collective)
and the step ends with the lnL all-gather + global log-sum-exp over NCCL.

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, Thomas solver,
OpenMP over all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

L, T, LENGTH, TIME, TOL, MAXIT = 128, 80000, 2000.0, 2000.0, 7, 10000
SIMPAR = [LENGTH, TIME, L, T, 1, (0,), TOL, MAXIT]
METRIC = "param-sample likelihoods/sec (3-curve power scan)"
UNIT = "likelihoods/s"
FLOP_STEP, FLOP_ITER = 29, 126          # per node: per time step / per Newton iteration


def workload_desc(S):
    return ("power_scan L=128 T=80000 Length=2000nm 3 curves x 80001 obs, default prior; "
            "batch of %d samples/GPU/step out of the 1M-sample config" % S)


def inputs(S, seed):
    from helpers import TRUTH, UC, power_scan_excitations, prior_samples
    X = prior_samples(S, seed=seed)
    return X, power_scan_excitations(), TRUTH * UC


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows),
                "reasons": reasons}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(n_samples, threads):
    """Oracle (port of the reference algorithm, Thomas solve) on `threads` host threads."""
    from oracle import oracle
    X, inis, _ = inputs(n_samples, seed=777)
    t0 = time.perf_counter()
    for c in range(3):
        oracle.solve(X[:, :12], SIMPAR, inis[c], solver="thomas", nthreads=threads)
    dt = time.perf_counter() - t0
    return n_samples / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    threads = host_threads()          # torchrun pins OMP_NUM_THREADS=1; use every core the process may run on
    n = max(threads, 1)
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_baseline(max(1, threads // 4), threads)
    times = []
    for _ in range(args.steps):
        v, dt = cpu_baseline(n, threads)
        times.append(dt)
    tot = sum(times)
    value = n * args.steps / tot
    sample = "%d samples x 3 curves at full T=80000 per step, oracle Thomas solver, %d OpenMP threads" % (n, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_desc(n)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--samples", type=int, default=0, help="samples per GPU per step (0 = 2 waves x resident / 3 curves... auto)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import bayesian_inference_trpl_b200 as trpl
    from bayesian_inference_trpl_b200 import distributed as D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    resident = trpl.engine.resident_sims(L, local)
    S = args.samples if args.samples > 0 else (2 * resident)          # S*3 sims = 6 waves
    X, inis, truth = inputs(S, seed=1234 + rank)

    # synthetic observations: PL of the truth sample on the full step grid, from the engine itself
    grid = np.linspace(0, TIME, T + 1)
    ts, vs, us = [], [], []
    for c in range(3):
        pl = np.empty((1, T + 1))
        trpl.pvSim(pl, None, None, None, truth[None, :12], SIMPAR, inis[c], (128,), 0, 1, init_mode="points")
        ts.append(grid.copy()); vs.append(np.log10(pl[0])); us.append(np.full(T + 1, 0.1))
    e_data = [(ts, vs, us)]
    problem = trpl.engine.Problem(SIMPAR, inis, e_data, device=local)

    X_pin = torch.from_numpy(X).pin_memory()
    Xd = torch.empty_like(X_pin, device=dev)
    Xd.copy_(X_pin)
    lnl_host = torch.empty((1, S), dtype=torch.float64).pin_memory()
    l2_flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def exchange(lnl):
        if world == 1:
            return lnl
        parts = [torch.empty_like(lnl) for _ in range(world)]
        dist.all_gather(parts, lnl)
        full = torch.cat(parts, dim=-1)
        D.global_logsumexp(trpl.engine.lse_partial(lnl[0].contiguous()))
        return full

    def step_resident(want_iters=False):
        lnl, status, iters = trpl.engine.solve_loglik(Xd, problem, want_iters=want_iters)
        exchange(lnl)
        return lnl, status, iters

    # end to end = the call a user of the reference makes: bayeslib.simulate(model, e_data, P, X, ...)
    # with HOST numpy arrays (bayeslib.py:83); per call it stages X, the excitations and the
    # bracketed observations to the device, runs the fused kernel and reads lnL back into P.
    sim_flags = {"load_PL_from_file": False, "log_pl": True, "self_normalize": False}
    gpu_info = {"has_GPU": True, "sims_per_gpu": S, "num_gpus": 1, "device": local,
                "threads_per_block": (128,), "max_sims_per_block": 1}
    P_host = np.zeros((1, S))
    obs_bytes = sum(len(t) * (4 + 8 + 8 + 8) for t in ts)
    h2d_bytes = int(X.size * 8 + inis.size * 8 + obs_bytes)
    d2h_bytes = int(S * 8 + S * 4)

    def step_e2e():
        P_host[:] = 0.0
        tm = [np.zeros(1), np.zeros(1), np.zeros(1)]
        trpl.bayeslib.simulate(trpl.pvSim, e_data, P_host, X, [None], [None], 3, list(SIMPAR), inis,
                               sim_flags, gpu_info, 0, tm[0], tm[1], tm[2])
        if world > 1:
            exchange(torch.from_numpy(P_host).to(dev))
        return P_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up
    for _ in range(max(args.warmup, 1)):
        l2_flush.zero_()
        lnl, status, iters = step_resident(want_iters=True)
    torch.cuda.synchronize(dev)
    n_bad = int((status != 0).sum().item())
    iters_total = iters.sum(dim=1).cpu().numpy().astype(np.float64)          # per curve
    flops_per_step = float(L * sum(FLOP_STEP * (T + 1) * S + FLOP_ITER * iters_total[c] for c in range(3)))

    # ---- timed region: device-resident inputs
    sampler = ClockSampler(local)
    sampler.start()
    kern_ms = []
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        l2_flush.zero_()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        lnl, status, _ = trpl.engine.solve_loglik(Xd, problem)
        k1.record()
        exchange(lnl)
        kern_ms.append((k0, k1))
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    kern = [a.elapsed_time(b) for a, b in kern_ms]
    clocks = sampler.summary()

    # ---- end to end: host buffers in, host lnL out, copies inside the timed region
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        l2_flush.zero_()
        step_e2e()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    t = torch.tensor([ms_total, ms_e2e, float(np.mean(kern))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, kern_mean = [float(v) for v in t.cpu()]

    if rank == 0:
        value = S * world * args.steps / (ms_total * 1e-3)
        e2e = S * world * args.steps / (ms_e2e * 1e-3)
        tf_peak, _ = trpl.engine.bench_dfma(20000, local)
        achieved = flops_per_step / (kern_mean * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None
        try:   # DRAM bytes of one bench-sized launch, from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_desc(S), "samples_per_gpu_per_step": S,
                       "sims_resident_per_gpu": resident, "l2": "256 MiB L2 flush between steps",
                       "nonconverged_samples": n_bad,
                       "mean_newton_iters_per_step": float(iters_total.sum() / (3.0 * S * (T + 1)))},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes,
                    "api": "bayeslib.simulate(model, e_data, P, X, ...) with host numpy arrays"},
            "gpu_launches": (2 + (3 if world > 1 else 0)) * args.steps,   # sim + finish (+ 3 lse kernels when sharded)
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                         "frac": achieved / tf_peak, "traffic": traffic,
                         "note": "neither HBM- nor tensor-bound: scalar FP64 with no dense contraction "
                                 "(north_star); HBM side given in hbm_* keys, see DESIGN.md section 4",
                         "kernel": "trpl_sim_kernel<4,false>", "kernel_ms": kern_mean,
                         "flops_per_launch": flops_per_step,
                         "peak_source": "DFMA microbenchmark (trpl_bench_dfma) measured in this run; "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "hbm_bytes_per_launch_algorithmic": int(S * (13 + 1) * 8),
                         "hbm_peak_gbs_measured": peaks.get("hbm_gbs")},
        }
        if not args.no_cpu_baseline and world == 1:
            threads = host_threads()
            n_cpu = 4 * threads                    # ~10 s of CPU work
            v, dt = cpu_baseline(n_cpu, threads)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                   "sample": "%d samples x 3 curves at full T=80000 (%.1f s), oracle "
                                             "Thomas solver, %d OpenMP threads" % (n_cpu, dt, threads)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
