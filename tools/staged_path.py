#!/usr/bin/env python3
"""Throughput of the reference-style STAGED pipeline through the drop-in callables
(pvSim -> fastlog -> host interpolation -> prob, float32 PL buffer, sims_per_gpu = 1024,
bayeslib.py:117-201) next to the fused path, same samples, same observations."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bayesian_inference_trpl_b200 as trpl
from helpers import TRUTH, UC, power_scan_excitations, prior_samples
L, T = 128, 80000
simPar = [2000.0, 2000.0, L, T, 1, (0,), 7, 10000]
inis = power_scan_excitations()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
X = prior_samples(S, seed=5)
grid = np.linspace(0, 2000.0, T + 1)
ts, vs, us = [], [], []
for c in range(3):
    pl = np.empty((1, T + 1))
    trpl.pvSim(pl, None, None, None, (TRUTH * UC)[None, :12], simPar, inis[c], (128,), 0, 1, init_mode="points")
    ts.append(grid[::10].copy()); vs.append(np.log10(pl[0][::10])); us.append(np.full(len(grid[::10]), 0.1))   # 8001 obs/curve
e_data = [(ts, vs, us)]
sim_flags = {"load_PL_from_file": False, "log_pl": True, "self_normalize": False}
res = {}
for fused in (True, False):
    gpu_info = {"has_GPU": True, "sims_per_gpu": 1024, "num_gpus": 1, "device": 0, "fused": fused,
                "threads_per_block": (128,), "max_sims_per_block": 1}
    P = np.zeros((1, S))
    tm = [np.zeros(1), np.zeros(1), np.zeros(1)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    trpl.bayeslib.simulate(trpl.pvSim, e_data, P, X, [None], [None], 3, list(simPar), inis, sim_flags, gpu_info, 0, tm[0], tm[1], tm[2])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[fused] = P.copy()
    print("%-7s S=%d: %.2f s -> %.1f likelihoods/s   (solver %.2f s, log/interp %.2f s, prob %.2f s)"
          % ("fused" if fused else "staged", S, dt, S / dt, tm[0][0], tm[2][0], tm[1][0]))
ok = np.isfinite(res[False]) & np.isfinite(res[True])
d = np.abs(res[True][ok] - res[False][ok]) / np.abs(res[False][ok])
print("fused (f64) vs staged (float32 PL buffer, log10f): max rel diff of lnL %.2e over %d samples; "
      "%d samples are -inf/NaN in the float32 pipeline (PL underflows float32, SURVEY Q5)" % (d.max(), ok.sum(), (~ok).sum()))
