#!/usr/bin/env python3
"""Short fused-likelihood launch for ncu / quick timing: power-scan shape with a reduced number
of time steps (same kernel, same per-step work).  usage: profile_case.py [T] [S] [reps] [L]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bayesian_inference_trpl_b200 as trpl  # noqa: E402
from helpers import TRUTH, UC, power_scan_excitations, prior_samples  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L = int(sys.argv[4]) if len(sys.argv) > 4 else 128
simPar = [2000.0, 0.025 * T, L, T, 1, (0,), 7, 10000]
if L == 128:
    inis = power_scan_excitations()
else:   # BASELINE config 5: dN_c(x) = A_c exp(-6e-3 x) at the cell centres
    xc = (np.arange(L) + 0.5) * (2000.0 / L)
    inis = np.stack([a * 1e-21 * np.exp(-6e-3 * xc) for a in (1.2738e16, 1.1539e17, 1.6485e18)])
res = trpl.engine.resident_sims(L, 0)
if S <= 0:
    S = res
X = prior_samples(S, seed=99)
grid = np.linspace(0, simPar[1], T + 1)
ts, vs, us = [], [], []
for c in range(3):
    pl = np.empty((1, T + 1))
    trpl.pvSim(pl, None, None, None, (TRUTH * UC)[None, :12], simPar, inis[c], (128,), 0, 1, init_mode="points")
    ts.append(grid.copy()); vs.append(np.log10(pl[0])); us.append(np.full(T + 1, 0.1))
prob = trpl.engine.Problem(simPar, inis, [(ts, vs, us)], device=0)
Xd = torch.from_numpy(X).cuda()
for r in range(reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lnl, st, it = trpl.engine.solve_loglik(Xd, prob, want_iters=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    iters = float(it.sum().item())
    flops = L * (29.0 * (T + 1) * 3 * S + 126.0 * iters)
    print("T=%d S=%d resident=%d: %.2f ms, %.1f sim-steps/us, %.2f TFLOP/s algorithmic, iters/step %.3f, bad %d"
          % (T, S, res, ms, 3 * S * (T + 1) / ms / 1e3, flops / ms / 1e9, iters / (3 * S * (T + 1)),
             int((st != 0).sum().item())), flush=True)
