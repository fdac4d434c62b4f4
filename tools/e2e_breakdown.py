#!/usr/bin/env python3
"""Where does the host-side time of one bayeslib.simulate call go? (bench e2e vs device-resident)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bayesian_inference_trpl_b200 as trpl
from helpers import TRUTH, UC, power_scan_excitations, prior_samples
L, T = 128, 80000
simPar = [2000.0, 2000.0, L, T, 1, (0,), 7, 10000]
inis = power_scan_excitations()
S = 64
X = prior_samples(S, seed=3)
grid = np.linspace(0, 2000.0, T + 1)
e_data = [([grid.copy() for _ in range(3)], [np.linspace(-7, -12, T + 1) for _ in range(3)], [np.full(T + 1, .1)] * 3)]
def tick(label, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); print("%-34s %8.2f ms" % (label, 1e3 * (t1 - t0))); return t1
for rep in range(2):
    t0 = time.perf_counter()
    prob = trpl.engine.Problem(simPar, inis, e_data, device=0); t0 = tick("Problem (obs prepare + upload)", t0)
    Xd = trpl.engine.to_device_f64(X, prob.dev); t0 = tick("X pinned staging + H2D", t0)
    lnl, st, _ = trpl.engine.solve_loglik(Xd, prob); t0 = tick("solve_loglik (S=64: latency of 1 wave)", t0)
    h = lnl.cpu().numpy(); t0 = tick("lnL D2H", t0)
    bad = int((st != 0).sum().item()); t0 = tick("status reduce", t0)
