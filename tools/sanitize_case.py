#!/usr/bin/env python3
"""Tiny fused + PL launches of both solver kernels (warp kernel L=128/96, CTA kernel L=260) for
compute-sanitizer."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bayesian_inference_trpl_b200 as trpl
from helpers import TRUTH, UC, prior_samples
for L, T, S in ((128, 40, 9), (96, 40, 5), (260, 24, 3), (8, 20, 4)):
    length = 2000.0
    simPar = [length, 0.025 * T, L, T, 1, (0,), 7, 10000]
    x = (np.arange(L) + 0.5) * (length / L)
    inis = np.stack([a * 1e-21 * np.exp(-6e-3 * x) for a in (1.2e16, 1.6e18)])
    X = prior_samples(S, seed=L, mag=True)
    grid = np.linspace(0, simPar[1], T + 1)
    e_data = [([grid[::2].copy(), grid.copy()], [np.full(len(grid[::2]), -7.0), np.full(T + 1, -6.0)], [None, None])]
    prob = trpl.engine.Problem(simPar, inis, e_data, device=0)
    lnl, st, it = trpl.engine.solve_loglik(torch.from_numpy(X).cuda(), prob, want_iters=True)
    pl = np.empty((S, T + 1), dtype=np.float32)
    trpl.pvSim(pl, None, None, None, X[:, :12], simPar, inis[0], (128,), 0, 1, init_mode="points")
    torch.cuda.synchronize()
    print(L, lnl.cpu().numpy()[0, :3], pl[0, :2], int(st.sum().item()))
