#!/usr/bin/env python3
"""Throughput of the fused path on the other BASELINE.json configurations (sub-batches sized in
whole waves of resident simulations; one B200):
  stiff     config 3: shipped Highfrontsurf/Highbacksurf/Balancedhighsurf observation files (3 files in
            one call, condensed fixtures), stiff prior Sf,Sb in [1,1e5] cm/s, L=128, dt=0.025 ns
  twothick  config 4: 6 curves, Length=[311,2000]x3, T=80000, synthetic observations
  finegrid  config 5: L=1000, T=20000 (500 ns), 3 curves, synthetic observations
usage: bench_configs.py [stiff twothick finegrid] [--waves W]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bayesian_inference_trpl_b200 as trpl
from helpers import TRUTH, UC, example_data, power_scan_excitations, prior_samples

args = [a for a in sys.argv[1:] if a in ("stiff", "twothick", "finegrid")] or ["stiff", "twothick", "finegrid"]
waves = int(sys.argv[sys.argv.index("--waves") + 1]) if "--waves" in sys.argv else 2
ex = example_data()


def synth_obs(simPar, inis, lengths):
    Time, T = simPar[1], simPar[3]
    grid = np.linspace(0, Time, T + 1)
    ts, vs, us = [], [], []
    for c in range(len(inis)):
        sp = list(simPar); sp[0] = lengths[c]
        pl = np.empty((1, T + 1))
        trpl.pvSim(pl, None, None, None, (TRUTH * UC)[None, :12], sp, inis[c], (128,), 0, 1, init_mode="points")
        ts.append(grid.copy()); vs.append(np.log10(pl[0])); us.append(np.full(T + 1, 0.1))
    return [(ts, vs, us)]


def run(name):
    if name == "stiff":
        L, T, C = 128, 80000, 3
        simPar = [2000.0, 2000.0, L, T, 1, (0,), 7, 10000]
        inis = power_scan_excitations(); lengths = [2000.0] * 3
        e_data = []
        for f in ("Highfrontsurf", "Highbacksurf", "Balancedhighsurf"):
            ts = [ex["%s_t%d" % (f, c)] for c in range(3)]
            vs = [np.log10(ex["%s_pl%d" % (f, c)] * 1e-23) for c in range(3)]
            e_data.append((ts, vs, [np.full(len(t), 0.1) for t in ts]))
        stiff = True
    elif name == "twothick":
        L, T, C = 128, 80000, 6
        lengths = [311.0, 2000.0] * 3
        simPar = [lengths, 2000.0, L, T, 1, (0,), 7, 10000]
        inis = ex["twothick_exc"] * 1e-21
        e_data = synth_obs(simPar, inis, lengths); stiff = False
    else:
        L, T, C = 1000, 20000, 3
        simPar = [2000.0, 500.0, L, T, 1, (0,), 7, 10000]
        xc = (np.arange(L) + 0.5) * (2000.0 / L)
        inis = np.stack([a * 1e-21 * np.exp(-6e-3 * xc) for a in (1.2738e16, 1.1539e17, 1.6485e18)])
        lengths = [2000.0] * 3
        e_data = synth_obs(simPar, inis, lengths); stiff = False
    res = trpl.engine.resident_sims(L, 0)
    S = max(1, waves * res // C)
    X = prior_samples(S, seed=17, stiff=stiff)
    prob = trpl.engine.Problem(simPar, inis, e_data, device=0)
    Xd = torch.from_numpy(X).cuda()
    best = None
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lnl, st, it = trpl.engine.solve_loglik(Xd, prob, want_iters=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    steps = prob.steps_per_sample()
    iters = float(it.sum().item())
    flops = L * (29.0 * steps * S + 126.0 * iters)
    print("%-9s L=%d C=%d E=%d  S=%d (%d waves of %d resident sims)  steps/sample %d  %.1f ms  -> %.1f likelihoods/s, "
          "%.2f TFLOP/s algorithmic, %.3f Newton iters/step, non-converged %d"
          % (name, L, C, len(e_data), S, waves, res, steps, best, S / best * 1e3, flops / best / 1e9,
             iters / (steps * S), int((st != 0).sum().item())), flush=True)


for n in args:
    run(n)
