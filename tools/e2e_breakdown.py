#!/usr/bin/env python3
"""Where does the time of one bayeslib.simulate call with host arrays go, next to the device-resident
launch of the same batch?  (bench.py `e2e` vs `value`)"""
import cProfile, io, os, pstats, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bayesian_inference_trpl_b200 as trpl
from helpers import TRUTH, UC, power_scan_excitations, prior_samples
L, T = 128, int(sys.argv[1]) if len(sys.argv) > 1 else 20000
simPar = [2000.0, 0.025 * T, L, T, 1, (0,), 7, 10000]
inis = power_scan_excitations()
S = 2 * trpl.engine.resident_sims(L, 0)
X = prior_samples(S, seed=3)
grid = np.linspace(0, simPar[1], T + 1)
e_data = [([grid.copy() for _ in range(3)], [np.linspace(-7, -12, T + 1) for _ in range(3)], [np.full(T + 1, .1)] * 3)]
flags = {"load_PL_from_file": False, "log_pl": True, "self_normalize": False}
info = {"has_GPU": True, "sims_per_gpu": S, "num_gpus": 1, "device": 0, "threads_per_block": (128,), "max_sims_per_block": 1}
P = np.zeros((1, S))
def e2e():
    P[:] = 0
    tm = [np.zeros(1), np.zeros(1), np.zeros(1)]
    trpl.bayeslib.simulate(trpl.pvSim, e_data, P, X, [None], [None], 3, list(simPar), inis, flags, info, 0, *tm)
    return tm[0][0]
prob = trpl.engine.Problem(simPar, inis, e_data, device=0)
Xd = torch.from_numpy(X).cuda()
def resident():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); trpl.engine.solve_loglik(Xd, prob); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
_orig = trpl.engine.solve_loglik
_ev = []
def _timed(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = _orig(*a, **k); e1.record(); _ev.append((e0, e1)); return r
trpl.engine.solve_loglik = _timed
e2e(); resident()
print("sequence test: e2e x3 then resident x3 then alternating")
for f in (e2e, e2e, e2e, resident, resident, resident):
    _ev.clear(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); w = time.perf_counter() - t0
    print("  %-9s wall %.2f ms, kernel (events) %.2f ms" % (f.__name__, 1e3 * w, _ev[-1][0].elapsed_time(_ev[-1][1])))
for rep in range(3):
    t0 = time.perf_counter(); inner = e2e(); torch.cuda.synchronize(); w = time.perf_counter() - t0
    r = resident()
    print("rep %d: simulate() wall %.2f ms (solve+copies inside %.2f ms) | resident launch %.2f ms" % (rep, 1e3 * w, 1e3 * inner, r))
pr = cProfile.Profile(); pr.enable(); e2e(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14); print(s.getvalue()[:3500])
