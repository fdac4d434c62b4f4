#!/usr/bin/env python3
"""How late does the host learn that a long kernel has finished?  (bench.py e2e vs value)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bayesian_inference_trpl_b200 as trpl
from helpers import power_scan_excitations, prior_samples
L, T = 128, 8000
simPar = [2000.0, 0.025 * T, L, T, 1, (0,), 7, 10000]
inis = power_scan_excitations()
S = 2 * trpl.engine.resident_sims(L, 0)
grid = np.linspace(0, simPar[1], T + 1)
e_data = [([grid.copy() for _ in range(3)], [np.linspace(-7, -12, T + 1) for _ in range(3)], [np.full(T + 1, .1)] * 3)]
prob = trpl.engine.Problem(simPar, inis, e_data, device=0)
Xd = torch.from_numpy(prior_samples(S, seed=3)).cuda()
st = torch.cuda.current_stream()
def wait_sync(ev): torch.cuda.synchronize()
def wait_event(ev): ev.synchronize()
def wait_stream(ev): st.synchronize()
def wait_spin(ev):
    while not ev.query():
        pass
print("cpus", len(os.sched_getaffinity(0)), "load", os.getloadavg())
for name, w in (("device sync", wait_sync), ("event sync", wait_event), ("stream sync", wait_stream), ("spin on query", wait_spin)):
    late = []
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); trpl.engine.solve_loglik(Xd, prob); e1.record(); w(e1); wall = time.perf_counter() - t0
        late.append(1e3 * wall - e0.elapsed_time(e1))
    print("%-14s host learns of completion late by (ms): %s" % (name, " ".join("%.1f" % v for v in late)))
