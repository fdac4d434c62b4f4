import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bayesian_inference_trpl_b200 as trpl
from helpers import TRUTH, UC, power_scan_excitations, prior_samples
L, T = 128, 20000
simPar = [2000.0, 500.0, L, T, 1, (0,), 7, 10000]
inis = power_scan_excitations()
S = 4736
X = prior_samples(S, seed=1234)
grid = np.linspace(0, 500.0, T + 1)
e_data = [([grid.copy() for _ in range(3)], [np.linspace(-7, -12, T + 1) for _ in range(3)], [np.full(T + 1, .1)] * 3)]
prob = trpl.engine.Problem(simPar, inis, e_data, device=0)
Xd = torch.from_numpy(X).cuda()
sim_flags = {"load_PL_from_file": False, "log_pl": True, "self_normalize": False}
gpu_info = {"has_GPU": True, "sims_per_gpu": S, "num_gpus": 1, "device": 0, "threads_per_block": (128,), "max_sims_per_block": 1}
P = np.zeros((1, S))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lnl, st, _ = trpl.engine.solve_loglik(Xd, prob)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    tm = [np.zeros(1), np.zeros(1), np.zeros(1)]
    trpl.bayeslib.simulate(trpl.pvSim, e_data, P, X, [None], [None], 3, list(simPar), inis, sim_flags, gpu_info, 0, tm[0], tm[1], tm[2])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("resident %.1f ms   simulate() %.1f ms  (inner solver_time %.1f ms)" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * tm[0][0]))
    print("   same result:", np.array_equal(P[0] / (rep + 1), lnl.cpu().numpy()[0]))
