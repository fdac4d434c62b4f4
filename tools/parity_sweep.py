#!/usr/bin/env python3
"""Parity statistics of the CUDA path against the CPU oracle on a larger random sample of the
default (and stiff) prior: PL curves, Newton iteration counts and lnL.  Test infrastructure."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bayesian_inference_trpl_b200 as trpl
from oracle import oracle
from helpers import TRUTH, UC, pl_noise_floor, power_scan_excitations, prior_samples

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
L = 128
inis = power_scan_excitations()
dev = torch.device("cuda", 0)
for label, length, stiff in (("default prior, 2000 nm", 2000.0, False), ("stiff prior (S up to 1e5 cm/s), 311 nm", 311.0, True)):
    simPar = [length, 0.025 * T, L, T, 1, (0,), 7, 10000]
    X = prior_samples(S, seed=2024, stiff=stiff, mag=True)
    X[0] = TRUTH * UC
    print("== %s: S=%d samples x 3 curves, L=%d, T=%d" % (label, S, L, T))
    truth_pl = []
    worst = 0.0; n_pts = 0; n_in = 0; it_equal = 0; it_tot = 0; it_maxdiff = 0
    yard_worst = 0.0; yard_in = 0
    for c in range(3):
        t0 = time.time()
        ref = oracle.solve(X[:, :12], simPar, inis[c], solver="pcr")
        t_cpu = time.time() - t0
        ref_t = oracle.solve(X[:, :12], simPar, inis[c], solver="thomas")     # yardstick: oracle vs oracle
        mat = torch.from_numpy(np.ascontiguousarray(X[:, :12])).to(dev)
        pl, st, it = trpl.engine.solve_pl(mat, torch.from_numpy(inis[c]).to(dev), length, simPar[1], L, T)
        torch.cuda.synchronize()
        pl = pl.cpu().numpy(); it = it.cpu().numpy()
        truth_pl.append(ref["pl"][0])
        floor = pl_noise_floor(X[:, :12], length, simPar[1], L, T)[:, None]
        sig = np.abs(ref["pl"]) > 1e3 * floor                 # points above the cancellation noise floor
        rel = np.abs(pl - ref["pl"]) / np.abs(ref["pl"])
        worst = max(worst, rel[sig].max()); n_pts += sig.sum(); n_in += (rel[sig] <= 1e-6).sum()
        rel_t = np.abs(ref_t["pl"] - ref["pl"]) / np.abs(ref["pl"])
        yard_worst = max(yard_worst, rel_t[sig].max()); yard_in += (rel_t[sig] <= 1e-6).sum()
        it_equal += (it == ref["iters"]).sum(); it_tot += S; it_maxdiff = max(it_maxdiff, np.abs(it - ref["iters"]).max())
        assert (st.cpu().numpy() == ref["status"]).all()
        print("   curve %d: max rel PL err %.3e (above noise floor), oracle %.1fs" % (c, rel[sig].max(), t_cpu))
    print("   PL: %d points compared, %.6f %% within 1e-6, worst %.3e" % (n_pts, 100.0 * n_in / n_pts, worst))
    print("   yardstick (oracle Thomas vs oracle PCR, both CPU FP64): %.6f %% within 1e-6, worst %.3e"
          % (100.0 * yard_in / n_pts, yard_worst))
    print("   Newton iteration totals identical for %d / %d simulations (max |diff| %d)" % (it_equal, it_tot, it_maxdiff))
    grid = np.linspace(0, simPar[1], T + 1)
    e_data = [([grid.copy()] * 3, [np.log10(p) for p in truth_pl], [np.full(T + 1, 0.1)] * 3)]
    ref_l = oracle.loglik(X, simPar, inis, e_data, solver="pcr")
    ref_lt = oracle.loglik(X, simPar, inis, e_data, solver="thomas")
    prob = trpl.engine.Problem(simPar, inis, e_data, device=0)
    lnl, st, _ = trpl.engine.solve_loglik(torch.from_numpy(X).to(dev), prob)
    got = lnl.cpu().numpy()
    ok = np.isfinite(ref_l[0]) & (np.abs(ref_l[0]) > 1e-6)
    rl = np.abs(got[0][ok] - ref_l[0][ok]) / np.abs(ref_l[0][ok])
    rt = np.abs(ref_lt[0][ok] - ref_l[0][ok]) / np.abs(ref_l[0][ok])
    print("   lnL yardstick (oracle Thomas vs PCR): max rel %.3e, within 1e-6: %.4f %%" % (rt.max(), 100.0 * (rt <= 1e-6).mean()))
    print("   lnL: %d samples, max rel err %.3e, median %.3e, within 1e-6: %.4f %%; non-finite in both: %d"
          % (ok.sum(), rl.max(), np.median(rl), 100.0 * (rl <= 1e-6).mean(), (~np.isfinite(ref_l[0]) & ~np.isfinite(got[0])).sum()))
