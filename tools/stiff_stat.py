#!/usr/bin/env python3
"""Stiff-regime agreement statistic (tests/test_gpu_stiff.py) for several library builds.
usage: stiff_stat.py name ...   (build/variants/libtrpl_<name>.so)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAR = os.path.join(ROOT, "build", "variants")
if len(sys.argv) > 2 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np, torch
    import bayesian_inference_trpl_b200 as trpl
    from helpers import TRUTH, UC, pl_noise_floor, power_scan_excitations, prior_samples
    S, T, L, length = 256, 8000, 128, 311.0
    simPar = [length, 0.025 * T, L, T, 1, (0,), 7, 10000]
    inis = power_scan_excitations()
    X = prior_samples(S, seed=2024, stiff=True, mag=True); X[0] = TRUTH * UC
    cache = "/tmp/stiff_oracle.npz"
    if not os.path.exists(cache):
        from oracle import oracle
        d = {}
        for c in range(3):
            r = oracle.solve(X[:, :12], simPar, inis[c], solver="pcr"); y = oracle.solve(X[:, :12], simPar, inis[c], solver="thomas")
            d["ref%d" % c] = r["pl"]; d["yard%d" % c] = y["pl"]; d["it%d" % c] = r["iters"]
        np.savez(cache, **d)
    d = np.load(cache)
    mat = torch.from_numpy(np.ascontiguousarray(X[:, :12])).cuda()
    floor = pl_noise_floor(X[:, :12], length, simPar[1], L, T)[:, None]
    n = nm = ny = ie = 0
    for c in range(3):
        pl, st, it = trpl.engine.solve_pl(mat, torch.from_numpy(inis[c]).cuda(), length, simPar[1], L, T)
        pl, it = pl.cpu().numpy(), it.cpu().numpy()
        ref, yard = d["ref%d" % c], d["yard%d" % c]
        sig = np.abs(ref) > 1e3 * floor
        n += sig.sum(); nm += (np.abs(pl - ref)[sig] <= 1e-6 * np.abs(ref)[sig]).sum()
        ny += (np.abs(yard - ref)[sig] <= 1e-6 * np.abs(ref)[sig]).sum(); ie += (it == d["it%d" % c]).sum()
    print("%-10s PL within 1e-6: CUDA %.4f | oracle Thomas-vs-PCR %.4f | Newton totals equal %d/%d" % (sys.argv[2], nm / n, ny / n, ie, 3 * S), flush=True)
    sys.exit(0)
for name in sys.argv[1:]:
    env = dict(os.environ, TRPL_LIB=os.path.join(VAR, "libtrpl_%s.so" % name))
    r = subprocess.run([sys.executable, __file__, "--child", name], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-1500:], flush=True)
