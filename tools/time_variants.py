#!/usr/bin/env python3
"""Time the fused power-scan launch for every library in build/variants/ (one subprocess each, TRPL_LIB
selects the library) and compare their lnL / Newton totals with the first one.
    python tools/time_variants.py [--T 20000] [--reps 3] [names...]"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAR = os.path.join(ROOT, "build", "variants")


def child(T, reps, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch
    import bayesian_inference_trpl_b200 as trpl
    from helpers import TRUTH, UC, power_scan_excitations, prior_samples
    L = 128
    simPar = [2000.0, 0.025 * T, L, T, 1, (0,), 7, 10000]
    inis = power_scan_excitations()
    S = 2 * trpl.engine.resident_sims(L, 0)
    X = prior_samples(S, seed=4321)
    truth = TRUTH * UC
    grid = np.linspace(0, simPar[1], T + 1)
    ts, vs, us = [], [], []
    for c in range(3):
        pl = np.empty((1, T + 1))
        trpl.pvSim(pl, None, None, None, truth[None, :12], simPar, inis[c], (128,), 0, 1, init_mode="points")
        ts.append(grid.copy()); vs.append(np.log10(pl[0])); us.append(np.full(T + 1, 0.1))
    prob = trpl.engine.Problem(simPar, inis, [(ts, vs, us)], device=0)
    Xd = torch.from_numpy(X).cuda()
    lnl, status, iters = trpl.engine.solve_loglik(Xd, prob, want_iters=True)
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        trpl.engine.solve_loglik(Xd, prob)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    np.savez(out, lnl=lnl.cpu().numpy(), iters=iters.cpu().numpy(), status=status.cpu().numpy(), v0=vs[0])
    print(json.dumps({"S": S, "T": T, "ms": ms, "lik_per_s_at_T": S / (min(ms) * 1e-3),
                      "lik_per_s_scaled_to_80000": S / (min(ms) * 1e-3) * (T + 1) / 80001.0,
                      "iters_per_step": float(iters.sum().item()) / (3.0 * S * (T + 1)),
                      "bad": int((status != 0).sum().item())}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
        sys.exit(0)
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=20000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("names", nargs="*")
    a = ap.parse_args()
    import numpy as np
    names = a.names or sorted(f[len("libtrpl_"):-3] for f in os.listdir(VAR) if f.endswith(".so"))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    base = None
    for n in names:
        out = os.path.join(ROOT, "gpurun_out", "variant_%s.npz" % n)
        env = dict(os.environ, TRPL_LIB=os.path.join(VAR, "libtrpl_%s.so" % n))
        r = subprocess.run([sys.executable, __file__, "--child", str(a.T), str(a.reps), out], env=env,
                           capture_output=True, text=True)
        line = (r.stdout.strip().splitlines() or ["{}"])[-1]
        if r.returncode != 0:
            print(n, "FAILED", r.stderr[-2000:])
            continue
        d = np.load(out)
        extra = ""
        if base is None:
            base = d
        else:
            ok = np.isfinite(base["lnl"]) & np.isfinite(d["lnl"])
            rel = np.abs(d["lnl"] - base["lnl"])[ok] / np.abs(base["lnl"])[ok]
            extra = " | vs %s: lnL max rel %.2e, iters equal on %.4f of sims" % (
                names[0], rel.max(), float((d["iters"] == base["iters"]).mean()))
        print("%-12s %s%s" % (n, line, extra), flush=True)
