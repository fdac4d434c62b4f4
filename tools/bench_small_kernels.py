#!/usr/bin/env python3
"""HBM-bound neighbours of the solver: achieved bandwidth of trpl_lnp_accumulate (probs.prob),
trpl_log10_clamp (probs.fastlog), weighted histogram / moments, sample generation."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bayesian_inference_trpl_b200 as trpl
E = trpl.engine
dev = torch.device("cuda", 0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
S, n = 4096, 80001                       # 4096 samples x 80001 log-PL points = 2.6 GB f64
pl = torch.rand((S, n), dtype=torch.float64, device=dev) * -5 - 5
val = torch.rand(n, dtype=torch.float64, device=dev) * -5 - 5
mag = torch.zeros(S, dtype=torch.float64, device=dev)
P = torch.zeros(S, dtype=torch.float64, device=dev)
ms = timeit(lambda: E.lnp_accumulate_(P, pl, val, mag))
print("trpl_lnp_accumulate   %7.3f ms  %7.1f GB/s  (%.2f of %.0f measured)" % (ms, S * n * 8 / ms / 1e6, S * n * 8 / ms / 1e6 / peak, peak))
x64 = torch.rand((S, n), dtype=torch.float64, device=dev) + 1e-3
ms = timeit(lambda: E.log10_clamp_(x64, 2.2e-308))
print("trpl_log10_clamp f64  %7.3f ms  %7.1f GB/s  (%.2f)" % (ms, 2 * S * n * 8 / ms / 1e6, 2 * S * n * 8 / ms / 1e6 / peak))
x32 = torch.rand((S, n), dtype=torch.float32, device=dev) + 1e-3
ms = timeit(lambda: E.log10_clamp_(x32, 2.2e-308))
print("trpl_log10_clamp f32  %7.3f ms  %7.1f GB/s  (%.2f)" % (ms, 2 * S * n * 4 / ms / 1e6, 2 * S * n * 4 / ms / 1e6 / peak))
del pl, x64, x32
N = 16 * 1024 * 1024
X = torch.rand((N, 13), dtype=torch.float64, device=dev)
w = torch.rand(N, dtype=torch.float64, device=dev)
ms = timeit(lambda: E.weighted_hist(X, 2, w, 0.0, 1.0, 96))
print("trpl_weighted_hist 1D %7.3f ms  %7.1f GB/s useful (16 B of every 112 B row; sector-granular traffic is higher)" % (ms, N * 16 / ms / 1e6))
ms = timeit(lambda: E.weighted_moments(X, w))
print("trpl_weighted_moments %7.3f ms  %7.1f GB/s  (%.2f)" % (ms, N * 14 * 8 / ms / 1e6, N * 14 * 8 / ms / 1e6 / peak))
lo = np.full(13, 1.0); hi = np.full(13, 10.0); dl = np.array([1, 1, 0, 0, 1, 1, 1, 1, 1, 0, 0, 1, 0])
ms = timeit(lambda: E.random_grid_device(lo, hi, dl, N, 42, out=X))
print("trpl_random_grid      %7.3f ms  %7.1f GB/s written (%.2f), %.1f Gsamples/s" % (ms, N * 13 * 8 / ms / 1e6, N * 13 * 8 / ms / 1e6 / peak, N / ms / 1e6))
