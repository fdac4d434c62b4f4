#!/usr/bin/env python3
"""Run the UNMODIFIED reference (baseline/_ref, numba-CUDA) natively on the GPU box, next to this
engine, on identical inputs.  SURVEY 8(c) "exact oracle, GPU box" / BASELINE.md "first GPU task".

    python tools/ref_on_b200.py [--S 1024] [--T 80000] [--curves 3] [--out gpurun_out/ref_gpu]

Writes <out>.json (throughput of the reference's numba kernels on this GPU, parity statistics of
trpl.pvSim against them at the full shape) and <out>_golden.npz (a sub-sampled slice of the
reference's PL curves: the fixture tests/test_reference_gpu.py falls back to when baseline/_ref
is not present on the box).
"""
import argparse
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--S", type=int, default=1024)
    ap.add_argument("--T", type=int, default=80000)
    ap.add_argument("--curves", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_gpu"))
    ap.add_argument("--skip-bayes", action="store_true")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    res = {"S": args.S, "T": args.T}

    from baseline import ref_runner as rr
    from helpers import TRUTH, UC, pl_noise_floor, power_scan_excitations, prior_samples
    import bayesian_inference_trpl_b200 as trpl

    L = 128
    inis = power_scan_excitations()
    X = prior_samples(args.S, seed=2024)
    X[0] = TRUTH * UC

    # ---- 1. does numba JIT the reference kernels for this device at all?
    try:
        from numba import cuda
        dev = cuda.get_current_device()
        res["numba_device"] = {"name": dev.name.decode() if isinstance(dev.name, bytes) else str(dev.name),
                               "cc": list(dev.compute_capability), "sms": dev.MULTIPROCESSOR_COUNT}
        sp = [2000.0, 0.025 * 64, L, 64, 1, (0,), 7, 10000]
        t0 = time.time()
        pl_small, _ = rr.ref_pvsim(X[:8], sp, inis[0])
        res["jit_seconds"] = time.time() - t0
        mine = np.empty_like(pl_small)
        trpl.pvSim(mine, None, None, None, X[:8, :12], sp, inis[0], (128,), 0, 1, init_mode="points")
        res["probe_max_rel"] = float(np.max(np.abs(mine - pl_small) / np.abs(pl_small)))
        print("numba JIT ok: %.1f s, probe max rel %.2e" % (res["jit_seconds"], res["probe_max_rel"]), flush=True)
    except Exception:
        res["numba_error"] = traceback.format_exc()
        print(res["numba_error"], flush=True)
        json.dump(res, open(args.out + ".json", "w"), indent=1)
        return 1

    # ---- 2. full shape: reference numba kernels vs trpl.pvSim, float64 PL buffers
    T = args.T
    Time = 0.025 * T
    sp = [2000.0, Time, L, T, 1, (0,), 7, 10000]
    floor = pl_noise_floor(X[:, :12], 2000.0, Time, L, T)        # cancellation noise of PL (tests/helpers.py)
    ref_secs, my_secs = [], []
    stats = []
    gold = {}
    keep_rows = np.arange(0, args.S, max(1, args.S // 16))[:16]
    keep_t = np.unique(np.concatenate([np.arange(0, 64), np.arange(64, T + 1, max(1, T // 512)), [T]]))
    lnl_ref = np.zeros(args.S)
    lnl_my = np.zeros(args.S)
    for c in range(args.curves):
        t0 = time.time()
        pl_ref, sec = rr.ref_pvsim(X, sp, inis[c])
        wall = time.time() - t0
        ref_secs.append((sec, wall))
        pl_my = np.empty_like(pl_ref)
        my_sec = trpl.pvSim(pl_my, None, None, None, X[:, :12], sp, inis[c], (128,), 0, 1, init_mode="points")
        my_secs.append(my_sec)
        above = np.abs(pl_ref) > floor[:, None]
        rel = np.abs(pl_my - pl_ref) / np.maximum(np.abs(pl_ref), 1e-300)
        st = {"curve": c, "ref_kernel_s": sec, "ref_wall_s": wall, "trpl_kernel_s": my_sec,
              "points": int(above.sum()), "frac_above_floor": float(above.mean()),
              "max_rel_above_floor": float(rel[above].max()),
              "frac_within_1e-6": float((rel[above] <= 1e-6).mean()),
              "frac_within_1e-9": float((rel[above] <= 1e-9).mean()),
              "max_abs_below_floor_over_floor": float((np.abs(pl_my - pl_ref) / floor[:, None])[~above].max()) if (~above).any() else 0.0,
              "ref_nonfinite_rows": int((~np.isfinite(pl_ref)).any(axis=1).sum()),
              "my_nonfinite_rows": int((~np.isfinite(pl_my)).any(axis=1).sum())}
        # f64 likelihood against the truth sample's curve (row 0), both sides evaluated identically
        with np.errstate(divide="ignore", invalid="ignore"):
            lr = np.log10(np.maximum(pl_ref, sys.float_info.min))
            lm = np.log10(np.maximum(pl_my, sys.float_info.min))
        lnl_ref -= ((lr - lr[0]) ** 2).sum(axis=1)
        lnl_my -= ((lm - lr[0]) ** 2).sum(axis=1)
        stats.append(st)
        gold["pl_ref_c%d" % c] = pl_ref[np.ix_(keep_rows, keep_t)]
        print(json.dumps(st), flush=True)
        del pl_ref, pl_my, rel, above, lr, lm
    ok = np.isfinite(lnl_ref) & (np.abs(lnl_ref) > 0)
    lrel = np.abs(lnl_my - lnl_ref)[ok] / np.abs(lnl_ref)[ok]
    res["curves"] = stats
    res["lnl"] = {"n": int(ok.sum()), "max_rel": float(lrel.max()), "frac_within_1e-6": float((lrel <= 1e-6).mean())}
    tot_ref = sum(s for s, _ in ref_secs)
    tot_wall = sum(w for _, w in ref_secs)
    res["reference_gpu"] = {"kind": "numba-cuda unmodified (pvSimPCR.pvSim, float64 PL buffer, BPG=8*SMs, TPB=128)",
                            "samples": args.S, "curves": args.curves,
                            "likelihoods_per_s_kernel": args.S / tot_ref * (args.curves / 3.0),
                            "likelihoods_per_s_pvsim_wall": args.S / tot_wall * (args.curves / 3.0),
                            "kernel_s": tot_ref, "wall_s": tot_wall}
    res["trpl_pvsim_same_call"] = {"likelihoods_per_s_kernel": args.S / sum(my_secs), "kernel_s": sum(my_secs)}
    gold.update(X=X[keep_rows], rows=keep_rows, t_idx=keep_t, simPar=np.array([2000.0, Time, L, T, 1, 7, 10000], float),
                inis=inis[:args.curves])
    np.savez_compressed(args.out + "_golden.npz", **gold)
    json.dump(res, open(args.out + ".json", "w"), indent=1)
    print(json.dumps(res["reference_gpu"]), json.dumps(res["lnl"]), flush=True)

    # ---- 3. the reference's own bayeslib.bayes: its kernels vs the drop-ins (INTEGRATION route A)
    if not args.skip_bayes:
        try:
            res["bayes_route_a"] = bayes_routes(rr, trpl, inis)
            print(json.dumps(res["bayes_route_a"]), flush=True)
        except Exception:
            res["bayes_route_a_error"] = traceback.format_exc()
            print(res["bayes_route_a_error"], flush=True)
        json.dump(res, open(args.out + ".json", "w"), indent=1)
    return 0


def bayes_routes(rr, trpl, inis, S=256, T=4000):
    from helpers import TRUTH, UC, route_a_case
    from oracle import oracle
    case = route_a_case(inis, S, T, truth_pl=lambda c: oracle.solve(
        (TRUTH * UC)[None, :12], [2000.0, 0.025 * T, 128, T, 1, (0,), 7, 10000], inis[c], solver="thomas")["pl"][0])
    out = {}
    t0 = time.time()
    N1, P1, X1 = rr.ref_bayes("reference", case["lo"], case["hi"], case["do_log"], inis, case["simPar"],
                              case["e_data"], case["flags"], case["info"])
    out["reference_s"] = time.time() - t0
    t0 = time.time()
    N2, P2, X2 = rr.ref_bayes("dropin", case["lo"], case["hi"], case["do_log"], inis, case["simPar"],
                              case["e_data"], case["flags"], case["info"])
    out["dropin_s"] = time.time() - t0
    assert np.array_equal(X1, X2)
    ok = np.isfinite(P1[0]) & np.isfinite(P2[0])
    rel = np.abs(P1[0] - P2[0])[ok] / np.maximum(np.abs(P1[0][ok]), 1e-300)
    out.update(S=S, T=T, finite=int(ok.sum()), max_rel=float(rel.max()), max_abs=float(np.abs(P1[0] - P2[0])[ok].max()),
               frac_within_1e_5=float((rel <= 1e-5).mean()),
               same_nonfinite=bool(np.array_equal(np.isfinite(P1[0]), np.isfinite(P2[0]))))
    e_t, e_v, e_u = case["e_data"][0]
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_gpu_bayes.npz"), P_ref=P1, P_dropin=P2, X=X1, S=S, T=T,
                        **{"v_obs%d" % c: e_v[c] for c in range(3)})
    return out


if __name__ == "__main__":
    sys.exit(main())
