#!/usr/bin/env python3
"""Generate golden vectors by running the UNMODIFIED reference (read from /root/reference)
under numba's CUDA simulator (NUMBA_ENABLE_CUDASIM=1).  Runs only in the build container;
the resulting small .npz files are committed so that the oracle can be pinned anywhere.

    python tests/golden/make_cudasim_golden.py [case ...]

Cases
  pvsim_points_f64   pvSimPCR.pvSim, init_mode="points", float64 PL buffer  (pvSimPCR.py:309)
  pvsim_points_f32   same with the float32 PL buffer bayeslib allocates      (bayeslib.py:137)
  pvsim_exp_f64      init_mode="exp"                                         (pvSimPCR.py:347-353)
  pvsim_stiff_f64    high surface recombination / short lifetime corner, L=16
  pvsim_L32_f64      L=32 (four PCR stages + 2x2 finish)
  pvsim_L128_f64     the production grid L=128 (six PCR stages), BDF ramp 1..5, 7 steps
  pvsim_long_f64     L=8, 241 steps at the highest excitation: BDF5 steady state with Auger on
  probs              probs.prob and probs.fastlog                            (probs.py:49-85)
  bayes              bayeslib.bayes end to end, 2 curves, 1 observation file (bayeslib.py:207)
  bayes_norm2        same with self_normalize=True and two observation files
  bayes_lin          same with log_pl=False (linear PL residuals)
  legacy             Legacy/pvSim.py (numba njit CPU solver, BDF2, no Auger) at L=128
"""
import os
import sys
import time

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

import numpy as np  # noqa: E402

UC = np.array([(1e7) ** -3, (1e7) ** -3,
               (1e7) ** 2 / (1e9) * .02569257, (1e7) ** 2 / (1e9) * .02569257,
               (1e7) ** 3 / (1e9), (1e7) / (1e9), (1e7) / (1e9),
               (1e7) ** 6 / (1e9), (1e7) ** 6 / (1e9), 1, 1, 704.3, 1])
TRUTH = np.array([1e8, 3e15, 20, 20, 4.8e-11, 10, 10, 4.4e-29, 4.4e-29, 511, 871, 0.1, 0])


def prior_samples(n, seed, stiff=False):
    rng = np.random.default_rng(seed)
    lo = np.array([1e8, 1e14, 0.5, 0.5, 1e-11, 0.1, 0.1, 1e-30, 1e-30, 1, 1, 0.1, 0])
    hi = np.array([1e8, 1e16, 50, 50, 1e-9, 100, 100, 1e-28, 1e-28, 1000, 2000, 0.1, 0])
    if stiff:
        lo[5:7] = 1e3
        hi[5:7] = 1e5
        hi[9:11] = 20
    do_log = np.array([1, 1, 0, 0, 1, 1, 1, 1, 1, 0, 0, 1, 0], bool)
    X = np.empty((n, 13))
    for j in range(13):
        if lo[j] == hi[j]:
            X[:, j] = lo[j]
        elif do_log[j]:
            X[:, j] = 10 ** rng.uniform(np.log10(lo[j]), np.log10(hi[j]), n)
        else:
            X[:, j] = rng.uniform(lo[j], hi[j], n)
    return X * UC


def excitation(L, amp_cm3=1.2155e16, alpha=6e-3, length=None):
    dx = length / L
    x = (np.arange(L) + 0.5) * dx
    return amp_cm3 * 1e-21 * np.exp(-alpha * x)


def run_pvsim(L, T, S, dtype, init_mode, seed, stiff=False, tol=7, plT=1, amp=1.2155e16):
    import pvSimPCR
    length = 15.625 * L
    Time = 0.025 * T
    simPar = [length, Time, L, T, plT, (0,), tol, 10000]
    X = prior_samples(S, seed, stiff)
    X[0] = TRUTH * UC
    mat = np.ascontiguousarray(X[:, :12])
    if init_mode == "points":
        ini = excitation(L, amp, length=length)
        ini_arg = ini.copy()
    else:
        ini = np.array([amp * 1e-21, 1 / 6e-3])
        ini_arg = list(ini)
    pl = np.zeros((S, T // plT + 1), dtype=dtype)
    plN = np.zeros((S, 2, L)); plP = np.zeros((S, 2, L)); plE = np.zeros((S, 2, L + 1))
    t0 = time.time()
    pvSimPCR.pvSim(pl, plN, plP, plE, mat.copy(), list(simPar), ini_arg, (L,), S,
                   max_sims_per_block=1, init_mode=init_mode)
    print("  pvSim L=%d T=%d S=%d took %.1fs" % (L, T, S, time.time() - t0), flush=True)
    return dict(matPar=mat, simPar=np.array([length, Time, L, T, plT, tol, 10000], float),
                iniPar=np.asarray(ini, float), pl=pl, init_mode=init_mode)


def case_probs():
    import probs
    rng = np.random.default_rng(7)
    S, n = 37, 53
    pli = rng.uniform(-30, 0, (S, n))
    values = rng.uniform(-30, 0, n)
    unc = rng.uniform(0.1, 1, n)
    mag = rng.uniform(-1, 1, S)
    P = rng.uniform(-5, 0, S)
    P0 = P.copy()
    probs.prob(P, pli, values, unc, mag, 16, 3)
    x64 = np.abs(rng.normal(0, 1, (5, 40))) * 10.0 ** rng.integers(-320, 20, (5, 40)).astype(float)
    x64[0, :4] = [0.0, -1.0, 1e-310, 2.5e-308]
    x32 = x64.astype(np.float32)
    x64_in, x32_in = x64.copy(), x32.copy()
    MIN = sys.float_info.min
    # the simulator's math.log10 raises on 0 -> only the f64 buffer can be run as is
    probs.fastlog(x64, MIN, 16, 2)
    x32_pos = np.where(x32_in >= np.float32(1e-37), x32_in, np.float32(1.0)).astype(np.float32)
    x32_pos_in = x32_pos.copy()
    probs.fastlog(x32_pos, MIN, 16, 2)
    return dict(pli=pli, values=values, unc=unc, mag=mag, P_in=P0, P_out=P,
                log_in64=x64_in, log_out64=x64, log_in32=x32_pos_in, log_out32=x32_pos, MIN=MIN)


def case_bayes(self_normalize=False, log_pl=True, n_exp=1, num_points=5):
    """bayeslib.bayes, unmodified, on the simulator (3-line get_current_device shim)."""
    from numba import cuda

    class _Dev:
        MULTIPROCESSOR_COUNT = 1
    if not hasattr(cuda, "get_current_device"):
        cuda.get_current_device = lambda: _Dev()
    os.environ["SLURM_ARRAY_TASK_ID"] = "0"
    import bayeslib
    import pvSimPCR
    L, T = 8, 12
    length = [15.625 * L * 0.5, 15.625 * L]      # two thicknesses, like Twothick
    Time = 0.025 * T
    simPar = [length, Time, L, T, 1, (0,), 7, 10000]
    iniPar = np.stack([excitation(L, 1.2e16, length=length[0]), excitation(L, 1.1e17, length=length[1])])
    rng = np.random.default_rng(3)
    # observations: every 2nd grid time, arbitrary smooth log curve
    e_data = []
    for e in range(n_exp):
        t_obs = [np.linspace(0, Time, T + 1)[::2 + e].copy() for _ in range(2)]
        if self_normalize:      # curves normalised to their own maximum, like bayes_io.get_data does
            v_obs = [np.sort(rng.uniform(0.05, 1.0, len(t)))[::-1].copy() for t in t_obs]
            for v in v_obs:
                v[0] = 1.0
            if log_pl:
                v_obs = [np.log10(v) for v in v_obs]
        else:
            v_obs = [rng.uniform(-8, -6, len(t)) if log_pl else 10 ** rng.uniform(-8, -6, len(t)) for t in t_obs]
        u_obs = [np.full(len(t), 0.1) for t in t_obs]
        e_data.append((t_obs, v_obs, u_obs))
    minX = np.array([1e8, 1e14, 1, 1, 1e-11, 0.1, 0.1, 1e-30, 1e-30, 1, 1, 0.1, -0.5]) * UC
    maxX = np.array([1e8, 1e16, 50, 50, 1e-9, 100, 100, 1e-28, 1e-28, 1000, 2000, 0.1, 0.5]) * UC
    do_log = np.array([1, 1, 0, 0, 1, 1, 1, 1, 1, 0, 0, 1, 0])
    sim_flags = {"load_PL_from_file": False, "override_equal_auger": False,
                 "override_equal_mu": False, "override_equal_s": True, "log_pl": log_pl,
                 "self_normalize": self_normalize, "random_sample": True, "num_points": num_points}
    gpu_info = {"sims_per_gpu": 2, "num_gpus": 1, "has_GPU": True,
                "threads_per_block": (L,), "max_sims_per_block": 1}
    np.random.seed(42)
    t0 = time.time()
    N, P, X = bayeslib.bayes(pvSimPCR.pvSim, np.array([0]), None, minX, maxX, do_log, iniPar,
                             list(simPar), e_data, sim_flags, gpu_info)
    print("  bayes took %.1fs" % (time.time() - t0), flush=True)
    out = dict(P=P, X=X, minX=minX, maxX=maxX, do_log=do_log, iniPar=iniPar,
               length=np.array(length), Time=Time, L=L, T=T, n_exp=n_exp,
               self_normalize=self_normalize, log_pl=log_pl)
    for e, (t_obs, v_obs, u_obs) in enumerate(e_data):
        sfx = "" if e == 0 else "_%d" % e
        out["t_obs" + sfx] = np.array(t_obs); out["v_obs" + sfx] = np.array(v_obs); out["u_obs" + sfx] = np.array(u_obs)
    return out


def case_legacy():
    """Legacy/pvSim.py (njit, Thomas, BDF1->2, no Auger, exp init) at the real grid."""
    sys.path.insert(0, os.path.join(REF, "Legacy"))
    os.environ["NUMBA_ENABLE_CUDASIM"] = "0"
    import importlib
    legacy = importlib.import_module("pvSim")
    L, T = 128, 4000
    length, Time = 2000.0, 100.0
    simPar = (length, Time, L, T, 1, (0,), 7, 10000)
    X = prior_samples(4, 11)
    X[0] = TRUTH * UC
    mat10 = np.ascontiguousarray(X[:, [0, 1, 2, 3, 4, 5, 6, 9, 10, 11]])
    out = {}
    for ci, amp in enumerate([1.2155e16, 1.6485e18]):
        ini = (amp * 1e-21, 1 / 6e-3)
        itrs, (plN, plP, plE, plI) = legacy.pvSim(mat10.copy(), simPar, ini)
        out["pl%d" % ci] = plI
        out["ini%d" % ci] = np.array(ini)
    out.update(mat10=mat10, simPar=np.array([length, Time, L, T, 1, 7, 10000], float))
    return out


CASES = {
    "pvsim_points_f64": lambda: run_pvsim(8, 24, 3, np.float64, "points", 1),
    "pvsim_points_f32": lambda: run_pvsim(8, 16, 2, np.float32, "points", 2),
    "pvsim_exp_f64": lambda: run_pvsim(8, 12, 2, np.float64, "exp", 3),
    "pvsim_stiff_f64": lambda: run_pvsim(16, 20, 3, np.float64, "points", 4, stiff=True, amp=1.6485e18),
    "pvsim_L32_f64": lambda: run_pvsim(32, 10, 2, np.float64, "points", 5, amp=1.1539e17),
    "pvsim_L128_f64": lambda: run_pvsim(128, 6, 1, np.float64, "points", 6, amp=1.6485e18),
    "pvsim_long_f64": lambda: run_pvsim(8, 240, 2, np.float64, "points", 7, amp=1.6485e18),
    "probs": case_probs,
    "bayes": case_bayes,
    "bayes_norm2": lambda: case_bayes(self_normalize=True, log_pl=True, n_exp=2, num_points=3),
    "bayes_lin": lambda: case_bayes(self_normalize=False, log_pl=False, n_exp=1, num_points=3),
    "legacy": case_legacy,
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for name in names:
        print("case", name, flush=True)
        data = CASES[name]()
        np.savez_compressed(os.path.join(HERE, "cudasim_%s.npz" % name if name != "legacy"
                                         else "legacy_pvsim.npz"), **data)
        print("  wrote", name, flush=True)
