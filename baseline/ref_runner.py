"""Run the UNMODIFIED reference from `baseline/_ref/` (a git-ignored copy of /root/reference that
travels to the GPU box with the snapshot).  Test / bench infrastructure only: nothing under
`bayesian_inference_trpl_b200/` imports this file.

  * `ref_pvsim`            pvSimPCR.pvSim natively through numba-CUDA       (pvSimPCR.py:309-401)
  * `ref_bayes`            the reference's own bayeslib.bayes               (bayeslib.py:207-252)
      - route "reference": its own probs.py + pvSimPCR.py kernels
      - route "dropin"   : INTEGRATION.md route A -- sys.modules["probs"] = trpl.probs and
                           model = trpl.pvSim, bayeslib.py itself untouched
  * `legacy_njit_rate`     Legacy/pvSim.py tEvol (numba njit, 1 thread per process) fanned over the
                           host cores with multiprocessing                  (Legacy/pvSim.py:90-173)
  * `fallback_rate`        pvSim_fallback.pvSim_cpu_fallback (SciPy BDF)    (pvSim_fallback.py:80-117)
"""
import importlib.util
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "pvSimPCR.py"))


def _load(name, alias=None, relpath=None):
    """Import baseline/_ref/<name>.py under `alias` (fresh module object, the reference's own
    top-level imports resolved through sys.path / sys.modules)."""
    alias = alias or name
    if alias in sys.modules:
        return sys.modules[alias]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    path = os.path.join(REF, relpath or (name + ".py"))
    spec = importlib.util.spec_from_file_location(alias, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


def ref_pvsim(matpar, simPar, iniPar, dtype=np.float64, init_mode="points", BPG=None):
    """PL [S, T//plT+1] from the reference's numba-CUDA solver; returns (pl, solver_seconds)."""
    from numba import cuda
    pv = _load("pvSimPCR")
    S = len(matpar)
    L, T, plT = int(simPar[2]), int(simPar[3]), int(simPar[4])
    if BPG is None:
        BPG = 8 * cuda.get_current_device().MULTIPROCESSOR_COUNT       # bayeslib.py:146
    pl = np.zeros((S, T // plT + 1), dtype=dtype)
    plN = np.zeros((S, 2, L)); plP = np.zeros((S, 2, L)); plE = np.zeros((S, 2, L + 1))
    ini = np.array(iniPar, dtype=np.float64) if init_mode == "points" else list(iniPar)
    sec = pv.pvSim(pl, plN, plP, plE, np.ascontiguousarray(matpar[:, :12]).copy(), list(simPar), ini,
                   (L,), BPG, max_sims_per_block=1, init_mode=init_mode)
    return pl, sec


def ref_bayes(route, minX, maxX, do_log, iniPar, simPar, e_data, sim_flags, gpu_info, seed=42):
    """(N, P, X) of the reference's bayeslib.bayes, unmodified.  route = "reference" | "dropin"."""
    os.environ.setdefault("SLURM_ARRAY_TASK_ID", "0")                  # bayeslib.py:231
    if route == "reference":
        _load("bayes_io")
        _load("probs")
        bl = _load("bayeslib", "bayeslib_ref_own")
        model = _load("pvSimPCR").pvSim
    else:
        import bayesian_inference_trpl_b200 as trpl
        _load("bayes_io")
        saved = sys.modules.get("probs")
        sys.modules["probs"] = trpl.probs                              # INTEGRATION.md route A
        try:
            bl = _load("bayeslib", "bayeslib_ref_dropin")
        finally:
            if saved is not None:
                sys.modules["probs"] = saved
            else:
                del sys.modules["probs"]
        assert bl.prob is trpl.probs.prob and bl.fastlog is trpl.probs.fastlog
        model = trpl.pvSim
    np.random.seed(seed)                                               # parallel_bayes_gpu.py:35
    N = np.array([0])
    P = None
    return bl.bayes(model, N, P, np.array(minX, float), np.array(maxX, float), do_log,
                    np.array(iniPar, float), list(simPar), e_data, dict(sim_flags), dict(gpu_info))


# ---------------------------------------------------------------------------------------------
# CPU paths of the reference, timed as baselines
# ---------------------------------------------------------------------------------------------
def _legacy_worker(job):
    """One process = one core: Legacy/pvSim.py on its share of the samples, all curves."""
    rows, simPar, amps_alpha, warm = job
    os.environ["NUMBA_NUM_THREADS"] = "1"
    import contextlib
    import io
    lp = _load("pvSim", "legacy_pvSim", os.path.join("Legacy", "pvSim.py"))
    with contextlib.redirect_stdout(io.StringIO()):
        if warm:                                           # JIT compile outside the timed part
            sp = list(simPar); sp[3] = 8; sp[1] = simPar[1] / simPar[3] * 8
            lp.pvSim(rows[:1].copy(), tuple(sp), tuple(amps_alpha[0]))
        t0 = time.perf_counter()
        for a, l in amps_alpha:
            lp.pvSim(rows.copy(), tuple(simPar), (a, l))   # heterogeneous tuple: what its njit typing needs
        return time.perf_counter() - t0


def legacy_njit_rate(X13, simPar, amps_alpha, procs):
    """likelihoods/s of Legacy/pvSim.py (BDF2, no Auger, exp-profile init: the reference's CPU twin
    of the GPU solver) with `procs` worker processes.  X13 rows in engine units; the 10 legacy
    columns are n0,p0,DN,DP,B,Sf,Sb,tauN,tauP,Lambda (Legacy/pvSim.py:40)."""
    import multiprocessing as mp
    cols = [0, 1, 2, 3, 4, 5, 6, 9, 10, 11]
    rows = np.ascontiguousarray(X13[:, cols])
    sp = list(simPar); sp[5] = (0,)
    chunks = [c for c in np.array_split(rows, procs) if len(c)]
    ctx = mp.get_context("spawn")
    with _one_thread_env(), ctx.Pool(len(chunks)) as pool:
        pool.map(_legacy_worker, [(c[:1], sp, amps_alpha[:1], True) for c in chunks])   # warm: JIT per worker
        t0 = time.perf_counter()
        pool.map(_legacy_worker, [(c, sp, amps_alpha, False) for c in chunks])
        dt = time.perf_counter() - t0
    return len(rows) / dt, dt


class _one_thread_env:
    """Children of a spawn pool inherit the environment: one BLAS/OpenMP thread per worker process,
    so `procs` workers use `procs` cores instead of oversubscribing them."""
    KEYS = ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMBA_NUM_THREADS")

    def __enter__(self):
        self.saved = {k: os.environ.get(k) for k in self.KEYS}
        for k in self.KEYS:
            os.environ[k] = "1"

    def __exit__(self, *exc):
        for k, v in self.saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _fallback_worker(job):
    rows, simPar, inis = job
    fb = _load("pvSim_fallback")
    T = int(simPar[3])
    t0 = time.perf_counter()
    for ini in inis:
        pl = np.empty((len(rows), T + 1))
        fb.pvSim_cpu_fallback(pl, rows, list(simPar), np.array(ini, float))
    return time.perf_counter() - t0


def fallback_rate(X13, simPar, inis, procs):
    """likelihoods/s of the reference's wired-in CPU model (SciPy BDF) with `procs` processes."""
    import multiprocessing as mp
    chunks = [c for c in np.array_split(np.ascontiguousarray(X13), procs) if len(c)]
    ctx = mp.get_context("spawn")
    with _one_thread_env(), ctx.Pool(len(chunks)) as pool:
        pool.map(_fallback_worker, [(c[:0], list(simPar), []) for c in chunks])          # imports
        t0 = time.perf_counter()
        pool.map(_fallback_worker, [(c, list(simPar), [np.array(i) for i in inis]) for c in chunks])
        dt = time.perf_counter() - t0
    return len(X13) / dt, dt
