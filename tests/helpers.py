"""Shared test inputs: unit conversions and priors of the reference entry script
(parallel_bayes_gpu.py:27-33,86,91-92), the shipped excitation profiles, golden loaders."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

UC = np.array([(1e7) ** -3, (1e7) ** -3,
               (1e7) ** 2 / (1e9) * .02569257, (1e7) ** 2 / (1e9) * .02569257,
               (1e7) ** 3 / (1e9), (1e7) / (1e9), (1e7) / (1e9),
               (1e7) ** 6 / (1e9), (1e7) ** 6 / (1e9), 1, 1, 704.3, 1])
TRUTH = np.array([1e8, 3e15, 20, 20, 4.8e-11, 10, 10, 4.4e-29, 4.4e-29, 511, 871, 0.1, 0])
DO_LOG = np.array([1, 1, 0, 0, 1, 1, 1, 1, 1, 0, 0, 1, 0])
MINX = np.array([1e8, 1e14, 0, 0, 1e-11, 0.1, 0.1, 1e-30, 1e-30, 1, 1, 10 ** -1, 0])
MAXX = np.array([1e8, 1e16, 50, 50, 1e-9, 100, 100, 1e-28, 1e-28, 1000, 2000, 10 ** -1, 0])


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=True)


def example_data():
    return golden("example_data.npz")


def power_scan_excitations():
    """[3,128] nm^-3 (bayes_io.get_initpoints scaling 1e-21)."""
    return example_data()["power_exc"] * 1e-21


def prior_samples(n, seed, stiff=False, mag=False):
    """Default prior of the entry script (physical engine units); mobilities start at 0.5 instead
    of 0 to keep D > 0."""
    rng = np.random.default_rng(seed)
    lo, hi = MINX.copy(), MAXX.copy()
    lo[2:4] = 0.5
    if stiff:
        lo[5:7], hi[5:7] = 1.0, 1e5
    if mag:
        lo[12], hi[12] = -0.5, 0.5
    X = np.empty((n, 13))
    for j in range(13):
        if lo[j] == hi[j]:
            X[:, j] = lo[j]
        elif DO_LOG[j]:
            X[:, j] = 10 ** rng.uniform(np.log10(lo[j]), np.log10(hi[j]), n)
        else:
            X[:, j] = rng.uniform(lo[j], hi[j], n)
    return X * UC


def pl_noise_floor(matpar_phys, length, time, L, T):
    """Absolute rounding floor of PL = rate*(sum N*P - L*N0*P0)/(dx^2 dt): the two terms cancel, so
    once the excess carriers have decayed the value is rounding noise of the equilibrium term
    T_eq = B*n0*p0*Length (pvSimPCR.py:278-281).  Two valid FP64 evaluation orders (the oracle's
    PCR and Thomas variants) already differ by ~100 eps*T_eq there; allow 2^12 eps*T_eq, which is
    < 1e-6*PL until PL has fallen ~13 decades below its initial value."""
    n0, p0, B = matpar_phys[:, 0], matpar_phys[:, 1], matpar_phys[:, 4]
    return 2.0 ** 12 * np.finfo(float).eps * B * n0 * p0 * length


def simpar_from_golden(g):
    length, Time, L, T, plT, tol, MAX = g["simPar"]
    return [float(length), float(Time), int(L), int(T), int(plT), (0,), int(tol), int(MAX)]


def route_a_case(inis, S=256, T=4000, truth_pl=None):
    """Inputs of the `bayeslib.bayes` comparison between the reference's own kernels and the drop-ins
    (tools/ref_on_b200.py, tests/test_reference_b200.py): default prior with a free mag_offset, 3
    power-scan curves at L=128, observation windows shorter than the simulation, the last one
    off the step grid.  `truth_pl(c)` -> PL of the truth sample for curve c on the full step grid."""
    L = 128
    Time = 0.025 * T
    simPar = [2000.0, Time, L, T, 1, (0,), 7, 10000]
    lo, hi = MINX * UC, MAXX * UC
    lo[2:4] = 0.5 * UC[2]
    lo[12], hi[12] = -0.2, 0.2
    grid = np.linspace(0, Time, T + 1)
    e_t, e_v, e_u = [], [], []
    for c in range(3):
        n = min([1501, 2201, 3001][c], (T * 3) // 4 + 1)
        tt = np.linspace(0, grid[n - 1], n)
        if c == 2:
            tt = tt[:-1] + 0.3 * 0.025          # off-grid
        e_t.append(tt)
        e_u.append(np.full(len(tt), 0.1))
        if truth_pl is not None:
            e_v.append(np.interp(tt, grid, np.log10(truth_pl(c))))
    flags = {"load_PL_from_file": False, "log_pl": True, "self_normalize": False, "random_sample": True,
             "num_points": S, "override_equal_mu": False, "override_equal_s": False,
             "override_equal_auger": False}
    info = {"has_GPU": True, "sims_per_gpu": 128, "num_gpus": 1, "threads_per_block": (128,),
            "max_sims_per_block": 1}
    return dict(simPar=simPar, lo=lo, hi=hi, do_log=DO_LOG, e_data=[(e_t, e_v, e_u)], flags=flags, info=info)
