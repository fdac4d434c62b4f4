"""Pin the CPU oracle (oracle/trpl_oracle.c) against outputs of the UNMODIFIED reference:
numba-CUDA kernels run under NUMBA_ENABLE_CUDASIM=1 (tests/golden/cudasim_*.npz, made by
tests/golden/make_cudasim_golden.py), the Legacy njit solver, and the t=0 PL integral pinned by
the shipped observation files (KAT-0)."""
import os
import sys

import numpy as np
import pytest

from helpers import GOLDEN, TRUTH, UC, example_data, golden, simpar_from_golden
from oracle import oracle


def _solve_like_golden(g, solver, simulator_pow=False):
    simPar = simpar_from_golden(g)
    mode = str(g["init_mode"])
    ini = g["iniPar"] if mode == "points" else tuple(g["iniPar"])
    f32 = g["pl"].dtype == np.float32
    r = oracle.solve(g["matPar"], simPar, ini, init_mode=mode, solver=solver, raw=f32,
                     simulator_pow=simulator_pow)
    pl = r["pl"]
    if f32:
        _, dx, dt = oracle.scales(simPar[0], simPar[1], simPar[2], simPar[3])
        pl = pl.astype(np.float32)
        pl /= dx ** 2 * dt
    return pl, r


@pytest.mark.parametrize("name", ["pvsim_points_f64", "pvsim_points_f32", "pvsim_exp_f64",
                                  "pvsim_stiff_f64", "pvsim_L32_f64", "pvsim_L128_f64", "pvsim_long_f64"])
def test_pcr_oracle_is_bit_exact_with_reference_kernels(name):
    path = os.path.join(GOLDEN, "cudasim_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    g = golden("cudasim_%s.npz" % name)
    # the simulator evaluates x**2 with libm pow(); numba itself compiles it to x*x
    pl, r = _solve_like_golden(g, "pcr", simulator_pow=True)
    assert r["status"].max() == 0
    assert pl.dtype == g["pl"].dtype
    np.testing.assert_array_equal(pl, g["pl"])
    pl2, _ = _solve_like_golden(g, "pcr", simulator_pow=False)
    np.testing.assert_allclose(pl2, g["pl"], rtol=1e-13 if pl2.dtype == np.float64 else 0)


@pytest.mark.parametrize("name", ["pvsim_points_f64", "pvsim_exp_f64", "pvsim_stiff_f64", "pvsim_L32_f64", "pvsim_L128_f64",
                                  "pvsim_long_f64"])
def test_thomas_oracle_matches_reference_kernels(name):
    path = os.path.join(GOLDEN, "cudasim_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    g = golden("cudasim_%s.npz" % name)
    pl, _ = _solve_like_golden(g, "thomas")
    np.testing.assert_allclose(pl, g["pl"], rtol=1e-12)


def test_oracle_matches_legacy_njit_solver():
    """Legacy/pvSim.py (Thomas, BDF1->2, no Auger, exp init) at L=128, T=4000."""
    g = golden("legacy_pvsim.npz")
    length, Time, L, T, plT, tol, MAX = g["simPar"]
    simPar = [length, Time, int(L), int(T), 1, (0,), int(tol), int(MAX)]
    m12 = np.zeros((len(g["mat10"]), 12))
    m12[:, [0, 1, 2, 3, 4, 5, 6, 9, 10, 11]] = g["mat10"]
    for ci in (0, 1):
        r = oracle.solve(m12, simPar, tuple(g["ini%d" % ci]), init_mode="exp", solver="thomas",
                         max_order=2)
        np.testing.assert_allclose(r["pl"], g["pl%d" % ci], rtol=1e-11)


def test_kat0_t0_pl_integral_matches_shipped_observations():
    """First row of every shipped *_Power_scan_Observations.csv curve = PL(t=0) of the three
    Power_scan excitations on a 2000 nm / 128-node film (units, scale factors, rectangle rule)."""
    ex = example_data()
    ini = ex["power_exc"] * 1e-21
    simPar = [2000.0, 0.025 * 4, 128, 4, 1, (0,), 7, 10000]
    mat = (TRUTH * UC)[None, :12]
    for c in range(3):
        pl0 = oracle.solve(mat, simPar, ini[c], solver="pcr")["pl"][0, 0]
        for name in ("Highfrontsurf", "Highbacksurf", "Balancedhighsurf"):
            obs0 = ex["%s_pl%d" % (name, c)][0] * 1e-23
            assert abs(pl0 - obs0) / obs0 < 5e-9


def test_kat1_stiff_curves_loose():
    """Whole shipped stiff curves vs the oracle at the identified truth values: a ~1e-2 sanity
    check in log10 (the generating simulator is not this solver)."""
    ex = example_data()
    ini = ex["power_exc"] * 1e-21
    truths = {"Highfrontsurf": (1e4, 10), "Highbacksurf": (10, 1e4), "Balancedhighsurf": (5e3, 5e3)}
    T = 8000   # 200 ns window
    simPar = [2000.0, 0.025 * T, 128, T, 1, (0,), 7, 10000]
    for name, (sf, sb) in truths.items():
        x = TRUTH.copy()
        x[5], x[6] = sf, sb
        x[7] = x[8] = 0.0          # the example set was generated without Auger terms
        mat = (x * UC)[None, :12]
        for c in (0, 2):
            pl = oracle.solve(mat, simPar, ini[c], solver="thomas")["pl"][0]
            t = ex["%s_t%d" % (name, c)]
            v = ex["%s_pl%d" % (name, c)] * 1e-23
            keep = t <= 0.025 * T
            idx = np.rint(t[keep] / 0.025).astype(int)
            d = np.abs(np.log10(pl[idx]) - np.log10(v[keep]))
            assert d.max() < 5e-2, (name, c, d.max())


def test_probs_oracle_matches_reference():
    g = golden("cudasim_probs.npz")
    P = g["P_in"].copy()
    oracle.prob(P, g["pli"], g["values"], g["mag"])
    np.testing.assert_allclose(P, g["P_out"], rtol=1e-14)
    x = g["log_in64"].copy()
    oracle.fastlog(x, float(g["MIN"]))
    np.testing.assert_allclose(x, g["log_out64"], rtol=1e-15)
    x32 = g["log_in32"].copy()
    oracle.fastlog(x32, float(g["MIN"]))
    # the simulator takes log10 in float64 and rounds to float32; log10f may differ by 1 ulp
    np.testing.assert_allclose(x32, g["log_out32"], rtol=3e-7, atol=1e-7)
    z = np.array([0.0, -1.0, 1e-310, 1.0], dtype=np.float32)
    oracle.fastlog(z, sys.float_info.min)
    assert np.isneginf(z[:3]).all() and z[3] == 0.0      # Q5: f32 clamp is 0 -> -inf


def test_oracle_pipeline_matches_reference_bayes():
    """bayeslib.bayes run unmodified on the simulator (2 curves, two thicknesses, f32 PL buffer,
    time interpolation, mag_offset) vs oracle.loglik(emulate_f32=True)."""
    path = os.path.join(GOLDEN, "cudasim_bayes.npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    g = golden("cudasim_bayes.npz")
    L, T = int(g["L"]), int(g["T"])
    simPar = [list(g["length"]), float(g["Time"]), L, T, 1, (0,), 7, 10000]
    e_data = [(list(g["t_obs"]), list(g["v_obs"]), list(g["u_obs"]))]
    P = oracle.loglik(g["X"], simPar, g["iniPar"], e_data, emulate_f32=True, solver="pcr")
    np.testing.assert_allclose(P, g["P"], rtol=2e-6)


def _edata_from_golden(g):
    e_data = []
    for e in range(int(g["n_exp"]) if "n_exp" in g.files else 1):
        sfx = "" if e == 0 else "_%d" % e
        e_data.append((list(g["t_obs" + sfx]), list(g["v_obs" + sfx]), list(g["u_obs" + sfx])))
    return e_data


@pytest.mark.parametrize("name", ["bayes_norm2", "bayes_lin"])
def test_oracle_pipeline_flags_match_reference_bayes(name):
    """self_normalize=True with two observation files, and log_pl=False, through the unmodified
    reference bayeslib.bayes on the simulator vs oracle.loglik(emulate_f32=True)."""
    path = os.path.join(GOLDEN, "cudasim_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    g = golden("cudasim_%s.npz" % name)
    L, T = int(g["L"]), int(g["T"])
    simPar = [list(g["length"]), float(g["Time"]), L, T, 1, (0,), 7, 10000]
    P = oracle.loglik(g["X"], simPar, g["iniPar"], _edata_from_golden(g), log_pl=bool(g["log_pl"]),
                      self_normalize=bool(g["self_normalize"]), emulate_f32=True, solver="pcr")
    assert P.shape == g["P"].shape
    np.testing.assert_allclose(P, g["P"], rtol=5e-6)
