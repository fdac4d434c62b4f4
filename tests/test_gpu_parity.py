"""GPU parity tests (run on the B200 box): the sm_100a kernels, reached through the C ABI, against
the CPU oracle (oracle/) and the committed reference goldens.  Tolerances:
  PL      rtol 1e-6 (north_star) + the absolute rounding floor of the PL cancellation; in
          practice agreement is ~1e-12 unless a Newton stop decision flips;
  lnL     rtol 1e-6 on samples whose PL stays above that floor.
"""
import os
import sys

import numpy as np
import pytest

from helpers import (GOLDEN, TRUTH, UC, example_data, golden, pl_noise_floor, power_scan_excitations,
                     prior_samples, simpar_from_golden)

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def trpl():
    import bayesian_inference_trpl_b200 as t
    assert os.path.exists(t._lib.LIB_PATH), "libtrpl_b200.so must ship with the snapshot"
    t._lib.lib()
    return t


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as o
    return o


def _assert_pl_close(pl, ref, mat, simPar, rtol=1e-6):
    floor = pl_noise_floor(mat, simPar[0], simPar[1], simPar[2], simPar[3])[:, None]
    err = np.abs(pl - ref)
    ok = err <= rtol * np.abs(ref) + floor
    assert ok.all(), "max rel err %.3e at %s" % (np.nanmax(err / np.abs(ref)), np.argwhere(~ok)[:5])


# ------------------------------------------------------------------------------------------------
# forward model: reference goldens (tiny shapes the numba simulator can run)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["pvsim_points_f64", "pvsim_points_f32", "pvsim_exp_f64",
                                  "pvsim_stiff_f64", "pvsim_L32_f64", "pvsim_L128_f64", "pvsim_long_f64"])
def test_pvsim_dropin_matches_reference_golden(trpl, name):
    path = os.path.join(GOLDEN, "cudasim_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    g = golden("cudasim_%s.npz" % name)
    simPar = simpar_from_golden(g)
    mode = str(g["init_mode"])
    ini = g["iniPar"].copy() if mode == "points" else list(g["iniPar"])
    ref = g["pl"]
    pl = np.empty_like(ref)
    mat = g["matPar"].copy()
    secs = trpl.pvSim(pl, None, None, None, mat, simPar, ini, (simPar[2],), 8, 1, init_mode=mode)
    assert isinstance(secs, float) and secs >= 0
    np.testing.assert_array_equal(mat, g["matPar"])          # inputs untouched
    if ref.dtype == np.float32:
        np.testing.assert_allclose(pl, ref, rtol=2.5e-7)      # <= 2 ulp of float32
    else:
        np.testing.assert_allclose(pl, ref, rtol=1e-9)


# ------------------------------------------------------------------------------------------------
# forward model vs oracle at real grid sizes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L,length", [(128, 2000.0), (128, 311.0), (64, 1000.0), (32, 500.0),
                                      (256, 2000.0), (96, 1500.0), (16, 250.0), (8, 125.0),
                                      (70, 1100.0), (33, 500.0), (131, 2000.0), (3, 47.0), (2, 31.0)])
def test_solve_pl_matches_oracle(trpl, oracle, L, length):
    T = 1500 if L <= 128 else 400
    Time = 0.025 * T
    simPar = [length, Time, L, T, 1, (0,), 7, 10000]
    X = prior_samples(6, seed=L, stiff=(length < 400))
    X[0] = TRUTH * UC
    x = (np.arange(L) + 0.5) * (length / L)
    for amp in (1.2738e16, 1.6485e18):
        ini = amp * 1e-21 * np.exp(-6e-3 * x)
        ref = oracle.solve(X[:, :12], simPar, ini, solver="pcr" if (L & (L - 1)) == 0 else "thomas")
        pl = np.empty((len(X), T + 1))
        st = np.zeros(len(X), dtype=np.int32)
        trpl.pvSim(pl, None, None, None, X[:, :12], simPar, ini, (128,), 0, 1, init_mode="points",
                   status_out=st)
        assert (st == 0).all() and (ref["status"] == 0).all()
        _assert_pl_close(pl, ref["pl"], X[:, :12], simPar)


def test_iteration_counts_and_status_match_oracle(trpl, oracle):
    L, T, length = 128, 800, 2000.0
    simPar = [length, 0.025 * T, L, T, 1, (0,), 7, 10000]
    X = prior_samples(16, seed=3)
    ini = power_scan_excitations()[2]
    ref = oracle.solve(X[:, :12], simPar, ini, solver="pcr")
    dev = torch.device("cuda", 0)
    mat = torch.from_numpy(np.ascontiguousarray(X[:, :12])).to(dev)
    pl, status, iters = trpl.engine.solve_pl(mat, torch.from_numpy(ini).to(dev), length, simPar[1],
                                            L, T)
    torch.cuda.synchronize()
    it = iters.cpu().numpy()
    assert (status.cpu().numpy() == 0).all()
    # identical stop decisions except (rarely) on a knife edge
    assert np.abs(it - ref["iters"]).max() <= 2, (it, ref["iters"])
    assert (it == ref["iters"]).mean() >= 0.8


def test_nonconvergence_is_reported_per_sample(trpl, oracle):
    L, T, length = 32, 40, 500.0
    simPar = [length, 0.025 * T, L, T, 1, (0,), 12, 3]     # tol 1e-12 within 3 iterations: fails
    X = prior_samples(4, seed=9)
    x = (np.arange(L) + 0.5) * (length / L)
    ini = 1.6e18 * 1e-21 * np.exp(-6e-3 * x)
    ref = oracle.solve(X[:, :12], simPar, ini, solver="pcr")
    pl = np.zeros((4, T + 1))
    st = np.zeros(4, dtype=np.int32)
    trpl.pvSim(pl, None, None, None, X[:, :12], simPar, ini, (32,), 0, 1, init_mode="points",
               status_out=st)
    np.testing.assert_array_equal(st & 1, ref["status"] & 1)
    assert (st & 1).any()
    np.testing.assert_array_equal(np.isnan(pl), np.isnan(ref["pl"]))
    good = ~np.isnan(pl)
    np.testing.assert_allclose(pl[good], ref["pl"][good], rtol=1e-9)


def test_plT_stride_and_legacy_order(trpl, oracle):
    L, T, plT, length = 64, 303, 4, 800.0
    simPar = [length, 0.025 * T, L, T, plT, (0,), 7, 10000]
    X = prior_samples(3, seed=5)
    x = (np.arange(L) + 0.5) * (length / L)
    ini = 1.1e17 * 1e-21 * np.exp(-6e-3 * x)
    ref = oracle.solve(X[:, :12], simPar, ini, solver="pcr")
    pl = np.empty((3, T // plT + 1))
    trpl.pvSim(pl, None, None, None, X[:, :12], simPar, ini, (64,), 0, 1, init_mode="points")
    np.testing.assert_allclose(pl, ref["pl"], rtol=1e-8)
    # BDF order cap 2 (Legacy/pvSim.py) through the device API
    dev = torch.device("cuda", 0)
    ref2 = oracle.solve(X[:, :12], simPar, ini, solver="thomas", max_order=2)
    pl2, _, _ = trpl.engine.solve_pl(torch.from_numpy(np.ascontiguousarray(X[:, :12])).to(dev),
                                     torch.from_numpy(ini).to(dev), length, simPar[1], L, T, plT,
                                     max_order=2)
    np.testing.assert_allclose(pl2.cpu().numpy(), ref2["pl"], rtol=1e-8)


def test_full_length_power_scan_curve(trpl, oracle):
    """BASELINE config shape: L=128, T=80000 (2000 ns), tol 7 -- truth sample, all 3 curves."""
    L, T, length = 128, 80000, 2000.0
    simPar = [length, 2000.0, L, T, 1, (0,), 7, 10000]
    X = prior_samples(2, seed=21)
    X[0] = TRUTH * UC
    inis = power_scan_excitations()
    for c in range(3):
        ref = oracle.solve(X[:, :12], simPar, inis[c], solver="thomas")
        pl = np.empty((2, T + 1))
        trpl.pvSim(pl, None, None, None, X[:, :12], simPar, inis[c], (128,), 0, 1, init_mode="points")
        _assert_pl_close(pl, ref["pl"], X[:, :12], simPar)
        # Testing/compare.py metric: relative L2 error of PL at 6 sample times
        m = T + 1
        tt = np.array([0 * m, 0.01 * m, 0.03 * m, 0.1 * m, 0.3 * m, m - 1], dtype=int)
        for s in range(2):
            nerr = np.linalg.norm(pl[s, tt] - ref["pl"][s, tt]) / np.linalg.norm(ref["pl"][s, tt])
            assert nerr < 1e-7


# ------------------------------------------------------------------------------------------------
# likelihood pieces
# ------------------------------------------------------------------------------------------------
def test_probs_dropins_match_reference_golden(trpl):
    g = golden("cudasim_probs.npz")
    P = g["P_in"].copy()
    secs = trpl.prob(P, g["pli"], g["values"], g["unc"], g["mag"], 128, 148)
    assert isinstance(secs, float)
    np.testing.assert_allclose(P, g["P_out"], rtol=1e-12)
    x = g["log_in64"].copy()
    trpl.fastlog(x, float(g["MIN"]), 128, 148)
    np.testing.assert_allclose(x, g["log_out64"], rtol=1e-14, atol=1e-13)
    x32 = g["log_in32"].copy()
    trpl.fastlog(x32, float(g["MIN"]), 128, 148)
    np.testing.assert_allclose(x32, g["log_out32"], rtol=3e-7, atol=1e-7)
    z = np.array([[0.0, -1.0, 1e-310, 1.0]], dtype=np.float32)
    trpl.fastlog(z, sys.float_info.min, 128, 148)
    assert np.isneginf(z[0, :3]).all() and z[0, 3] == 0.0


def test_prob_large_random(trpl, oracle):
    rng = np.random.default_rng(2)
    S, n = 300, 5001
    pli = rng.uniform(-12, -5, (S, n))
    values = rng.uniform(-12, -5, n)
    mag = rng.uniform(-1, 1, S)
    P = np.zeros(S)
    Pr = np.zeros(S)
    trpl.prob(P, pli, values, np.ones(n), mag, 128, 148)
    oracle.prob(Pr, pli, values, mag)
    np.testing.assert_allclose(P, Pr, rtol=1e-12)


def _synthetic_edata(oracle, simPar, inis, lengths, rng, n_exp=1, every=(1, 3, 7), offgrid=False):
    """Observations = oracle PL of the truth sample (+ small deterministic wiggle), log10."""
    Time, L, T = simPar[1], simPar[2], simPar[3]
    e_data = []
    for e in range(n_exp):
        ts, vs, us = [], [], []
        for c in range(len(inis)):
            sp = list(simPar); sp[0] = lengths[c]
            pl = oracle.solve((TRUTH * UC)[None, :12], sp, inis[c], solver="thomas")["pl"][0]
            grid = np.linspace(0, Time, T + 1)
            idx = np.arange(0, T + 1 - 5 * e, every[(c + e) % len(every)])
            t = grid[idx].copy()
            v = np.log10(pl[idx]) + 0.01 * np.sin(idx / 50.0 + e)
            if offgrid:
                t = np.sort(np.clip(t + rng.uniform(-0.4, 0.4, len(t)) * Time / T, 0, Time))
            ts.append(t); vs.append(v); us.append(np.full(len(t), 0.1))
        e_data.append((ts, vs, us))
    return e_data


@pytest.mark.parametrize("emulate_f32,normalize,log_pl,offgrid,n_exp",
                         [(False, False, True, False, 1), (True, False, True, False, 1),
                          (False, True, True, True, 2), (True, True, True, False, 1),
                          (False, False, False, True, 1)])
def test_fused_loglik_matches_oracle_pipeline(trpl, oracle, emulate_f32, normalize, log_pl, offgrid,
                                              n_exp):
    rng = np.random.default_rng(4)
    L, T = 128, 600
    lengths = [2000.0, 311.0, 2000.0]
    simPar = [lengths, 0.025 * T, L, T, 1, (0,), 7, 10000]
    inis = power_scan_excitations()
    e_data = _synthetic_edata(oracle, simPar, inis, lengths, rng, n_exp=n_exp, offgrid=offgrid)
    if not log_pl:
        e_data = [(ts, [10 ** v for v in vs], us) for ts, vs, us in e_data]
    X = prior_samples(24, seed=8, mag=True)
    X[0] = TRUTH * UC
    ref = oracle.loglik(X, simPar, inis, e_data, log_pl=log_pl, self_normalize=normalize,
                        emulate_f32=emulate_f32, solver="pcr")
    dev = torch.device("cuda", 0)
    prob = trpl.engine.Problem(simPar, inis, e_data, device=0)
    Xd = torch.from_numpy(X).to(dev)
    lnl, status, iters = trpl.engine.solve_loglik(Xd, prob, log_pl=log_pl, self_normalize=normalize,
                                                  emulate_f32=emulate_f32, want_iters=True)
    torch.cuda.synchronize()
    assert (status.cpu().numpy() == 0).all()
    got = lnl.cpu().numpy()
    assert got.shape == ref.shape == (n_exp, len(X))
    rtol = 1e-6 if not emulate_f32 else 2e-5       # log10f vs float64 log10 rounded (1 ulp f32)
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=1e-9)
    # accumulate semantics (probs.py:60): a second call adds to the table
    lnl2, _, _ = trpl.engine.solve_loglik(Xd, prob, log_pl=log_pl, self_normalize=normalize,
                                          emulate_f32=emulate_f32, lnl=lnl.clone())
    np.testing.assert_allclose(lnl2.cpu().numpy(), 2 * got, rtol=1e-12)


def test_bayes_mirror_matches_reference_bayes_golden(trpl):
    """Unmodified reference bayeslib.bayes (simulator run) vs this package's bayes(): same seed,
    same sample matrix, likelihood table within float32-pipeline tolerance."""
    path = os.path.join(GOLDEN, "cudasim_bayes.npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    g = golden("cudasim_bayes.npz")
    L, T = int(g["L"]), int(g["T"])
    simPar = [list(g["length"]), float(g["Time"]), L, T, 1, (0,), 7, 10000]
    e_data = [(list(g["t_obs"]), list(g["v_obs"]), list(g["u_obs"]))]
    sim_flags = {"load_PL_from_file": False, "override_equal_auger": False,
                 "override_equal_mu": False, "override_equal_s": True, "log_pl": True,
                 "self_normalize": False, "random_sample": True, "num_points": len(g["X"])}
    for fused, emu in ((True, True), (False, False), (True, False)):
        gpu_info = {"sims_per_gpu": 2, "num_gpus": 1, "fused": fused, "emulate_f32": emu}
        trpl.bayes_validate.connect_to_gpu(gpu_info, nthreads=128, sims_per_block=1)
        np.random.seed(42)
        N, P, X = trpl.bayeslib.bayes(trpl.pvSim, np.array([0]), None, g["minX"], g["maxX"],
                                      g["do_log"], g["iniPar"], list(simPar), e_data, sim_flags,
                                      gpu_info)
        np.testing.assert_array_equal(X, g["X"])
        np.testing.assert_allclose(P, g["P"], rtol=2e-5 if (emu or not fused) else 1e-4)


def test_lse_partial_matches_numpy(trpl):
    rng = np.random.default_rng(6)
    x = rng.normal(-5000, 2000, 100003)
    x[5] = np.nan
    out = trpl.engine.lse_partial(torch.from_numpy(x).cuda()).cpu().numpy()
    m = np.nanmax(x)
    assert out[0] == m
    np.testing.assert_allclose(out[1], np.nansum(np.exp(x - m)), rtol=1e-12)


def test_dfma_microbenchmark_runs(trpl):
    tf, ms = trpl.engine.bench_dfma(2000)
    assert 1.0 < tf < 100.0


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs 3 and 4 at reduced sample counts
# ------------------------------------------------------------------------------------------------
def _shipped_observations(name, t_max):
    """(t_list, log10 PL list, unc list) from the condensed fixture of a shipped stiff-regime
    observation file (every 25th point; values scaled like bayes_io.get_data, bayes_io.py:15)."""
    ex = example_data()
    ts, vs, us = [], [], []
    for c in range(3):
        t = ex["%s_t%d" % (name, c)]
        v = ex["%s_pl%d" % (name, c)] * 1e-23
        keep = t <= t_max
        ts.append(t[keep]); vs.append(np.log10(v[keep])); us.append(np.full(keep.sum(), 0.1))
    return (ts, vs, us)


def test_config3_stiff_surface_regime_shipped_observations(trpl, oracle):
    """Highfrontsurf / Highbacksurf / Balancedhighsurf observation files (3 files = 3 experiments
    in one fused call), stiff prior Sf,Sb in [1, 1e5] cm/s, 100 ns window."""
    L, T = 128, 4000
    simPar = [2000.0, 0.025 * T, L, T, 1, (0,), 7, 10000]
    inis = power_scan_excitations()
    e_data = [_shipped_observations(n, simPar[1]) for n in
              ("Highfrontsurf", "Highbacksurf", "Balancedhighsurf")]
    X = prior_samples(20, seed=31, stiff=True)
    for k, (sf, sb) in enumerate(((1e4, 10), (10, 1e4), (5e3, 5e3))):      # identified truths
        x = TRUTH.copy(); x[5], x[6] = sf, sb
        X[k] = x * UC
    ref = oracle.loglik(X, simPar, inis, e_data, solver="pcr")
    prob = trpl.engine.Problem(simPar, inis, e_data, device=0)
    lnl, status, _ = trpl.engine.solve_loglik(torch.from_numpy(X).cuda(), prob)
    torch.cuda.synchronize()
    assert (status.cpu().numpy() == 0).all()
    got = lnl.cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-9)
    # each truth sample is the most likely one for "its" file among the 20
    for e in range(3):
        assert np.argmax(got[e]) == e


def test_config4_two_thickness_six_curves(trpl, oracle):
    """Twothick excitations: 6 curves alternating 311 nm / 2000 nm."""
    ex = example_data()
    inis = ex["twothick_exc"] * 1e-21
    lengths = [311.0, 2000.0] * 3
    L, T = 128, 500
    simPar = [lengths, 0.025 * T, L, T, 1, (0,), 7, 10000]
    rng = np.random.default_rng(12)
    e_data = _synthetic_edata(oracle, simPar, inis, lengths, rng, n_exp=1, every=(1, 2, 5))
    X = prior_samples(10, seed=41)
    X[0] = TRUTH * UC
    ref = oracle.loglik(X, simPar, inis, e_data, solver="pcr")
    prob = trpl.engine.Problem(simPar, inis, e_data, device=0)
    lnl, status, iters = trpl.engine.solve_loglik(torch.from_numpy(X).cuda(), prob, want_iters=True)
    torch.cuda.synchronize()
    assert (status.cpu().numpy() == 0).all()
    assert iters.shape == (6, 10)
    np.testing.assert_allclose(lnl.cpu().numpy(), ref, rtol=1e-6, atol=1e-9)
    assert np.argmax(lnl.cpu().numpy()[0]) == 0


def test_likelihood_properties_at_full_size(trpl):
    """BASELINE-size property checks that need no oracle run: lnL <= 0, the generating sample
    scores exactly like itself (lnL = 0 up to rounding), mag_offset shifts add n*m^2-type terms
    consistently, and two identical rows give bit-identical results (determinism)."""
    L, T = 128, 80000
    simPar = [2000.0, 2000.0, L, T, 1, (0,), 7, 10000]
    inis = power_scan_excitations()
    truth = (TRUTH * UC)
    grid = np.linspace(0, 2000.0, T + 1)
    ts, vs, us = [], [], []
    for c in range(3):
        pl = np.empty((1, T + 1))
        trpl.pvSim(pl, None, None, None, truth[None, :12], simPar, inis[c], (128,), 0, 1, init_mode="points")
        ts.append(grid.copy()); vs.append(np.log10(pl[0])); us.append(np.full(T + 1, 0.1))
    prob = trpl.engine.Problem(simPar, inis, [(ts, vs, us)], device=0)
    X = prior_samples(6, seed=77)
    X[0] = truth
    X[1] = truth; X[1, 12] = 0.25            # same curves, shifted by mag_offset
    X[3] = X[2]
    lnl, status, _ = trpl.engine.solve_loglik(torch.from_numpy(X).cuda(), prob)
    got = lnl.cpu().numpy()[0]
    assert (status.cpu().numpy() == 0).all()
    assert (got <= 0).all()
    assert abs(got[0]) < 1e-12
    np.testing.assert_allclose(got[1], -3 * (T + 1) * 0.25 ** 2, rtol=1e-9)
    assert got[2] == got[3]
    assert got[2] < got[0]


# ------------------------------------------------------------------------------------------------
# fine grids: one CTA of W warps per simulation (BASELINE config 5 shape: L = 1000)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L,T", [(1000, 250), (512, 300), (1024, 120), (2048, 60), (260, 300),
                                 (257, 200), (1001, 120), (774, 150), (264, 200), (1528, 80)])
def test_fine_grid_cta_kernel_matches_oracle(trpl, oracle, L, T):
    length = 2000.0
    simPar = [length, 0.025 * T, L, T, 1, (0,), 7, 10000]
    X = prior_samples(5, seed=L + 1)
    X[0] = TRUTH * UC
    x = (np.arange(L) + 0.5) * (length / L)
    for amp in (1.2738e16, 1.6485e18):
        ini = amp * 1e-21 * np.exp(-6e-3 * x)
        ref = oracle.solve(X[:, :12], simPar, ini, solver="pcr" if (L & (L - 1)) == 0 else "thomas")
        pl = np.empty((len(X), T + 1))
        st = np.zeros(len(X), dtype=np.int32)
        trpl.pvSim(pl, None, None, None, X[:, :12], simPar, ini, (128,), 0, 1, init_mode="points",
                   status_out=st)
        assert (st == 0).all() and (ref["status"] == 0).all()
        _assert_pl_close(pl, ref["pl"], X[:, :12], simPar)


def test_fine_grid_fused_likelihood(trpl, oracle):
    """Config 5 shape at reduced size: L=1000, 3 curves dN_c(x) = A_c exp(-6e-3 x)."""
    L, T, length = 1000, 160, 2000.0
    simPar = [length, 0.025 * T, L, T, 1, (0,), 7, 10000]
    x = (np.arange(L) + 0.5) * (length / L)
    inis = np.stack([a * 1e-21 * np.exp(-6e-3 * x) for a in (1.2738e16, 1.1539e17, 1.6485e18)])
    rng = np.random.default_rng(2)
    e_data = _synthetic_edata(oracle, simPar, inis, [length] * 3, rng, n_exp=1, every=(1, 2, 4))
    X = prior_samples(7, seed=1000, mag=True)
    X[0] = TRUTH * UC
    ref = oracle.loglik(X, simPar, inis, e_data, solver="thomas")
    prob = trpl.engine.Problem(simPar, inis, e_data, device=0)
    lnl, status, _ = trpl.engine.solve_loglik(torch.from_numpy(X).cuda(), prob)
    torch.cuda.synchronize()
    assert (status.cpu().numpy() == 0).all()
    np.testing.assert_allclose(lnl.cpu().numpy(), ref, rtol=1e-6, atol=1e-9)


def test_entry_script_end_to_end(trpl, oracle, tmp_path):
    """EXC.csv + OBS.csv in, BAYRAN_P/X out: the whole reference workflow (parallel_bayes_gpu.py)
    on files written in the reference formats; likelihood table checked against the oracle."""
    L, T = 128, 300
    Time = 0.025 * T
    ex = example_data()
    exc_path = str(tmp_path / "exc.csv")
    with open(exc_path, "w") as fh:
        for row in ex["power_exc"]:
            fh.write(",".join("%.8E" % v for v in row) + ",\n")
    inis = trpl.bayes_io.get_initpoints(exc_path, {"select_obs_sets": None})
    simPar = [2000.0, Time, L, T, 1, (0,), 7, 10000]
    grid = np.linspace(0, Time, T + 1)
    pls = [oracle.solve((TRUTH * UC)[None, :12], simPar, inis[c], solver="thomas")["pl"][0] for c in range(3)]
    obs_path = str(tmp_path / "obs.csv")
    trpl.bayes_io.write_observations(obs_path, [grid] * 3, pls)
    cfg = trpl.parallel_bayes_gpu.default_config()
    cfg.update(Length=2000.0, Time=Time, T=T)
    cfg["ic_flags"]["time_cutoff"] = Time
    cfg["sim_flags"]["num_points"] = 12
    cfg["gpu_info"]["sims_per_gpu"] = 5
    out = str(tmp_path / "RUN")
    P, X = trpl.parallel_bayes_gpu.run(exc_path, [obs_path], [out], cfg=cfg, posterior=True)
    Pf = np.load(os.path.join(out, "RUN_BAYRAN_P.npy"))
    Xf = np.load(os.path.join(out, "RUN_BAYRAN_X.npy"))
    W = np.load(os.path.join(out, "RUN_BAYRAN_W.npy"))
    assert Pf.shape == (12,) and Xf.shape == (12, 13)
    np.testing.assert_array_equal(Pf, P[0])
    # same draw as the reference sampler with seed 42
    np.random.seed(42)
    Xr = trpl.bayeslib.random_grid(cfg["minX"] * trpl.parallel_bayes_gpu.unit_conversions,
                                   cfg["maxX"] * trpl.parallel_bayes_gpu.unit_conversions,
                                   cfg["do_log"], 12)
    np.testing.assert_allclose(Xf * trpl.parallel_bayes_gpu.unit_conversions, Xr, rtol=1e-15)
    e_data = trpl.bayes_io.get_data([obs_path], cfg["ic_flags"], cfg["sim_flags"])
    ref = oracle.loglik(Xr, simPar, inis, e_data, solver="pcr")
    np.testing.assert_allclose(Pf, ref[0], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(W.sum(), 1.0, rtol=1e-12)
    np.testing.assert_allclose(W, np.exp(Pf - Pf.max()) / np.exp(Pf - Pf.max()).sum(), rtol=1e-10)


# ------------------------------------------------------------------------------------------------
# neighbours of the path: device-side sampling and posterior products
# ------------------------------------------------------------------------------------------------
def _philox4x32_10(c, k):
    """numpy restatement of Philox4x32-10 (Salmon et al., Random123): c [n,4] uint32, k (k0,k1)."""
    c = c.astype(np.uint64)
    k0, k1 = np.uint64(k[0]), np.uint64(k[1])
    M0, M1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & mask, (k1 + np.uint64(0xBB67AE85)) & mask
    return np.stack([c0, c1, c2, c3], axis=1)


def test_device_random_grid_matches_numpy_philox(trpl):
    from helpers import DO_LOG, MAXX, MINX
    lo, hi = MINX * UC, MAXX * UC
    lo[2:4] = 0.5 * UC[2:4]
    S, seed, first = 5000, 0x1234567812345678, 7 * 2 ** 32 + 11
    X = trpl.engine.random_grid_device(lo, hi, DO_LOG, S, seed, first_sample=first, override_flags=2).cpu().numpy()
    ids = first + np.arange(S, dtype=np.uint64)
    ref = np.empty((S, 13))
    for j in range(13):
        src = 5 if j == 6 else j                       # override_equal_s
        c = np.stack([ids & np.uint64(0xFFFFFFFF), ids >> np.uint64(32),
                      np.full(S, src, np.uint64), np.zeros(S, np.uint64)], axis=1)
        r = _philox4x32_10(c, (seed & 0xFFFFFFFF, seed >> 32))
        u = (((r[:, 0] << np.uint64(32)) | r[:, 1]) >> np.uint64(11)).astype(np.float64) / 2.0 ** 53
        if lo[src] == hi[src]:
            ref[:, j] = lo[src]
        elif DO_LOG[src]:
            a, b = np.log10(lo[src]), np.log10(hi[src])
            ref[:, j] = 10 ** (a + (b - a) * u)
        else:
            ref[:, j] = lo[src] + (hi[src] - lo[src]) * u
    np.testing.assert_allclose(X, ref, rtol=4e-15)
    np.testing.assert_array_equal(X[:, 6], X[:, 5])
    assert ((X >= lo * (1 - 1e-14)) & (X <= hi * (1 + 1e-14))).all()
    # shards of one global draw are reproducible independently
    Xb = trpl.engine.random_grid_device(lo, hi, DO_LOG, 100, seed, first_sample=first + 400, override_flags=2)
    np.testing.assert_array_equal(Xb.cpu().numpy(), X[400:500])


def test_posterior_products_match_numpy(trpl):
    rng = np.random.default_rng(10)
    S = 200003
    X = rng.normal(size=(S, 13)) * np.arange(1, 14) + 3.0
    lnP = -0.5 * ((X[:, 2] - 3.5) ** 2 + (X[:, 5] - 2.0) ** 2 / 4) - 4000.0
    lnP[17] = np.nan
    Xd, Ld = torch.from_numpy(X).cuda(), torch.from_numpy(lnP).cuda()
    w = trpl.posterior.normalize(Ld)
    wr = np.exp(lnP - np.nanmax(lnP)); wr[np.isnan(wr)] = 0; wr /= wr.sum()
    np.testing.assert_allclose(w.cpu().numpy(), wr, rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(float(trpl.posterior.log_evidence(Ld)),
                               np.nanmax(lnP) + np.log(np.nansum(np.exp(lnP - np.nanmax(lnP)))), rtol=1e-13)
    dens, bins = trpl.posterior.marginalize_1D(w, Xd, 2, -9.0, 15.0, 96)
    ref, rb = np.histogram(X[:, 2], weights=wr, bins=np.linspace(-9, 15, 97), density=True)
    # numpy.histogram accumulates weights through a cumulative sum (absolute error ~eps*total), so
    # its tail bins are noise; the device result is compared with an absolute floor for those
    np.testing.assert_allclose(dens.cpu().numpy(), ref, rtol=1e-9, atol=1e-12 * ref.max())
    exact = np.bincount(np.clip(((X[:, 2] + 9.0) / 0.25).astype(int), 0, 95)[(X[:, 2] >= -9) & (X[:, 2] <= 15)],
                        weights=wr[(X[:, 2] >= -9) & (X[:, 2] <= 15)], minlength=96)
    np.testing.assert_allclose(dens.cpu().numpy(), exact / (exact.sum() * 0.25), rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(bins.cpu().numpy(), rb, rtol=1e-14, atol=1e-14)
    d2 = trpl.posterior.marginalize_2D(w, Xd, 2, 5, -9.0, 15.0, -20.0, 25.0, 32)
    r2 = np.histogram2d(X[:, 2], X[:, 5], bins=[np.linspace(-9, 15, 33), np.linspace(-20, 25, 33)],
                        weights=wr, density=True)[0]
    np.testing.assert_allclose(d2.cpu().numpy(), r2, rtol=1e-9, atol=1e-300)
    cnt = trpl.engine.weighted_hist(Xd, 0, None, -3.0, 9.0, 50).cpu().numpy()
    np.testing.assert_array_equal(cnt, np.histogram(X[:, 0], bins=np.linspace(-3, 9, 51))[0])
    mean, cov = trpl.posterior.moments(w, Xd)
    np.testing.assert_allclose(mean.cpu().numpy(), np.average(X, axis=0, weights=wr), rtol=1e-10)
    np.testing.assert_allclose(cov.cpu().numpy(), np.cov(X.T, aweights=wr, ddof=0), rtol=1e-7, atol=1e-9)


def test_raw_cabi_binding_as_documented_in_integration_md(trpl, oracle):
    """Call libtrpl_b200.so exactly the way INTEGRATION.md tells a reference maintainer to: plain
    ctypes, raw device pointers, no helper from this package in between."""
    import ctypes
    lib = ctypes.CDLL(trpl._lib.LIB_PATH)
    lib.trpl_solve_pl.restype = ctypes.c_int
    lib.trpl_solve_pl.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                  ctypes.c_double, ctypes.c_double] + [ctypes.c_int] * 7 + \
                                 [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    lib.trpl_error_string.restype = ctypes.c_char_p
    L, T, length, Time = 64, 200, 900.0, 5.0
    X = prior_samples(5, seed=64)
    x = (np.arange(L) + 0.5) * (length / L)
    ini = 3e17 * 1e-21 * np.exp(-6e-3 * x)
    d_mat = torch.from_numpy(np.ascontiguousarray(X[:, :12])).cuda()
    d_ini = torch.from_numpy(ini).cuda()
    d_pl = torch.empty((5, T + 1), dtype=torch.float64, device="cuda")
    d_st = torch.zeros(5, dtype=torch.int32, device="cuda")
    rc = lib.trpl_solve_pl(d_mat.data_ptr(), 5, 12, d_ini.data_ptr(), length, Time, L, T, 1, 7, 10000, 5, 0,
                           d_pl.data_ptr(), 0, T + 1, d_st.data_ptr(), None, 0, None)
    assert rc == 0, lib.trpl_error_string(rc)
    torch.cuda.synchronize()
    ref = oracle.solve(X[:, :12], [length, Time, L, T, 1, (0,), 7, 10000], ini, solver="pcr")
    np.testing.assert_allclose(d_pl.cpu().numpy(), ref["pl"], rtol=1e-8)
    # argument errors come back as codes, not exceptions or crashes
    assert lib.trpl_solve_pl(d_mat.data_ptr(), 5, 11, d_ini.data_ptr(), length, Time, L, T, 1, 7, 10000, 5, 0,
                             d_pl.data_ptr(), 0, T + 1, None, None, 0, None) == -1
    assert lib.trpl_solve_pl(d_mat.data_ptr(), 5, 12, d_ini.data_ptr(), length, Time, 4096, T, 1, 7, 10000, 5, 0,
                             d_pl.data_ptr(), 0, T + 1, None, None, 0, None) == -2
    assert lib.trpl_solve_pl(d_mat.data_ptr(), 5, 12, d_ini.data_ptr(), length, Time, L, T, 1, 7, 10000, 5, 0,
                             d_pl.data_ptr(), 0, T + 1, None, None, 99, None) == -4


def test_degenerate_parameters_match_oracle(trpl, oracle):
    """Corners of the reference prior: zero mobility (the entry script's minX, parallel_bayes_gpu.py:91),
    zero surface velocity, no Auger, extreme lifetimes."""
    L, T, length = 128, 600, 2000.0
    simPar = [length, 0.025 * T, L, T, 1, (0,), 7, 10000]
    X = np.tile(TRUTH * UC, (8, 1))
    X[0, 2] = 0.0                       # mu_n = 0
    X[1, 3] = 0.0                       # mu_p = 0
    X[2, 2:4] = 0.0                     # no transport at all: diagonal systems
    X[3, 5:7] = 0.0                     # S = 0
    X[4, 7:9] = 0.0                     # no Auger
    X[5, 9:11] = [1.0, 1.0]             # 1 ns lifetimes
    X[6, 9:11] = [1e6, 1e6]             # essentially no SRH
    X[7, 4] = 1e-15 * UC[4]             # negligible radiative rate
    ini = power_scan_excitations()[2]
    ref = oracle.solve(X[:, :12], simPar, ini, solver="pcr")
    pl = np.empty((8, T + 1))
    st = np.ones(8, dtype=np.int32)
    trpl.pvSim(pl, None, None, None, X[:, :12], simPar, ini, (128,), 0, 1, init_mode="points", status_out=st)
    np.testing.assert_array_equal(st, ref["status"])
    assert (st == 0).all()
    _assert_pl_close(pl, ref["pl"], X[:, :12], simPar, rtol=1e-8)


_NCCL_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
import bayesian_inference_trpl_b200 as trpl
from bayesian_inference_trpl_b200 import distributed as D
from helpers import TRUTH, UC, MINX, MAXX, DO_LOG, power_scan_excitations
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
L, T = 128, 200
simPar = [2000.0, 0.025 * T, L, T, 1, (0,), 7, 10000]
inis = power_scan_excitations()
grid = np.linspace(0, simPar[1], T + 1)
e_data = [([grid.copy()] * 3, [np.linspace(-6.5, -7.5, T + 1)] * 3, [np.full(T + 1, .1)] * 3)]
flags = {"load_PL_from_file": False, "override_equal_auger": False, "override_equal_mu": False,
         "override_equal_s": False, "log_pl": True, "self_normalize": False, "random_sample": True, "num_points": 37}
lo, hi = MINX * UC, MAXX * UC
lo[2:4] = 0.5 * UC[2:4]
def run(num_gpus):
    gi = {"sims_per_gpu": 5, "num_gpus": num_gpus}
    trpl.bayes_validate.connect_to_gpu(gi)
    np.random.seed(42)
    return trpl.bayeslib.bayes(trpl.pvSim, np.array([0]), None, lo, hi, DO_LOG, inis, list(simPar), e_data, flags, gi)
N, P, X = run(world)                       # block-cyclic share of this rank (RANK from torchrun)
full = D.merge_block_cyclic(P).cpu().numpy()
os.environ.pop("RANK"); N1, P1, X1 = run(1); os.environ["RANK"] = str(rank)   # whole table on one GPU
assert np.array_equal(X, X1)
assert np.allclose(full, P1, rtol=1e-12, atol=0), np.abs(full - P1).max()
own = np.zeros(37, bool)
for b in range(rank * 5, 37, world * 5): own[b:b + 5] = True
assert (P[0][~own] == 0).all() and (P[0][own] != 0).all()
# contiguous shards + NCCL gather + global log-sum-exp
a, b = D.shard_bounds(37, rank, world)
loc = torch.from_numpy(P1[:, a:b].copy()).cuda()
got = D.gather_rows(loc, 37).cpu().numpy()
assert np.array_equal(got, P1)
lse = D.global_logsumexp(trpl.engine.lse_partial(loc[0].contiguous()))
m = P1[0].max(); ref = m + np.log(np.exp(P1[0] - m).sum())
assert abs(float(lse) - ref) < 1e-10 * abs(ref)
w = trpl.posterior.normalize(loc[0].contiguous())
tot = w.sum(); dist.all_reduce(tot)
assert abs(float(tot) - 1.0) < 1e-12
# Philox sampler on the product path: every rank draws only its own rows on its GPU, tables gathered at the end
pf = dict(flags, sampler="philox", seed=11, num_points=41)
Nq, Pq, Xq = trpl.bayeslib.bayes(trpl.pvSim, np.array([0]), None, lo, hi, DO_LOG, inis, list(simPar), e_data, pf,
                                 {"sims_per_gpu": 5, "num_gpus": world, "has_GPU": True})
Xall = trpl.engine.random_grid_device(lo, hi, DO_LOG, 41, 11)
assert np.array_equal(Xq, Xall.cpu().numpy()), "shards do not concatenate to the single-rank draw"
prob = trpl.engine.Problem(simPar, inis, e_data, device=rank)
ref_l, st, _ = trpl.engine.solve_loglik(Xall, prob)
assert np.allclose(Pq, ref_l.cpu().numpy(), rtol=1e-12, atol=0)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_gpu_nccl_sharding_and_merge(trpl, tmp_path):
    """Two ranks on two GPUs: block-cyclic bayes() shares merged over NCCL equal the one-GPU table;
    all-gather of contiguous shards and the global log-sum-exp (skipped on a one-GPU box)."""
    import subprocess
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "nccl_worker.py"
    script.write_text(_NCCL_WORKER % {"root": root})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_fused_path_marks_nonconverged_samples(trpl, oracle):
    """max_iter too small for some samples: their lnL is NaN and status bit0 is set, the others
    are unaffected (the reference would abort the whole launch, pvSimPCR.py:269-274)."""
    L, T, length = 128, 120, 2000.0
    simPar = [length, 0.025 * T, L, T, 1, (0,), 7, 80]           # too few for the first steps of some samples
    inis = power_scan_excitations()[0:2]
    X = prior_samples(12, seed=123)
    grid = np.linspace(0, simPar[1], T + 1)
    e_data = [([grid.copy()] * 2, [np.full(T + 1, -7.0)] * 2, [np.full(T + 1, 0.1)] * 2)]
    ref = oracle.loglik(X, simPar, inis, e_data, solver="pcr")
    prob = trpl.engine.Problem(simPar, inis, e_data, device=0)
    lnl, status, _ = trpl.engine.solve_loglik(torch.from_numpy(X).cuda(), prob)
    got, st = lnl.cpu().numpy()[0], status.cpu().numpy()
    assert (st != 0).any() and (st == 0).any()
    np.testing.assert_array_equal(np.isnan(got), st != 0)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref[0]))
    ok = st == 0
    np.testing.assert_allclose(got[ok], ref[0][ok], rtol=1e-6)


def test_more_curves_and_files_than_one_launch_holds(trpl, oracle):
    """10 curves x 5 observation files: tiled over launches of <= 8 curves x <= 4 files."""
    L, T = 64, 60
    rng = np.random.default_rng(8)
    lengths = [float(v) for v in rng.uniform(300, 2000, 10)]
    simPar = [lengths, 0.025 * T, L, T, 1, (0,), 7, 10000]
    inis = np.stack([a * 1e-21 * np.exp(-6e-3 * (np.arange(L) + 0.5) * (lengths[c] / L))
                     for c, a in enumerate(10 ** rng.uniform(16, 18, 10))])
    grid = np.linspace(0, simPar[1], T + 1)
    e_data = []
    for e in range(5):
        idx = [np.sort(rng.choice(T + 1, size=rng.integers(5, T), replace=False)) for _ in range(10)]
        e_data.append(([grid[i] for i in idx], [rng.uniform(-8, -5, len(i)) for i in idx],
                       [np.full(len(i), 0.1) for i in idx]))
    X = prior_samples(6, seed=55, mag=True)
    ref = oracle.loglik(X, simPar, inis, e_data, solver="pcr")
    prob = trpl.engine.Problem(simPar, inis, e_data, device=0)
    assert len(prob.parts) == 4
    lnl, status, iters = trpl.engine.solve_loglik(torch.from_numpy(X).cuda(), prob, want_iters=True)
    assert (status.cpu().numpy() == 0).all() and iters.shape == (10, 6) and (iters.cpu().numpy() > 0).all()
    np.testing.assert_allclose(lnl.cpu().numpy(), ref, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("name", ["bayes_norm2", "bayes_lin"])
def test_bayes_mirror_flags_match_reference_golden(trpl, name):
    """self_normalize / two observation files / log_pl=False: this package's bayes() (fused, float32
    emulation) vs the unmodified reference bayeslib.bayes recorded on the simulator."""
    path = os.path.join(GOLDEN, "cudasim_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    g = golden("cudasim_%s.npz" % name)
    L, T = int(g["L"]), int(g["T"])
    simPar = [list(g["length"]), float(g["Time"]), L, T, 1, (0,), 7, 10000]
    e_data = []
    for e in range(int(g["n_exp"])):
        sfx = "" if e == 0 else "_%d" % e
        e_data.append((list(g["t_obs" + sfx]), list(g["v_obs" + sfx]), list(g["u_obs" + sfx])))
    sim_flags = {"load_PL_from_file": False, "override_equal_auger": False, "override_equal_mu": False,
                 "override_equal_s": True, "log_pl": bool(g["log_pl"]), "self_normalize": bool(g["self_normalize"]),
                 "random_sample": True, "num_points": len(g["X"])}
    for fused in (True, False):
        gpu_info = {"sims_per_gpu": 2, "num_gpus": 1, "fused": fused, "emulate_f32": True}
        trpl.bayes_validate.connect_to_gpu(gpu_info, nthreads=128, sims_per_block=1)
        np.random.seed(42)
        N, P, X = trpl.bayeslib.bayes(trpl.pvSim, np.array([0]), None, g["minX"], g["maxX"], g["do_log"],
                                      g["iniPar"], list(simPar), e_data, sim_flags, gpu_info)
        np.testing.assert_array_equal(X, g["X"])
        np.testing.assert_allclose(P, g["P"], rtol=3e-5)
