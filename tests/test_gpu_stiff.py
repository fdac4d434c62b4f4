"""Stiff surface-recombination regime on a thin film (BASELINE config 3's hard corner: prior widened to
Sf,Sb in [1,1e5] cm/s, Length = 311 nm as in the reference entry script, parallel_bayes_gpu.py:72).
PL collapses by >13 decades within nanoseconds; what is left of the window is cancellation noise of
rate*(sum N*P - L*N0*P0) (pvSimPCR.py:278-281), in which two valid FP64 evaluations of the REFERENCE
algorithm (oracle PCR = the reference's kernels, oracle Thomas = Legacy/pvSim.py's solve) already
disagree.  The CUDA path must (a) take the same Newton iterations as the oracle (pvSimPCR.py:213-216)
and (b) agree with the oracle at least as often as the oracle agrees with itself."""
import numpy as np
import pytest

from helpers import TRUTH, UC, pl_noise_floor, power_scan_excitations, prior_samples

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_stiff_regime_parity_is_as_good_as_the_oracles_own():
    import bayesian_inference_trpl_b200 as trpl
    from oracle import oracle
    S, T, L, length = 256, 8000, 128, 311.0
    simPar = [length, 0.025 * T, L, T, 1, (0,), 7, 10000]
    inis = power_scan_excitations()
    X = prior_samples(S, seed=2024, stiff=True, mag=True)
    X[0] = TRUTH * UC
    dev = torch.device("cuda", 0)
    mat = torch.from_numpy(np.ascontiguousarray(X[:, :12])).to(dev)
    floor = pl_noise_floor(X[:, :12], length, simPar[1], L, T)[:, None]
    n_pts = n_mine = n_yard = 0
    it_equal = 0
    truth_pl = []
    for c in range(3):
        ref = oracle.solve(X[:, :12], simPar, inis[c], solver="pcr")
        yard = oracle.solve(X[:, :12], simPar, inis[c], solver="thomas")
        pl, st, it = trpl.engine.solve_pl(mat, torch.from_numpy(inis[c]).to(dev), length, simPar[1], L, T)
        pl, it = pl.cpu().numpy(), it.cpu().numpy()
        assert (st.cpu().numpy() == ref["status"]).all()
        truth_pl.append(ref["pl"][0])
        sig = np.abs(ref["pl"]) > 1e3 * floor
        n_pts += sig.sum()
        n_mine += (np.abs(pl - ref["pl"])[sig] <= 1e-6 * np.abs(ref["pl"])[sig]).sum()
        n_yard += (np.abs(yard["pl"] - ref["pl"])[sig] <= 1e-6 * np.abs(ref["pl"])[sig]).sum()
        it_equal += (it == ref["iters"]).sum()
        assert np.abs(it - ref["iters"]).max() <= 2
    f_mine, f_yard = n_mine / n_pts, n_yard / n_pts
    print("PL within 1e-6: CUDA %.4f, oracle Thomas-vs-PCR %.4f; Newton totals equal on %d/%d" % (f_mine, f_yard, it_equal, 3 * S))
    assert it_equal >= 0.99 * 3 * S                      # identical stop decisions (pvSimPCR.py:213-216)
    assert f_mine >= f_yard - 0.005

    grid = np.linspace(0, simPar[1], T + 1)
    e_data = [([grid.copy()] * 3, [np.log10(p) for p in truth_pl], [np.full(T + 1, 0.1)] * 3)]
    ref_l = oracle.loglik(X, simPar, inis, e_data, solver="pcr")[0]
    yard_l = oracle.loglik(X, simPar, inis, e_data, solver="thomas")[0]
    prob = trpl.engine.Problem(simPar, inis, e_data, device=0)
    lnl, st, _ = trpl.engine.solve_loglik(torch.from_numpy(X).to(dev), prob)
    got = lnl.cpu().numpy()[0]
    ok = np.isfinite(ref_l) & (np.abs(ref_l) > 1e-6)
    assert np.array_equal(np.isfinite(got), np.isfinite(ref_l))
    f_mine = (np.abs(got - ref_l)[ok] <= 1e-6 * np.abs(ref_l)[ok]).mean()
    f_yard = (np.abs(yard_l - ref_l)[ok] <= 1e-6 * np.abs(ref_l)[ok]).mean()
    print("lnL within 1e-6: CUDA %.4f, oracle Thomas-vs-PCR %.4f" % (f_mine, f_yard))
    assert f_mine >= f_yard - 0.005
    # the samples whose curves stay clear of the noise agree tightly
    clean = ok & (np.abs(yard_l - ref_l) <= 1e-9 * np.abs(ref_l))
    assert clean.sum() > 0.3 * S
    np.testing.assert_allclose(got[clean], ref_l[clean], rtol=1e-6)
