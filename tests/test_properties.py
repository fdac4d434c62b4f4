"""Property-based CPU tests (hypothesis) of the host-side pieces around the hot path."""
import ctypes

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle


@settings(max_examples=60, deadline=None)
@given(T=st.integers(1, 400), time=st.floats(0.1, 5000.0), n=st.integers(0, 60), seed=st.integers(0, 2 ** 31 - 1))
def test_obs_prepare_brackets_every_time(T, time, n, seed):
    """hi is the searchsorted-left bracket, weights are in [0,1], sum to 1 and reproduce the time."""
    import bayesian_inference_trpl_b200 as trpl
    lib = trpl._lib.lib()
    rng = np.random.default_rng(seed)
    grid = np.linspace(0, time, T + 1)
    times = np.sort(np.concatenate([rng.uniform(0, time, n), grid[rng.integers(0, T + 1, 5)]]))
    m = len(times)
    hi = np.empty(m, np.int32); whi = np.empty(m); wlo = np.empty(m)
    rc = lib.trpl_obs_prepare(times.ctypes.data_as(ctypes.c_void_p), m, float(time), T,
                              hi.ctypes.data_as(ctypes.c_void_p), whi.ctypes.data_as(ctypes.c_void_p),
                              wlo.ctypes.data_as(ctypes.c_void_p))
    assert rc >= 1
    np.testing.assert_array_equal(hi, np.clip(np.searchsorted(grid, times), 1, T))
    assert (whi >= -1e-12).all() and (whi <= 1 + 1e-12).all()
    np.testing.assert_allclose(whi + wlo, 1.0, atol=1e-12)
    np.testing.assert_allclose(whi * grid[hi] + wlo * grid[hi - 1], times, atol=1e-9 * max(time, 1.0))
    assert (np.diff(hi) >= 0).all() and rc == hi.max()


@settings(max_examples=40, deadline=None)
@given(n=st.integers(0, 10 ** 7), world=st.integers(1, 16))
def test_shards_partition_the_samples(n, world):
    from bayesian_inference_trpl_b200.distributed import shard_bounds
    b = [shard_bounds(n, r, world) for r in range(world)]
    assert b[0][0] == 0 and b[-1][1] == n
    assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
    sizes = [hi - lo for lo, hi in b]
    assert max(sizes) - min(sizes) <= 1


@settings(max_examples=25, deadline=None)
@given(S=st.integers(1, 12), n=st.integers(1, 200), seed=st.integers(0, 2 ** 31 - 1))
def test_oracle_lnp_is_minus_sum_of_squares(S, n, seed):
    rng = np.random.default_rng(seed)
    pl = rng.uniform(-20, 0, (S, n)); val = rng.uniform(-20, 0, n); mag = rng.uniform(-1, 1, S)
    P = np.zeros(S)
    oracle.prob(P, pl, val, mag)
    np.testing.assert_allclose(P, -np.sum((pl + mag[:, None] - val) ** 2, axis=1), rtol=1e-12)


@settings(max_examples=10, deadline=None)
@given(seed=st.integers(0, 10 ** 6), L=st.sampled_from([4, 8, 16, 32]))
def test_oracle_pcr_and_thomas_agree(seed, L):
    """The two tridiagonal solvers of the oracle (reference PCR vs Legacy Thomas) give the same PL."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import prior_samples
    X = prior_samples(2, seed=seed)
    T = 25
    simPar = [15.625 * L, 0.025 * T, L, T, 1, (0,), 7, 10000]
    ini = 1e-5 * np.exp(-np.arange(L) / 5.0)
    a = oracle.solve(X[:, :12], simPar, ini, solver="pcr")
    b = oracle.solve(X[:, :12], simPar, ini, solver="thomas")
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    np.testing.assert_allclose(a["pl"], b["pl"], rtol=1e-9)
    np.testing.assert_array_equal(a["iters"], b["iters"])


@settings(max_examples=40, deadline=None)
@given(curves=st.lists(st.integers(1, 40), min_size=1, max_size=5), seed=st.integers(0, 2 ** 31 - 1),
       crlf=st.booleans(), blanks=st.booleans())
def test_vectorised_observation_reader_equals_row_by_row_parsing(tmp_path_factory, curves, seed, crlf, blanks):
    """bayes_io.get_data (one vectorised pass) vs a literal row-by-row parse of the reference's format
    (bayes_io.py:15-60): same curves, same values to the last bit, whatever the line endings."""
    from bayesian_inference_trpl_b200 import bayes_io
    rng = np.random.default_rng(seed)
    nl = "\r\n" if crlf else "\n"
    lines, expect = [], []
    for n in curves:
        t = np.concatenate([[0.0], np.cumsum(rng.uniform(0.01, 1.0, n - 1))]) if n > 1 else np.array([0.0])
        pl = 10.0 ** rng.uniform(10, 22, n)
        un = 10.0 ** rng.uniform(8, 15, n)
        rows = ["%.10G,%.9E,%G" % (a, b, c) for a, b, c in zip(t, pl, un)]
        expect.append(np.array([[float(x) for x in r.split(",")] for r in rows]))
        for r in rows:
            lines.append(r)
            if blanks and rng.uniform() < 0.1:
                lines.append("")
    path = tmp_path_factory.mktemp("obs") / "o.csv"
    path.write_text(nl.join(lines) + nl + "END" + nl + "1,2,3" + nl, newline="")
    e = bayes_io.get_data([str(path)], {"time_cutoff": None, "select_obs_sets": None, "noise_level": None},
                          {"log_pl": False, "self_normalize": False}, scale_f=1.0)[0]
    assert len(e[0]) == len(curves)
    for k, ref in enumerate(expect):
        np.testing.assert_array_equal(e[0][k], ref[:, 0])
        np.testing.assert_array_equal(e[1][k], ref[:, 1])
        np.testing.assert_array_equal(e[2][k], ref[:, 2])
