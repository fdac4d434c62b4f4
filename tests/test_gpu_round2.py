"""Round-2 host-side behaviour on the GPU: the Philox sampler on the product path, the staged-problem
cache behind bayeslib.simulate, argument guards, the documented command line."""
import os

import numpy as np
import pytest

from helpers import TRUTH, UC, example_data, power_scan_excitations, prior_samples

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def trpl():
    import bayesian_inference_trpl_b200 as t
    return t


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as o
    return o


def _files(trpl, oracle, tmp_path, T=240):
    L = 128
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
    return exc_path, obs_path, inis, simPar


def test_philox_sampler_on_the_product_path(trpl, oracle, tmp_path):
    """sim_flags["sampler"]="philox": X is drawn on the device (engine units), never uploaded, and the
    exported table equals the oracle's likelihood of exactly those rows (bayeslib.py:18-76 semantics)."""
    exc_path, obs_path, inis, simPar = _files(trpl, oracle, tmp_path)
    cfg = trpl.parallel_bayes_gpu.default_config()
    cfg.update(Length=2000.0, Time=simPar[1], T=simPar[3])
    cfg["ic_flags"]["time_cutoff"] = simPar[1]
    cfg["sim_flags"].update(num_points=21, sampler="philox", override_equal_s=True)
    cfg["minX"][2:4] = 0.5
    out = str(tmp_path / "PHX")
    P, X = trpl.parallel_bayes_gpu.run(exc_path, [obs_path], [out], cfg=cfg, seed=7)
    uc = trpl.parallel_bayes_gpu.unit_conversions
    Xd = trpl.engine.random_grid_device(cfg["minX"] * uc, cfg["maxX"] * uc, cfg["do_log"], 21, 7, override_flags=2)
    np.testing.assert_array_equal(X * uc, (Xd.cpu().numpy() / uc) * uc)
    assert (X[:, 6] == X[:, 5]).all() and (X[:, 1] >= 1e14).all() and (X[:, 1] <= 1e16).all()
    e_data = trpl.bayes_io.get_data([obs_path], cfg["ic_flags"], cfg["sim_flags"])
    ref = oracle.loglik(Xd.cpu().numpy(), simPar, inis, e_data, solver="pcr")
    np.testing.assert_allclose(P[0], ref[0], rtol=1e-6, atol=1e-9)
    # shards of the draw concatenate to the single-rank draw bit for bit
    parts = [trpl.engine.random_grid_device(cfg["minX"] * uc, cfg["maxX"] * uc, cfg["do_log"], hi - lo, 7,
                                            first_sample=lo, override_flags=2).cpu().numpy()
             for lo, hi in (trpl.distributed.shard_bounds(21, r, 4) for r in range(4))]
    np.testing.assert_array_equal(np.concatenate(parts), Xd.cpu().numpy())


def test_command_line_main(trpl, oracle, tmp_path):
    """README.md:29 documents `OBS EXC OUT` arguments the reference never parses; main() does."""
    exc_path, obs_path, inis, simPar = _files(trpl, oracle, tmp_path)
    out = str(tmp_path / "CLI")
    trpl.parallel_bayes_gpu.main([exc_path, out, obs_path, "--num-points", "9", "--length", "2000",
                                  "--time-steps", str(simPar[3]), "--final-time", str(simPar[1]),
                                  "--sims-per-gpu", "4", "--posterior"])
    P = np.load(os.path.join(out, "CLI_BAYRAN_P.npy"))
    X = np.load(os.path.join(out, "CLI_BAYRAN_X.npy"))
    W = np.load(os.path.join(out, "CLI_BAYRAN_W.npy"))
    assert P.shape == (9,) and X.shape == (9, 13) and np.isfinite(P).all()
    np.testing.assert_allclose(W.sum(), 1.0, rtol=1e-12)
    cfg = trpl.parallel_bayes_gpu.default_config()
    cfg["ic_flags"]["time_cutoff"] = simPar[1]
    e_data = trpl.bayes_io.get_data([obs_path], cfg["ic_flags"], cfg["sim_flags"])
    ref = oracle.loglik(X * trpl.parallel_bayes_gpu.unit_conversions, simPar, inis, e_data, solver="thomas")
    np.testing.assert_allclose(P, ref[0], rtol=1e-6, atol=1e-9)


def test_simulate_reuses_the_staged_problem(trpl):
    T, L = 64, 128
    simPar = [2000.0, 0.025 * T, L, T, 1, (0,), 7, 10000]
    inis = power_scan_excitations()
    grid = np.linspace(0, simPar[1], T + 1)
    e_data = [([grid.copy()] * 3, [np.linspace(-6.5, -7.0, T + 1)] * 3, [np.full(T + 1, .1)] * 3)]
    p1 = trpl.engine.cached_problem(simPar, inis, e_data, device=0)
    p2 = trpl.engine.cached_problem(list(simPar), inis.copy(), [tuple(list(a) for a in e_data[0])], device=0)
    assert p1 is p2                                           # same content -> same staged arrays
    e2 = [([grid.copy()] * 3, [np.linspace(-6.5, -7.1, T + 1)] * 3, [np.full(T + 1, .1)] * 3)]
    assert trpl.engine.cached_problem(simPar, inis, e2, device=0) is not p1
    # and simulate() gives identical tables on repeated calls with host arrays
    X = prior_samples(10, seed=3, mag=True)
    flags = {"load_PL_from_file": False, "log_pl": True, "self_normalize": False}
    info = {"has_GPU": True, "sims_per_gpu": 10, "num_gpus": 1, "threads_per_block": (128,), "max_sims_per_block": 1}
    tabs = []
    for _ in range(2):
        P = np.zeros((1, 10))
        tm = [np.zeros(1), np.zeros(1), np.zeros(1)]
        trpl.bayeslib.simulate(trpl.pvSim, e_data, P, X, [None], [None], 3, list(simPar), inis, flags, info, 0, *tm)
        tabs.append(P.copy())
    np.testing.assert_array_equal(tabs[0], tabs[1])
    assert np.isfinite(tabs[0]).all()


def test_fused_path_rejects_plT_other_than_one(trpl):
    """the fused call has no plT argument; a caller asking for plT != 1 must not silently get
    plT = 1 numbers (the staged path refuses the buffer shape, bayeslib.py:137)."""
    T, L = 64, 128
    inis = power_scan_excitations()
    grid = np.linspace(0, 0.025 * T, T + 1)
    e_data = [([grid] * 3, [np.zeros(T + 1)] * 3, [np.ones(T + 1)] * 3)]
    with pytest.raises(ValueError):
        trpl.engine.Problem([2000.0, 0.025 * T, L, T, 2, (0,), 7, 10000], inis, e_data, device=0)


def test_on_grid_observation_next_to_nonpositive_pl_gives_minus_inf_not_nan(trpl):
    """an observation exactly on a step-grid point has weight 0 on the other neighbour; if that
    neighbour is log10(0) = -inf (float32 clamp, probs.py:72-75) the product 0 * -inf must not poison lnL."""
    T, L = 96, 128
    simPar = [2000.0, 0.025 * T, L, T, 1, (0,), 7, 10000]
    inis = power_scan_excitations()[:1]
    grid = np.linspace(0, simPar[1], T + 1)
    X = prior_samples(4, seed=5, mag=True)
    X[:, 0] = 1e-2                                   # huge n0: PL - equilibrium cancels to <= 0 within a few steps
    e_data = [([grid[:40].copy()], [np.full(40, -7.0)], [np.full(40, .1)])]
    prob = trpl.engine.Problem(simPar, inis, e_data, device=0)
    lnl, st, _ = trpl.engine.solve_loglik(torch.from_numpy(X).cuda(), prob, emulate_f32=True)
    got = lnl.cpu().numpy()[0]
    assert not np.isnan(got[st.cpu().numpy() == 0]).any()


def test_calls_leave_the_callers_current_device_alone(trpl):
    """a C-ABI call on device k must not change the runtime's current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    torch.cuda.set_device(0)
    x = torch.zeros(1000, dtype=torch.float64, device="cuda:1")
    trpl.engine.lse_partial(x)
    assert torch.cuda.current_device() == 0
    assert trpl.engine.resident_sims(128, 1) > 0
    assert torch.cuda.current_device() == 0
    y = torch.ones(8, device="cuda")
    assert y.device.index == 0


def test_export_straight_from_device_buffers(trpl, tmp_path):
    """BAYRAN_P / BAYRAN_X written from CUDA tensors through a pinned staging chunk;
    np.load (what the reference's GUI does, marginalization_visual.py) reads them back bit for bit."""
    X = trpl.engine.random_grid_device(np.full(13, 1.0), np.full(13, 2.0), np.zeros(13, dtype=np.int32), 70001, 5)
    P = -X[:, 0].contiguous() * 3.0
    out = str(tmp_path / "DEV")
    trpl.bayes_io.export_from_device(out, P, X, chunk_rows=4096)
    np.testing.assert_array_equal(np.load(os.path.join(out, "DEV_BAYRAN_P.npy")), P.cpu().numpy())
    np.testing.assert_array_equal(np.load(os.path.join(out, "DEV_BAYRAN_X.npy")), X.cpu().numpy())


def test_debug_build_reports_itself(trpl):
    """The library built with -DTRPL_DEBUG=1 (device-side bounds checks + ring canaries) has an odd version."""
    v = trpl._lib.lib().trpl_version()
    assert v >= 200
    if "debug" in os.path.basename(trpl._lib.LIB_PATH):
        assert v % 2 == 1


def test_solver_reciprocal_is_one_ulp_over_its_domain_and_loud_outside(trpl):
    """rcp64 has no slow path.  Inside the solver's domain (finite, normal, far from the
    exponent limits) it must be within 1 ulp of 1/x; outside (0, denormal, inf, NaN) the result must be
    non-finite so that the solver's per-sample non-finite status catches it."""
    import ctypes
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(1, 2, 200000) * 2.0 ** rng.integers(-1000, 1000, 200000),
                        -rng.uniform(1, 2, 1000) * 2.0 ** rng.integers(-1000, 1000, 1000),
                        [1.0, 2.0, 0.5, 3.0, 1e300, 1e-300, np.nextafter(1.0, 2.0), np.nextafter(2.0, 1.0)]])
    bad = np.array([0.0, -0.0, 5e-324, 1e-310, np.inf, -np.inf, np.nan])
    xd = torch.from_numpy(np.concatenate([x, bad])).cuda()
    yd = torch.empty_like(xd)
    rc = trpl._lib.lib().trpl_selftest_rcp(ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(yd.data_ptr()),
                                           xd.numel(), 0, None)
    assert rc == 0
    y = yd.cpu().numpy()
    good, weird = y[:len(x)], y[len(x):]
    exact = 1.0 / x
    ulp = np.abs(good - exact) / np.spacing(np.abs(exact))
    assert ulp.max() <= 1.0, ulp.max()
    assert (good[-8:-4] == [1.0, 0.5, 2.0, 1.0 / 3.0]).all()
    assert not np.isfinite(weird[[0, 1, 2, 3, 6]]).all() and not np.isfinite(weird[[0, 1, 2, 3, 6]]).any()
    assert (weird[4:6] == 0.0).all() or not np.isfinite(weird[4:6]).any()      # 1/inf = 0 is fine too


def test_slurm_array_tasks_export_their_own_columns_and_merge(trpl, oracle, tmp_path, monkeypatch):
    """The reference's launcher: independent SLURM array tasks (bayeslib.py:231).  Every task exports, foreign
    columns are NaN (not 0 = best likelihood), and the merged table equals the single-task run."""
    exc_path, obs_path, inis, simPar = _files(trpl, oracle, tmp_path, T=160)
    entry = trpl.parallel_bayes_gpu

    def cfg_for(num_gpus):
        cfg = entry.default_config()
        cfg.update(Length=2000.0, Time=simPar[1], T=simPar[3])
        cfg["ic_flags"]["time_cutoff"] = simPar[1]
        cfg["sim_flags"]["num_points"] = 23
        cfg["gpu_info"].update(sims_per_gpu=4, num_gpus=num_gpus)
        cfg["minX"][2:4] = 0.5
        return cfg

    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("SLURM_ARRAY_TASK_ID", raising=False)
    P1, X1 = entry.run(exc_path, [obs_path], [str(tmp_path / "ONE")], cfg=cfg_for(1))
    out = str(tmp_path / "ARR")
    monkeypatch.setenv("SLURM_ARRAY_TASK_COUNT", "3")
    for k in range(3):
        monkeypatch.setenv("SLURM_ARRAY_TASK_ID", str(k))
        Pk, Xk = entry.run(exc_path, [obs_path], [out], cfg=cfg_for(3))
        mine = trpl.bayeslib.owned_columns(23, {"sims_per_gpu": 4, "num_gpus": 3}, k)
        assert np.isnan(Pk[0][~mine]).all() and np.isfinite(Pk[0][mine]).all()
        np.testing.assert_array_equal(Xk, X1)
        assert os.path.exists(os.path.join("%s_task%d" % (out, k), "ARR_task%d_BAYRAN_P.npy" % k))
    Pm, Xm = entry.merge_task_exports(out, 3)
    np.testing.assert_allclose(Pm, P1[0], rtol=1e-12)
    # a task count that disagrees with the configuration must not silently leave columns uncomputed
    monkeypatch.setenv("SLURM_ARRAY_TASK_COUNT", "2")
    monkeypatch.setenv("SLURM_ARRAY_TASK_ID", "0")
    with pytest.raises(RuntimeError):
        entry.run(exc_path, [obs_path], [out], cfg=cfg_for(3))
