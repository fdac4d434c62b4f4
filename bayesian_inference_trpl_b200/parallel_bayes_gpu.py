#!/usr/bin/env python3
"""Entry point mirroring the reference's `parallel_bayes_gpu.py` (config block :72-131, unit
conversion :27-33,183-194, bayes call :189, export :197-198) with the documented-but-never-
implemented command line (`README.md:4,29`: OBS EXC OUT) actually parsed.

    python -m bayesian_inference_trpl_b200.parallel_bayes_gpu EXC.csv OUT_DIR OBS.csv [OBS2.csv ...]
    torchrun --nproc-per-node 8 -m bayesian_inference_trpl_b200.parallel_bayes_gpu ...

Every rank draws the same sample matrix (global seed 42, parallel_bayes_gpu.py:35), evaluates
its block-cyclic share on its GPU with the fused kernel, the likelihood tables are merged over
NCCL and rank 0 writes <OUT>_BAYRAN_P.npy / _BAYRAN_X.npy (plus the normalised posterior
weights when --posterior is given)."""
import argparse
import logging
import os
import sys
from time import perf_counter

import numpy as np

from . import bayes_io, bayes_validate, bayeslib, distributed
from .pvsim import pvSim

lambda0 = 704.3                           # q^2/(eps0*k_B T=25C) [nm]
param_names = ["n0", "p0", "mun", "mup", "B", "Sf", "Sb", "CN", "CP", "taun", "taup", "lambda",
               "mag_offset"]
# common units -> [V, nm, ns]
unit_conversions = np.array([(1e7) ** -3, (1e7) ** -3,
                             (1e7) ** 2 / (1e9) * .02569257, (1e7) ** 2 / (1e9) * .02569257,
                             (1e7) ** 3 / (1e9), (1e7) / (1e9), (1e7) / (1e9),
                             (1e7) ** 6 / (1e9), (1e7) ** 6 / (1e9), 1, 1, lambda0, 1])


def default_config():
    """The reference's hard-coded configuration (parallel_bayes_gpu.py:72-124)."""
    return {
        "Length": 311, "L": 2 ** 7, "Time": 2000, "T": 80000, "plT": 1, "pT": (0, 1, 3, 10, 30, 100),
        "tol": 7, "MAX": 10000,
        "do_log": np.array([1, 1, 0, 0, 1, 1, 1, 1, 1, 0, 0, 1, 0]),
        "minX": np.array([1e8, 1e14, 0, 0, 1e-11, 0.1, 0.1, 1e-30, 1e-30, 1, 1, 10 ** -1, 0]),
        "maxX": np.array([1e8, 1e16, 50, 50, 1e-9, 100, 100, 1e-28, 1e-28, 1000, 2000, 10 ** -1, 0]),
        "ic_flags": {"time_cutoff": 2000, "select_obs_sets": None, "noise_level": None},
        "gpu_info": {"sims_per_gpu": 2 ** 10, "num_gpus": 1},
        "sim_flags": {"load_PL_from_file": False, "override_equal_auger": False,
                      "override_equal_mu": False, "override_equal_s": False, "log_pl": True,
                      "self_normalize": False, "random_sample": True, "num_points": 2 ** 17},
    }


def run(init_filename, experimental_data_filename, out_filename, cfg=None, seed=42, logger=None,
        posterior=False):
    """Same sequence as the reference `__main__` block; returns (P, X in common units).

    Launchers: (a) torchrun, one rank per GPU: the per-rank tables are merged over NCCL and rank 0
    exports; (b) SLURM array tasks like the reference (no communication): task k fills its
    block-cyclic columns and EVERY task exports, into `<out>_task<k>` when there is more than one
    task, with NaN in the columns it does not own (`merge_task_exports` joins them afterwards);
    the task count must agree with gpu_info["num_gpus"]."""
    import torch
    import torch.distributed as dist
    cfg = cfg or default_config()
    rank, world = bayeslib.rank_and_world()
    slurm = os.getenv("SLURM_ARRAY_TASK_ID") is not None
    if not slurm and world is not None and world > 1 and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl")
    np.random.seed(seed)
    simPar = [cfg["Length"], cfg["Time"], cfg["L"], cfg["T"], cfg["plT"], cfg["pT"], cfg["tol"], cfg["MAX"]]
    ic_flags, sim_flags, gpu_info = cfg["ic_flags"], dict(cfg["sim_flags"]), dict(cfg["gpu_info"])
    sim_flags.setdefault("seed", seed)
    minX, maxX, do_log = cfg["minX"].astype(float).copy(), cfg["maxX"].astype(float).copy(), cfg["do_log"]

    iniPar = bayes_io.get_initpoints(init_filename, ic_flags)
    e_data = bayes_io.get_data(experimental_data_filename, ic_flags, sim_flags, logger=logger,
                               scale_f=1e-23)
    for exp in e_data:
        assert len(iniPar) == len(exp[0]), "Num. ICs mismatch num. datasets"
    bayes_validate.validate_ic_flags(ic_flags)
    bayes_validate.validate_IC(iniPar, cfg["L"])
    bayes_validate.validate_gpu_info(gpu_info)
    bayes_validate.validate_params(len(param_names), unit_conversions, do_log, minX, maxX)
    bayes_validate.connect_to_gpu(gpu_info, nthreads=128, sims_per_block=1)
    if not gpu_info["has_GPU"]:
        raise RuntimeError("no GPU: the B200 engine has no CPU fallback")
    if slurm:
        if world is not None and world != gpu_info["num_gpus"]:
            raise RuntimeError("SLURM_ARRAY_TASK_COUNT=%d disagrees with gpu_info['num_gpus']=%d: the tasks "
                               "would leave columns of P uncomputed" % (world, gpu_info["num_gpus"]))
        if rank >= gpu_info["num_gpus"]:
            raise RuntimeError("SLURM_ARRAY_TASK_ID=%d but gpu_info['num_gpus']=%d" % (rank, gpu_info["num_gpus"]))
        if sim_flags.get("sampler", "numpy") == "philox" and gpu_info["num_gpus"] > 1:
            raise RuntimeError("the philox sampler shards over torch.distributed ranks; use torchrun")
    elif world is not None:
        gpu_info["num_gpus"] = world

    minX *= unit_conversions
    maxX *= unit_conversions
    clock0 = perf_counter()
    N, P, X = bayeslib.bayes(pvSim, np.array([0]), None, minX, maxX, do_log, iniPar, simPar, e_data,
                             sim_flags, gpu_info, logger=logger)
    complete = sim_flags.get("sampler", "numpy") == "philox"
    if not complete and not slurm and world and world > 1:
        P = np.asarray(distributed.merge_block_cyclic(P).cpu())
    if logger is not None:
        logger.info("Bayesim took %.3f s", perf_counter() - clock0)
    X = X / unit_conversions
    outs = list(out_filename)
    if slurm and gpu_info["num_gpus"] > 1:
        mine = bayeslib.owned_columns(P.shape[1], gpu_info, rank)
        P = np.where(mine[None, :], P, np.nan)             # 0 would read as the best possible likelihood
        outs = ["%s_task%d" % (of.rstrip("/\\"), rank) for of in outs]
    if slurm or rank == 0:
        for i, of in enumerate(outs):
            bayes_io.export(of, P[i], X, logger=logger)
            if posterior:
                t = torch.from_numpy(np.nan_to_num(P[i], nan=-np.inf)).cuda()
                from . import engine
                ms = engine.lse_partial(t)
                lse = ms[0] + torch.log(ms[1])
                w = distributed.normalize_posterior(t, lse).cpu().numpy()
                np.save(os.path.join(of, os.path.basename(os.path.normpath(of)) + "_BAYRAN_W.npy"), w)
    return P, X


def merge_task_exports(out_filename, num_tasks):
    """Join the `<out>_task<k>` exports of a SLURM array run into `<out>`: every column is taken from
    the task that owns it (the others hold NaN there); a column nobody computed stays NaN."""
    base = os.path.basename(os.path.normpath(out_filename))
    P, X = None, None
    for k in range(num_tasks):
        d = "%s_task%d" % (out_filename.rstrip("/\\"), k)
        b = os.path.basename(os.path.normpath(d))
        Pk = np.load(os.path.join(d, b + "_BAYRAN_P.npy"))
        X = np.load(os.path.join(d, b + "_BAYRAN_X.npy"))
        P = Pk.copy() if P is None else np.where(np.isnan(P), Pk, P)
    bayes_io.export(out_filename, P, X)
    return P, X


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("excitations")
    ap.add_argument("out")
    ap.add_argument("observations", nargs="+")
    ap.add_argument("--num-points", type=int, default=None)
    ap.add_argument("--length", type=float, nargs="+", default=None, help="film thickness(es) in nm")
    ap.add_argument("--time-steps", type=int, default=None)
    ap.add_argument("--final-time", type=float, default=None)
    ap.add_argument("--posterior", action="store_true", help="also write normalised posterior weights")
    ap.add_argument("--sampler", choices=["numpy", "philox"], default="numpy",
                    help="numpy: the reference's host draw (np.random.seed(42), bit-compatible); philox: every rank "
                         "draws only its own rows on its GPU")
    ap.add_argument("--sims-per-gpu", type=int, default=None)
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)s: %(message)s")
    logger = logging.getLogger("Bayes Logger Main")
    cfg = default_config()
    if args.num_points:
        cfg["sim_flags"]["num_points"] = args.num_points
    cfg["sim_flags"]["sampler"] = args.sampler
    if args.sims_per_gpu:
        cfg["gpu_info"]["sims_per_gpu"] = args.sims_per_gpu
    if args.length:
        cfg["Length"] = args.length[0] if len(args.length) == 1 else list(args.length)
    if args.time_steps:
        cfg["T"] = args.time_steps
    if args.final_time:
        cfg["Time"] = args.final_time
        cfg["ic_flags"]["time_cutoff"] = args.final_time
    outs = [args.out] if len(args.observations) == 1 else \
        [args.out + "_%d" % i for i in range(len(args.observations))]
    run(args.excitations, args.observations, outs, cfg=cfg, logger=logger, posterior=args.posterior)


if __name__ == "__main__":
    main(sys.argv[1:])
