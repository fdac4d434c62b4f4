#!/usr/bin/env python3
"""BASELINE config 2 at its NOMINAL size through the reference-style entry point: 2^20 random
parameter samples, 3-curve power scan, L=128, T=80000, on all ranks of a torchrun launch.
Every rank writes the (synthetic) excitation / observation CSVs in the reference formats, calls
bayesian_inference_trpl_b200.parallel_bayes_gpu.run(), rank 0 reports throughput and a posterior
sanity check.   torchrun --nproc-per-node 8 tools/nominal_power_scan.py [log2 num_points] [numpy|philox]"""
import json, os, sys, tempfile, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bayesian_inference_trpl_b200 as trpl
from bayesian_inference_trpl_b200 import parallel_bayes_gpu as entry
from helpers import TRUTH, UC, example_data

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
n_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sampler = sys.argv[2] if len(sys.argv) > 2 else "numpy"        # "philox": every rank draws only its own rows on its GPU
L, T, Time, Length = 128, 80000, 2000.0, 2000.0
tmp = tempfile.mkdtemp(prefix="trpl_nominal_r%d_" % rank)
exc = os.path.join(tmp, "Power_scan_Excitations.csv")
with open(exc, "w") as fh:
    for row in example_data()["power_exc"]:
        fh.write(",".join("%.8E" % v for v in row) + "\n")
inis = trpl.bayes_io.get_initpoints(exc, {"select_obs_sets": None})
simPar = [Length, Time, L, T, 1, (0,), 7, 10000]
grid = np.linspace(0, Time, T + 1)
pls = []
for c in range(3):
    pl = np.empty((1, T + 1))
    trpl.pvSim(pl, None, None, None, (TRUTH * UC)[None, :12], simPar, inis[c], (128,), 0, 1, init_mode="points")
    pls.append(pl[0])
obs = os.path.join(tmp, "Power_scan_Observations.csv")
trpl.bayes_io.write_observations(obs, [grid] * 3, pls)
cfg = entry.default_config()
cfg.update(Length=Length, Time=Time, T=T)
cfg["sim_flags"]["num_points"] = 2 ** n_log2
cfg["sim_flags"]["sampler"] = sampler
if sampler == "philox":
    cfg["minX"][2:4] = 0.5          # mobilities from 0.5 instead of 0 (the numpy path of this tool keeps the entry script's 0)
out = os.path.join(tmp, "NOMINAL")
t0 = time.perf_counter()
P, X = entry.run(exc, [obs], [out], cfg=cfg, posterior=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if rank == 0:
    S = P.shape[1]
    lnp = P[0]
    best = int(np.nanargmax(lnp))
    w = np.exp(lnp - np.nanmax(lnp)); w[~np.isfinite(w)] = 0; w /= w.sum()
    free = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10]
    names = entry.param_names
    rep = {"num_points": S, "n_gpus": world, "sampler": sampler, "wall_s": dt, "likelihoods_per_s_incl_io": S / dt,
           "nonfinite_lnL": int((~np.isfinite(lnp)).sum()), "best_lnL": float(lnp[best]),
           "effective_sample_size": float(1.0 / np.sum(w ** 2)),
           "best_sample": {names[j]: float(X[best, j]) for j in free},
           "truth": {names[j]: float(TRUTH[j]) for j in free}}
    print(json.dumps(rep, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "nominal_power_scan.json"), "w"), indent=1)
    assert os.path.exists(os.path.join(out, "NOMINAL_BAYRAN_P.npy"))
import torch.distributed as dist
if dist.is_initialized():
    dist.destroy_process_group()
