#!/usr/bin/env python3
"""Golden for the file readers: a small synthetic observation CSV and excitation CSV (written
here, committed) parsed by the REFERENCE's bayes_io.get_data / get_initpoints (read from
/root/reference, build container only) under several flag combinations."""
import os, sys
import numpy as np
sys.path.insert(0, "/root/reference")
import bayes_io as ref_io                       # noqa: E402  (the reference module)
HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(0)
obs = os.path.join(HERE, "io_obs.csv")
with open(obs, "w") as fh:
    for c, n in enumerate((9, 6, 12)):
        t = np.arange(n) * 0.025
        pl = 10 ** (17 + c - 0.3 * t * 40) * (1 + 0.01 * rng.normal(size=n))
        if c == 1:
            pl[3] = -pl[3]                       # a noisy negative value (|.| is taken, bayes_io.py:72)
        for ti, pi in zip(t, pl):
            fh.write("%.10G,%.9E,%G\n" % (ti, pi, 1e14))
    fh.write("END\n")
exc = os.path.join(HERE, "io_exc.csv")
with open(exc, "w") as fh:
    for c in range(3):
        fh.write(",".join("%.8E" % v for v in 10 ** (16 + c) * np.exp(-np.arange(8) / 3.0)) + "\n")
out = {}
cases = {"log": ({"time_cutoff": None, "select_obs_sets": None, "noise_level": None}, {"log_pl": True, "self_normalize": False}),
         "cut_sel": ({"time_cutoff": 0.126, "select_obs_sets": [0, 2], "noise_level": None}, {"log_pl": True, "self_normalize": False}),
         "norm_lin": ({"time_cutoff": None, "select_obs_sets": None, "noise_level": None}, {"log_pl": False, "self_normalize": True})}
for name, (ic, sf) in cases.items():
    e = ref_io.get_data([obs], ic, sf, scale_f=1e-23)[0]
    for k, part in enumerate(("t", "pl", "unc")):
        for c in range(len(e[0])):
            out["%s_%s%d" % (name, part, c)] = np.asarray(e[k][c])
    out[name + "_n"] = len(e[0])
    out[name + "_ini"] = ref_io.get_initpoints(exc, ic)
np.savez_compressed(os.path.join(HERE, "io_golden.npz"), **out)
print(sorted(out)[:8], len(out))
