#!/usr/bin/env python3
"""Condense the reference's `Example Data/*.csv` (read from /root/reference, build container only)
into one small fixture, tests/golden/example_data.npz:

  power_exc   [3,128]  Power_scan_Excitations.csv rows, cm^-3 (file units)
  twothick_exc[6,128]  Twothick_Excitations.csv rows, cm^-3
  <set>_t<c>, <set>_pl<c>  every 25th point (+ the first 40) of curve c of the three shipped
                       *_Power_scan_Observations.csv files, file units (ns, cm^-2 s^-1)
  <set>_n     [3]      original number of points of each curve
"""
import os
import numpy as np

REF = "/root/reference/Example Data"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_obs(path):
    curves, cur = [], []
    for line in open(path):
        p = line.strip().split(",")
        if p[0] == "END":
            break
        t = float(p[0])
        if t == 0 and cur:
            curves.append(np.array(cur))
            cur = []
        cur.append((t, float(p[1]), float(p[2])))
    curves.append(np.array(cur))
    return curves


out = {}
out["power_exc"] = np.loadtxt(os.path.join(REF, "Power_scan_Excitations.csv"), delimiter=",",
                              usecols=range(128))
out["twothick_exc"] = np.loadtxt(os.path.join(REF, "Twothick_Excitations.csv"), delimiter=",",
                                 usecols=range(128))
for name in ("Highfrontsurf", "Highbacksurf", "Balancedhighsurf"):
    curves = read_obs(os.path.join(REF, name + "_Power_scan_Observations.csv"))
    out[name + "_n"] = np.array([len(c) for c in curves])
    for c, arr in enumerate(curves):
        keep = np.unique(np.concatenate([np.arange(40), np.arange(0, len(arr), 25), [len(arr) - 1]]))
        out["%s_t%d" % (name, c)] = arr[keep, 0]
        out["%s_pl%d" % (name, c)] = arr[keep, 1]
np.savez_compressed(os.path.join(HERE, "example_data.npz"), **out)
print({k: v.shape for k, v in out.items()})
