"""Drop-ins for the reference's `probs.prob` and `probs.fastlog` (probs.py:49-85): host numpy
buffers in, in-place results, elapsed seconds returned."""
import time

import numpy as np
import torch

from . import engine


def fastlog(plI, MIN, TPB=None, BPG=None, device=None):
    """In-place log10(max(plI, MIN)) with the reference's dtype semantics (float32 buffers clamp
    to float32(MIN) == 0 and use log10f; probs.py:64-75)."""
    clock0 = time.time()
    dev = engine.require_cuda(device)
    if plI.dtype not in (np.float32, np.float64):
        raise TypeError("plI must be float32 or float64")
    t = torch.from_numpy(np.ascontiguousarray(plI)).to(dev)
    engine.log10_clamp_(t, MIN)
    plI[:] = t.cpu().numpy()
    return time.time() - clock0


def prob(P, plI, values, uncertainty, mag_grid, TPB=None, BPG=None, device=None):
    """P[j] -= sum_i (plI[j,i] + mag_grid[j] - values[i])**2, in place (probs.py:20-62).
    `uncertainty` is accepted and unused, exactly like the reference kernel (probs.py:40)."""
    clock0 = time.time()
    dev = engine.require_cuda(device)
    pl = engine.to_device_f64(plI, dev)
    v = engine.to_device_f64(values, dev)
    m = engine.to_device_f64(mag_grid, dev)
    acc = torch.zeros(pl.shape[0], dtype=torch.float64, device=dev)
    engine.lnp_accumulate_(acc, pl, v, m)
    P += acc.cpu().numpy()
    return time.time() - clock0
