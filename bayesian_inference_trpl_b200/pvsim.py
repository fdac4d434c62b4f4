"""Drop-in for the reference's `pvSimPCR.pvSim` (pvSimPCR.py:309-401): same signature, same
in-place output buffer, same return value (solver seconds) -- the work is done by
trpl_solve_pl on the B200."""
import time

import numpy as np
import torch

from . import engine


def pvSim(plI_main, plN_main, plP_main, plE_main, matPar, simPar, iniPar, TPB, BPG,
          max_sims_per_block=1, init_mode="exp", device=None, status_out=None):
    """Simulate PL(t) for every row of `matPar` ([S,12], nm/ns/V units) and ONE excitation curve.

    plI_main [S, T//plT+1] (float32 or float64) is filled in place; plN/plP/plE, TPB, BPG and
    max_sims_per_block are accepted for signature compatibility and ignored (the reference
    ignores the first three too, pvSimPCR.py:385-387; launch geometry is the engine's business).
    `matPar` and `iniPar` are not modified.  Returns the solver wall time in seconds."""
    dev = engine.require_cuda(device)
    Length, Time, L, T, plT, pT, tol, MAX = simPar
    L, T, plT = int(L), int(T), int(plT)
    mat = engine.to_device_f64(np.asarray(matPar, dtype=np.float64)[:, :12], dev)
    if init_mode == "points":
        prof = np.asarray(iniPar, dtype=np.float64)
        grid_units = False
    elif init_mode == "exp":
        # pvSimPCR.py:347-353: profile evaluated in grid units at the cell centres
        dx = Length / L
        a, l = iniPar
        prof = (a * dx ** 3) * np.exp(-(np.arange(L) + 0.5) / (l / dx))
        grid_units = True
    else:
        raise ValueError("init_mode must be 'points' or 'exp' ('continue' is a stub in the "
                         "reference, pvSimPCR.py:357-358)")
    if prof.shape != (L,):
        raise ValueError("initial profile must have L=%d points" % L)
    init = engine.to_device_f64(prof, dev)
    if plI_main.dtype == np.float32:
        odt = torch.float32
    elif plI_main.dtype == np.float64:
        odt = torch.float64
    else:
        raise TypeError("plI_main must be float32 or float64")
    if plI_main.shape != (mat.shape[0], T // plT + 1):
        raise ValueError("plI_main must have shape (len(matPar), T//plT+1)")
    engine.host_wait(dev)
    clock0 = time.time()
    pl, status, _ = engine.solve_pl(mat, init, float(Length), float(Time), L, T, plT, int(tol),
                                    int(MAX), out_dtype=odt, init_grid_units=grid_units,
                                    want_iters=False)
    engine.host_wait(dev)
    solver_time = time.time() - clock0
    plI_main[:] = pl.cpu().numpy()
    if status_out is not None:
        status_out[:] = status.cpu().numpy()
    return solver_time
