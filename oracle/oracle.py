"""CPU ORACLE bindings -- test infrastructure, NOT the product.

ctypes front-end of oracle/libtrpl_oracle.so (oracle/trpl_oracle.c) plus a numpy
restatement of the host-side glue that sits between the reference's hot-path calls
(bayeslib.py:117-201: f32 PL buffer -> optional self-normalise -> log10 -> time
interpolation -> sum of squared log residuals, accumulated over curves).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtrpl_oracle.so")
_lib = None

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_float_p = ctypes.POINTER(ctypes.c_float)
_c_int64_p = ctypes.POINTER(ctypes.c_int64)
_c_int32_p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "trpl_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libtrpl_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.trpl_oracle_solve.restype = ctypes.c_int
        L.trpl_oracle_solve.argtypes = [
            _c_double_p, ctypes.c_int, _c_double_p, ctypes.c_double, ctypes.c_double,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, _c_double_p, _c_int64_p, _c_int32_p, _c_double_p,
            ctypes.c_int]
        L.trpl_oracle_scales.restype = None
        L.trpl_oracle_scales.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                         ctypes.c_int, _c_double_p, _c_double_p, _c_double_p]
        L.trpl_oracle_log10_clamp_f64.restype = None
        L.trpl_oracle_log10_clamp_f64.argtypes = [_c_double_p, ctypes.c_int64, ctypes.c_double]
        L.trpl_oracle_log10_clamp_f32.restype = None
        L.trpl_oracle_log10_clamp_f32.argtypes = [_c_float_p, ctypes.c_int64, ctypes.c_double]
        L.trpl_oracle_lnp.restype = None
        L.trpl_oracle_lnp.argtypes = [_c_double_p, _c_double_p, _c_double_p, _c_double_p,
                                      ctypes.c_int64, ctypes.c_int64]
        L.trpl_oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(_c_double_p)


def num_threads():
    return int(lib().trpl_oracle_num_threads())


def scales(length, time, L, T):
    """pvSimPCR.py:327-331 -> (scales[12], dx, dt)."""
    s = np.empty(12)
    dx = ctypes.c_double()
    dt = ctypes.c_double()
    lib().trpl_oracle_scales(float(length), float(time), int(L), int(T), _dp(s),
                             ctypes.byref(dx), ctypes.byref(dt))
    return s, dx.value, dt.value


def solve(matPar, simPar, iniPar, init_mode="points", solver="pcr", max_order=5, nthreads=0,
          return_state=False, raw=False, simulator_pow=False):
    """Oracle of pvSimPCR.pvSim (pvSimPCR.py:309-401) for ONE curve.

    matPar [S,12] physical units; simPar = [Length, Time, L, T, plT, pT, tol, MAX].
    Returns dict(pl [S,T//plT+1] f64, iters [S] i64, status [S] i32[, state [S,3,L+1]]).
    """
    Length, Time, L, T, plT, _pT, tol, MAX = simPar
    mp = np.ascontiguousarray(np.asarray(matPar, dtype=np.float64)[:, :12])
    S = mp.shape[0]
    flags = (2 if raw else 0) | (4 if simulator_pow else 0)
    if init_mode == "exp":
        # The reference scales a by dx^3 and evaluates the profile in grid units
        # (pvSimPCR.py:347-353); pass it through unscaled (flag bit0).
        dx = Length / L
        a, l = iniPar
        dn = (a * dx ** 3) * np.exp(-(np.arange(L) + 0.5) / (l / dx))
        flags |= 1
    else:
        dn = np.asarray(iniPar, dtype=np.float64)
    dn = np.ascontiguousarray(dn, dtype=np.float64)
    assert dn.shape == (L,), "initial profile must have L points"
    npl = T // plT + 1
    pl = np.empty((S, npl))
    iters = np.zeros(S, dtype=np.int64)
    status = np.zeros(S, dtype=np.int32)
    state = np.zeros((S, 3, L + 1)) if return_state else None
    rc = lib().trpl_oracle_solve(
        _dp(mp), S, _dp(dn), float(Length), float(Time), int(L), int(T), int(plT), int(tol),
        int(MAX), 0 if solver == "pcr" else 1, int(max_order), flags, _dp(pl),
        iters.ctypes.data_as(_c_int64_p), status.ctypes.data_as(_c_int32_p),
        _dp(state) if return_state else None, int(nthreads))
    if rc != 0:
        raise ValueError("trpl_oracle_solve failed with code %d" % rc)
    out = {"pl": pl, "iters": iters, "status": status}
    if return_state:
        out["state"] = state
    return out


def fastlog(plI, MIN):
    """Oracle of probs.fastlog (probs.py:64-85): in-place log10(max(x, MIN))."""
    if plI.dtype == np.float32:
        assert plI.flags.c_contiguous
        lib().trpl_oracle_log10_clamp_f32(plI.ctypes.data_as(_c_float_p), plI.size, float(MIN))
    elif plI.dtype == np.float64:
        assert plI.flags.c_contiguous
        lib().trpl_oracle_log10_clamp_f64(_dp(plI), plI.size, float(MIN))
    else:
        raise TypeError("plI must be float32 or float64")


def prob(P, plI, values, mag_grid):
    """Oracle of probs.prob (probs.py:20-62): P[j] -= sum_i (plI[j,i] + mag[j] - values[i])^2."""
    plI = np.ascontiguousarray(plI, dtype=np.float64)
    values = np.ascontiguousarray(values, dtype=np.float64)
    mag = np.ascontiguousarray(mag_grid, dtype=np.float64)
    acc = np.zeros(plI.shape[0])
    lib().trpl_oracle_lnp(_dp(acc), _dp(plI), _dp(values), _dp(mag), plI.shape[0], plI.shape[1])
    P += acc


def interp_linear(sim_times, y, times):
    """1-D linear interpolation with the index/weight rule of scipy.interpolate.interp1d
    (_call_linear of the installed SciPy 1.18; the reference pins no version), which is what
    scipy.interpolate.griddata runs for 1-D float32 rows (bayeslib.py:186-189): hi = searchsorted
    (left) clipped to [1, n-1], lo = hi-1, y = w_hi*y_hi + w_lo*y_lo.  Out-of-range -> NaN."""
    sim_times = np.asarray(sim_times, dtype=np.float64)
    times = np.asarray(times, dtype=np.float64)
    hi = np.clip(np.searchsorted(sim_times, times), 1, len(sim_times) - 1)
    lo = hi - 1
    x_lo, x_hi = sim_times[lo], sim_times[hi]
    y = np.asarray(y)
    y_lo, y_hi = y[..., lo], y[..., hi]
    out = ((times - x_lo) / (x_hi - x_lo)) * y_hi + ((x_hi - times) / (x_hi - x_lo)) * y_lo
    oob = (times < sim_times[0]) | (times > sim_times[-1])
    if np.any(oob):
        out = np.array(out, dtype=np.float64)
        out[..., oob] = np.nan
    return out


def loglik(X, simPar, iniPars, e_data, log_pl=True, self_normalize=False, emulate_f32=False,
           solver="pcr", nthreads=0, thicknesses=None):
    """Oracle of the whole per-sample pipeline of bayeslib.simulate (bayeslib.py:117-201).

    X [S,13] (last column mag_offset); iniPars [C,L]; e_data = [(t_list, logPL_list, unc_list)].
    emulate_f32=True reproduces the float32 PL buffer of bayeslib.py:137 (Q4 in SURVEY.md).
    Returns P [E,S].
    """
    X = np.asarray(X, dtype=np.float64)
    S = X.shape[0]
    Length, Time, L, T, plT, pT, tol, MAX = simPar
    C = len(iniPars)
    if thicknesses is None:
        thicknesses = list(Length) if isinstance(Length, (list, tuple)) else [Length] * C
    P = np.zeros((len(e_data), S))
    sim_times = np.linspace(0, Time, T + 1)
    MIN = sys.float_info.min
    for c in range(C):
        sp = [thicknesses[c], Time, L, T, plT, pT, tol, MAX]
        res = solve(X[:, :12], sp, iniPars[c], solver=solver, nthreads=nthreads, raw=emulate_f32)
        pl = res["pl"]
        if emulate_f32:
            # f64 kernel result rounded on store, then /= dx^2*dt in float32 (pvSimPCR.py:384,393)
            _, dx, dt = scales(thicknesses[c], Time, L, T)
            pl = pl.astype(np.float32)
            pl /= dx ** 2 * dt
        if self_normalize:
            pl = (pl.T / pl.T[0]).T
        if log_pl:
            pl = np.ascontiguousarray(pl)
            fastlog(pl, MIN)
        for e, exp in enumerate(e_data):
            times, values = exp[0][c], exp[1][c]
            pli = interp_linear(sim_times, pl, times)
            prob(P[e], pli, values, X[:, 12])
    return P
