"""ctypes binding of libtrpl_b200.so (C ABI declared in include/trpl_b200.h).

The shared library is built in-tree by `build()` (nvcc, sm_100a only).  There is no CPU
fallback: if the library is missing or no B200 is visible, calls raise.
"""
import ctypes
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("TRPL_LIB", os.path.join(_PKG, "libtrpl_b200.so"))   # TRPL_LIB: A/B-test builds
SRC = os.path.join(_PKG, "csrc", "trpl_kernels.cu")
INCLUDE = os.path.join(_ROOT, "include")

NPAR = 12
MAX_CURVES = 8
MAX_EXP = 4
F64, F32 = 0, 1
ST_NOCONV, ST_NONFINITE = 1, 2
F_INIT_GRID_UNITS, F_LOG_PL, F_SELF_NORMALIZE, F_EMULATE_F32 = 1, 2, 4, 8

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class TrplError(RuntimeError):
    pass


class Obs(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("hi_max", ctypes.c_int32),
                ("d_hi", ctypes.c_void_p), ("d_whi", ctypes.c_void_p),
                ("d_wlo", ctypes.c_void_p), ("d_val", ctypes.c_void_p)]


class Curve(ctypes.Structure):
    _fields_ = [("d_init", ctypes.c_void_p), ("length", ctypes.c_double),
                ("obs", Obs * MAX_EXP)]


def build(force=False, verbose=False):
    """Compile csrc/trpl_kernels.cu into libtrpl_b200.so for sm_100a (cross-compiles without a GPU)."""
    csrc = os.path.dirname(SRC)
    deps = [os.path.join(INCLUDE, "trpl_b200.h")] + [os.path.join(csrc, f) for f in os.listdir(csrc)
                                                     if f.endswith((".cu", ".cuh"))]
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(d) for d in deps)):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-o", LIB_PATH, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return LIB_PATH


_lib = None


def lib():
    """Load the library (never builds implicitly on a GPU box: the .so ships with the snapshot)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TrplError("libtrpl_b200.so is not built: run `python -c 'import __graft_entry__ as g; "
                        "g.build()'` (needs nvcc). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    L.trpl_version.restype = i32
    L.trpl_error_string.restype = ctypes.c_char_p
    L.trpl_error_string.argtypes = [i32]
    L.trpl_last_cuda_error.restype = ctypes.c_char_p
    L.trpl_resident_sims.restype = i32
    L.trpl_resident_sims.argtypes = [i32, i32]
    L.trpl_solve_pl.restype = i32
    L.trpl_solve_pl.argtypes = [vp, i64, i64, vp, dbl, dbl, i32, i32, i32, i32, i32, i32, i32,
                                vp, i32, i64, vp, vp, i32, vp]
    L.trpl_solve_loglik.restype = i32
    L.trpl_solve_loglik.argtypes = [vp, i64, i64, i32, ctypes.POINTER(Curve), i32, i32, dbl, i32,
                                    i32, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp]
    L.trpl_log10_clamp.restype = i32
    L.trpl_log10_clamp.argtypes = [vp, i32, i64, dbl, i32, vp]
    L.trpl_lnp_accumulate.restype = i32
    L.trpl_lnp_accumulate.argtypes = [vp, vp, i64, i64, i64, vp, vp, i32, vp]
    L.trpl_obs_prepare.restype = i32
    L.trpl_obs_prepare.argtypes = [vp, ctypes.c_int32, dbl, i32, vp, vp, vp]
    L.trpl_lse_partial.restype = i32
    L.trpl_lse_partial.argtypes = [vp, i64, vp, i32, vp]
    u64 = ctypes.c_uint64
    L.trpl_random_grid.restype = i32
    L.trpl_random_grid.argtypes = [vp, i64, i64, vp, vp, vp, i32, i32, u64, u64, i32, vp]
    L.trpl_posterior_weights.restype = i32
    L.trpl_posterior_weights.argtypes = [vp, i64, dbl, vp, i32, vp]
    L.trpl_weighted_hist.restype = i32
    L.trpl_weighted_hist.argtypes = [vp, i64, i64, i32, i32, vp, dbl, dbl, i32, dbl, dbl, i32, vp, i32, vp]
    L.trpl_weighted_moments.restype = i32
    L.trpl_weighted_moments.argtypes = [vp, i64, i64, i32, vp, vp, i32, vp]
    L.trpl_selftest_rcp.restype = i32
    L.trpl_selftest_rcp.argtypes = [vp, vp, i64, i32, vp]
    L.trpl_bench_dfma.restype = i32
    L.trpl_bench_dfma.argtypes = [i32, i32, ctypes.POINTER(dbl), ctypes.POINTER(dbl)]
    _lib = L
    return L


EXPORTS = ["trpl_version", "trpl_error_string", "trpl_last_cuda_error", "trpl_resident_sims",
           "trpl_solve_pl", "trpl_solve_loglik", "trpl_log10_clamp", "trpl_lnp_accumulate",
           "trpl_obs_prepare", "trpl_lse_partial", "trpl_bench_dfma", "trpl_random_grid",
           "trpl_posterior_weights", "trpl_weighted_hist", "trpl_weighted_moments", "trpl_selftest_rcp"]


def check(rc, what):
    if rc < 0:
        L = lib()
        msg = L.trpl_error_string(rc).decode()
        if rc == -3:
            msg += " [" + L.trpl_last_cuda_error().decode() + "]"
        raise TrplError("%s failed: %s (code %d)" % (what, msg, rc))
    return rc
