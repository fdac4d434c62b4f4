"""Configuration checks and device attach for the B200 engine.

Same entry points and failure behaviour (AssertionError) as the reference's
`bayes_validate.py:10-55`, so its entry script keeps working; `connect_to_gpu` fills the same
`gpu_info` keys (`has_GPU`, `threads_per_block`, `max_sims_per_block`) but asks torch/CUDA
instead of numba and also records the device name."""
import numbers

import torch


def _require(ok, message):
    if not ok:
        raise AssertionError(message)


def _is_number(x):
    return isinstance(x, numbers.Real) and not isinstance(x, bool)


def validate_IC(ics, L):
    """Every excitation profile must have one value per grid node."""
    for k, profile in enumerate(ics):
        _require(len(profile) == L,
                 "excitation curve {} has {} points but L = {}".format(k, len(profile), L))


def validate_ic_flags(ic_flags):
    cutoff = ic_flags["time_cutoff"]
    _require(cutoff is None or (_is_number(cutoff) and cutoff > 0),
             "time_cutoff must be None or a positive number")
    chosen = ic_flags["select_obs_sets"]
    _require(chosen is None or isinstance(chosen, list),
             "select_obs_sets must be None or a list of curve indices")
    noise = ic_flags["noise_level"]
    _require(noise is None or _is_number(noise), "noise_level must be None or a number")


def validate_gpu_info(gpu_info):
    for key in ("num_gpus", "sims_per_gpu"):
        value = gpu_info[key]
        _require(isinstance(value, int) and not isinstance(value, bool) and value > 0,
                 "{} must be a positive integer".format(key))


def validate_params(num_params, unit_conversions, do_log, minX, maxX):
    named = {"unit_conversions": unit_conversions, "do_log": do_log, "minX": minX, "maxX": maxX}
    for name, arr in named.items():
        _require(len(arr) == num_params,
                 "{} has {} entries, expected {}".format(name, len(arr), num_params))
    _require(all(lo <= hi for lo, hi in zip(minX, maxX)), "minX exceeds maxX for some parameter")


def connect_to_gpu(gpu_info, nthreads=128, sims_per_block=1):
    gpu_info["has_GPU"] = bool(torch.cuda.is_available())
    if not gpu_info["has_GPU"]:
        return
    gpu_info["threads_per_block"] = (nthreads,)
    gpu_info["max_sims_per_block"] = sims_per_block
    gpu_info["device_name"] = torch.cuda.get_device_name(torch.cuda.current_device())
