"""Configuration checks and device attach, mirroring bayes_validate.py:10-55 of the reference.
`connect_to_gpu` fills the same gpu_info keys (has_GPU, threads_per_block,
max_sims_per_block) so that reference-style driver code keeps working; it asks torch/CUDA
instead of numba."""
import torch


def validate_IC(ics, L):
    for i, ic in enumerate(ics):
        assert len(ic) == L, "Error: IC #{} length:{}, declared L:{}".format(i, len(ic), L)


def validate_ic_flags(ic_flags):
    tc = ic_flags["time_cutoff"]
    assert tc is None or (isinstance(tc, (float, int)) and tc > 0), "invalid time cutoff"
    sel = ic_flags["select_obs_sets"]
    assert sel is None or isinstance(sel, list), "invalid observation set selection"
    nl = ic_flags["noise_level"]
    assert nl is None or isinstance(nl, (float, int)), "invalid noise level"


def validate_gpu_info(gpu_info):
    assert isinstance(gpu_info["num_gpus"], int) and gpu_info["num_gpus"] > 0, "invalid num_gpus"
    assert isinstance(gpu_info["sims_per_gpu"], int) and gpu_info["sims_per_gpu"] > 0, \
        "invalid sims per gpu"


def validate_params(num_params, unit_conversions, do_log, minX, maxX):
    for name, arr in (("Unit conversion array", unit_conversions), ("do_log mask", do_log),
                      ("min param values", minX), ("max param values", maxX)):
        assert len(arr) == num_params, name + " is missing entries"
    assert all(minX <= maxX), "Min params larger than max params"


def connect_to_gpu(gpu_info, nthreads=128, sims_per_block=1):
    gpu_info["has_GPU"] = bool(torch.cuda.is_available())
    if gpu_info["has_GPU"]:
        gpu_info["threads_per_block"] = (nthreads,)
        gpu_info["max_sims_per_block"] = sims_per_block
        gpu_info["device_name"] = torch.cuda.get_device_name(torch.cuda.current_device())
