"""Posterior products on the GPU -- the step right after the hot path: the
computations of the reference GUI back end, `Visualization/utils.py:157-285` (normalize,
marginalize_1D, marginalize_2D, w_mean, w_variance, covariance), on device tensors, with the
raw sums all-reduced over the ranks so every rank gets the statistics of ALL samples."""
import torch
import torch.distributed as dist

from . import distributed, engine


def _allreduce_sum(t):
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def log_evidence(lnP):
    """log sum_i exp(lnP_i) over every rank's samples (device tensor [S] -> 0-d tensor)."""
    return distributed.global_logsumexp(engine.lse_partial(lnP.contiguous()))


def normalize(lnP):
    """Posterior weights summing to one over all ranks (utils.normalize, shift-invariant form)."""
    return engine.posterior_weights(lnP.contiguous(), float(log_evidence(lnP)))


def marginalize_1D(P, X, col, lo, hi, bin_count, correct_sampling=False):
    """Density-normalised weighted histogram of column `col` (utils.marginalize_1D); with
    correct_sampling the weighted counts are divided by the raw counts first."""
    h = _allreduce_sum(engine.weighted_hist(X, col, P, lo, hi, bin_count))
    bins = lo + (hi - lo) * torch.arange(bin_count + 1, dtype=torch.float64, device=X.device) / bin_count
    width = torch.diff(bins)
    dens = h / (h.sum() * width)
    if correct_sampling:
        cnt = _allreduce_sum(engine.weighted_hist(X, col, None, lo, hi, bin_count))
        dens = torch.where(cnt != 0, dens / cnt.clamp(min=1), torch.zeros_like(dens))
        dens = dens / (width * dens).sum()
    return dens, bins


def marginalize_2D(P, X, colx, coly, lox, hix, loy, hiy, bin_count):
    """Density-normalised weighted 2-D histogram (numpy.histogram2d(..., density=True))."""
    h = _allreduce_sum(engine.weighted_hist(X, colx, P, lox, hix, bin_count, coly=coly, loy=loy,
                                            hiy=hiy, nby=bin_count))
    area = ((hix - lox) / bin_count) * ((hiy - loy) / bin_count)
    return h / (h.sum() * area)


def moments(P, X, ncol=None):
    """(mean [ncol], covariance [ncol,ncol]) with weights P: utils.w_mean / w_variance / covariance."""
    ncol = X.shape[1] if ncol is None else ncol
    raw = _allreduce_sum(engine.weighted_moments(X, P, ncol))
    sw = raw[0]
    mean = raw[1:1 + ncol] / sw
    second = raw[1 + ncol:].reshape(ncol, ncol) / sw
    return mean, second - torch.outer(mean, mean)
