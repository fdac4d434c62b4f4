"""Reader/writer for the reference's file formats (host side, runs once per job):
observation CSVs (`t, PL, uncertainty` rows, curves separated by a return to t=0, closed by an
`END` row -- bayes_io.py:15-104), excitation CSVs (one row of L densities per curve --
bayes_io.py:106-119) and the `<name>_BAYRAN_P.npy` / `<name>_BAYRAN_X.npy` pair
(bayes_io.py:121-140).  Pre-processing follows the reference: scale, optional noise,
optional self-normalisation, |PL| clamped to sys.float_info.min and log10, uncertainty
converted to log10 units."""
import os
import sys

import numpy as np


def _read_rows(path):
    """All `t, PL, uncertainty` rows of an observation file up to its END row as one [n,3] array, parsed
    in a single vectorised pass (the row-by-row Python loop of the reference costs ~1 s per 240 k rows)."""
    with open(path, newline="") as fh:
        text = fh.read()
    end = text.find("END")
    if end >= 0:
        text = text[:end]
    lines = [ln for ln in text.replace("\r", "").split("\n") if ln.strip() != ""]
    try:
        vals = np.array(",".join(lines).split(","), dtype=np.float64)
        if vals.size == 3 * len(lines):
            return vals.reshape(-1, 3)
    except ValueError:
        pass
    rows = []                                   # ragged rows (extra columns, trailing commas): first three fields
    for ln in lines:
        parts = ln.strip().split(",")
        rows.append((float(parts[0]), float(parts[1]), float(parts[2])))
    return np.array(rows, dtype=np.float64).reshape(-1, 3)


def _split_curves(rows):
    """Split the [n,3] row array into curves: a new curve starts whenever t returns to 0."""
    if len(rows) == 0:
        return []
    starts = np.flatnonzero(rows[:, 0] == 0)
    starts = starts[starts > 0]
    return np.split(rows, starts)


def get_data(exp_files, ic_flags, sim_flags, logger=None, scale_f=1e-23):
    """-> [(t_list, PL_list, unc_list)] with one entry per file and one array per curve."""
    cutoff = sys.float_info.min
    t_max = ic_flags.get("time_cutoff")
    select = ic_flags.get("select_obs_sets")
    noise = ic_flags.get("noise_level")
    log_pl = sim_flags["log_pl"]
    normalize = sim_flags["self_normalize"]
    out = []
    for path in exp_files:
        rows = _read_rows(path)
        ts, pls, uncs = [], [], []
        for k, arr in enumerate(_split_curves(rows)):
            if t_max is not None:
                arr = arr[arr[:, 0] <= t_max]
            t, pl, unc = arr[:, 0].copy(), arr[:, 1].copy(), arr[:, 2].copy()
            if noise is not None:
                pl = pl + noise * np.random.normal(0, 1, len(pl))
            pl *= scale_f
            unc *= scale_f
            if normalize:
                pl /= pl.max()
            if logger is not None:
                logger.info("PL curve #%d: %d points, t in [%g, %g]", k + 1, len(t), t[0], t[-1])
            if log_pl:
                pl = np.abs(pl)
                pl[pl < cutoff] = cutoff
                unc = unc / pl / 2.3
                pl = np.log10(pl)
            ts.append(t)
            pls.append(pl)
            uncs.append(unc)
        if select is not None:
            ts = [ts[i] for i in select]
            pls = [pls[i] for i in select]
            uncs = [uncs[i] for i in select]
        out.append((ts, pls, uncs))
    return out


def get_initpoints(init_file, ic_flags, scale_f=1e-21):
    """Excitation profiles [C, L] in nm^-3 (file values are cm^-3)."""
    rows = []
    with open(init_file, newline="") as fh:
        for line in fh:
            parts = [p for p in line.strip().split(",") if p != ""]
            if parts:
                rows.append([float(p) for p in parts])
    select = ic_flags.get("select_obs_sets")
    if select is not None:
        rows = [rows[i] for i in select]
    return np.array(rows, dtype=np.float64) * scale_f


def export(out_filename, P, X, logger=None):
    """Write <out>/<base>_BAYRAN_P.npy and <out>/<base>_BAYRAN_X.npy."""
    os.makedirs(out_filename, exist_ok=True)
    base = os.path.basename(os.path.normpath(out_filename))
    np.save(os.path.join(out_filename, base + "_BAYRAN_P.npy"), P)
    np.save(os.path.join(out_filename, base + "_BAYRAN_X.npy"), X)
    if logger is not None:
        logger.info("wrote %s_BAYRAN_[PX].npy to %s", base, out_filename)


def export_from_device(out_filename, P_dev, X_dev, logger=None, chunk_rows=1 << 16):
    """`export` for tables that live on the GPU: the .npy headers are written by numpy's own format
    module and the device buffers are streamed through one pinned staging chunk straight into the files,
    so a 1M-sample run (8 MB of P, 104 MB of X) never materialises a second host copy."""
    import torch
    from numpy.lib import format as npy_format
    os.makedirs(out_filename, exist_ok=True)
    base = os.path.basename(os.path.normpath(out_filename))
    for suffix, t in (("_BAYRAN_P.npy", P_dev), ("_BAYRAN_X.npy", X_dev)):
        t = t.detach()
        if t.dtype != torch.float64:
            t = t.double()
        t = t.contiguous()
        flat = t.reshape(t.shape[0], -1) if t.dim() > 1 else t.reshape(-1, 1)
        rows = min(chunk_rows, max(1, flat.shape[0]))
        stage = torch.empty((rows, flat.shape[1]), dtype=torch.float64).pin_memory() if t.is_cuda else None
        with open(os.path.join(out_filename, base + suffix), "wb") as fh:
            npy_format.write_array_header_1_0(fh, {"descr": "<f8", "fortran_order": False, "shape": tuple(t.shape)})
            for lo in range(0, flat.shape[0], rows):
                part = flat[lo:lo + rows]
                if stage is not None:
                    stage[:part.shape[0]].copy_(part)
                    fh.write(stage[:part.shape[0]].numpy().tobytes())
                else:
                    fh.write(part.numpy().tobytes())
    if logger is not None:
        logger.info("wrote %s_BAYRAN_[PX].npy to %s", base, out_filename)


def write_observations(path, times, pl_values, uncertainty=1e14, scale_f=1e-23):
    """Write simulated curves in the reference's observation format (used to synthesise the
    Power_scan/Twothick observation files that are missing from the reference checkout)."""
    with open(path, "w", newline="") as fh:
        for t, pl in zip(times, pl_values):
            unc = np.broadcast_to(uncertainty, np.shape(t))
            for ti, pi, ui in zip(t, np.asarray(pl) / scale_f, unc):
                fh.write("%.10G,%.9E,%G\n" % (ti, pi, ui))
        fh.write("END\n")
