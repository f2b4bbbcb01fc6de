"""On-disk chain formats of the reference, byte-compatible.

  <root>.stats               one line  ``log(Z)   : <v>   +/-   <e>``           (writer cli.py:292-295)
  <root>_equal_weights.txt   rows ``[weight, -2*logL, theta...]`` via np.savetxt (writer cli.py:312-325)

``read_chains`` returns what ``pc_analyzer`` (hires_fitter.py:704-747) returns for the same files, including
the redshift sort of the components and the NaN fill above the active count (vectorised over the samples); ``logl_of_chain`` re-evaluates a chain's samples in
one batch through the CUDA path (what the plotting code does point by point, cli.py:414-418).
"""
import re

import numpy as np


def write_stats(filesbasename, log_z, log_z_err):
    with open(filesbasename + ".stats", "w") as f:
        f.write('log(Z)   : {}   +/-   {}\n'.format(float(log_z), float(log_z_err)))


def write_equal_weights(filesbasename, logl, samples):
    logl = np.asarray(logl, dtype=float).reshape(-1, 1)
    samples = np.asarray(samples, dtype=float).reshape(logl.shape[0], -1)
    out = np.hstack([np.ones((logl.shape[0], 1)), -2.0 * logl, samples])
    np.savetxt(filesbasename + "_equal_weights.txt", out)


_STATS_LINE = re.compile(r"^log\(Z\)\s*:\s*(\S+)\s*\+/-\s*(\S+)")


def read_stats(filesbasename):
    """(log Z, its error) from ``<root>.stats``; the last ``log(Z)`` line wins, as in the reference reader."""
    lnz = lnz_err = None
    with open(filesbasename + ".stats", "r") as fh:
        for m in filter(None, map(_STATS_LINE.match, fh)):
            lnz, lnz_err = float(m.group(1)), float(m.group(2))
    return lnz, lnz_err


def read_chains(filesbasename, return_sorted=True):
    """-> (lnz, lnz_err, logL of every sample, samples): what ``pc_analyzer`` (hires_fitter.py:704-747) returns
    for the same files.  With ``return_sorted`` every row's active components (the first ``int(ncomp)``
    [N, z, b] triples after the ncomp column) are reordered by increasing redshift and everything behind them
    is NaN.  The column of the ncomp slot follows from the row length alone: (ncols - 1) mod 3."""
    lnz, lnz_err = read_stats(filesbasename)
    table = np.loadtxt(filesbasename + "_equal_weights.txt", ndmin=2)
    logl = -0.5 * table[:, 1]
    theta = table[:, 2:]
    if not return_sorted:
        return lnz, lnz_err, logl, theta
    n, ncols = theta.shape
    first = (ncols - 1) % 3 + 1                        # first component column; the ncomp slot sits before it
    ntrip = (ncols - first) // 3
    trip = theta[:, first:].reshape(n, ntrip, 3)
    nact = np.clip(theta[:, first - 1].astype(np.int64), 0, ntrip)
    slot = np.arange(ntrip)[None, :]
    active = slot < nact[:, None]
    order = np.argsort(np.where(active, trip[:, :, 1], np.inf), axis=1, kind="stable")    # active ones first, by z
    trip = np.take_along_axis(trip, order[:, :, None], axis=1)
    trip[~active] = np.nan                             # after the sort the first nact slots are the active ones
    out = theta.copy()
    out[:, first:] = trip.reshape(n, 3 * ntrip)
    return lnz, lnz_err, logl, out


def logl_of_chain(fitter, filesbasename):
    """Re-evaluate every sample of a chain in one batched launch; returns (stored logL, recomputed logL)."""
    _, _, lhood, samples = read_chains(filesbasename, return_sorted=False)
    return lhood, fitter.lnlhood_batch(np.ascontiguousarray(samples))
