"""On-disk chain formats of the reference, byte-compatible.

  <root>.stats               one line  ``log(Z)   : <v>   +/-   <e>``           (writer cli.py:292-295)
  <root>_equal_weights.txt   rows ``[weight, -2*logL, theta...]`` via np.savetxt (writer cli.py:312-325)

``read_chains`` follows ``pc_analyzer`` (hires_fitter.py:704-747), including the redshift sort of the
components and the NaN fill above ``thisncomp``; ``logl_of_chain`` re-evaluates a chain's samples in
one batch through the CUDA path (what the plotting code does point by point, cli.py:414-418).
"""
import numpy as np


def write_stats(filesbasename, log_z, log_z_err):
    with open(filesbasename + ".stats", "w") as f:
        f.write('log(Z)   : {}   +/-   {}\n'.format(float(log_z), float(log_z_err)))


def write_equal_weights(filesbasename, logl, samples):
    logl = np.asarray(logl, dtype=float).reshape(-1, 1)
    samples = np.asarray(samples, dtype=float).reshape(logl.shape[0], -1)
    out = np.hstack([np.ones((logl.shape[0], 1)), -2.0 * logl, samples])
    np.savetxt(filesbasename + "_equal_weights.txt", out)


def read_chains(filesbasename, return_sorted=True):
    """-> (lnz, lnz_err, lhoodsamples, samples) exactly as ``pc_analyzer`` returns them."""
    lnz = lnz_err = None
    with open(filesbasename + ".stats", "r") as f:
        for line in f:
            if line[:6] == 'log(Z)':
                items = line.split()
                lnz, lnz_err = float(items[2]), float(items[4])
    allsamples = np.loadtxt(filesbasename + "_equal_weights.txt", ndmin=2)
    lhoodsamples = -0.5 * allsamples[:, 1]
    postsamples = allsamples[:, 2:]
    if not return_sorted:
        return lnz, lnz_err, lhoodsamples, postsamples
    postsorted = np.copy(postsamples)
    ncols = postsorted.shape[1]
    startind = (ncols - 1) % 3
    for ii in range(postsamples.shape[0]):
        thisncomp = int(postsamples[ii, startind])
        thisendind = startind + 1 + 3 * thisncomp
        postsamples[ii, thisendind:] = 99
        postsorted[ii, thisendind:] = 99
        zsort = np.argsort(postsamples[ii, startind + 2:startind + 1 + 3 * thisncomp:3])
        for jj in range(len(zsort)):
            postsorted[ii, 3 * jj + startind + 1:3 * jj + 3 + startind + 1] = \
                postsamples[ii, 3 * zsort[jj] + np.array([0, 1, 2]) + startind + 1]
        postsorted[postsorted == 99] = np.nan
    return lnz, lnz_err, lhoodsamples, postsorted


def logl_of_chain(fitter, filesbasename):
    """Re-evaluate every sample of a chain in one batched launch; returns (stored logL, recomputed logL)."""
    _, _, lhood, samples = read_chains(filesbasename, return_sorted=False)
    return lhood, fitter.lnlhood_batch(np.ascontiguousarray(samples))
