"""CPU: chain files round-trip and match what the reference's own reader makes of them."""
import os

import numpy as np
import pytest

from mcalf_b200 import chains
from oracle import refshim


def _fake_chain(tmp_path, ndim=11, n=40, startind=1):
    rng = np.random.default_rng(0)
    samples = rng.random((n, ndim))
    samples[:, startind] = rng.integers(0, 4, n)
    logl = -rng.random(n) * 100
    base = os.path.join(tmp_path, "chain_0")
    chains.write_stats(base, -123.456, 0.789)
    chains.write_equal_weights(base, logl, samples)
    return base, logl, samples


def test_round_trip(tmp_path):
    base, logl, samples = _fake_chain(str(tmp_path))
    assert open(base + ".stats").read() == "log(Z)   : -123.456   +/-   0.789\n"
    lnz, err, lh, post = chains.read_chains(base, return_sorted=False)
    assert (lnz, err) == (-123.456, 0.789)
    assert np.allclose(lh, logl, rtol=1e-15) and np.allclose(post, samples, rtol=1e-15)
    first = open(base + "_equal_weights.txt").readline().split()
    assert first[0] == "1.000000000000000000e+00" and len(first) == 2 + samples.shape[1]
    _, _, _, srt = chains.read_chains(base)
    for row, orig in zip(srt, samples):
        nc = int(orig[1])
        z = row[3:2 + 3 * nc:3]
        assert np.all(np.diff(z) >= 0) and np.isnan(row[2 + 3 * nc:]).all()


@pytest.mark.skipif(not refshim.available(), reason="reference tree absent (GPU box)")
def test_reader_matches_reference_pc_analyzer(tmp_path):
    hf = refshim.install()
    base, _, _ = _fake_chain(str(tmp_path))
    ref = hf.pc_analyzer(base)
    got = chains.read_chains(base)
    assert ref[0] == got[0] and ref[1] == got[1]
    assert np.array_equal(ref[2], got[2]) and np.array_equal(ref[3], got[3], equal_nan=True)
