"""Helpers of the STFT-consumer parity tests: golden cases (tests/golden/stft.npz, written from the
real reference by oracle/gen_golden_stft.py), their regenerated inputs and the tolerances."""
import json
import os

import numpy as np

from pypevoc_b200 import signals

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

with open(os.path.join(GOLD, "stft_cases.json")) as _fh:
    CASES = json.load(_fh)

# fp32 signal + fp32 FFT against the reference's fp64: relative to the largest value of the table
# (band energies and RMS are sums of squares -> 2x the spectrum's relative error)
TOL_BANK = 2e-5
TOL_RMS = 1e-5
# flux is a difference of magnitudes: the error scales with the spectrum, not with the flux
TOL_FLUX = 5e-5


def golden():
    return np.load(os.path.join(GOLD, "stft.npz"))


def signal(name):
    gen, kw = CASES["signals"][name]
    out = getattr(signals, gen)(**kw)
    x, sr = out if isinstance(out, tuple) else (out, kw["sr"])
    return np.asarray(x, dtype=np.float32), sr


def windfunc(kw):
    kw = dict(kw)
    if "windfunc" in kw:
        kw["windfunc"] = getattr(np, kw["windfunc"])
    return kw


def close(got, ref, tol, scale=None):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    s = np.max(np.abs(ref)) if scale is None else scale
    err = np.max(np.abs(got - ref)) / s if got.size else 0.0
    assert err < tol, "max error %.3g (relative to %.3g) exceeds %.3g" % (err, s, tol)
    return err


def band_norm(x, wind, hop, minbin, maxbin):
    """Largest norm of a frame's magnitude spectrum over bins [minbin, maxbin): the scale of the
    flux tolerance (the flux is a difference of magnitudes)."""
    n = len(wind)
    best = 0.0
    for i in range(0, len(x) - n, hop):
        spec = np.abs(np.fft.fft(x[i:i + n].astype(np.float64) * wind))
        best = max(best, float(np.sqrt(np.sum(spec[minbin:maxbin] ** 2))))
    return best
