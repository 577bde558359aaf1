"""Comparison helpers shared by the emulator (CPU) and GPU parity tests.

Tolerances are the north star's (BASELINE.json): |df| < 1e-3 * sr/nfft, relative magnitude
1e-4, phase 1e-4 rad; realph = ph + pi*df/fstep inherits both, so its bound is
1e-4 + pi*1e-3 rad.  Peak bins must be bit-exact in every frame whose decision margin (oracle
``peak_margin``: smallest gap of any comparison that shapes the row, relative to max|fx|)
exceeds MARGIN_FP32; frames under it are counted and reported.
"""
import numpy as np

TOL_F = 1e-3        # in units of fstep = sr/nfft
TOL_MAG = 1e-4      # relative
TOL_PH = 1e-4       # rad
TOL_REALPH = 1e-4 + np.pi * 1e-3
MARGIN_FP32 = 2e-6  # relative to max|fx| of the frame


def angdiff(a, b):
    return np.abs(np.angle(np.exp(1j * (a - b))))


def compare_analysis(got, ref, sr, nfft, margin=None):
    """got/ref: dicts with f mag ph realph binno [F, K] (+ totalmag).  Returns a report dict;
    raises AssertionError on a violation."""
    gb, rb = np.asarray(got["binno"]), np.asarray(ref["binno"])
    assert gb.shape == rb.shape, (gb.shape, rb.shape)
    if rb.ndim != 2:
        return dict(frames=0, mismatched=0)
    bad = (gb != rb).any(axis=1)
    if margin is not None:
        must = margin > MARGIN_FP32
        assert not (bad & must).any(), "peak bins differ in %d frames with margin > %g (first %s)" % (
            int((bad & must).sum()), MARGIN_FP32, np.flatnonzero(bad & must)[:5])
    else:
        assert not bad.any(), "peak bins differ in frames %s" % np.flatnonzero(bad)[:5]
    ok = ~bad
    fstep = sr / float(nfft)
    rep = dict(frames=len(rb), mismatched=int(bad.sum()))
    if ok.any():
        valid = rb[ok] > 0
        rep["df"] = float(np.abs(got["f"][ok] - ref["f"][ok]).max() / fstep)
        den = np.where(valid, np.abs(ref["mag"][ok]), 1.0)
        rep["dmag"] = float((np.abs(got["mag"][ok] - ref["mag"][ok]) / den).max())
        rep["dph"] = float(angdiff(got["ph"][ok], ref["ph"][ok]).max())
        rep["drealph"] = float(angdiff(got["realph"][ok], ref["realph"][ok]).max())
        assert rep["df"] < TOL_F, rep
        assert rep["dmag"] < TOL_MAG, rep
        assert rep["dph"] < TOL_PH, rep
        assert rep["drealph"] < TOL_REALPH, rep
    if "totalmag" in got and "totalmag" in ref:
        tg, tr = np.asarray(got["totalmag"], dtype=float), np.asarray(ref["totalmag"], dtype=float)
        rep["dtotalmag"] = float((np.abs(tg - tr) / np.maximum(tr, 1e-30)).max()) if len(tr) else 0.0
        assert rep["dtotalmag"] < TOL_MAG, rep
    return rep


def compare_exact_on_spectrum(got, orc_out, rtol=1e-11):
    """Kernel logic vs oracle run on the kernel's own fp32 spectrum: integer results must be
    bit-exact in every frame, floating results agree to fp64 rounding."""
    assert np.array_equal(got["binno"], orc_out["binno"])
    assert np.array_equal(got["npk"], orc_out["npk"])
    for k in ("f", "mag", "ph", "realph"):
        d = np.abs(got[k] - orc_out[k]) / (np.abs(orc_out[k]) + 1.0)
        m = float(np.nanmax(d)) if d.size else 0.0
        assert m <= rtol, (k, m)
    assert np.allclose(got["totalmag"], np.asarray(orc_out["totalmag"]), rtol=1e-6, atol=0)


def snr_db(x, ref):
    num = float(np.sum(np.asarray(ref, dtype=float) ** 2))
    den = float(np.sum((np.asarray(x, dtype=float) - np.asarray(ref, dtype=float)) ** 2))
    return 10 * np.log10(num / max(den, 1e-300))
