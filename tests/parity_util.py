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


def compare_exact_on_spectrum(got, orc_out):
    """Kernel logic vs the oracle run on the kernel's own fp32 spectrum.  Integer results (bins,
    peak counts, i.e. every decision incl. the freq > 0 filter and the unwrap candidate) must be
    bit-exact in every frame.  The per-peak values come from fp32 angles / magnitudes (atan2f,
    sqrtf on the fp32 spectrum; round 1 used fp64 here and compared at 1e-11):
        |d ph|   <= 2e-6 rad                       (atan2f: <= 2 ulp of pi ~ 5e-7)
        |d mag|  <= 1e-6 relative                  (two fp32 adds + sqrtf)
        |d f|    <= 1e-6 * (nfft/hop) * fstep      (a phase-difference error e moves f by e/(2 pi) * nfft/hop bins)
        |d realph| <= 2e-6 * (1 + nfft/hop)        (ph + pi*df/fstep)
    all >= 50x inside the north-star tolerances at the configs' overlaps (nfft/hop <= 8)."""
    assert np.array_equal(got["binno"], orc_out["binno"])
    assert np.array_equal(got["npk"], orc_out["npk"])
    ratio = float(orc_out["nfft"]) / float(orc_out["hop"])
    fstep = float(orc_out["sr"]) / float(orc_out["nfft"])
    rep = {}
    if np.asarray(orc_out["f"]).size:
        rep["df"] = float(np.nanmax(np.abs(got["f"] - orc_out["f"]))) / fstep
        rep["dmag"] = float(np.nanmax(np.abs(got["mag"] - orc_out["mag"]) / np.maximum(np.abs(orc_out["mag"]), 1e-300)))
        rep["dph"] = float(np.nanmax(np.abs(got["ph"] - orc_out["ph"])))
        rep["drealph"] = float(np.nanmax(np.abs(got["realph"] - orc_out["realph"])))
        assert rep["df"] <= 1e-6 * ratio, rep
        assert rep["dmag"] <= 1e-6, rep
        assert rep["dph"] <= 2e-6, rep
        assert rep["drealph"] <= 2e-6 * (1.0 + ratio), rep
        for k in ("f", "mag", "ph", "realph"):
            assert np.array_equal(np.isnan(got[k]), np.isnan(orc_out[k])), k
    assert np.allclose(got["totalmag"], np.asarray(orc_out["totalmag"]), rtol=1e-6, atol=0)
    return rep


def snr_db(x, ref):
    num = float(np.sum(np.asarray(ref, dtype=float) ** 2))
    den = float(np.sum((np.asarray(x, dtype=float) - np.asarray(ref, dtype=float)) ** 2))
    return 10 * np.log10(num / max(den, 1e-300))


def compare_harmonic(got, ref, f0, sr, nfft, totalmag, exact=False):
    """PVHarmonic tables: got/ref dicts with f mag ph [F, K] and residuals [F].

    Skipped frames (f0 <= 0 or NaN) must be zero rows with a NaN residual, bit for bit.
    ``exact``: ref was computed from the kernel's own spectrum -> every entry to fp64 rounding.
    Otherwise ref is the fp64 reference: north-star tolerances on the harmonics that carry
    energy (mag > 1e-2 of the frame's ``totalmag`` >= max|fx|; the fp32 FFT's relative error on
    a bin grows as max/|bin|), and the residual through its square (power), which is what is subtracted:
    |residual^2 - ref^2| <= 2e-5 (1e-12 when exact) of the frame's total power ``totalmag^2``."""
    f0 = np.asarray(f0, dtype=float)[:len(ref["f"])]
    skip = ~(f0 > 0)
    for k in ("f", "mag", "ph"):
        assert got[k].shape == ref[k].shape, (k, got[k].shape, ref[k].shape)
        assert not np.any(got[k][skip]), k
    assert np.all(np.isnan(got["residuals"][skip]))
    run = ~skip
    rep = dict(frames=int(len(f0)), processed=int(run.sum()))
    if not run.any():
        return rep
    gf, rf = got["f"][run], ref["f"][run]
    gm, rm = got["mag"][run], ref["mag"][run]
    gp, rp = got["ph"][run], ref["ph"][run]
    assert np.array_equal(np.isnan(gf), np.isnan(rf))
    fstep = sr / float(nfft)
    if exact:
        for k, a, b in (("f", gf, rf), ("mag", gm, rm), ("ph", gp, rp)):
            d = np.abs(a - b) / (np.abs(b) + 1.0)
            m = float(np.nanmax(d)) if d.size else 0.0
            assert m <= 1e-11, (k, m)
        strong = np.ones_like(rm, dtype=bool)
    else:
        strong = rm > 1e-2 * np.asarray(totalmag, dtype=float)[:len(f0)][run][:, None]
        strong &= ~np.isnan(rf)
        rep["df"] = float(np.abs(gf - rf)[strong].max() / fstep) if strong.any() else 0.0
        rep["dmag"] = float((np.abs(gm - rm)[strong] / rm[strong]).max()) if strong.any() else 0.0
        rep["dph"] = float(angdiff(gp, rp)[strong].max()) if strong.any() else 0.0
        assert rep["df"] < TOL_F and rep["dmag"] < TOL_MAG and rep["dph"] < TOL_PH, rep
    # residual^2 = total power - harmonic power
    g2, r2 = got["residuals"][run], ref["residuals"][run]
    power = np.asarray(totalmag, dtype=float)[:len(f0)][run] ** 2
    tol = (1e-12 if exact else 2e-5) * np.maximum(power, 1e-300)
    both = ~np.isnan(g2) & ~np.isnan(r2)
    assert np.all(np.abs(g2[both] ** 2 - r2[both] ** 2) <= tol[both]), "residual power differs"
    one = np.isnan(g2) != np.isnan(r2)          # the difference changed sign within rounding
    fin = np.where(np.isnan(g2), r2, g2)
    assert np.all(fin[one] ** 2 <= tol[one]), "residual NaN pattern differs"
    rep["strong"] = int(strong.sum())
    return rep


def wide_rows(K, mode, F=9):
    """Random peak tables [F, K] for the link kernels: ascending gap-free rows ("asc", "asc_dense"),
    rows with holes and one empty frame ("gaps"), out-of-order columns ("shuffled"), many equal
    magnitudes ("ties")."""
    rng = np.random.RandomState(K)
    f = np.zeros((F, K))
    mag = np.zeros((F, K))
    base = np.sort(rng.uniform(200., 18000., K)) if mode != "asc_dense" else 200. + 30. * np.arange(K)
    for j in range(F):
        n = K if mode in ("asc", "asc_dense", "ties") else rng.randint(K // 2, K + 1)
        fr = np.sort(base[:n] * (1.0 + 0.004 * rng.randn(n)))
        mg = rng.uniform(0.01, 1.0, n)
        if mode == "ties":
            mg = np.round(mg, 1) + 0.1                       # many equal magnitudes
        if mode == "shuffled":
            perm = rng.permutation(n)
            fr, mg = fr[perm], mg[perm]
        cols = np.arange(n)
        if mode == "gaps":
            cols = np.sort(rng.choice(K, n, replace=False))
        f[j, cols], mag[j, cols] = fr, mg
    if mode == "gaps":
        f[4] = 0.0; mag[4] = 0.0                            # an empty frame: everything restarts
    return f, mag
