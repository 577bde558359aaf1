"""TEST INFRASTRUCTURE ONLY -- numpy restatement of PyPeVoc's phase-vocoder hot path.

This is the CPU oracle the CUDA path is checked against (and the ``cpu_baseline`` that
bench.py times).  It is *not* part of the product and the product never imports it.

It restates, in vectorised numpy, what the reference computes with python loops; every
function cites the reference lines (relative to /root/reference/pypevoc/) it follows.
The arithmetic that decides results (np.fft.fft, np.angle, np.round, np.argmin,
np.argsort, np.interp, np.cumsum, np.mod, np.linspace, python ``sum``) is the same numpy
call the reference makes, on the same operands, in the same order -- numpy is the
reference's (un-vendored, un-pinned: setup.py has no install_requires) arithmetic library.

Parity status: PINNED against the real reference run in the build container
(oracle/gen_golden.py -> tests/golden/*.npz; tests/test_oracle_golden.py), bit-for-bit
for analysis and tracking, max-abs-diff 0.0 for resynthesis.
"""
import numpy as np

pi2 = 2.0 * np.pi


# --------------------------------------------------------------------------- tables
def pv_tables(sr, nfft, hop, wind=np.hanning):
    """Window / normalisation / per-bin tables of ``PV.__init__`` (PVAnalysis.py:97-118)."""
    win = wind(nfft)
    wsum2 = sum(win ** 2)                      # python sum, left to right (:99)
    wfact = np.sqrt(wsum2 * nfft) / 2.0        # :102
    fstep = float(sr) / float(nfft)            # :105
    dt = float(hop) / float(sr)                # :108
    fbin = np.arange(float(nfft)) * fstep      # :114
    wfbin = np.round(pi2 * fbin * dt / pi2) * pi2   # :116-118
    return dict(win=win, wfact=wfact, fstep=fstep, dt=dt, fbin=fbin, wfbin=wfbin)


def n_frames(nsamp, nfft, hop):
    """Frame count of ``run_pv``'s ``while curpos < nsamp-nfft`` loop (:223-225,249)."""
    span = nsamp - nfft
    return 0 if span <= 0 else -(-span // hop)


# --------------------------------------------------------------------------- peaks
def peak_select(y, npks, pkthresh):
    """Top-``npks`` peak bins *before* the salience filter.

    Restates ``PeakFinder.__init__`` + ``findpos`` (PeakFinder.py:57-70,155-194):
    interior local maxima ``y[i-1] < y[i] >= y[i+1]`` scored ``y[i]-min(y)``, strict
    threshold ``max(y)*pkthresh - min(y)`` (``min(y)`` replaces a zero threshold, :69-70),
    repeated arg-max == order by (score desc, index asc); result sorted by bin (:189).
    """
    y = np.asarray(y)
    miny = np.min(y)
    minamp = y.max() * pkthresh
    if not minamp:
        minamp = miny
    th = minamp - miny
    inner = y[1:-1]
    mask = (y[:-2] < inner) & (inner >= y[2:])
    score = mask * (inner - miny)
    cand = np.flatnonzero(score > th)
    order = np.lexsort((cand, -score[cand]))
    return np.sort(cand[order[:npks]] + 1), th, miny


def salience_keep(y, bins, rad=5):
    """``filter_by_salience(rad)`` (PeakFinder.py:113-134): drop a peak when any value in
    ``y[max(i-rad,1) : min(i+rad,len(y))+1]`` is strictly larger."""
    n = len(y)
    keep = np.ones(len(bins), dtype=bool)
    for q, i in enumerate(bins):
        lo, hi = max(i - rad, 1), min(i + rad, n)
        keep[q] = not np.any(y[lo:hi + 1] > y[i])
    return keep


def peak_pick(y, npks, pkthresh, rad=5):
    """Peak bins as ``PV.calc_pv_frame`` obtains them (PVAnalysis.py:175-178).
    ``boundaries()`` (:176) has no effect on the result and is skipped."""
    sel, _, _ = peak_select(y, npks, pkthresh)
    return sel[salience_keep(y, sel, rad)]


def peak_margin(y, npks, pkthresh, rad=5):
    """Smallest gap (relative to max(y)) of any comparison that decides the peak row.

    A frame whose margin exceeds the fp32 FFT error must give bit-identical bins on the
    GPU; frames below it are reported, not required (north star / SURVEY H4).
    """
    y = np.asarray(y, dtype=np.float64)
    ymax = y.max()
    if ymax <= 0:
        return np.inf
    sel, th, miny = peak_select(y, npks, pkthresh)
    inner = y[1:-1]
    score_all = inner - miny
    # bins that are, or could become under a tiny perturbation, candidates
    near = np.flatnonzero(score_all > th - 1e-3 * ymax)
    gaps = [np.inf]
    if len(near):
        i = near + 1
        gaps.append(np.min(np.abs(y[i] - y[i - 1])))
        gaps.append(np.min(np.abs(y[i] - y[i + 1])))
        gaps.append(np.min(np.abs(score_all[near] - th)))
    mask = (y[:-2] < inner) & (inner >= y[2:])
    sc = mask * score_all
    cand = np.flatnonzero(sc > th)
    if len(cand) > npks:
        s = np.sort(sc[cand])[::-1]
        gaps.append(s[npks - 1] - s[npks])
    n = len(y)
    for i in sel:
        lo, hi = max(i - rad, 1), min(i + rad, n)
        w = np.delete(y[lo:hi + 1], i - lo)
        if len(w):
            gaps.append(np.min(np.abs(w - y[i])))
    return min(gaps) / ymax


# --------------------------------------------------------------------------- analysis
def dphase2freq(dph, bins, tb):
    """``PV.dphase2freq`` (PVAnalysis.py:133-147), vectorised over the peaks of a frame."""
    dphw = (dph + tb["wfbin"][bins])[:, None] + (pi2 * np.arange(-1, 2))[None, :]
    freq = dphw / tb["dt"] / pi2
    df = tb["fbin"][bins][:, None] - freq
    ii = np.argmin(np.abs(df), axis=1)
    r = np.arange(len(bins))
    return freq[r, ii], df[r, ii]


def analyze(x, sr, nfft=1024, hop=None, npks=20, pkthresh=0.005, wind=np.hanning,
            fft_dtype=np.float64, spectra=False, margins=False, fx_given=None, old0=None):
    """``PV(...).run_pv()`` (PVAnalysis.py:72-131,150-264).

    Returns a dict with the reference's attributes: ``f mag ph realph binno`` float64
    ``[nframes, npks]`` zero padded, ``t``, ``totalmag`` (list), ``nframes`` plus ``npk``
    (valid entries per row).  ``fft_dtype=np.float32`` emulates a single-precision FFT
    (feasibility probe only).  ``spectra=True`` adds ``fx`` ``[nframes, nfft/2]``
    (``calc_fft_frame(pos)[:nfft/2]``, :150-158,169).

    ``fx_given`` (complex64 ``[nframes, nfft/2]``): skip the FFT and run everything after
    PVAnalysis.py:169 on these spectra, with ``famp = sqrt(float64(re*re + im*im))`` on the
    float32 power exactly as the CUDA kernel forms it (the kernel compares powers, which is
    order-isomorphic to comparing this famp) -- used to check the kernel's
    integer / per-peak logic bit-for-bit on the kernel's own spectrum.  ``old0`` replaces
    the all-zero spectrum before frame 0 (segment warm-up).
    """
    x = np.array(x, dtype=np.float64)
    nsamp = len(x)
    nfft2 = int(nfft / 2)
    if hop is None:
        hop = int(nfft / 2)
    tb = pv_tables(sr, nfft, hop, wind)
    win, wfact, fstep = tb["win"], tb["wfact"], tb["fstep"]
    old = np.zeros(nfft2) if old0 is None else np.asarray(old0, dtype=np.complex128)   # :121
    nfr = n_frames(nsamp, nfft, hop) if fx_given is None else len(fx_given)
    K = npks
    out = {k: np.zeros((nfr, K)) for k in ("f", "mag", "ph", "realph", "binno")}
    npk = np.zeros(nfr, dtype=np.int32)
    totalmag, t = [], []
    fxs = np.zeros((nfr, nfft2), dtype=np.complex128) if spectra else None
    marg = np.zeros(nfr) if margins else None
    with np.errstate(all="ignore"):
        for j in range(nfr):
            pos = j * hop
            if fx_given is not None:
                g = np.asarray(fx_given[j], dtype=np.complex64)
                fx = g.astype(np.complex128)
                pw = g.real * g.real + g.imag * g.imag     # float32 ops, as the kernel forms |fx|^2
                famp = np.sqrt(pw.astype(np.float64))
            else:
                xw = x[pos:pos + nfft] * win               # :155-156
                if fft_dtype == np.float32:
                    spec = np.fft.fft(xw.astype(np.float32)).astype(np.complex128)
                else:
                    spec = np.fft.fft(xw)
                fx = (spec / wfact)[:nfft2]                # :157,169
                famp = abs(fx)                             # :173
            frat = fx / old                                # :171
            bins = peak_pick(famp, K, pkthresh)            # :175-178
            if margins:
                marg[j] = peak_margin(famp, K, pkthresh)
            if len(bins):
                thisph = np.angle(fx[bins])                # :188
                dph = np.angle(frat[bins])                 # :190
                freq, df = dphase2freq(dph, bins, tb)      # :191
                ok = freq > 0.0                            # :193
                b = bins[ok]
                lo = np.maximum(b - 1, 1)                  # :197-199, left-to-right sum
                m2 = np.where(lo < b, famp[lo] ** 2, 0.0)
                m2 = np.where(lo < b, m2 + famp[b] ** 2, famp[b] ** 2)
                hi_ok = (b + 1) < len(famp)
                m2 = np.where(hi_ok, m2 + famp[np.minimum(b + 1, len(famp) - 1)] ** 2, m2)
                n = len(b)
                out["f"][j, :n] = freq[ok]
                out["mag"][j, :n] = np.sqrt(m2)
                out["ph"][j, :n] = thisph[ok]
                out["realph"][j, :n] = thisph[ok] + np.pi * df[ok] / fstep   # :207
                out["binno"][j, :n] = b
                npk[j] = n
            old = fx                                       # :209
            totalmag.append(np.sqrt(np.sum(famp ** 2)))    # :210
            t.append((pos + nfft / 2.0) / sr)              # :247
            if spectra:
                fxs[j] = fx
    out.update(t=np.array(t), totalmag=totalmag, nframes=nfr, npk=npk,
               nfft=nfft, hop=hop, sr=sr, tables=tb)
    if nfr == 0:                                           # np.array([]) in the reference
        for k in ("f", "mag", "ph", "realph", "binno"):
            out[k] = np.array([])
    if spectra:
        out["fx"] = fxs
    if margins:
        out["margin"] = marg
    return out


# --------------------------------------------------------------------------- f0-guided analysis
def harmonic_bins(f0bin, nfft2):
    """``np.round(np.arange(f0bin, nfft2 - 1, f0bin)).astype('int')`` (PVAnalysis.py:462)."""
    return np.round(np.arange(f0bin, nfft2 - 1, f0bin)).astype('int')


def analyze_harmonic(x, sr, f0, nfft=1024, hop=None, npks=20, wind=np.hanning, fmin=30.0,
                     fx_given=None):
    """``PVHarmonic(...).run_pv()`` (PVAnalysis.py:419-538) after ``set_f0(f0)`` (:424-441,
    ``t=None``: one f0 value per frame).

    Per frame with ``f0[j] > 0`` and not NaN (:509): bins at multiples of ``f0bin =
    f0/sr*nfft`` (:461-462), harmonics 2.. re-centred on the measured first harmonic when it
    exceeds ``fmin`` (:464-468), ``dphase2freq`` (:474), 3-bin magnitude (:479-483), phase;
    ``residual = sqrt(sum(famp**2) - sum of all harmonics' 3-bin powers)`` (:490).  The
    previous spectrum ``oldfft`` is only replaced by processed frames (:491), rows are cut /
    zero padded to ``npks`` (:512-516), skipped frames give zero rows and a NaN residual.
    ``fx_given``: as in :func:`analyze` (spectra of ALL frames; skipped ones are ignored).
    Returns ``f mag ph`` ``[nframes, npks]``, ``residuals``, ``t``, ``nframes``, ``nharm`` and
    (not a reference attribute, used to scale tolerances) ``totalmag = sqrt(sum(famp**2))``.
    """
    x = np.array(x, dtype=np.float64)
    nfft2 = int(nfft / 2)
    if hop is None:
        hop = int(nfft / 2)
    tb = pv_tables(sr, nfft, hop, wind)
    win, wfact = tb["win"], tb["wfact"]
    old = np.zeros(nfft2)
    nfr = n_frames(len(x), nfft, hop) if fx_given is None else len(fx_given)
    K = npks
    out = {k: np.zeros((nfr, K)) for k in ("f", "mag", "ph")}
    res = np.full(nfr, np.nan)
    tot = np.full(nfr, np.nan)
    nharm = np.zeros(nfr, dtype=np.int64)
    t = []
    with np.errstate(all="ignore"):
        for j in range(nfr):
            pos = j * hop
            thisf = f0[int(pos / hop)]                                    # :506
            t.append((pos + nfft / 2.0) / sr)                             # :523
            if not (thisf > 0 and not np.isnan(thisf)):                   # :509
                continue
            if fx_given is not None:
                g = np.asarray(fx_given[j], dtype=np.complex64)
                fx = g.astype(np.complex128)
                famp = np.sqrt((g.real * g.real + g.imag * g.imag).astype(np.float64))
            else:
                fx = (np.fft.fft(x[pos:pos + nfft] * win) / wfact)[:nfft2]    # :150-158,452
                famp = abs(fx)                                            # :456
            frat = fx / old                                               # :454
            f0bin = thisf / sr * nfft                                     # :461
            bins = harmonic_bins(f0bin, nfft2)
            ff, mm, pp = [], [], []
            cummagsq = 0
            for ipk, nbin in enumerate(bins):
                if ipk > 0:
                    if ff[0] > fmin:                                      # :465
                        corrbin = ff[0] / sr * nfft * (ipk + 1)           # :466
                        if corrbin < nfft2 - 1:
                            nbin = int(round(corrbin))                    # :468
                thisph = np.angle(fx[nbin])                               # :471
                dph = np.angle(frat[nbin])                                # :473
                fr_, _ = dphase2freq(np.array([dph]), np.array([nbin]), tb)   # :474
                ff.append(fr_[0])
                imin = max(nbin - 1, 1)                                   # :479
                imax = min(nbin + 1, len(famp))                           # :480
                thismagsq = sum(famp[imin:imax + 1] ** 2)                 # :481
                cummagsq += thismagsq
                mm.append(np.sqrt(thismagsq))
                pp.append(thisph)
            res[j] = np.sqrt(np.sum(famp ** 2) - cummagsq)                # :490
            tot[j] = np.sqrt(np.sum(famp ** 2))
            old = fx                                                      # :491
            nh = min(len(ff), K)                                          # :512
            out["f"][j, :nh] = ff[:nh]
            out["mag"][j, :nh] = mm[:nh]
            out["ph"][j, :nh] = pp[:nh]
            nharm[j] = len(ff)
    out.update(residuals=res, t=np.array(t), nframes=nfr, nharm=nharm, nfft=nfft, hop=hop, sr=sr,
               totalmag=tot)
    return out


# --------------------------------------------------------------------------- consumers
def calc_f0(f, mag, fmin=50, fmax=10000, thr=0.1):
    """``PV.calc_f0`` (PVAnalysis.py:371-391): per frame the lowest-frequency peak with
    ``fmin < f < fmax`` and ``mag > max(mag)*thr``.  Returns ``(fm, fundamental_idx)``."""
    f = np.asarray(f, dtype=np.float64)
    mag = np.asarray(mag, dtype=np.float64)
    fm = np.zeros(f.shape[0])
    im = np.zeros(f.shape[0], dtype='i')
    for ii in range(len(fm)):
        ff, mm = f[ii, :], mag[ii, :]
        maxmag = np.max(mm)
        in0 = np.flatnonzero(np.all((ff > fmin, ff < fmax, mm > maxmag * thr), axis=0))
        if len(in0) > 0:
            isel = np.argmin(ff[in0])                                     # :385
            fm[ii] = ff[in0][isel]
            im[ii] = in0[isel]
    return fm, im


def partial_sum_magnitude(mag):
    """``PV.partial_sum_magnitude`` (PVAnalysis.py:411-413)."""
    return np.sqrt(np.sum(np.asarray(mag, dtype=np.float64) ** 2, axis=1))


def calc_harmonic_power(f, mag, f_threshold=0.01):
    """PV.calc_harmonic_power (PVAnalysis.py:266-297) restated, INCLUDING the behaviour of :278:
    ``valid_mag = self.mag[valid_idx]`` indexes rows (frames) of the mag table with the column
    numbers of the frame's valid peaks, so ``valid_mag[harmonic_comp]**2`` summed (:286) is the sum
    of squares of whole table rows.  Returns (hpower, nharmonics) float64 [F, K]; raises IndexError
    like the reference when a valid peak sits in a column >= F."""
    f, mag = np.asarray(f), np.asarray(mag)
    hpower, nharm = np.zeros(f.shape), np.zeros(f.shape)
    for nfr in range(f.shape[0]):
        valid_idx = np.flatnonzero(f[nfr] > 0)                               # :276
        valid_f = f[nfr, valid_idx]
        valid_rows = mag[valid_idx]                                          # :278 (rows, not columns)
        for c, fv in zip(valid_idx, valid_f):
            nbr = np.round(valid_f / fv)                                     # :282
            nbr[nbr == 0] = 1
            comp = np.flatnonzero(np.abs(valid_f / nbr / fv - 1) < f_threshold)   # :284-285
            hpower[nfr, c] = np.sum(valid_rows[comp] ** 2)                   # :286
            nharm[nfr, c] = len(comp)
    return hpower, nharm


def refine_peaks(y, bins):
    """``PeakFinder.refine`` (PeakFinder.py:331-372, ``fun=None``, default ``x = arange``) for
    the peaks at integer ``bins`` of ``y``: parabola through the three samples around a
    peak -> (fine position, fine value); a non-peak returns (pos, y[pos])."""
    y = np.asarray(y, dtype=np.float64)
    fpos = np.zeros(len(bins))
    fval = np.zeros(len(bins))
    with np.errstate(all="ignore"):
        for q, pos in enumerate(bins):
            sur = y[pos - 1:pos + 2]
            if sur[1] > sur[0] and sur[1] >= sur[2]:                      # :354
                c = sur[1]
                b = (sur[2] - sur[0]) / 2
                a = (sur[2] + sur[0]) / 2 - c
                lpos = - b / 2 / a
                fpos[q] = float(pos) + lpos
                fval[q] = a * lpos * lpos + b * lpos + c                  # :365
            else:
                fpos[q] = pos
                fval[q] = sur[1]
    return fpos, fval


# --------------------------------------------------------------------------- tracking
def track(f, mag, maxpitchjmp=0.5):
    """``PV.toSinSum`` -> ``SinSum.add_frame`` (PVAnalysis.py:299-322,871-957), restated as
    a frame-pair-local greedy link plus id numbering.

    Returns ``tid`` int32 ``[F, K]`` (track id per peak slot, -1 = not a point),
    ``link`` int32 ``[F, K]`` (column of the continued peak in frame j-1, -1 = new /
    none), ``st``, ``end`` (first / last frame per track, :827-828,950).
    Track ids are numbered in ``add_empty_partial`` call order (:819-830).

    Exact ties: see the comment at the argsort below (current frame) -- previous-frame ties
    keep the reference's deterministic (magnitude, track index) descending order (:893).
    """
    f = np.asarray(f, dtype=np.float64)
    mag = np.asarray(mag, dtype=np.float64)
    F = f.shape[0] if f.ndim == 2 else 0
    K = f.shape[1] if f.ndim == 2 else 0
    tid = -np.ones((F, K), dtype=np.int32)
    link = -np.ones((F, K), dtype=np.int32)
    st, end = [], []
    for fr in range(F):
        # :874-875.  The reference calls np.argsort(mag)[::-1] with numpy's default *unstable*
        # sort, so the order of peaks whose magnitudes are exactly equal (e.g. the 1e-17 hash
        # of a frame leaving digital silence) depends on the numpy build / CPU SIMD dispatch.
        # The oracle pins that one undefined case to "reversed stable sort" (= what numpy's
        # insertion sort gives for <= 16 entries): among exact ties the higher column first.
        idx = np.argsort(mag[fr], kind="stable")[::-1]
        idx = idx[np.logical_and(f[fr][idx] > 0, mag[fr][idx] > 0)]       # :876
        if fr > 0:
            pcols = np.flatnonzero(tid[fr - 1] >= 0)                      # partials ending at fr-1 (:887,984-994)
        else:
            pcols = np.zeros(0, dtype=int)
        if len(pcols):
            # sorted(zip(pmag, pidx), reverse=True) (:893): by magnitude, then track index, descending
            keys = sorted(zip(mag[fr - 1][pcols].tolist(), tid[fr - 1][pcols].tolist(),
                              pcols.tolist()), reverse=True)
            pc = np.array([k[2] for k in keys])
            pf = f[fr - 1][pc]
            unused = np.ones(len(pc), dtype=bool)
        for c in idx:
            fc = float(f[fr, c])
            matched = -1
            if len(pcols) and unused.any():
                u = np.flatnonzero(unused)
                stonediff = abs(17.312 * (fc / pf[u] - 1.0))              # dpitch2st :62-68, :914
                nearest = np.argmin(stonediff)                            # :920
                if stonediff[nearest] < maxpitchjmp:                      # :923
                    matched = u[nearest]
            if matched >= 0:
                unused[matched] = False
                link[fr, c] = pc[matched]
                tid[fr, c] = tid[fr - 1, pc[matched]]
                end[tid[fr, c]] = fr                                      # :950
            else:
                tid[fr, c] = len(st)                                      # add_empty_partial :819-830
                st.append(fr)
                end.append(fr)
    return dict(tid=tid, link=link, st=np.array(st, dtype=np.int64),
                end=np.array(end, dtype=np.int64))


def partials_from_tracks(tr, f, mag, ph, realph):
    """Per-track value lists (= ``ss.partial[i].f/mag/ph/realph``, RegPartial.append_point
    :616-626) gathered from the frame tables."""
    tid = tr["tid"]
    n = len(tr["st"])
    parts = []
    fr_idx, col_idx = np.nonzero(tid >= 0)
    ids = tid[fr_idx, col_idx]
    order = np.lexsort((fr_idx, ids))
    fr_idx, col_idx, ids = fr_idx[order], col_idx[order], ids[order]
    bounds = np.searchsorted(ids, np.arange(n + 1))
    for i in range(n):
        a, b = bounds[i], bounds[i + 1]
        r, c = fr_idx[a:b], col_idx[a:b]
        parts.append(dict(start_idx=int(tr["st"][i]), f=f[r, c], mag=mag[r, c],
                          ph=ph[r, c], realph=realph[r, c]))
    return parts


# --------------------------------------------------------------------------- resynthesis
def synth_partial(pf, pmag, prealph, sr, hop, overlap, fstep, edge=1.0):
    """``RegPartial.synth`` (PVAnalysis.py:684-756) with all ``nfr`` blocks evaluated at once
    as rows of a ``[nfr, hop]`` array (blocks are independent: the cumsum of :707 never
    crosses a block).  Returns ``(signal, start_sample)`` like the reference."""
    pf = np.asarray(pf, dtype=np.float64)
    pmag = np.asarray(pmag, dtype=np.float64)
    prealph = np.asarray(prealph, dtype=np.float64)
    nfr = len(pf)
    hop = int(hop)
    dfr = 1. / overlap / 2.                                               # :687
    newt = np.arange(hop * (nfr + dfr))                                   # :689
    fsig = np.interp(newt, hop * (dfr + .5 + np.arange(nfr)), pf)         # :701
    msig = np.interp(newt, hop * (dfr + np.arange(nfr)), pmag)            # :702
    rows = np.arange(nfr)
    ph = np.zeros((nfr, hop))
    if hop > 1:
        fs = fsig[:hop * nfr].reshape(nfr, hop)[:, :hop - 1]
        ph[:, 1:] = pi2 * np.cumsum(fs / float(sr), axis=1)               # :705-708
    b0, b1 = fsig[hop * rows], fsig[hop * (rows + 1)]
    phcor = np.pi * (b1 - b0) / fstep / 2.                                # :715
    ph += (prealph + phcor)[:, None]                                      # :721-722
    if nfr > 1:
        b2 = fsig[hop * (rows[:-1] + 2)]
        phcornext = np.pi * (b2 - b1[:-1]) / fstep / 2.                   # :717-718
        phend = ph[:-1, -1] + pi2 * b1[:-1] / float(sr)                   # :726
        dph = np.mod(prealph[1:] + phcornext - phend + np.pi, pi2) - np.pi   # :727-728
        ph[:-1] += np.linspace(0.0, dph, num=hop + 1)[:-1].T              # :729
    body = msig[:hop * nfr] * np.cos(ph.ravel())                          # :734-736
    edgsam = int(dfr * hop * edge)                                        # :740
    q = np.arange(edgsam)
    hmag = msig[0] * (1 - np.cos(np.pi * q / float(edgsam))) / 2.         # :742
    hph = np.flipud(prealph[0] - pi2 * np.cumsum(pf[0] * np.ones(edgsam) / float(sr)))   # :743-744
    tmag = msig[hop * nfr] * (1 + np.cos(np.pi * q / float(edgsam))) / 2.  # :748-749
    tph = ph[-1, -1] + pi2 * np.cumsum(pf[-1] * np.ones(edgsam) / float(sr))            # :750
    sig = np.concatenate((hmag * np.cos(hph), body, tmag * np.cos(tph)))  # :745,751
    return sig, edgsam


def synth(parts, sr, hop, nfft, hop_an, edge=1.0, minframes=3, max_end=None):
    """``SinSum.synth`` overlap-add (PVAnalysis.py:1053-1070) over ``parts`` (list of dicts
    from :func:`partials_from_tracks`), integer ``hop`` (the py3 shim)."""
    hop = int(hop)
    dfr = nfft / hop_an / 2.                                              # :1055
    edgsamp = int(edge * hop * dfr)                                       # :1056
    if max_end is None:
        max_end = max(p["start_idx"] + len(p["f"]) - 1 for p in parts)
    w = np.zeros((max_end + 2) * hop + 2 * edgsamp)                       # :1059
    overlap = hop_an / float(nfft)                                        # :824
    fstep = sr / float(nfft)                                              # :825
    for p in parts:
        if len(p["f"]) >= minframes:                                      # :1061
            wi, e = synth_partial(p["f"], p["mag"], p["realph"], sr, hop, overlap, fstep, edge)
            s = int(p["start_idx"] * hop - e) + edgsamp                   # :756,1066
            if s >= 0:
                w[s:s + len(wi)] += wi                                    # :1067-1069
    return w[edgsamp:]                                                    # :1070


def partial_samples(parts, hop, nfft, hop_an, edge=1.0, minframes=3):
    """Work units of resynthesis: sum over rendered tracks of (hop*nfr + 2*edgsam)."""
    dfr = nfft / hop_an / 2.
    e = int(dfr * hop * edge)
    return sum(hop * len(p["f"]) + 2 * e for p in parts if len(p["f"]) >= minframes)
