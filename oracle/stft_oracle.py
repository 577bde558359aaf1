"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's frame-wise spectral consumers
(SURVEY 8f row 4): FFTFilters.FilterBank.specout and its filter construction, SoundUtils.RMSWind
and SoundUtils.SpecFlux.  Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import
this module; the product (pypevoc_b200/stft.py -> libpvk.so) never does.

Parity status: PINNED -- tests/test_oracle_golden.py checks every function here against
tests/golden/stft.npz, written by oracle/gen_golden_stft.py from the unmodified reference
(/root/reference/pypevoc/FFTFilters.py, SoundUtils.py) in the build container.  All file:line
references are relative to /root/reference/pypevoc/.
"""
import numpy as np


# ------------------------------------------------------------------ filter specifications
def spec_bands(mode, freq, gain=None, sr=1.0):
    """bandf [nb, 2] (fractions of sr), bandg [nb, 2] of one PiecewiseFilterSpec
    (FFTFilters.py:97-168).  Presets take ``freq`` already divided by sr (:119-126)."""
    m = mode.lower()
    if m in ("lp", "lowpass"):                                    # :136-139
        f = freq / float(sr)
        return np.array([[0.0, f], [f, 0.5]]), np.array([[1.0, 1.0], [0.0, 0.0]])
    if m in ("hp", "hipass", "highpass"):                         # :141-144
        f = freq / float(sr)
        return np.array([[0.0, f], [f, 0.5]]), np.array([[0.0, 0.0], [1.0, 1.0]])
    if m in ("bp", "bandpass"):                                   # :146-149
        f1, f2 = freq[0] / float(sr), freq[-1] / float(sr)
        return np.array([[0.0, f1], [f1, f2], [f2, 0.5]]), np.array([[0.0, 0.0], [1.0, 1.0], [0.0, 0.0]])
    if m in ("bs", "bandstop"):                                   # :151-154
        f1, f2 = freq[0] / float(sr), freq[-1] / float(sr)
        return np.array([[0.0, f1], [f1, f2], [f2, 0.5]]), np.array([[1.0, 1.0], [0.0, 0.0], [1.0, 1.0]])
    # vertex list (:157-167): consecutive vertices in ascending frequency order; the dtype of
    # ``freq`` is kept (TriangularFilterBank hands float32 vertices in, :328)
    idx = np.argsort(freq)
    bands, gains = [], []
    for a, b in zip(idx[:-1], idx[1:]):
        bands.append([freq[a] / sr, freq[b] / sr])
        gains.append([gain[a], gain[b]])
    return np.array(bands), np.array(gains)


def filter_mask(bandf, bandg, sr, fvec, align_edges=True):
    """PiecewiseFilterSpec.apply_to_freq_vector (FFTFilters.py:200-229): gain at every frequency
    of ``fvec``; band edges snapped to the nearest ``fvec`` entry when ``align_edges``; later
    bands overwrite earlier ones on shared edges; a band of zero width raises."""
    fvec = np.array(fvec)
    freqs = bandf * sr
    edges = np.unique(np.array(bandf).flatten() * sr)             # :193-198
    snap = {}
    for ff in edges:
        snap[ff] = fvec[np.argmin(np.abs(fvec - ff))] if align_edges else ff
    mask = np.zeros(len(fvec))
    for f, g in zip(freqs, bandg):
        fst, fend = snap[f[0]], snap[f[1]]
        idx = np.logical_and(fvec >= fst, fvec <= fend)
        if fend == fst:
            raise ValueError("Band is too narrow: try increasing nwind")   # BandError :226
        mask[idx] = (fvec[idx] - fst) / (fend - fst) * (g[1] - g[0]) + g[0]
    return mask


def bank_matrix(specs, sr, nwind, align_edges=True):
    """FilterBank.__init__ (FFTFilters.py:244-272): fb [nfilt, nwind] on fvec = linspace(0, sr,
    nwind) (the endpoint sr is included: bin k sits at k*sr/(nwind-1), as the reference has it).
    ``specs``: list of (bandf, bandg)."""
    fvec = np.linspace(0., sr, nwind)
    fb = np.zeros((len(specs), len(fvec)))
    for i, (bf, bg) in enumerate(specs):
        fb[i, :] = filter_mask(bf, bg, sr, fvec, align_edges)
    return fb


def triangular_specs(flim, sr=1.):
    """TriangularFilterBank.__init__ (FFTFilters.py:309-334): one triangle per interior limit,
    vertices in float32 (:328)."""
    flim = np.sort(flim).astype('f')
    out = []
    for n in range(len(flim) - 2):
        out.append(spec_bands('', flim[n:n + 3], np.array([0.0, 1.0, 0.0]), sr))
    return out


def mel_limits(n=26, fmin=300., fmax=8000.):
    """MelFilterBank.__init__ (FFTFilters.py:343-350) with the reference's own mel map
    mel = 1125 + ln(1 + f/700) (:61-66, a sum, not the textbook product)."""
    melmin = 1125. + np.log(1. + fmin / 700.)
    melmax = 1125. + np.log(1. + fmax / 700.)
    return 700. * (np.exp(np.linspace(melmin, melmax, n + 2) - 1125.) - 1)


def mel_geometry(twind=.025, sr=44100., thop=.01):
    return int(2 ** np.round(np.log2(twind * sr))), int(thop * sr)          # :344-345


# ------------------------------------------------------------------ frame-wise consumers
def specout(w, fb, wind, hop, sr):
    """FilterBank.specout (FFTFilters.py:274-292): bankout [F, nfilt] = sum_k |FFT(frame*wind)|^2
    * fb[i, k] over ALL nwind bins, frames while n < len(w) - nwind; t = (n + nwind/2)/sr."""
    nwind = len(wind)
    out, t = [], []
    n = 0
    while n < len(w) - nwind:
        Sww = np.abs(np.fft.fft(w[n:n + nwind] * wind)) ** 2
        out.append([sum(Sww * fb[i, :]) for i in range(fb.shape[0])])
        t.append((float(n) + nwind / 2.) / float(sr))
        n += hop
    return np.array(out), np.array(t)


def rms_wind(x, sr=1, nwind=1024, nhop=512, windfunc=np.blackman):
    """SoundUtils.RMSWind (SoundUtils.py:74-103)."""
    wind = windfunc(nwind)
    wsum2 = np.sum(wind ** 2)
    ret, t = [], []
    ist, iend = 0, nwind
    while iend < len(x):
        xw = x[ist:iend] * wind
        ret.append(np.sum(xw * xw / wsum2))
        t.append(float(ist + iend) / 2.0 / float(sr))
        ist += nhop
        iend = ist + nwind
    return np.sqrt(np.array(ret)), np.array(t)


def spec_flux(x, sr=1, nwind=1024, nhop=512, minf=0, maxf=np.inf, windfunc=np.blackman):
    """SoundUtils.SpecFlux (SoundUtils.py:196-231): Euclidean distance between the magnitude
    spectra of frames hop apart, over bins [minbin, maxbin) of the full nwind-point FFT."""
    wind = windfunc(nwind)
    minbin = int(minf / sr * nwind)
    maxbinf = float(maxf) / sr * nwind
    maxbin = nwind if maxbinf > nwind else int(maxbinf)
    res, t = [], []
    ist, iend = 0, nwind
    while iend < len(x) - nhop:
        ff = np.abs(np.fft.fft(x[ist:iend] * wind))
        fl = np.abs(np.fft.fft(x[ist + nhop:iend + nhop] * wind))
        res.append(np.sqrt(sum((ff[minbin:maxbin] - fl[minbin:maxbin]) ** 2)))
        t.append(float(ist + iend + nhop) / 2.0 / float(sr))
        ist += nhop
        iend = ist + nwind
    return np.array(res), np.array(t)
