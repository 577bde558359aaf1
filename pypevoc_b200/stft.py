"""Frame-wise spectral consumers on the analysis front end.

Host-side mirrors of the reference's other frame loops, backed by ``pvk_stft_bank`` (the framing +
window + FFT of the PV kernel, include/pvk.h):

  * ``PiecewiseFilterSpec``, ``FilterBank``, ``TriangularFilterBank``, ``MelFilterBank`` with
    ``specout`` / ``mfcc`` / ``mfcc_and_mel``          (pypevoc/FFTFilters.py:88-374)
  * ``RMSWind`` and ``SpecFlux``                         (pypevoc/SoundUtils.py:74-103, 196-231)

Same names, arguments and return values (float64 numpy arrays).  Building the filter matrix is
host logic (numpy, once per bank); every per-frame loop runs on the GPU -- there is no CPU
fallback.  Differences from the reference: the window length must be a power of two in
[64, 8192] (``MelFilterBank`` always makes one), the signal is rounded to float32 and the FFT is
fp32 (sums in fp64).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .pv import _analysis_tables, _device, _ptr, _stream, n_frames


class BandError(Exception):
    """A filter band narrower than the frequency grid (FFTFilters.py:27-37)."""

    def __init__(self, message):
        self.message = message
        Exception.__init__(self, message)


def _f_to_mel_py(freq):
    # the reference's own map (FFTFilters.py:61-63): a SUM, kept as is for drop-in results
    return 1125. + np.log(1. + freq / 700.)


def _mel_to_f_py(mel):
    return 700. * (np.exp(mel - 1125.) - 1)                      # FFTFilters.py:65-66


f_to_mel = np.vectorize(_f_to_mel_py)
mel_to_f = np.vectorize(_mel_to_f_py)


def nextpow2(x):
    return 2 ** (np.ceil(np.log2(x)))                            # FFTFilters.py:337-338


class PiecewiseFilterSpec(object):
    """Piecewise-linear gain over frequency (FFTFilters.py:88-229): ``bandf`` [nb, 2] band edges as
    fractions of ``sr``, ``bandg`` [nb, 2] the gains at those edges."""

    def __init__(self, mode='', cutoff=0.5, freq=np.array([0.0, 0.5]), gain=np.array([1.0, 1.0]), sr=1.0, label=''):
        self.sr = sr
        self.label = ''
        self.bandf = np.array([0.0, 0.5])
        self.bandg = np.array([1.0, 1.0])
        m = mode.lower()
        if m in ('lp', 'lowpass'):
            self.set_lowpass_cutoff(freq / float(sr))
        elif m in ('hp', 'hipass', 'highpass'):
            self.set_hipass_cutoff(freq / float(sr))
        elif m in ('bp', 'bandpass'):
            self.set_bandpass_freqs(freq[0] / float(sr), freq[-1] / float(sr))
        elif m in ('bs', 'bandstop'):
            self.set_bandstop_freqs(freq[0] / float(sr), freq[-1] / float(sr))
        else:
            assert len(freq) == len(gain)
            self.set_triangular_filter(freq, gain)
            self.label = label
        if not self.label:
            self.label = 'Piecewise filter with {} bands'.format(len(self.bandf) - 1)

    def _set(self, edges, gains, label):
        self.bandf = np.array([[a, b] for a, b in zip(edges[:-1], edges[1:])])
        self.bandg = np.array(gains)
        self.label = label

    def set_lowpass_cutoff(self, f):
        self._set([0.0, f, 0.5], [[1.0, 1.0], [0.0, 0.0]], 'Lowpass filter, fc={}'.format(f * self.sr))

    def set_hipass_cutoff(self, f):
        self._set([0.0, f, 0.5], [[0.0, 0.0], [1.0, 1.0]], 'Hipass filter, fc={}'.format(f * self.sr))

    def set_bandpass_freqs(self, f1, f2):
        self._set([0.0, f1, f2, 0.5], [[0.0, 0.0], [1.0, 1.0], [0.0, 0.0]],
                  'Bandpass filter, fc={}'.format((f1 / 2 + f2 / 2) * self.sr))

    def set_bandstop_freqs(self, f1, f2):
        self._set([0.0, f1, f2, 0.5], [[1.0, 1.0], [0.0, 0.0], [1.0, 1.0]],
                  'Bandstop filter, fc={}'.format((f1 / 2 + f2 / 2) * self.sr))

    def set_triangular_filter(self, freq, gain):
        """Vertex list -> one band per pair of neighbouring vertices (ascending frequency); the
        dtype of ``freq`` carries through, as in the reference (:157-167)."""
        order = np.argsort(freq)
        lo, hi = order[:-1], order[1:]
        self.bandf = np.array([[freq[a] / self.sr, freq[b] / self.sr] for a, b in zip(lo, hi)])
        self.bandg = np.array([[gain[a], gain[b]] for a, b in zip(lo, hi)])

    def __repr__(self):
        rep = '{}:\n'.format(self.label)
        for f, g in zip(self.bandf, self.bandg):
            span = '[{},{}]'.format(f[0] * self.sr, f[1] * self.sr)
            rep += ('  Freq = {}: gain = {}\n'.format(span, g[0]) if g[0] == g[1]
                    else '  Freq = {}: gain = [{},{}]\n'.format(span, g[0], g[1]))
        return rep

    def get_frequency_gains(self):
        return np.array(self.bandf) * self.sr, np.array(self.bandg)

    def get_frequency_edges(self):
        return np.unique(np.array(self.bandf).flatten() * self.sr)

    def apply_to_freq_vector(self, fvec, align_edges=False):
        """Gain at every frequency of ``fvec`` (:200-229); with ``align_edges`` every band edge
        moves to the nearest entry of ``fvec`` first.  Bands are applied in order (a later band
        wins on a shared edge); a band that collapses to zero width raises BandError."""
        fvec = np.array(fvec)
        edges = self.get_frequency_edges()
        where = {ff: (fvec[np.argmin(np.abs(fvec - ff))] if align_edges else ff) for ff in edges}
        mask = np.zeros(len(fvec))
        for f, g in zip(self.bandf * self.sr, self.bandg):
            a, b = where[f[0]], where[f[1]]
            if b == a:
                raise BandError('Band is too narrow: try increasing nwind')
            inside = np.logical_and(fvec >= a, fvec <= b)
            mask[inside] = (fvec[inside] - a) / (b - a) * (g[1] - g[0]) + g[0]
        return mask


# --------------------------------------------------------------------------- device plumbing
def _signal_device(w, dev):
    """1-D float32 CUDA tensor of the signal (numpy / torch, host or device)."""
    if isinstance(w, torch.Tensor):
        t = w.detach()
        if t.dim() != 1:
            raise ValueError("signal must be one-dimensional")
        return t.to(device=dev, dtype=torch.float32).contiguous()
    a = np.asarray(w)
    if a.ndim != 1:
        raise ValueError("signal must be one-dimensional")
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)


def _check_nwind(nwind):
    n = int(nwind)
    if n != nwind or n < 64 or n > 8192 or (n & (n - 1)):
        raise ValueError("nwind=%r is not supported by the CUDA kernels: it must be a power of two in [64, 8192]"
                         % (nwind,))
    return n


def stft_bank_device(xd, win, nwind, hop, nframes, fb_folded=None, fb_lo=None, fb_hi=None, flux_bins=None,
                     inv_wsum2=None, run_frames=0):
    """Launch pvk_stft_bank on a device signal; returns dict(bank=[F, nfilt], flux=[F-1], rms=[F])
    with the requested float64 CUDA tensors (asynchronous).  ``win``: float32 CUDA window."""
    L = _lib.lib()
    dev = xd.device
    tab = _analysis_tables(nwind, dev)
    out = {}
    nfilt = 0
    if fb_folded is not None:
        nfilt = int(fb_folded.shape[0])
        out["bank"] = torch.empty((nframes, nfilt), dtype=torch.float64, device=dev)
    if flux_bins is not None:
        out["flux"] = torch.empty((max(nframes - 1, 0),), dtype=torch.float64, device=dev)
    if inv_wsum2 is not None:
        out["rms"] = torch.empty((nframes,), dtype=torch.float64, device=dev)
    if nframes <= 0:
        return out
    lo, hi = (int(flux_bins[0]), int(flux_bins[1])) if flux_bins is not None else (0, 0)
    with torch.cuda.device(dev):
        _lib.check(L.pvk_stft_bank(_ptr(xd), xd.numel(), _ptr(win), _ptr(tab), nwind, int(hop), int(nframes),
                                   int(run_frames), _ptr(fb_folded), _ptr(fb_lo), _ptr(fb_hi), nfilt,
                                   _ptr(out.get("bank")), lo, hi, _ptr(out.get("flux")),
                                   float(inv_wsum2 if inv_wsum2 is not None else 0.0), _ptr(out.get("rms")),
                                   _stream()), "pvk_stft_bank")
    return out


def fold_bank(fb):
    """Fold a filter matrix [nfilt, nwind] (weights on all FFT bins) onto bins 0..nwind/2 of a real
    frame's spectrum: |X[nwind-h]| = |X[h]|.  Returns (folded float64 [nfilt, nwind/2+1], lo, hi)
    with the support [lo, hi) of every row."""
    fb = np.asarray(fb, dtype=np.float64)
    nf, n = fb.shape
    m = n // 2
    fold = fb[:, :m + 1].copy()
    fold[:, 1:m] += fb[:, :m:-1]                                   # columns n-1 ... m+1 -> 1 ... m-1
    lo = np.zeros(nf, dtype=np.int32)
    hi = np.zeros(nf, dtype=np.int32)
    for i in range(nf):
        nz = np.flatnonzero(fold[i])
        if len(nz):
            lo[i], hi[i] = nz[0], nz[-1] + 1
    return fold, lo, hi


# --------------------------------------------------------------------------- filter banks
class FilterBank(object):
    """FFT-based filter bank (FFTFilters.py:235-298).  ``fb`` [nfilt, nwind] holds the gain of
    every filter on ``fvec = linspace(0, sr, nwind)``; ``specout`` runs on the GPU."""

    def __init__(self, fspec_list=None, sr=1.0, nwind=256, windfunc=np.hanning, nhop=None, align_edges=True,
                 device=None):
        self.sr = sr
        self.nwind = int(nwind)
        self.wind = windfunc(nwind)
        self.hop = nhop if nhop else int(nwind / 2)
        self.fvec = np.linspace(0., sr, nwind)
        if not fspec_list:
            fc = 0.25
            fspec_list = [PiecewiseFilterSpec(mode='lowpass', freq=fc, sr=sr),
                          PiecewiseFilterSpec(mode='hipass', freq=fc, sr=sr)]
        self.fb = np.zeros((len(fspec_list), len(self.fvec)))
        self.label = []
        for i, spec in enumerate(fspec_list):
            self.fb[i, :] = spec.apply_to_freq_vector(self.fvec, align_edges=align_edges)
            self.label.append(spec.label)
        self._device = device
        self._dev_state = None

    def _state(self):
        """Device copies of the folded filter matrix, its supports and the window (built once)."""
        if self._dev_state is None:
            _check_nwind(self.nwind)
            dev = _device(self._device)
            fold, lo, hi = fold_bank(self.fb)
            self._dev_state = dict(dev=dev, fold=torch.from_numpy(fold).to(dev), lo=torch.from_numpy(lo).to(dev),
                                   hi=torch.from_numpy(hi).to(dev),
                                   win=torch.from_numpy(np.asarray(self.wind, dtype=np.float32)).to(dev))
        return self._dev_state

    def specout_device(self, w):
        """specout with the result left on the GPU: (float64 CUDA [F, nfilt], frame count)."""
        st = self._state()
        xd = _signal_device(w, st["dev"])
        F = n_frames(xd.numel(), self.nwind, int(self.hop))      # while n < len(w) - nwind (:281)
        out = stft_bank_device(xd, st["win"], self.nwind, int(self.hop), F, st["fold"], st["lo"], st["hi"])
        return out["bank"], F

    def specout(self, w):
        """Output of the filterbank applied to ``w`` (:274-292): (bankout [F, nfilt], tout [F])."""
        bank, F = self.specout_device(w)
        tout = (np.arange(F) * float(self.hop) + self.nwind / 2.) / float(self.sr)
        return bank.cpu().numpy(), tout

    def __repr__(self):
        return 'FilterBank with filters:\n' + ''.join('  ' + ll + '\n' for ll in self.label)


class TriangularFilterBank(FilterBank):
    """Triangles between consecutive frequency limits (FFTFilters.py:300-334); the limits are
    rounded to float32 as the reference does (:328)."""

    def __init__(self, flim=[0, .5, 1.], nwind=256, sr=1., nhop=None, device=None):
        unit = 'Hz' if sr > 1.0 else ''
        flim = np.sort(flim).astype('f')
        specs = []
        for n, cc in enumerate(flim[1:-1]):
            lab = '{}{} band ({}-{}{})'.format(cc, unit, flim[n], flim[n + 2], unit)
            specs.append(PiecewiseFilterSpec(freq=flim[n:n + 3], gain=np.array([0.0, 1.0, 0.0]), label=lab, sr=sr))
        super(TriangularFilterBank, self).__init__(fspec_list=specs, nwind=nwind, sr=sr, nhop=nhop, device=device)


class MelFilterBank(TriangularFilterBank):
    """Mel-spaced triangular bank + MFCCs (FFTFilters.py:342-374)."""

    def __init__(self, n=26, fmin=300., fmax=8000., twind=.025, sr=44100., thop=.01, device=None):
        nwind = int(2 ** np.round(np.log2(twind * sr)))
        nhop = int(thop * sr)
        fc = mel_to_f(np.linspace(f_to_mel(fmin), f_to_mel(fmax), n + 2))
        super(MelFilterBank, self).__init__(flim=fc, nwind=nwind, sr=sr, nhop=nhop, device=device)

    def _cepstrum(self, logs, mode):
        if mode[:3] == 'DCT':
            from scipy.fftpack import dct
            return dct(logs, type=int(mode[3]))
        if mode == 'IFFT':
            return np.fft.ifft(logs)
        raise NotImplementedError

    def mfcc(self, w, mode='DCT2'):
        spec, tspec = self.specout(w)
        return self._cepstrum(np.log(spec), mode), tspec

    def mfcc_and_mel(self, w, mode='DCT2'):
        spec, tspec = self.specout(w)
        return self._cepstrum(np.log(spec), mode), spec, tspec


# --------------------------------------------------------------------------- SoundUtils
def RMSWind(x, sr=1, nwind=1024, nhop=512, windfunc=np.blackman, device=None, to_host=True):
    """RMS amplitude of ``x`` in windows of ``nwind`` samples every ``nhop`` (SoundUtils.py:74-103):
    (rms [F], t [F]); frames while ist + nwind < len(x)."""
    nwind = _check_nwind(nwind)
    dev = _device(device)
    xd = _signal_device(x, dev)
    wind = windfunc(nwind)
    wsum2 = np.sum(wind ** 2)
    F = n_frames(xd.numel(), nwind, int(nhop))
    win = torch.from_numpy(np.asarray(wind, dtype=np.float32)).to(dev)
    out = stft_bank_device(xd, win, nwind, int(nhop), F, inv_wsum2=1.0 / wsum2)["rms"]
    t = (np.arange(F) * float(nhop) * 2 + nwind) / 2.0 / float(sr)
    return (out.cpu().numpy() if to_host else out), t


def SpecFlux(x, sr=1, nwind=1024, nhop=512, minf=0, maxf=np.inf, windfunc=np.blackman, device=None, to_host=True):
    """Spectral flux (SoundUtils.py:196-231): distance between the magnitude spectra of windows
    ``nhop`` apart over bins [int(minf/sr*nwind), int(maxf/sr*nwind)) of the nwind-point FFT;
    (flux [F], t [F]); frames while ist + nwind < len(x) - nhop."""
    nwind = _check_nwind(nwind)
    dev = _device(device)
    xd = _signal_device(x, dev)
    wind = windfunc(nwind)
    minbin = int(minf / sr * nwind)
    maxbinf = float(maxf) / sr * nwind
    maxbin = nwind if maxbinf > nwind else int(maxbinf)
    F = n_frames(xd.numel() - int(nhop), nwind, int(nhop))       # pairs (j, j+1)
    win = torch.from_numpy(np.asarray(wind, dtype=np.float32)).to(dev)
    if F > 0:
        out = stft_bank_device(xd, win, nwind, int(nhop), F + 1, flux_bins=(max(minbin, 0), maxbin))["flux"]
    else:
        out = torch.zeros((0,), dtype=torch.float64, device=dev)
    t = (np.arange(F) * float(nhop) * 2 + nwind + nhop) / 2.0 / float(sr)
    return (out.cpu().numpy() if to_host else out), t
