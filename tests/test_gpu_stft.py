"""-m gpu: the STFT-consumer row through the product classes -> ctypes ->
libpvk.so (pvk_stft_bank): FilterBank / TriangularFilterBank / MelFilterBank.specout, mfcc, RMSWind
and SpecFlux against the goldens of the real reference and against the numpy oracle on larger
seeded inputs, edge cases and size-independent properties at full size."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import stft_oracle as so
import stft_util as su
from test_stft_cpu import _product_bank, BANKS

pytestmark = pytest.mark.gpu
G = su.golden()


@pytest.mark.parametrize("name", sorted(BANKS))
def test_specout_vs_reference_golden(name):
    fbk = _product_bank(name)
    x, _ = su.signal(BANKS[name][0])
    spec, t = fbk.specout(x)
    assert spec.dtype == np.float64
    su.close(spec, G[name + "/spec"], su.TOL_BANK)
    assert np.array_equal(t, G[name + "/t"])
    # float64 numpy input, torch CPU and CUDA tensors give the same result
    import torch
    for w in (x.astype(np.float64), torch.from_numpy(x), torch.from_numpy(x).cuda()):
        assert np.array_equal(fbk.specout(w)[0], spec)


def test_mfcc_vs_reference_pipeline():
    from scipy.fftpack import dct
    name = "mel_44k"
    fbk = _product_bank(name)
    x, _ = su.signal(BANKS[name][0])
    c, t = fbk.mfcc(x)
    ref = dct(np.log(G[name + "/spec"]), type=2)               # FFTFilters.py:352-358 on the reference's spec
    assert c.shape == ref.shape and np.max(np.abs(c - ref)) < 1e-3
    c2, spec, t2 = fbk.mfcc_and_mel(x, mode="IFFT")
    assert np.iscomplexobj(c2) and spec.shape == G[name + "/spec"].shape and np.array_equal(t, t2)
    with pytest.raises(NotImplementedError):
        fbk.mfcc(x, mode="nope")


def test_rms_and_flux_vs_reference_golden():
    from pypevoc_b200 import stft
    for name, (sig, kw) in su.CASES["rms"].items():
        x, sr = su.signal(sig)
        v, t = stft.RMSWind(x, sr=sr, **su.windfunc(kw))
        su.close(v, G[name + "/v"], su.TOL_RMS)
        assert np.array_equal(t, G[name + "/t"]), name
    for name, (sig, kw) in su.CASES["flux"].items():
        x, sr = su.signal(sig)
        kw = su.windfunc(kw)
        v, t = stft.SpecFlux(x, sr=sr, **kw)
        wind = kw.get("windfunc", np.blackman)(kw["nwind"])
        minbin = int(kw.get("minf", 0) / sr * kw["nwind"])
        mb = float(kw.get("maxf", np.inf)) / sr * kw["nwind"]
        maxbin = kw["nwind"] if mb > kw["nwind"] else int(mb)
        su.close(v, G[name + "/v"], su.TOL_FLUX, scale=su.band_norm(x, wind, kw["nhop"], minbin, maxbin))
        assert np.array_equal(t, G[name + "/t"]), name


@pytest.mark.parametrize("nwind,hop", [(64, 17), (128, 64), (1024, 300), (4096, 1024), (8192, 2048)])
def test_all_window_sizes_vs_oracle(nwind, hop):
    """Every FFT size of the kernels, odd hops: bank, RMS and flux against the numpy oracle."""
    from pypevoc_b200 import signals, stft
    sr = 22050
    x = signals.harm(sr, (nwind * 3 + hop * 9) / sr, 300, 20, 1.0, 0.1, 21).astype(np.float32)
    fbk = stft.TriangularFilterBank(flim=[0., 900., 2500., 6000., 11000.], nwind=nwind, sr=float(sr), nhop=hop)
    spec, t = fbk.specout(x)
    ref, tr = so.specout(x.astype(np.float64), fbk.fb, fbk.wind, hop, float(sr))
    su.close(spec, ref, su.TOL_BANK)
    assert np.array_equal(t, tr)
    v, t = stft.RMSWind(x, sr=sr, nwind=nwind, nhop=hop)
    ref, tr = so.rms_wind(x.astype(np.float64), sr=sr, nwind=nwind, nhop=hop)
    su.close(v, ref, su.TOL_RMS)
    assert np.array_equal(t, tr)
    v, t = stft.SpecFlux(x, sr=sr, nwind=nwind, nhop=hop, minf=100., maxf=9000.)
    ref, tr = so.spec_flux(x.astype(np.float64), sr=sr, nwind=nwind, nhop=hop, minf=100., maxf=9000.)
    su.close(v, ref, su.TOL_FLUX, scale=su.band_norm(x, np.blackman(nwind), hop, int(100. / sr * nwind), int(9000. / sr * nwind)))
    assert np.array_equal(t, tr)


def test_edge_cases():
    from pypevoc_b200 import stft
    fbk = stft.MelFilterBank(n=12, fmin=200., fmax=6000., twind=.016, sr=16000., thop=.005)
    for n in (0, fbk.nwind - 1, fbk.nwind):                    # no frame: n < len(w) - nwind never holds
        spec, t = fbk.specout(np.zeros(n, dtype=np.float32))
        assert spec.shape == (0, 12) and t.shape == (0,)
        assert stft.RMSWind(np.zeros(n), nwind=fbk.nwind, nhop=64)[0].shape == (0,)
        assert stft.SpecFlux(np.zeros(n), nwind=fbk.nwind, nhop=64)[0].shape == (0,)
    spec, _ = fbk.specout(np.zeros(fbk.nwind + 1, dtype=np.float32))
    assert spec.shape == (1, 12) and not spec.any()            # one frame of silence
    v, _ = stft.SpecFlux(np.zeros(fbk.nwind + 64 + 1), nwind=fbk.nwind, nhop=64)
    assert v.tolist() == [0.0]
    v, _ = stft.SpecFlux(np.ones(2000), nwind=256, nhop=64, minf=0.3, maxf=0.2)   # empty band
    assert not v.any()
    with pytest.raises(ValueError):
        stft.RMSWind(np.zeros(5000), nwind=1000)
    with pytest.raises(ValueError):
        stft.SpecFlux(np.zeros((4, 500)), nwind=256)


def test_full_size_properties():
    """10 minutes at 44.1 kHz (BASELINE configs[1] length): the frame count of the reference's
    loop, Parseval (a lowpass + highpass pair sums to one up to Nyquist and to zero above -- the
    reference's bands end at sr/2 -- so it recovers half of N * sum((x w)^2)), scaling (2x -> 4x band energy, 2x RMS and flux) and run-length independence."""
    import torch
    from pypevoc_b200 import signals, stft
    from pypevoc_b200.pv import n_frames
    sr, nwind, hop = 44100, 1024, 441
    x = signals.harm_torch(sr, sr * 600, 220.0, 90, 0.5, 0.01, 1, torch.device("cuda"), scale=0.25)
    specs = [stft.PiecewiseFilterSpec(mode="lowpass", freq=5000., sr=float(sr)),
             stft.PiecewiseFilterSpec(mode="hipass", freq=5000., sr=float(sr))]
    fbk = stft.FilterBank(fspec_list=specs, sr=float(sr), nwind=nwind, nhop=hop)
    ones = fbk.fb.sum(axis=0)
    bank, F = fbk.specout_device(x)
    assert F == n_frames(x.numel(), nwind, hop) and bank.shape == (F, 2)
    rms, _ = stft.RMSWind(x, sr=sr, nwind=nwind, nhop=hop, windfunc=np.hanning, to_host=False)
    assert np.count_nonzero(ones[:nwind // 2] != 1.0) <= 2 and not ones[nwind // 2 + 2:].any()
    total = bank.sum(dim=1)
    wsum2 = float(np.sum(np.hanning(nwind) ** 2))
    par = rms ** 2 * wsum2 * nwind / 2.0
    rel = ((total - par).abs() / par.clamp_min(1e-30)).max().item()
    assert rel < 2e-2                                          # DC / Nyquist / edge bins carry little energy
    bank2, _ = fbk.specout_device(2.0 * x)
    assert torch.allclose(bank2, 4.0 * bank, rtol=1e-12, atol=0.0)
    flux, _ = stft.SpecFlux(x, sr=sr, nwind=nwind, nhop=hop, to_host=False)
    flux2, _ = stft.SpecFlux(2.0 * x, sr=sr, nwind=nwind, nhop=hop, to_host=False)
    assert flux.shape == (n_frames(x.numel() - hop, nwind, hop),)
    assert torch.allclose(flux2, 2.0 * flux, rtol=1e-12, atol=0.0)
    st = fbk._state()
    alt = stft.stft_bank_device(x, st["win"], nwind, hop, F, st["fold"], st["lo"], st["hi"], run_frames=33)["bank"]
    assert torch.equal(alt, bank)
