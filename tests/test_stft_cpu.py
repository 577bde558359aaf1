"""CPU: the STFT-consumer row.  (1) the numpy oracle reproduces the goldens of
the real reference bit for bit; (2) the host-side filter classes of pypevoc_b200.stft build the
reference's filter matrices exactly; (3) the bank kernel, compiled for the SIMT emulator, meets
the goldens within the stated fp32 tolerances."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

from oracle import stft_oracle as so
import stft_util as su

G = su.golden()
BANKS = su.CASES["banks"]


def _oracle_bank(name):
    sig, kind, kw = BANKS[name]
    kw = dict(kw)
    if kind == "FilterBank":
        sr, nwind = kw["sr"], kw["nwind"]
        hop = kw.get("nhop") or int(nwind / 2)
        specs = kw.get("specs")
        if specs:
            sp = [so.spec_bands(m, np.asarray(f) if isinstance(f, list) else f, sr=sr) for m, f in specs]
        else:
            sp = [so.spec_bands("lowpass", 0.25, sr=sr), so.spec_bands("hipass", 0.25, sr=sr)]
    elif kind == "TriangularFilterBank":
        sr, nwind = kw["sr"], kw["nwind"]
        hop = int(nwind / 2)
        sp = so.triangular_specs(kw["flim"], sr)
    else:
        sr = kw["sr"]
        nwind, hop = so.mel_geometry(kw["twind"], sr, kw["thop"])
        sp = so.triangular_specs(so.mel_limits(kw["n"], kw["fmin"], kw["fmax"]), sr)
    return so.bank_matrix(sp, sr, nwind), nwind, hop, sr


def _product_bank(name, **extra):
    from pypevoc_b200 import stft
    sig, kind, kw = BANKS[name]
    kw = dict(kw)
    if kind == "FilterBank":
        specs = kw.pop("specs", None)
        fsl = None
        if specs:
            fsl = [stft.PiecewiseFilterSpec(mode=m, freq=np.asarray(f) if isinstance(f, list) else f, sr=kw["sr"])
                   for m, f in specs]
        return stft.FilterBank(fspec_list=fsl, **kw, **extra)
    return getattr(stft, kind)(**kw, **extra)


@pytest.mark.parametrize("name", sorted(BANKS))
def test_oracle_filterbank_matches_reference(name):
    fb, nwind, hop, sr = _oracle_bank(name)
    assert [nwind, hop] == G[name + "/geom"].tolist()
    assert np.array_equal(fb, G[name + "/fb"])
    x, _ = su.signal(BANKS[name][0])
    spec, t = so.specout(x.astype(np.float64), fb, G[name + "/wind"], hop, sr)
    assert np.array_equal(spec, G[name + "/spec"]) and np.array_equal(t, G[name + "/t"])


def test_oracle_rms_and_flux_match_reference():
    for name, (sig, kw) in su.CASES["rms"].items():
        x, sr = su.signal(sig)
        v, t = so.rms_wind(x.astype(np.float64), sr=sr, **su.windfunc(kw))
        assert np.array_equal(v, G[name + "/v"]) and np.array_equal(t, G[name + "/t"]), name
    for name, (sig, kw) in su.CASES["flux"].items():
        x, sr = su.signal(sig)
        v, t = so.spec_flux(x.astype(np.float64), sr=sr, **su.windfunc(kw))
        assert np.array_equal(v, G[name + "/v"]) and np.array_equal(t, G[name + "/t"]), name


@pytest.mark.parametrize("name", sorted(BANKS))
def test_host_filter_classes_build_the_reference_matrix(name):
    fbk = _product_bank(name)
    assert [fbk.nwind, fbk.hop] == G[name + "/geom"].tolist()
    assert np.array_equal(fbk.fb, G[name + "/fb"])
    assert np.array_equal(fbk.wind, G[name + "/wind"])
    assert len(fbk.label) == fbk.fb.shape[0] and "FilterBank with filters" in repr(fbk)


def test_host_filter_spec_api():
    from pypevoc_b200 import stft
    sp = stft.PiecewiseFilterSpec(mode="bp", freq=np.array([1000., 3000.]), sr=16000.)
    fr, g = sp.get_frequency_gains()
    assert np.allclose(fr, [[0, 1000], [1000, 3000], [3000, 8000]]) and g.shape == (3, 2)
    assert np.allclose(sp.get_frequency_edges(), [0, 1000, 3000, 8000])
    m = sp.apply_to_freq_vector(np.array([0., 500., 1000., 2000., 3000., 5000.]))
    assert m.tolist() == [0, 0, 1, 1, 0, 0]                    # later bands win on shared edges
    assert "Bandpass filter" in repr(sp)
    with pytest.raises(stft.BandError):
        stft.FilterBank(sr=16000., nwind=256)                  # the default 0.25 Hz cutoff collapses at sr > 1
    with pytest.raises(ValueError):
        stft.FilterBank(sr=1.0, nwind=300)._state()            # not a power of two: no CUDA kernel
    fold, lo, hi = stft.fold_bank(np.arange(16.).reshape(2, 8))
    assert fold[0].tolist() == [0, 1 + 7, 2 + 6, 3 + 5, 4] and (lo[0], hi[0]) == (1, 5)
    assert float(stft.mel_to_f(stft.f_to_mel(440.))) == pytest.approx(440.)


@pytest.mark.parametrize("name", sorted(BANKS))
def test_emu_bank_kernel_vs_reference_golden(name):
    eh = pytest.importorskip("emu_harness")
    eh.build()
    from pypevoc_b200 import stft
    from pypevoc_b200.pv import n_frames
    fbk = _product_bank(name)
    x, _ = su.signal(BANKS[name][0])
    fold, lo, hi = stft.fold_bank(fbk.fb)
    F = n_frames(len(x), fbk.nwind, fbk.hop)
    out = eh.stft_bank(x, fbk.wind, fbk.nwind, fbk.hop, F, fold, lo, hi, run_frames=5)
    su.close(out["bank"], G[name + "/spec"], su.TOL_BANK)


def test_emu_rms_and_flux_vs_reference_golden():
    eh = pytest.importorskip("emu_harness")
    eh.build()
    from pypevoc_b200.pv import n_frames
    for name, (sig, kw) in su.CASES["rms"].items():
        x, sr = su.signal(sig)
        kw = su.windfunc(kw)
        wind = kw.get("windfunc", np.blackman)(kw["nwind"])
        F = n_frames(len(x), kw["nwind"], kw["nhop"])
        out = eh.stft_bank(x, wind, kw["nwind"], kw["nhop"], F, inv_wsum2=1.0 / np.sum(wind ** 2), run_frames=7)
        su.close(out["rms"], G[name + "/v"], su.TOL_RMS)
    for name, (sig, kw) in su.CASES["flux"].items():
        x, sr = su.signal(sig)
        kw = su.windfunc(kw)
        nwind, nhop = kw["nwind"], kw["nhop"]
        wind = kw.get("windfunc", np.blackman)(nwind)
        minbin = int(kw.get("minf", 0) / sr * nwind)
        mb = float(kw.get("maxf", np.inf)) / sr * nwind
        maxbin = nwind if mb > nwind else int(mb)
        F = n_frames(len(x) - nhop, nwind, nhop)
        assert F == len(G[name + "/v"])
        out = eh.stft_bank(x, wind, nwind, nhop, F + 1, flux_bins=(minbin, maxbin), run_frames=4)
        su.close(out["flux"], G[name + "/v"], su.TOL_FLUX, scale=su.band_norm(x, wind, nhop, minbin, maxbin))
