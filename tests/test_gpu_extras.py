"""GPU parity tests of the SURVEY 8f rows (run on the B200 box), all through the C ABI:
PVHarmonic (pvk_harmonic), device-side calc_f0 / partial_sum_magnitude (pvk_frame_stats) and
the opt-in PeakFinder.refine output of the peak kernel (pvk_analyze_ex) -- against goldens
produced by the unmodified reference and against the oracle run on the kernel's own spectrum."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import pv_oracle as orc
from golden_util import CASES, GOLD, case_golden, case_signal, pv_kwargs
from pypevoc_b200 import signals
import parity_util as pu

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLD, "harmonic_cases.json")) as _fh:
    HCASES = json.load(_fh)
HG = np.load(os.path.join(GOLD, "harmonic.npz"))
CG = np.load(os.path.join(GOLD, "consumers.npz"))


@pytest.fixture(scope="module")
def pvmod():
    import pypevoc_b200
    from pypevoc_b200 import _lib
    _lib.lib()           # fails loudly if libpvk.so is missing
    return pypevoc_b200


def hsignal(name):
    c = HCASES[name]
    out = getattr(signals, c["generator"])(**c["gen_kwargs"])
    x, sr = out if isinstance(out, tuple) else (out, c["gen_kwargs"]["sr"])
    return np.asarray(x, dtype=np.float32), sr


def own_spectrum(x, sr, nfft, hop):
    from pypevoc_b200.pv import analyze_device, host_tables
    o = analyze_device(torch.from_numpy(x).cuda(), sr, nfft, hop, 4, 0.005, host_tables(sr, nfft, hop), spectra=True)
    return o["fx"][0].cpu().numpy().astype(np.complex64)


@pytest.mark.parametrize("name", sorted(HCASES))
def test_pvharmonic_vs_reference_golden_and_own_spectrum(pvmod, name):
    x, sr = hsignal(name)
    kw = HCASES[name]["pv_kwargs"]
    g = {k: HG["%s.%s" % (name, k)] for k in ("f", "mag", "ph", "residuals", "t", "f0")}
    pv = pvmod.PVHarmonic(x, sr, progress=False, **kw)
    pv.set_f0(g["f0"])
    pv.run_pv()
    assert pv.nframes == len(g["t"]) and np.allclose(pv.t, g["t"], rtol=0, atol=1e-15)
    got = dict(f=pv.f, mag=pv.mag, ph=pv.ph, residuals=pv.residuals)
    o = orc.analyze_harmonic(np.zeros(1), sr, g["f0"], nfft=kw["nfft"], hop=kw["hop"], npks=kw["npks"],
                             fx_given=own_spectrum(x, sr, kw["nfft"], kw["hop"]))
    rep = pu.compare_harmonic(got, g, g["f0"], sr, kw["nfft"], o["totalmag"])
    pu.compare_harmonic(got, o, g["f0"], sr, kw["nfft"], o["totalmag"], exact=True)
    assert np.array_equal(pv.nharmonics, o["nharm"])
    print(name, rep)
    # frames per CTA (and with it the backward search for the last processed frame) is immaterial
    pv2 = pvmod.PVHarmonic(x, sr, progress=False, **kw)
    pv2.set_f0(torch.from_numpy(g["f0"]))
    pv2.run_pv(run_frames=3)
    for k in ("f", "mag", "ph", "residuals"):
        assert np.array_equal(getattr(pv, k), getattr(pv2, k), equal_nan=True), k


def test_pvharmonic_api(pvmod):
    x, sr = hsignal("h_readme")
    pv = pvmod.PVHarmonic(x, sr, nfft=2048, hop=512, npks=8, progress=False)
    assert pv.fmin == 30.0
    with pytest.raises(AttributeError):
        pv.run_pv()
    F = orc.n_frames(len(x), 2048, 512)
    # set_f0 with a time axis interpolates onto the frame times (PVAnalysis.py:433-441)
    t = np.linspace(0, 1, 11)
    pv.set_f0(500.0 + 10 * t, t=t)
    tint = np.arange(round(512 + 1024), len(x), 512) / float(sr)
    assert np.array_equal(pv.f0, np.interp(tint, t, 500.0 + 10 * t))
    pv.set_f0(np.full(F - 1, 500.0))
    with pytest.raises(IndexError):
        pv.run_pv()
    f0 = np.full(F, 500.0)
    pv.set_f0(f0)
    pv.run_pv()
    # single-frame accessor == row of the full run (frame 5, previous frame processed)
    ff, mm, pp, res = pv.calc_pv_frame(5 * 512, 500.0)
    n = len(ff)
    assert n == min(int(pv.nharmonics[5]), 8)
    assert np.array_equal(ff, pv.f[5, :n]) and np.array_equal(mm, pv.mag[5, :n]) and np.array_equal(pp, pv.ph[5, :n])
    assert res == pv.residuals[5]
    with pytest.raises(AttributeError):
        pv.toSinSum()
    # all-unvoiced track: nothing is processed
    pv.set_f0(np.zeros(F))
    pv.run_pv()
    assert not pv.f.any() and np.all(np.isnan(pv.residuals))


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_calc_f0_and_partial_sums(pvmod, name):
    """pvk_frame_stats on the reference's tables == the reference's calc_f0 (bit exact) and
    partial_sum_magnitude (fp64 rounding); PV.calc_f0 on the GPU tables == oracle on the same."""
    from pypevoc_b200.pv import frame_stats_device
    g = case_golden(name)
    fd, md = torch.from_numpy(g["f"]).cuda(), torch.from_numpy(g["mag"]).cuda()
    for args in ((50, 10000, 0.1), (200, 3000, 0.5)):
        fm, idx, ps = frame_stats_device(fd, md, *args)
        tag = "%s.%d_%d_%g" % ((name,) + args)
        assert np.array_equal(fm.cpu().numpy(), CG[tag + ".fm"])
        assert np.array_equal(idx.cpu().numpy(), CG[tag + ".idx"])
        assert np.allclose(ps.cpu().numpy(), CG[name + ".psm"], rtol=1e-13, atol=0)
    x, sr = case_signal(name)
    pv = pvmod.PV(x, sr, progress=False, **pv_kwargs(name))
    pv.run_pv()
    assert pv._on_device()
    fm = pv.calc_f0()                       # device path: the tables are not downloaded
    idx = pv.fundamental_idx
    psm = pv.partial_sum_magnitude
    assert pv._on_device()
    ofm, oidx = orc.calc_f0(pv.f, pv.mag)   # (fetches the tables)
    assert np.array_equal(fm, ofm) and np.array_equal(idx, oidx)
    assert np.allclose(psm, orc.partial_sum_magnitude(pv.mag), rtol=1e-13, atol=0)
    assert np.array_equal(pv.calc_f0(), ofm)                              # host path, same answer
    assert np.array_equal(pv.fundamental_frequency, pv.f[np.arange(pv.nframes), idx])
    with np.errstate(all="ignore"):
        assert np.allclose(pv.partial_magnitude_ratio, psm / np.asarray(pv.totalmag), rtol=1e-12, equal_nan=True)


def test_pitchjumps_detect_pitch_flow(pvmod):
    """The in-package consumer of PV (speech/PitchJumps.py:167-176): nfft / hop from nextpow2,
    default npks, sqrt(sum(mag^2)), get_time_vector, calc_f0 -- against the oracle."""
    x, sr = case_signal("cfg3_clip")
    nfft = int(2 ** np.ceil(np.log2(sr / 70.0 * 2)))
    hop = int(2 ** np.ceil(np.log2(sr * 0.01)))
    pv = pvmod.PV(x, sr, nfft=nfft, hop=hop)
    pv.run_pv()
    mag = np.sqrt(np.sum(pv.mag ** 2, axis=1))
    t = pv.get_time_vector()
    f0 = pv.calc_f0()
    o = orc.analyze(x, sr, nfft=nfft, hop=hop, npks=20, margins=True)
    safe = o["margin"] > pu.MARGIN_FP32
    of0, _ = orc.calc_f0(o["f"], o["mag"])
    assert np.allclose(t, o["t"], rtol=0, atol=1e-15)
    assert np.abs(f0 - of0)[safe].max() < pu.TOL_F * sr / nfft
    omag = orc.partial_sum_magnitude(o["mag"])
    assert (np.abs(mag - omag)[safe] / np.maximum(omag[safe], 1e-30)).max() < pu.TOL_MAG


@pytest.mark.parametrize("name", ["metric_1s", "noisy_odd_hop", "cfg3_clip", "cfg5_like"])
def test_refine_output_exact_on_own_spectrum(pvmod, name):
    """run_pv(refine=True): fine_pos / fine_val == PeakFinder.refine (oracle.refine_peaks, pinned to
    the reference by tests/golden/refine.npz) on the kernel's own |fx|; the other tables unchanged."""
    from pypevoc_b200.pv import analyze_device, host_tables
    x, sr = case_signal(name)
    kw = pv_kwargs(name)
    hop = kw["hop"] or kw["nfft"] // 2
    pv = pvmod.PV(x, sr, progress=False, **kw)
    pv.run_pv(refine=True)
    plain = pvmod.PV(x, sr, progress=False, **kw)
    plain.run_pv()
    for k in ("f", "mag", "ph", "realph", "binno"):
        assert np.array_equal(getattr(pv, k), getattr(plain, k)), k
    with pytest.raises(AttributeError):
        plain.fine_pos
    fx = analyze_device(torch.from_numpy(x).cuda(), sr, kw["nfft"], hop, kw["npks"], 0.005,
                        host_tables(sr, kw["nfft"], hop), spectra=True)["fx"][0].cpu().numpy().astype(np.complex64)
    npk = pv.device_tables["npk"].cpu().numpy()
    fp, fv, binno = pv.fine_pos, pv.fine_val, pv.binno
    for j in range(pv.nframes):
        pw = fx[j].real * fx[j].real + fx[j].imag * fx[j].imag
        y = np.sqrt(pw.astype(np.float64))
        n = npk[j]
        ofp, ofv = orc.refine_peaks(y, binno[j, :n].astype(int))
        assert np.array_equal(ofp, fp[j, :n]) and np.array_equal(ofv, fv[j, :n]), j
        assert not fp[j, n:].any() and not fv[j, n:].any()
    # streamed variant computes the same
    hb = {}
    pvs = pvmod.PV(torch.from_numpy(x).pin_memory(), sr, progress=False, **kw)
    pvs.run_pv(hostbuf=hb, refine=True)
    assert np.array_equal(pvs.fine_pos, fp) and np.array_equal(pvs.fine_val, fv) and np.array_equal(pvs.f, pv.f)


@pytest.mark.parametrize("K,mode", [(150, "asc"), (300, "asc_dense"), (200, "gaps"), (160, "shuffled"), (130, "ties"),
                                    (1024, "asc"), (100, "ties"), (50, "gaps"), (400, "asc_dense"), (512, "ties"), (257, "gaps")])
def test_link_kernels_on_random_wide_rows(K, mode):
    """pvk_track on random tables, bit for bit against the oracle's sequential greedy loop: rows
    wider than 128 peaks (sorted ranks + binary-searched window, full-scan fallback for rows with
    holes or out of order) and the propose/commit kernel below that."""
    import torch
    import parity_util as pu
    from oracle import pv_oracle as orc
    from pypevoc_b200 import pv as P
    f, mag = pu.wide_rows(K, mode, F=40)
    tr = P.track_device(torch.from_numpy(f).cuda(), torch.from_numpy(mag).cuda())
    ref = orc.track(f, mag)
    assert np.array_equal(tr["tid"].cpu().numpy(), ref["tid"])
    assert P.track_counts(tr)[0] == int(ref["tid"].max()) + 1


def test_pvbatch_equals_single_clips():
    """PVBatch (one launch over [nclips, nsamp]) == a PV per clip, bit for bit, incl. tracking and
    the per-clip PV views; clips shard across ranks by dist.clip_range without overlap."""
    import pypevoc_b200 as pb200
    from pypevoc_b200 import dist as D
    clips = np.stack([signals.speech_like_clip(seed=1000 + i, sr=16000, dur=0.8) for i in range(5)]).astype(np.float32)
    pb = pb200.PVBatch(clips, 16000, nfft=512, hop=128, npks=20)
    pb.run_pv()
    assert len(pb) == 5 and pb.f.shape == (5, pb.nframes, 20)
    tr = pb.track()
    ssb = pb.toSinSum()
    wb = ssb.synth(16000, 128)                       # ONE link / pack / resynthesis over the flattened table
    wb200 = ssb.synth(16000, 200, edge=0.5, minframes=2)
    assert len(ssb) == 5 and len(wb) == 5
    for i in range(5):
        pv = pb200.PV(clips[i], 16000, nfft=512, hop=128, npks=20, progress=False)
        pv.run_pv()
        for k in ("f", "mag", "ph", "realph", "binno"):
            assert np.array_equal(getattr(pb, k)[i], getattr(pv, k)), (i, k)
        assert np.array_equal(pb.totalmag[i], np.asarray(pv.totalmag))
        ss = pv.toSinSum()
        assert np.array_equal(tr["tid"][i].cpu().numpy(), ss.track_ids)
        view = pb[i]
        assert np.array_equal(view.f, pv.f) and np.array_equal(view.t, pv.t)
        assert np.array_equal(view.toSinSum().synth(16000, 128), ss.synth(16000, 128))
        assert int(tr["ntracks"][i]) == len(ss.st) == int(ssb.ntracks[i])
        assert wb[i].shape == ss.synth(16000, 128).shape and np.array_equal(wb[i], ss.synth(16000, 128))
        assert np.array_equal(wb200[i], ss.synth(16000, 200, edge=0.5, minframes=2))
    got = []
    for r in range(3):
        c0, c1 = D.clip_range(5, r, 3)
        part = pb200.PVBatch(clips[c0:c1], 16000, nfft=512, hop=128, npks=20)
        part.run_pv()
        got.append(part.f)
    assert np.array_equal(np.concatenate(got), pb.f)
    with pytest.raises(ValueError):
        pb200.PVBatch(clips[0], 16000)
    with pytest.raises(ValueError):
        ssb.synth(16000, 128, edge=2.0)               # beyond the guard rows (PVBatch(max_edge=...))
    # a silent clip in the batch: no partials, empty signal, neighbours untouched
    mixed = clips[:3].copy()
    mixed[1] = 0.0
    pm = pb200.PVBatch(mixed, 16000, nfft=512, hop=128, npks=20)
    pm.run_pv()
    wm = pm.toSinSum().synth(16000, 128)
    assert wm[1].size == 0 and np.array_equal(wm[0], wb[0]) and np.array_equal(wm[2], wb[2])
    assert pm.toSinSum().ntracks.tolist()[1] == 0


HPG = np.load(os.path.join(GOLD, "hpower.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_calc_harmonic_power(pvmod, name):
    """pvk_harmonic_power on the reference's tables == the reference's calc_harmonic_power
    (PVAnalysis.py:266-297, row indexing of :278 included): counts bit exact, hpower to fp64
    summation order; IndexError where the reference raises it; PV.calc_harmonic_power on the GPU's
    own tables == the oracle on the same tables."""
    from pypevoc_b200.pv import harmonic_power_device
    g = case_golden(name)
    fd, md = torch.from_numpy(g["f"]).cuda(), torch.from_numpy(g["mag"]).cuda()
    x, sr = case_signal(name)
    pv = pvmod.PV(x, sr, progress=False, **pv_kwargs(name))
    pv.run_pv()
    if name + ".indexerror" in HPG.files:
        assert int(harmonic_power_device(fd, md)[2].item()) == 1
        with pytest.raises(IndexError):
            pv.calc_harmonic_power()
        return
    for thr in (0.01, 0.05):
        tag = "%s.%g" % (name, thr)
        hp, nh, err = harmonic_power_device(fd, md, thr)
        assert int(err.item()) == 0 and np.array_equal(nh.cpu().numpy(), HPG[tag + ".nharmonics"])
        assert np.allclose(hp.cpu().numpy(), HPG[tag + ".hpower"], rtol=1e-13, atol=0)
    pv.calc_harmonic_power(0.02)
    ohp, onh = orc.calc_harmonic_power(pv.f, pv.mag, 0.02)
    assert np.array_equal(pv.nharmonics, onh) and np.allclose(pv.hpower, ohp, rtol=1e-13, atol=0)
    pv.mag = pv.mag * 2.0                                   # host-edited tables are honoured
    pv.calc_harmonic_power(0.02)
    assert np.allclose(pv.hpower, 4.0 * ohp, rtol=1e-13, atol=0)
