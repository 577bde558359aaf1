"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the REAL reference.

Run in the build container (where /root/reference is mounted):

    python -m oracle.gen_golden

Inputs are not stored: each case names a generator in ``pypevoc_b200.signals`` and its
arguments, the tests regenerate the identical fp32-representable samples from the seed.
Outputs are what the unmodified reference returns (``PV.run_pv``, ``PV.toSinSum``,
``PeakFinder``) plus the resynthesis obtained through the documented py3 shim
(``oracle.ref_loader.ref_sinsum_synth``: unmodified ``RegPartial.synth`` per partial).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from pypevoc_b200 import signals  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name -> (generator, gen kwargs, PV kwargs, synth hops)
CASES = {
    "two_sines": ("two_sines", {}, dict(nfft=1024, hop=512, npks=20), [512]),
    "readme_vibrato": ("readme_vibrato", {}, dict(nfft=2048, hop=None, npks=3), [1024]),
    "metric_1s": ("harm", dict(sr=44100, dur=1.0, f0=220, nharm=90, p=0.5, sigma=0.01, seed=1),
                  dict(nfft=2048, hop=512, npks=50), [512, 700]),
    "cfg2_like": ("harm", dict(sr=44100, dur=0.6, f0=110, nharm=150, p=0.5, sigma=0.01, seed=2),
                  dict(nfft=4096, hop=512, npks=50), []),
    "cfg3_clip": ("speech_like_clip", dict(seed=1000, sr=16000, dur=3.0),
                  dict(nfft=512, hop=128, npks=20), [128]),
    "cfg4_like": ("harm", dict(sr=44100, dur=0.5, f0=200, nharm=100, p=0.4, sigma=0.01, seed=4000),
                  dict(nfft=2048, hop=256, npks=100), []),
    "cfg5_like": ("harm", dict(sr=48000, dur=0.8, f0=55, nharm=420, p=0.3, sigma=0.001, seed=5),
                  dict(nfft=8192, hop=1024, npks=400), []),
    "noisy_odd_hop": ("harm", dict(sr=22050, dur=0.5, f0=300, nharm=20, p=1.0, sigma=0.2, seed=7),
                      dict(nfft=1024, hop=300, npks=30), [300]),
}


def make_signal(gen, kw):
    out = getattr(signals, gen)(**kw)
    if isinstance(out, tuple):
        x, sr = out
    else:
        x, sr = out, kw["sr"]
    return np.asarray(x, dtype=np.float32), sr


def run_case(name):
    gen, gkw, pkw, synth_hops = CASES[name]
    x, sr = make_signal(gen, gkw)
    pv = ref_loader.ref_run_pv(x.astype(np.float64), sr, **pkw)
    ss = pv.toSinSum()
    F, K = pv.nframes, pv.npeaks
    out = dict(f=pv.f, mag=pv.mag, ph=pv.ph, realph=pv.realph, binno=pv.binno, t=pv.t,
               totalmag=np.array(pv.totalmag), nframes=np.int64(F), hop=np.int64(pv.hop),
               st=np.array(ss.st, dtype=np.int64), end=np.array(ss.end, dtype=np.int64))
    # track id per slot: find every partial's points back in the frame tables
    tid = -np.ones((F, K), dtype=np.int32)
    for i, part in enumerate(ss.partial):
        for k in range(len(part.f)):
            fr = part.start_idx + k
            cols = np.flatnonzero((pv.f[fr] == part.f[k]) & (pv.mag[fr] == part.mag[k]) & (tid[fr] < 0))
            assert len(cols) >= 1, (name, i, k)
            tid[fr, cols[0]] = i
    out["tid"] = tid
    for h in synth_hops:
        out["synth_%d" % h] = ref_loader.ref_sinsum_synth(ss, sr, h)
    # first / middle spectrum frames from calc_fft_frame (PVAnalysis.py:150-158)
    nf2 = pv.nfft2
    frames = sorted(set([0, F // 2, F - 1])) if F else []
    out["fx_frames"] = np.array(frames, dtype=np.int64)
    out["fx"] = np.array([pv.calc_fft_frame(j * pv.hop)[:nf2] for j in frames])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    return F, len(ss.partial)


def peakfinder_cases():
    """Random spectra (Rayleigh, heavy tailed, quantised plateaus, flat / th<0) through the
    reference PeakFinder exactly as PV drives it (PVAnalysis.py:175-178)."""
    pfm = ref_loader.load_peakfinder()
    rng = np.random.RandomState(1234)
    ys, ks, ths, sels, keeps = [], [], [], [], []
    for trial in range(400):
        n = int(rng.choice([16, 64, 256, 1024]))
        kind = trial % 5
        if kind == 0:
            y = rng.rayleigh(1.0, n)
        elif kind == 1:
            y = np.abs(rng.standard_cauchy(n))
        elif kind == 2:
            y = np.round(rng.rayleigh(1.0, n) * 4) / 4.0          # plateaus / ties
        elif kind == 3:
            y = 1.0 + 0.01 * rng.rand(n)                          # flat: th < 0
        else:
            y = np.abs(np.sinc(np.linspace(-8, 8, n))) + 0.001 * rng.rand(n)
        if trial % 37 == 0:
            y[:] = 0.0
        k = int(rng.choice([1, 3, 20, 50, 400]))
        th = float(rng.choice([0.005, 0.0, 0.1, 0.9]))
        pk = pfm.PeakFinder(y, npeaks=k, minrattomax=th)
        pk.boundaries()
        sel = np.array(pk._idx, dtype=np.int64)
        pk.filter_by_salience(rad=5)
        keep = np.array(pk.get_pos(), dtype=np.int64)
        ys.append(y), ks.append(k), ths.append(th), sels.append(sel), keeps.append(keep)
    np.savez_compressed(
        os.path.join(GOLD, "peakfinder.npz"),
        y=np.array(ys, dtype=object), npks=np.array(ks), pkthresh=np.array(ths),
        sel=np.array(sels, dtype=object), keep=np.array(keeps, dtype=object))


# name -> (generator, gen kwargs, PVHarmonic kwargs, nominal f0)
HARM_CASES = {
    "h_readme": ("readme_vibrato", {}, dict(nfft=2048, hop=512, npks=8), 500.0),
    "h_metric": ("harm", dict(sr=44100, dur=0.5, f0=220, nharm=90, p=0.5, sigma=0.01, seed=1),
                 dict(nfft=2048, hop=512, npks=50), 220.0),
    "h_odd_hop": ("harm", dict(sr=22050, dur=0.4, f0=300, nharm=20, p=1.0, sigma=0.2, seed=7),
                  dict(nfft=1024, hop=300, npks=5), 300.0),
    "h_speech": ("speech_like_clip", dict(seed=1000, sr=16000, dur=3.0), dict(nfft=512, hop=128, npks=20), 150.0),
    "h_tiny_f0": ("harm", dict(sr=44100, dur=0.3, f0=110, nharm=150, p=0.5, sigma=0.01, seed=2),
                  dict(nfft=4096, hop=512, npks=30), 7.0),
    "h_high_f0": ("harm", dict(sr=8000, dur=0.5, f0=100, nharm=10, p=0.5, sigma=0.01, seed=2),
                  dict(nfft=256, hop=64, npks=4), 3900.0),
}


def harmonic_f0(name, nframes):
    """The f0 track of a harmonic case: nominal f0 with 2 % jitter, ~15 % unvoiced (0) and ~5 % NaN
    frames (both are skipped by PVHarmonic.run_pv, PVAnalysis.py:509)."""
    base = HARM_CASES[name][3]
    rng = np.random.RandomState(sum(map(ord, name)))
    f0 = base * (1 + 0.02 * rng.randn(nframes))
    f0[rng.rand(nframes) < 0.15] = 0.0
    f0[rng.rand(nframes) < 0.05] = np.nan
    if name == "h_high_f0":
        f0[3] = 4100.0          # f0bin beyond nfft/2 - 1: no harmonics at all
    return f0


def harmonic_cases():
    """PVHarmonic.set_f0 + run_pv of the unmodified reference (progress=True: its run_pv calls
    self.progress unconditionally, :528-530; the bar's output is swallowed)."""
    import contextlib
    import io
    mod = ref_loader.load()
    out = {}
    for name, (gen, gkw, pkw, _) in HARM_CASES.items():
        x, sr = make_signal(gen, gkw)
        nfr = -(-(len(x) - pkw["nfft"]) // pkw["hop"])
        f0 = harmonic_f0(name, nfr)
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            pv = mod.PVHarmonic(x.astype(np.float64), sr, progress=True, **pkw)
            pv.set_f0(f0)
            pv.run_pv()
        assert pv.nframes == nfr
        for k in ("f", "mag", "ph", "residuals", "t"):
            out["%s.%s" % (name, k)] = getattr(pv, k)
        out["%s.f0" % name] = f0
        print(name, nfr, int(np.isnan(pv.residuals).sum()))
    np.savez_compressed(os.path.join(GOLD, "harmonic.npz"), **out)
    with open(os.path.join(GOLD, "harmonic_cases.json"), "w") as fh:
        json.dump({k: dict(generator=v[0], gen_kwargs=v[1], pv_kwargs=v[2]) for k, v in HARM_CASES.items()},
                  fh, indent=1, sort_keys=True)


def consumer_cases():
    """PV.calc_f0 / fundamental_idx / partial_sum_magnitude / partial_magnitude_ratio
    (PVAnalysis.py:371-417) of the reference on its own tables, for every analysis case."""
    out = {}
    for name in CASES:
        gen, gkw, pkw, _ = CASES[name]
        x, sr = make_signal(gen, gkw)
        pv = ref_loader.ref_run_pv(x.astype(np.float64), sr, **pkw)
        for args in ((50, 10000, 0.1), (200, 3000, 0.5)):
            fm = pv.calc_f0(*args)
            tag = "%s.%d_%d_%g" % ((name,) + args)
            out[tag + ".fm"] = fm
            out[tag + ".idx"] = np.array(pv.fundamental_idx)
        out[name + ".psm"] = pv.partial_sum_magnitude
        out[name + ".pmr"] = pv.partial_magnitude_ratio
    np.savez_compressed(os.path.join(GOLD, "consumers.npz"), **out)


def hpower_cases():
    """PV.calc_harmonic_power (PVAnalysis.py:266-297) of the unmodified reference on its own tables,
    for every analysis case and two thresholds.  Cases in which a peak sits in a column >= nframes
    raise IndexError in the reference (:278 indexes rows of mag with column numbers): recorded as
    ``<case>.indexerror = 1``."""
    out = {}
    for name in CASES:
        gen, gkw, pkw, _ = CASES[name]
        x, sr = make_signal(gen, gkw)
        pv = ref_loader.ref_run_pv(x.astype(np.float64), sr, **pkw)
        for thr in (0.01, 0.05):
            tag = "%s.%g" % (name, thr)
            try:
                with np.errstate(all="ignore"):
                    pv.calc_harmonic_power(thr)
            except IndexError:
                out[name + ".indexerror"] = np.array(1)
                break
            out[tag + ".hpower"] = pv.hpower
            out[tag + ".nharmonics"] = pv.nharmonics
        print(name, pv.f.shape, name + ".indexerror" in out)
    np.savez_compressed(os.path.join(GOLD, "hpower.npz"), **out)


def refine_cases():
    """PeakFinder.refine_all (PeakFinder.py:331-406) of the reference: the known answers of its
    tests/test_peak_finder.py:22-48 and random spectra."""
    pfm = ref_loader.load_peakfinder()

    def parabolic_peak(max_pos=1.0, max_val=1.0, n=3, a=-1.):      # tests/test_peak_finder.py:7-11
        x = np.arange(n)
        b = -max_pos * 2 * a
        c = max_val - a * max_pos * (b + max_pos)
        return a * x * x + b * x + c
    ys, idxs, fps, fvs = [], [], [], []
    known = [(1.0, 3), (1.2, 4), (1.5, 4), (1.499, 4)]
    rng = np.random.RandomState(99)
    for trial in range(60):
        if trial < len(known):
            y = parabolic_peak(max_pos=known[trial][0], n=known[trial][1])
        elif trial % 2:
            y = rng.rayleigh(1.0, int(rng.choice([32, 256, 1024])))
        else:
            y = np.abs(np.sinc(np.linspace(-8, 8, 512) + rng.rand())) + 0.001 * rng.rand(512)
        pk = pfm.PeakFinder(y, npeaks=50)
        pk.refine_all()
        ys.append(y), idxs.append(np.array(pk._idx, dtype=np.int64))
        fps.append(np.array(pk._fine_pos)), fvs.append(np.array(pk._fine_val))
    assert [float(v[0]) for v in fps[:3]] == [1.0, 1.2, 1.5] and abs(fps[3][0] - 1.499) < 1e-7
    np.savez_compressed(os.path.join(GOLD, "refine.npz"), y=np.array(ys, dtype=object), idx=np.array(idxs, dtype=object),
                        fine_pos=np.array(fps, dtype=object), fine_val=np.array(fvs, dtype=object))


def main():
    os.makedirs(GOLD, exist_ok=True)
    if "--hpower-only" in sys.argv:
        hpower_cases()
        return
    if "--extras-only" in sys.argv:
        harmonic_cases(), consumer_cases(), refine_cases(), hpower_cases()
        return
    meta = {}
    for name in CASES:
        F, P = run_case(name)
        meta[name] = dict(generator=CASES[name][0], gen_kwargs=CASES[name][1],
                          pv_kwargs=CASES[name][2], synth_hops=CASES[name][3],
                          nframes=int(F), npartials=int(P))
        print(name, F, P)
    peakfinder_cases()
    harmonic_cases(), consumer_cases(), refine_cases(), hpower_cases()
    with open(os.path.join(GOLD, "cases.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
