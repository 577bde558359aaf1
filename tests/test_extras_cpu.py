"""CPU: the "next" rows of SURVEY 8f -- PVHarmonic, calc_f0 / partial_sum_magnitude, PeakFinder.refine.

* the numpy oracle against goldens produced by the unmodified reference
  (tests/golden/harmonic.npz, consumers.npz, refine.npz <- oracle/gen_golden.py), bit for bit;
* the CUDA sources of the new kernels (pvk_harmonic, pvk_frame_stats, pvk_analyze_ex) compiled
  for the SIMT emulator against the goldens and against the oracle run on the kernel's own
  spectrum.  The emulator is test infrastructure; the GPU tests proper are in test_gpu_extras.py."""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

from oracle import pv_oracle as orc
from golden_util import CASES, GOLD, case_golden, case_signal, pv_kwargs
from pypevoc_b200 import signals
import parity_util as pu

with open(os.path.join(GOLD, "harmonic_cases.json")) as _fh:
    HCASES = json.load(_fh)
HG = np.load(os.path.join(GOLD, "harmonic.npz"))
CG = np.load(os.path.join(GOLD, "consumers.npz"))
RG = np.load(os.path.join(GOLD, "refine.npz"), allow_pickle=True)


def hsignal(name):
    c = HCASES[name]
    out = getattr(signals, c["generator"])(**c["gen_kwargs"])
    x, sr = out if isinstance(out, tuple) else (out, c["gen_kwargs"]["sr"])
    return np.asarray(x, dtype=np.float32), sr


def hgolden(name):
    return {k: HG["%s.%s" % (name, k)] for k in ("f", "mag", "ph", "residuals", "t", "f0")}


@pytest.mark.parametrize("name", sorted(HCASES))
def test_oracle_harmonic_bit_exact_vs_reference(name):
    x, sr = hsignal(name)
    g = hgolden(name)
    o = orc.analyze_harmonic(x.astype(np.float64), sr, g["f0"], **HCASES[name]["pv_kwargs"])
    for k in ("f", "mag", "ph", "residuals", "t"):
        assert np.array_equal(o[k], g[k], equal_nan=True), k


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_consumers_bit_exact_vs_reference(name):
    g = case_golden(name)
    for args in ((50, 10000, 0.1), (200, 3000, 0.5)):
        fm, im = orc.calc_f0(g["f"], g["mag"], *args)
        tag = "%s.%d_%d_%g" % ((name,) + args)
        assert np.array_equal(fm, CG[tag + ".fm"]) and np.array_equal(im, CG[tag + ".idx"])
    psm = orc.partial_sum_magnitude(g["mag"])
    assert np.array_equal(psm, CG[name + ".psm"])
    with np.errstate(all="ignore"):
        assert np.array_equal(psm / g["totalmag"], CG[name + ".pmr"], equal_nan=True)


def test_oracle_refine_bit_exact_vs_reference():
    for y, idx, fp, fv in zip(RG["y"], RG["idx"], RG["fine_pos"], RG["fine_val"]):
        ofp, ofv = orc.refine_peaks(np.asarray(y, dtype=float), np.asarray(idx, dtype=int))
        assert np.array_equal(ofp, fp) and np.array_equal(ofv, fv)
    # the reference's own known answers (tests/test_peak_finder.py:22-48)
    assert [float(v[0]) for v in RG["fine_pos"][:3]] == [1.0, 1.2, 1.5]
    assert abs(float(RG["fine_pos"][3][0]) - 1.499) < 1e-7


# ------------------------------------------------------------------ emulator (same kernel sources)
eh = pytest.importorskip("emu_harness")
EMU_H = ["h_odd_hop", "h_high_f0", "h_readme"]


@pytest.fixture(scope="module", autouse=True)
def _build():
    eh.build()


@pytest.mark.parametrize("name", EMU_H)
def test_emu_harmonic(name):
    x, sr = hsignal(name)
    g = hgolden(name)
    kw = HCASES[name]["pv_kwargs"]
    got = eh.harmonic(x, sr, g["f0"], kw["nfft"], kw["hop"], kw["npks"])
    a = eh.analyze(x, sr, kw["nfft"], kw["hop"], 4, spectra=True)
    o = orc.analyze_harmonic(np.zeros(1), sr, g["f0"], nfft=kw["nfft"], hop=kw["hop"], npks=kw["npks"],
                             fx_given=a["fx"][0].astype(np.complex64))
    pu.compare_harmonic(got, g, g["f0"], sr, kw["nfft"], o["totalmag"])
    pu.compare_harmonic(got, o, g["f0"], sr, kw["nfft"], o["totalmag"], exact=True)
    assert np.array_equal(got["nharm"], o["nharm"])
    # the run length (frames per CTA, backward search for the last processed frame) does not matter
    got3 = eh.harmonic(x, sr, g["f0"], kw["nfft"], kw["hop"], kw["npks"], run_frames=3)
    for k in ("f", "mag", "ph", "residuals"):
        assert np.array_equal(got[k], got3[k], equal_nan=True), k


def test_emu_refine_and_frame_stats():
    name = "noisy_odd_hop"
    x, sr = case_signal(name)
    kw = pv_kwargs(name)
    a = eh.analyze(x, sr, kw["nfft"], kw["hop"], kw["npks"], spectra=True, refine=True)
    plain = eh.analyze(x, sr, kw["nfft"], kw["hop"], kw["npks"])
    for k in ("f", "mag", "ph", "realph", "binno"):
        assert np.array_equal(a[k], plain[k]), k
    fx = a["fx"][0].astype(np.complex64)
    for j in range(a["nframes"]):
        pw = fx[j].real * fx[j].real + fx[j].imag * fx[j].imag
        y = np.sqrt(pw.astype(np.float64))
        n = a["npk"][0, j]
        fp, fv = orc.refine_peaks(y, a["binno"][0, j, :n].astype(int))
        assert np.array_equal(fp, a["fine_pos"][0, j, :n]) and np.array_equal(fv, a["fine_val"][0, j, :n])
        assert not a["fine_pos"][0, j, n:].any() and not a["fine_val"][0, j, n:].any()
        assert np.all(np.abs(fp - a["binno"][0, j, :n]) <= 0.5)
    g = case_golden(name)
    for args in ((50, 10000, 0.1), (200, 3000, 0.5)):
        fm, idx, ps = eh.frame_stats(g["f"], g["mag"], *args)
        tag = "%s.%d_%d_%g" % ((name,) + args)
        assert np.array_equal(fm, CG[tag + ".fm"]) and np.array_equal(idx, CG[tag + ".idx"])
        assert np.allclose(ps, CG[name + ".psm"], rtol=1e-13, atol=0)


HPG = np.load(os.path.join(GOLD, "hpower.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_harmonic_power_oracle_and_emulated_kernel_vs_reference(name):
    """calc_harmonic_power (PVAnalysis.py:266-297, incl. the row indexing of :278): the oracle equals
    the unmodified reference bit for bit; the emulated pvk_harmonic_power kernel gives identical
    harmonic counts and hpower to fp64 summation-order round-off; the reference's IndexError cases
    (a peak in a column >= nframes) raise in the oracle and set the kernel's flag."""
    g = case_golden(name)
    if name + ".indexerror" in HPG.files:
        with pytest.raises(IndexError):
            orc.calc_harmonic_power(g["f"], g["mag"])
        assert eh.harmonic_power(g["f"], g["mag"])[2] == 1
        return
    for thr in (0.01, 0.05):
        tag = "%s.%g" % (name, thr)
        hp, nh = orc.calc_harmonic_power(g["f"], g["mag"], thr)
        assert np.array_equal(hp, HPG[tag + ".hpower"]) and np.array_equal(nh, HPG[tag + ".nharmonics"])
        ehp, enh, err = eh.harmonic_power(g["f"], g["mag"], thr)
        assert err == 0 and np.array_equal(enh, HPG[tag + ".nharmonics"])
        assert np.allclose(ehp, HPG[tag + ".hpower"], rtol=1e-13, atol=0)
