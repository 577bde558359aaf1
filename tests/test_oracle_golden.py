"""CPU: the numpy oracle is pinned bit-for-bit against outputs of the real reference
(tests/golden/*.npz, made by oracle/gen_golden.py) and the reference's own PeakFinder
known answers (reference tests/test_peak_finder.py:15-20)."""
import numpy as np
import pytest

from oracle import pv_oracle as orc
from golden_util import CASES, case_golden, case_signal, pv_kwargs

NAMES = sorted(CASES)


@pytest.mark.parametrize("name", NAMES)
def test_analysis_bit_exact(name):
    x, sr = case_signal(name)
    g = case_golden(name)
    o = orc.analyze(x, sr, spectra=True, **pv_kwargs(name))
    assert o["nframes"] == int(g["nframes"])
    for k in ("f", "mag", "ph", "realph", "binno", "t"):
        assert np.array_equal(o[k], g[k], equal_nan=True), k
    assert np.array_equal(np.array(o["totalmag"]), g["totalmag"])
    assert np.array_equal(o["fx"][g["fx_frames"]], g["fx"])


@pytest.mark.parametrize("name", NAMES)
def test_tracking_bit_exact(name):
    g = case_golden(name)
    tr = orc.track(g["f"], g["mag"])
    assert np.array_equal(tr["tid"], g["tid"])
    assert np.array_equal(tr["st"], g["st"])
    assert np.array_equal(tr["end"], g["end"])


@pytest.mark.parametrize("name", [n for n in NAMES if CASES[n]["synth_hops"]])
def test_resynthesis_bit_exact(name):
    g = case_golden(name)
    _, sr = case_signal(name)
    kw = pv_kwargs(name)
    tr = orc.track(g["f"], g["mag"])
    parts = orc.partials_from_tracks(tr, g["f"], g["mag"], g["ph"], g["realph"])
    for h in CASES[name]["synth_hops"]:
        w = orc.synth(parts, sr, h, kw["nfft"], int(g["hop"]))
        ref = g["synth_%d" % h]
        assert w.shape == ref.shape
        assert np.max(np.abs(w - ref)) == 0.0


def test_peakfinder_golden():
    g = case_golden("peakfinder")
    for y, k, th, sel, keep in zip(g["y"], g["npks"], g["pkthresh"], g["sel"], g["keep"]):
        y = np.asarray(y, dtype=np.float64)
        s, _, _ = orc.peak_select(y, int(k), float(th))
        assert np.array_equal(s, sel)
        assert np.array_equal(orc.peak_pick(y, int(k), float(th)), keep)


def test_peakfinder_reference_known_answer():
    # reference tests/test_peak_finder.py:15-20 (testFindOnePeak): single peak at index 9
    x = np.concatenate((np.linspace(0, 1, 10), np.linspace(.9, 1, 9)))
    sel, _, _ = orc.peak_select(x, len(x), None or 0.0)
    assert sel.tolist() == [9]


def test_two_sines_known_answer():
    # printed summary of reference tests/test_pypevoc.py
    g = case_golden("two_sines")
    tr = orc.track(g["f"], g["mag"])
    parts = orc.partials_from_tracks(tr, g["f"], g["mag"], g["ph"], g["realph"])
    summ = [(p["start_idx"], len(p["f"]), float(np.mean(p["f"])), float(np.mean(p["mag"]))) for p in parts]
    assert summ[1][0] == 0 and summ[1][1] == 85 and abs(summ[1][2] - 1199.689) < 1e-2
    assert summ[2][0] == 1 and summ[2][1] == 84 and abs(summ[2][2] - 400.0) < 1e-3
    assert abs(summ[2][3] - 0.099773) < 1e-5


def test_frame_count_and_empty():
    assert orc.n_frames(2048 + 3 * 512, 2048, 512) == 3
    o = orc.analyze(np.zeros(100), 44100, nfft=1024)
    assert o["nframes"] == 0 and o["f"].shape == (0,)
