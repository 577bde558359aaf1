"""TEST INFRASTRUCTURE ONLY -- load the real PyPeVoc reference (build container only).

The reference (pure Python) does ``import pylab`` at module import time
(/root/reference/pypevoc/PVAnalysis.py:29) and evaluates ``pl.cm.rainbow`` while the
``SinSum`` class body is executed (:1033); ``Periodicity.py:33`` and
``TransferFunctions.py:6`` import matplotlib.  Neither is installed here, so inert
stub modules are seeded into ``sys.modules`` *before* the import.  The reference source
itself is never modified or copied: it is imported from where it is mounted.

Two documented shims are needed for the resynthesis half, which is Python-2-only:
  * ``PVAnalysis.xrange = range``                      (PVAnalysis.py:703 uses xrange)
  * the 12-line overlap-add of ``SinSum.synth`` (:1053-1070) is re-done in
    :func:`ref_sinsum_synth` with ``int()`` casts, calling the *unmodified*
    ``RegPartial.synth`` for every partial (the original raises TypeError at :1059
    because ``np.zeros`` gets a float size).

``/root/reference`` does not exist on the GPU box.  The one thing that travels is the offline
``pip install --target baseline/_ref`` of the unmodified reference (done by
``__graft_entry__.build()`` in the build container; git-ignored): ``bench.py --impl reference``
times it there; where neither exists ``available()`` is False and callers fall back / skip.
"""
import importlib
import os
import sys
import types
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """$PYPEVOC_REFERENCE, else /root/reference (build container), else the offline pip install of
    the unmodified reference under baseline/_ref (git-ignored; it travels to the GPU box with the
    snapshot, where bench.py --impl reference times it)."""
    cands = [os.environ.get("PYPEVOC_REFERENCE"), "/root/reference",
             os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "pypevoc", "PVAnalysis.py")):
            return c
    return cands[0] or "/root/reference"


REF_ROOT = _find_root()


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "pypevoc", "PVAnalysis.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


_cached = None


def load():
    """Return the reference's ``pypevoc.PVAnalysis`` module (and PeakFinder via .pf)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("PyPeVoc reference not found at %s" % REF_ROOT)
    try:
        import matplotlib  # noqa: F401  (use the real one if it ever exists)
        import pylab  # noqa: F401
    except Exception:
        cm = types.SimpleNamespace(rainbow=None)
        _stub("pylab", cm=cm)
        mpl = _stub("matplotlib")
        colors = _stub("matplotlib.colors", hsv_to_rgb=lambda x: x)
        pyplot = _stub("matplotlib.pyplot")
        mlab = _stub("matplotlib.mlab", psd=None, csd=None, cohere=None, specgram=None)
        mpl.colors, mpl.pyplot, mpl.mlab = colors, pyplot, mlab
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("pypevoc.PVAnalysis")
    mod.xrange = range  # py2 shim for RegPartial.synth (PVAnalysis.py:703)
    _cached = mod
    return mod


def load_peakfinder():
    load()
    return importlib.import_module("pypevoc.PeakFinder")


def ref_run_pv(x, sr, nfft, hop=None, npks=20, pkthresh=0.005, wind=np.hanning):
    """Run the unmodified reference analysis; returns the PV object."""
    mod = load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # divide by zero on frame 0 (PVAnalysis.py:171)
        with np.errstate(all="ignore"):
            pv = mod.PV(x, sr, nfft=nfft, hop=hop, npks=npks, pkthresh=pkthresh,
                        wind=wind, progress=False)
            pv.run_pv()
    return pv


def ref_sinsum_synth(ss, sr, hop, edge=1.0, minframes=3):
    """``SinSum.synth`` (PVAnalysis.py:1053-1070) with the int() casts Python 3 needs.

    Every partial is rendered by the reference's own, unmodified ``RegPartial.synth``.
    """
    load()
    hop = int(hop)
    dfr = ss.nfft / ss.hop / 2.0
    edgsamp = int(edge * hop * dfr)
    w = np.zeros((max(ss.end) + 2) * hop + 2 * edgsamp)
    for part in ss.partial:
        if len(part.f) >= minframes:
            wi, spl_st = part.synth(sr, hop, edge=edge)
            spl_st = int(spl_st + edgsamp)
            if spl_st >= 0:
                w[spl_st:spl_st + len(wi)] += wi
    return w[edgsamp:]
