"""TEST INFRASTRUCTURE ONLY -- tests/golden/real_wav.npz from the REAL reference (build container):

    python -m oracle.gen_golden_wav

Real-signal regression (SURVEY 8c iv): 1.5 s of the reference's bundled examples/pepperSx.wav
(22.05 kHz, 16-bit mono; a saxophone phrase), scaled as examples/WavResynth.py:17-18 does and analysed
with that script's parameters (nfft 4096, npks 100, hop 1024, :25), then tracked and resynthesised.
The input excerpt cannot be regenerated from a seed, so its int16 samples are stored in the fixture
next to the reference's outputs.
"""
import os
import sys
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.gen_golden_params import tid_table  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
WAV = os.path.join(ref_loader.REF_ROOT, "examples", "pepperSx.wav")
START, COUNT = 22050, 33075                    # 1.0 s ... 2.5 s
PVKW = dict(nfft=4096, hop=1024, npks=100)


# more recordings of the reference's examples/ (round 2), each with parameters of a BASELINE config:
# (output file, wav, first sample, samples, PV kwargs)
EXTRA = [
    ("real_guitar.npz", "SoloGuitarArpegi.wav", 44100, 44100, dict(nfft=2048, hop=512, npks=50)),       # the metric's parameters
    ("real_speech.npz", "ProtectMarraigeInAmerica.wav", 22050, 33075, dict(nfft=512, hop=128, npks=20)),  # cfg3's (clip batches)
]


def one(outname, wavname, start, count, pvkw):
    w = wave.open(os.path.join(ref_loader.REF_ROOT, "examples", wavname), "r")
    sr = w.getframerate()
    w.setpos(start)
    pcm = np.frombuffer(w.readframes(count), dtype="<i2").copy()
    w.close()
    x = (pcm / float(np.iinfo(np.int16).max)).astype(np.float32).astype(np.float64)
    pv = ref_loader.ref_run_pv(x, sr, **pvkw)
    ss = pv.toSinSum()
    out = dict(pcm=pcm, sr=np.int64(sr), f=pv.f, mag=pv.mag, ph=pv.ph, realph=pv.realph, binno=pv.binno,
               totalmag=np.array(pv.totalmag), tid=tid_table(pv, ss), st=np.array(ss.st, dtype=np.int64),
               end=np.array(ss.end, dtype=np.int64), synth=ref_loader.ref_sinsum_synth(ss, sr, pvkw["hop"]),
               nfft=np.int64(pvkw["nfft"]), hop=np.int64(pvkw["hop"]), npks=np.int64(pvkw["npks"]))
    print(outname, "frames", pv.nframes, "partials", len(ss.partial), "peaks/frame %.1f" % (pv.f > 0).sum(1).mean())
    np.savez_compressed(os.path.join(GOLD, outname), **out)


def main():
    if "--extra-only" in sys.argv:
        for case in EXTRA:
            one(*case)
        return
    w = wave.open(WAV, "r")
    sr = w.getframerate()
    w.setpos(START)
    pcm = np.frombuffer(w.readframes(COUNT), dtype="<i2").copy()
    w.close()
    x = (pcm / float(np.iinfo(np.int16).max)).astype(np.float32).astype(np.float64)
    pv = ref_loader.ref_run_pv(x, sr, **PVKW)
    ss = pv.toSinSum()
    out = dict(pcm=pcm, sr=np.int64(sr), f=pv.f, mag=pv.mag, ph=pv.ph, realph=pv.realph, binno=pv.binno,
               totalmag=np.array(pv.totalmag), tid=tid_table(pv, ss), st=np.array(ss.st, dtype=np.int64),
               end=np.array(ss.end, dtype=np.int64), synth=ref_loader.ref_sinsum_synth(ss, sr, PVKW["hop"]))
    print("frames", pv.nframes, "partials", len(ss.partial), "peaks/frame %.1f" % (pv.f > 0).sum(1).mean())
    np.savez_compressed(os.path.join(GOLD, "real_wav.npz"), **out)
    for case in EXTRA:
        one(*case)


if __name__ == "__main__":
    main()
