"""WAV ingest for the phase-vocoder path (reference pypevoc/AudioInterface.py:8-37).

``wavInfo`` / ``wavLoad`` keep the reference's names and return values (integer PCM samples, mono
1-D); ``wav_pinned`` is what the GPU path wants: the same samples scaled to [-1, 1] as in the
reference's example (examples/WavResynth.py:17-18) in a *pinned* float32 host tensor, which
``PV(x, ...).run_pv(hostbuf=...)`` uploads in chunks that overlap the analysis.  File reading is host
work by nature; nothing here touches the GPU.
"""
import wave

import numpy as np


def wavInfo(fname):
    """(nchannels, sampwidth, framerate, nframes, comptype, compname) (AudioInterface.py:8-13)."""
    wav = wave.open(fname, "r")
    try:
        return wav.getparams()
    finally:
        wav.close()


def wavLoad(fname, startTime=0.0, endTime=None):
    """(framerate, samples): 16-bit PCM as a numpy integer array, one row per channel for stereo
    (AudioInterface.py:15-37; the reference's stereo branch raises, mono is what it supports).
    ``startTime`` / ``endTime`` in seconds."""
    wav = wave.open(fname, "r")
    try:
        nchannels, sampwidth, framerate, nframes = wav.getparams()[:4]
        if sampwidth != 2:
            raise ValueError("wavLoad reads 16-bit PCM only (sample width %d bytes)" % sampwidth)
        first = int(startTime * float(framerate)) if startTime > 0.0 else 0
        first = min(first, nframes)
        wav.setpos(first)
        count = nframes - first if not endTime else max(0, min(int((endTime - startTime) * float(framerate)), nframes - first))
        data = np.frombuffer(wav.readframes(count), dtype="<i2")
    finally:
        wav.close()
    if nchannels == 1:
        return framerate, data.astype(np.int64)
    return framerate, data.reshape(-1, nchannels).T.astype(np.int64)


def wav_pinned(fname, startTime=0.0, endTime=None, channel=0):
    """(framerate, pinned float32 torch tensor in [-1, 1]) of one channel: the input format of the
    streamed GPU path (falls back to pageable memory when no CUDA runtime is present)."""
    import torch
    sr, data = wavLoad(fname, startTime, endTime)
    if data.ndim == 2:
        data = data[channel]
    x = torch.from_numpy((data / float(np.iinfo(np.int16).max)).astype(np.float32))
    if torch.cuda.is_available():
        x = x.pin_memory()
    return sr, x
