"""CPU: WAV ingest (pypevoc_b200/audio.py) against the reference's AudioInterface.wavLoad / wavInfo on a
generated 16-bit mono file (reference run only where /root/reference is mounted), plus time ranges and
stereo, which the reference does not handle."""
import importlib
import os
import sys
import wave

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pypevoc_b200 import audio


def _write(path, data, sr, nch):
    w = wave.open(path, "w")
    w.setnchannels(nch); w.setsampwidth(2); w.setframerate(sr)
    w.writeframes(np.asarray(data, dtype="<i2").tobytes())
    w.close()


def test_wavload_mono_matches_reference(tmp_path):
    rng = np.random.RandomState(3)
    pcm = rng.randint(-32768, 32768, 22050).astype(np.int16)
    path = str(tmp_path / "m.wav")
    _write(path, pcm, 22050, 1)
    sr, x = audio.wavLoad(path)
    assert sr == 22050 and np.array_equal(x, pcm.astype(np.int64))
    assert tuple(audio.wavInfo(path))[:4] == (1, 2, 22050, 22050)
    from oracle import ref_loader
    if ref_loader.available():
        ref_loader.load()
        ai = importlib.import_module("pypevoc.AudioInterface")
        rsr, rx = ai.wavLoad(path)
        assert rsr == sr and np.array_equal(np.asarray(rx), x)
        assert tuple(ai.wavInfo(path))[:4] == tuple(audio.wavInfo(path))[:4]
    sr2, seg = audio.wavLoad(path, startTime=0.25, endTime=0.5)
    assert np.array_equal(seg, pcm[5512:5512 + 5512].astype(np.int64))
    sr3, xp = audio.wav_pinned(path)
    assert xp.dtype.is_floating_point and xp.shape == (22050,) and float(xp.abs().max()) <= 1.0001
    assert np.allclose(xp.numpy(), pcm / 32767.0, atol=1e-7)


def test_wavload_stereo_and_errors(tmp_path):
    pcm = np.arange(-50, 50, dtype=np.int16)
    path = str(tmp_path / "s.wav")
    _write(path, pcm, 8000, 2)
    sr, x = audio.wavLoad(path)
    assert x.shape == (2, 50) and np.array_equal(x[0], pcm[0::2]) and np.array_equal(x[1], pcm[1::2])
    assert audio.wav_pinned(path, channel=1)[1].shape == (50,)
    w = wave.open(str(tmp_path / "b.wav"), "w")
    w.setnchannels(1); w.setsampwidth(1); w.setframerate(8000); w.writeframes(b"\x00" * 10); w.close()
    with pytest.raises(ValueError):
        audio.wavLoad(str(tmp_path / "b.wav"))
