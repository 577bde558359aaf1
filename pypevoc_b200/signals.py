"""Synthetic test / benchmark signals.

All generators return **fp32-representable** samples (generated in fp64, scaled to a
0.9 peak, rounded to fp32) so that the CPU oracle (which works on the float64 view of
the same values) and the GPU path (fp32 tensor) see identical inputs.

``harm``        numpy, for tests and small configs
``harm_torch``  same recipe on a torch device (used by bench.py for the 10-minute and
                8-hour configs, where a numpy generator would take minutes)
"""
import numpy as np


def _finish(x):
    x = 0.9 * x / np.max(np.abs(x))
    return x.astype(np.float32)


def harm(sr, dur, f0, nharm, p, sigma, seed, t0=0.0):
    """Harmonic tone with 0.5 % / 5 Hz vibrato plus Gaussian noise (numpy).

    x = sigma*randn + sum_h 0.5*h**-p * sin(h*phi + 2*pi*u_h), phi = 2*pi*cumsum(f0(t))/sr,
    harmonics limited to h*f0*1.005 < 0.49*sr.
    """
    rng = np.random.RandomState(seed)
    n = int(round(sr * dur))
    x = sigma * rng.randn(n)
    t = t0 + np.arange(n) / float(sr)
    f0t = f0 * (1.0 + 0.005 * np.sin(2 * np.pi * 5.0 * t))
    phi = 2 * np.pi * np.cumsum(f0t) / sr
    for h in range(1, nharm + 1):
        if h * f0 * 1.005 >= 0.49 * sr:
            break
        x += 0.5 * h ** (-p) * np.sin(h * phi + 2 * np.pi * rng.rand())
    return _finish(x)


def readme_vibrato(seed=0, sr=44100, dur=1.0):
    """The 3-harmonic vibrato tone of the reference README (README.md:24-64)."""
    rng = np.random.RandomState(seed)
    vibfreq = 5.0
    hamp0 = 0.1 * np.array([1, .5, .3])
    hvib = 1.0 * np.array([.5, 0.1, .9])
    hph = np.array([0, np.pi / 2, np.pi])
    f0, f0vib = 500, 0.01
    n = int(sr * dur)
    sig = np.zeros(n) + 0.01 * (rng.rand(n) - .5)
    t = np.arange(0, dur, 1. / sr)[:n]
    vibsig = np.sin(2 * np.pi * vibfreq * t)
    f0sig = f0 * (1 + f0vib * vibsig)
    for k, ha in enumerate(hamp0):
        phsig = np.cumsum(2 * np.pi * f0sig * (k + 1) / sr)
        sig += ha * (1 + hvib[k] * np.sin(2 * np.pi * vibfreq * t + hph[k])) * np.sin(phsig)
    return sig.astype(np.float32), sr


def two_sines(sr=44100):
    """Signal of the reference's tests/test_pypevoc.py:4-16 (400 Hz + 1200 Hz, 1 s)."""
    t = np.arange(sr) / float(sr)
    xx = 0.1 * np.sin(2.0 * np.pi * 400. * t) + 0.05 * np.sin(2.0 * np.pi * 1200. * t)
    return xx.astype(np.float32), sr


def speech_like_clip(seed, sr=16000, dur=3.0):
    """One speech-like clip of config 3: gliding f0, 3 formant resonances, 4 Hz syllable
    envelope with exact-zero gaps (exercises the zero-frame / divide-by-zero paths)."""
    rng = np.random.RandomState(seed)
    n = int(sr * dur)
    t = np.arange(n) / float(sr)
    f0a = rng.uniform(90, 250)
    glide = rng.uniform(-0.2, 0.2)
    f0t = f0a * (1 + glide * (t / dur - 0.5))
    phi = 2 * np.pi * np.cumsum(f0t) / sr
    nh = int(3800 // f0a)
    x = np.zeros(n)
    for h in range(1, nh + 1):
        fh = h * f0a
        a = 0.0
        for fc in (500., 1500., 2500.):
            q = 5.0
            a += 1.0 / np.sqrt(1 + (q * (fh / fc - fc / fh)) ** 2)
        x += a * np.sin(h * phi + 2 * np.pi * rng.rand())
    x += 0.003 * rng.randn(n) * np.max(np.abs(x))
    # syllable envelope: 4 Hz, raised-cosine bursts separated by >= 100 ms of exact zeros
    ph = (t * 4.0 + rng.rand()) % 1.0
    env = np.where(ph < 0.55, 0.5 * (1 - np.cos(2 * np.pi * ph / 0.55)), 0.0)
    x *= env
    return _finish(x)


def harm_torch(sr, nsamp, f0, nharm, p, sigma, seed, device, t0_samples=0, chunk=1 << 22,
               scale=None):
    """``harm`` recipe evaluated on a torch device in chunks (fp64 phase), returns fp32.

    The vibrato phase is evaluated in closed form so that chunks (and the segments that
    different ranks generate) are phase-continuous: phi(t) = 2*pi*f0*(t - 0.005/(2*pi*5)
    * cos(2*pi*5*t)).
    """
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    hs = [h for h in range(1, nharm + 1) if h * f0 * 1.005 < 0.49 * sr]
    ph0 = torch.rand(len(hs), generator=g, device=device, dtype=torch.float64)
    amp = torch.tensor([0.5 * h ** (-p) for h in hs], device=device, dtype=torch.float64)
    if scale is None:
        # deterministic normalisation (no global max pass): RMS-based head-room
        scale = 0.9 / (3.0 * float(torch.sqrt((amp ** 2).sum() / 2 + sigma ** 2)))
    out = torch.empty(nsamp, device=device, dtype=torch.float32)
    for s in range(0, nsamp, chunk):
        e = min(nsamp, s + chunk)
        t = (torch.arange(s, e, device=device, dtype=torch.float64) + t0_samples) / sr
        cyc = f0 * (t - 0.005 / (2 * np.pi * 5.0) * torch.cos(2 * np.pi * 5.0 * t))
        acc = sigma * torch.randn(e - s, generator=g, device=device, dtype=torch.float64)
        for i, h in enumerate(hs):
            fr = h * cyc + ph0[i]
            fr = fr - torch.floor(fr)
            acc += amp[i] * torch.sin(2 * np.pi * fr)
        out[s:e] = (acc * scale).clamp_(-0.999, 0.999).to(torch.float32)
    return out
