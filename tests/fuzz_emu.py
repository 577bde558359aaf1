"""TEST INFRASTRUCTURE -- randomized comparison of the kernels (SIMT-emulator build) with the oracle.

    python tests/fuzz_emu.py track <seed> <cases>       link / scan / resolve kernels vs orc.track
    python tests/fuzz_emu.py analyze <seed> <cases>     analysis kernel vs the oracle run on the kernel's own spectrum
    python tests/fuzz_emu.py synth <seed> <cases>       pack + resynthesis kernels vs orc.synth (random hop, edge, minframes)

tests/test_fuzz_cpu.py runs a short, fixed-seed slice of both.  Known, documented non-bugs are left
out of the generators: simultaneous exact ties of magnitude AND distance in one frame (the reference
orders those by partial index, the kernels by column) and the sign of a zero
imaginary part on an exactly real negative bin (+pi vs -pi)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "emu"))


def track_case(rng):
    K = int(rng.choice([1, 2, 3, 5, 8, 16, 31, 32, 33, 50, 63, 64, 65, 100, 127, 128, 129, 140, 200, 256, 257, 300, 400, 512]))
    F = int(rng.randint(2, 10))
    f = np.zeros((F, K))
    mag = np.zeros((F, K))
    nbase = rng.randint(1, K + 1)
    base = np.sort(rng.uniform(100, 10000, nbase))
    if rng.rand() < 0.3:
        base = 100 + rng.uniform(5, 40) * np.arange(nbase)
    for j in range(F):
        if rng.rand() < 0.1:
            continue
        n = rng.randint(0, nbase + 1)
        sel = np.sort(rng.choice(nbase, n, replace=False))
        fr = base[sel] * (1 + rng.choice([0.0005, 0.005, 0.02]) * rng.randn(n))
        mg = rng.uniform(0.01, 1, n)
        tie = rng.rand()
        if tie < 0.3:
            mg = np.round(mg, 1) + 0.1                      # magnitude ties
        elif tie < 0.5:
            fr = np.round(fr, -1)                           # distance ties (magnitudes distinct)
        order = np.argsort(fr) if rng.rand() < 0.7 else rng.permutation(n)
        fr, mg = fr[order], mg[order]
        cols = np.arange(n) if rng.rand() < 0.6 else np.sort(rng.choice(K, n, replace=False))
        f[j, cols], mag[j, cols] = fr, mg
        if rng.rand() < 0.1 and n:
            mag[j, cols[0]] = 0.0
        if rng.rand() < 0.1 and n:
            f[j, cols[-1]] = -5.0
    return f, mag, float(rng.choice([0.5, 0.5, 0.2, 1.0, 3.0]))


def analyze_case(rng, lognfft=(6, 12)):
    nfft = int(2 ** rng.randint(*lognfft))
    hop = int(rng.randint(1, nfft + 1)) if rng.rand() < 0.5 else nfft // int(rng.choice([1, 2, 4, 8]))
    npks = min(int(rng.choice([1, 2, 3, 5, 20, 50, 64, 100, nfft // 2])), 1024)
    th = float(rng.choice([0.005, 0.0, 0.1, 0.5, 0.9, -0.1]))
    n = nfft + hop * int(rng.randint(1, 7)) + int(rng.randint(0, hop + 1))
    kind = rng.randint(0, 6)
    t = np.arange(n)
    if kind == 0:
        x = rng.randn(n)
    elif kind == 1:
        x = sum(rng.rand() * np.sin(2 * np.pi * rng.uniform(0.001, 0.49) * t + rng.rand() * 6)
                for _ in range(rng.randint(1, 12))) + 1e-3 * rng.randn(n)
    elif kind == 2:
        x = np.zeros(n); x[rng.randint(0, n, size=rng.randint(1, 4))] = 1.0
    elif kind == 3:
        x = np.ones(n) * rng.randn()
    elif kind == 4:
        x = rng.randn(n); x[:n // 2] = 0.0
    else:
        x = np.round(rng.randn(n) * 4) / 4
    x = (x * rng.choice([1e-6, 1.0, 1e3])).astype(np.float32)
    return x, int(rng.choice([8000, 22050, 44100])), nfft, hop, npks, th, int(rng.choice([0, 1, 2, 3]))


def synth_case(rng):
    f, mag, _ = track_case(rng)
    reps = int(rng.randint(1, 5))                           # longer tables so that partials reach minframes
    f = np.concatenate([f * (1 + 0.001 * r) for r in range(reps)])
    mag = np.concatenate([mag] * reps)
    F, K = f.shape
    ph = rng.uniform(-np.pi, np.pi, (F, K))
    realph = ph + rng.uniform(-0.5, 0.5, (F, K))
    nfft = int(rng.choice([256, 512, 1024, 2048]))
    hop_an = nfft // int(rng.choice([1, 2, 4, 8]))
    hop = int(rng.choice([hop_an, hop_an, max(hop_an // 2, 8), hop_an + 37, 64, 200]))
    sr = int(rng.choice([8000, 22050, 44100]))
    f = f * min(1.0, 0.45 * sr / max(f.max(), 1.0))
    return dict(f=f, mag=mag, ph=ph, realph=realph, nfft=nfft, hop_an=hop_an, hop=hop, sr=sr,
                edge=float(rng.choice([1.0, 1.0, 0.5, 0.25, 2.0])), minframes=int(rng.choice([3, 3, 1, 2, 5])))


def check_synth(eh, orc, pu, c):
    tr = eh.track(c["f"], c["mag"])
    ref = orc.track(c["f"], c["mag"])
    assert np.array_equal(tr["tid"][0], ref["tid"])
    nt = int(tr["ntracks"][0])
    if nt == 0:
        return
    pk = eh.track_pack(c["f"], c["mag"], c["ph"], c["realph"], tr["tid"][0], tr["link"][0], nt)
    parts = orc.partials_from_tracks(ref, c["f"], c["mag"], c["ph"], c["realph"])
    w = eh.resynth(tr["tid"][0], pk, c["sr"], c["hop"], c["nfft"], c["hop_an"], edge=c["edge"], minframes=c["minframes"])
    r = orc.synth(parts, c["sr"], c["hop"], c["nfft"], c["hop_an"], edge=c["edge"], minframes=c["minframes"])
    assert w.shape == r.shape, (w.shape, r.shape)
    if np.any(r):
        assert pu.snr_db(w, r) > 100.0
    else:
        assert not np.any(w)


def check_track(eh, orc, case):
    f, mag, mj = case
    return np.array_equal(eh.track(f, mag, maxpitchjmp=mj)["tid"][0], orc.track(f, mag, maxpitchjmp=mj)["tid"])


def check_analyze(eh, orc, pu, case):
    x, sr, nfft, hop, npks, th, run = case
    o = eh.analyze(x, sr, nfft, hop, npks, pkthresh=th, spectra=True, run_frames=run)
    got = {k: o[k][0] for k in ("f", "mag", "ph", "realph", "binno", "totalmag", "npk")}
    oo = orc.analyze(np.zeros(1), sr, nfft=nfft, hop=hop, npks=npks, pkthresh=th, fx_given=o["fx"][0].astype(np.complex64))
    for k in ("ph", "realph"):                               # +-pi on an exactly real negative bin
        flip = np.abs(np.abs(got[k] - oo[k]) - 2 * np.pi) < 1e-6
        got[k] = np.where(flip, oo[k], got[k])
    pu.compare_exact_on_spectrum(got, oo)


if __name__ == "__main__":
    import emu_harness as eh
    from oracle import pv_oracle as orc
    import parity_util as pu
    eh.build()
    what, seed, cases = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rng = np.random.RandomState(seed)
    bad = 0
    for it in range(cases):
        if what == "track":
            c = track_case(rng)
            if not check_track(eh, orc, c):
                bad += 1
                print("MISMATCH case", it, "K", c[0].shape[1], "maxpitchjmp", c[2])
        elif what == "synth":
            c = synth_case(rng)
            try:
                check_synth(eh, orc, pu, c)
            except AssertionError as e:
                bad += 1
                print("MISMATCH case", it, {k: v for k, v in c.items() if np.isscalar(v)}, str(e)[:160])
        else:
            c = analyze_case(rng)
            try:
                check_analyze(eh, orc, pu, c)
            except AssertionError as e:
                bad += 1
                print("MISMATCH case", it, c[2:], str(e)[:160])
    print("done: %d cases, %d mismatches" % (cases, bad))
