"""GPU parity tests (run on the B200 box): pypevoc_b200 -> libpvk.so (C ABI) vs the reference's
golden outputs (tests/golden, produced by the real PyPeVoc) and vs the CPU oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import pv_oracle as orc
from golden_util import CASES, case_golden, case_signal, pv_kwargs
import parity_util as pu

pytestmark = pytest.mark.gpu
NAMES = sorted(CASES)


@pytest.fixture(scope="module")
def pvmod():
    import pypevoc_b200
    from pypevoc_b200 import _lib
    _lib.lib()           # fails loudly if libpvk.so is missing
    return pypevoc_b200


def _run(pvmod, name, **extra):
    x, sr = case_signal(name)
    kw = pv_kwargs(name)
    kw.update(extra)
    pv = pvmod.PV(x, sr, progress=False, **kw)
    pv.run_pv()
    return pv, x, sr, kw


@pytest.mark.parametrize("name", NAMES)
def test_analysis_vs_reference_golden(pvmod, name):
    pv, x, sr, kw = _run(pvmod, name)
    g = case_golden(name)
    assert pv.nframes == int(g["nframes"])
    o = orc.analyze(x, sr, margins=True, **kw)
    got = dict(f=pv.f, mag=pv.mag, ph=pv.ph, realph=pv.realph, binno=pv.binno, totalmag=pv.totalmag)
    ref = {k: g[k] for k in ("f", "mag", "ph", "realph", "binno", "totalmag")}
    rep = pu.compare_analysis(got, ref, sr, kw["nfft"], margin=o["margin"])
    assert np.allclose(pv.t, g["t"], rtol=0, atol=1e-15)
    print(name, rep)


@pytest.mark.parametrize("name", NAMES)
def test_kernel_logic_exact_on_own_spectrum(pvmod, name):
    from pypevoc_b200.pv import analyze_device
    x, sr = case_signal(name)
    kw = pv_kwargs(name)
    pv = pvmod.PV(x, sr, progress=False, **kw)
    d = analyze_device(pv._xd, sr, pv.nfft, pv.hop, pv.npeaks, pv.peakthresh, pv._tb, spectra=True)
    got = {k: d[k][0].cpu().numpy() for k in ("f", "mag", "ph", "realph", "binno", "npk", "totalmag")}
    fx = d["fx"][0].cpu().numpy()
    o = orc.analyze(np.zeros(1), sr, nfft=pv.nfft, hop=pv.hop, npks=pv.npeaks, pkthresh=pv.peakthresh,
                    fx_given=fx)
    pu.compare_exact_on_spectrum(got, o)
    # FFT accuracy against the reference's calc_fft_frame (fp64): fp32 round-off only
    g = case_golden(name)
    err = np.abs(fx[g["fx_frames"]] - g["fx"]).max()
    assert err < 3e-6 * max(np.abs(g["fx"]).max(), 1e-30), err


@pytest.mark.parametrize("name", NAMES)
def test_tracking_bit_exact_on_reference_tables(pvmod, name):
    from pypevoc_b200.pv import track_device, pack_device
    g = case_golden(name)
    dev = torch.device("cuda")
    f, mag, ph, rph = (torch.from_numpy(np.ascontiguousarray(g[k])).to(dev) for k in ("f", "mag", "ph", "realph"))
    tr = track_device(f, mag)
    assert np.array_equal(tr["tid"].cpu().numpy(), g["tid"])
    nt = int(tr["ntracks"][0].item())
    assert nt == len(g["st"])
    pk = pack_device(f, mag, ph, rph, tr["tid"], tr["link"], nt)
    assert np.array_equal(pk["tstart"].cpu().numpy(), g["st"])
    assert np.array_equal((pk["tstart"] + pk["tlen"] - 1).cpu().numpy(), g["end"])


@pytest.mark.parametrize("name", NAMES)
def test_tosinsum_matches_oracle_on_gpu_tables(pvmod, name):
    pv, x, sr, kw = _run(pvmod, name)
    ss = pv.toSinSum()
    o = orc.track(pv.f, pv.mag)
    assert np.array_equal(ss.track_ids, o["tid"])
    assert ss.st == o["st"].tolist() and ss.end == o["end"].tolist()
    parts = orc.partials_from_tracks(o, pv.f, pv.mag, pv.ph, pv.realph)
    assert len(ss.partial) == len(parts)
    for i in (0, len(parts) // 2, len(parts) - 1):
        p = ss.partial[i]
        assert p.start_idx == parts[i]["start_idx"]
        for k in ("f", "mag", "ph", "realph"):
            assert np.array_equal(getattr(p, k), parts[i][k])


@pytest.mark.parametrize("name", NAMES)
def test_resynthesis_vs_reference(pvmod, name):
    from pypevoc_b200.pv import track_device, pack_device, resynth_device
    g = case_golden(name)
    _, sr = case_signal(name)
    kw = pv_kwargs(name)
    dev = torch.device("cuda")
    f, mag, ph, rph = (torch.from_numpy(np.ascontiguousarray(g[k])).to(dev) for k in ("f", "mag", "ph", "realph"))
    tr = track_device(f, mag)
    pk = pack_device(f, mag, ph, rph, tr["tid"], tr["link"], int(tr["ntracks"][0].item()))
    hop_an = int(g["hop"])
    hops = CASES[name]["synth_hops"]
    if hops:
        refs = {h: g["synth_%d" % h] for h in hops}
    else:
        o = orc.track(g["f"], g["mag"])
        parts = orc.partials_from_tracks(o, g["f"], g["mag"], g["ph"], g["realph"])
        refs = {hop_an: orc.synth(parts, sr, hop_an, kw["nfft"], hop_an)}
    for h, ref in refs.items():
        w = resynth_device(tr["tid"], pk, sr, h, kw["nfft"], hop_an).cpu().numpy()
        assert w.shape == ref.shape
        snr = pu.snr_db(w, ref)
        print(name, h, "SNR vs reference resynthesis: %.1f dB" % snr)
        assert snr > 90.0, snr       # stated bound: > 90 dB (fp32 cosine, MUFU)


def test_end_to_end_pipeline_snr(pvmod):
    """PV -> toSinSum -> synth on the GPU vs the same chain in the oracle (fp64 numpy)."""
    name = "metric_1s"
    pv, x, sr, kw = _run(pvmod, name)
    w = pv.toSinSum().synth(sr, pv.hop)
    g = case_golden(name)
    ref = g["synth_%d" % pv.hop]
    assert w.shape == ref.shape and w.dtype == np.float64
    snr = pu.snr_db(w, ref)
    print("end-to-end SNR vs reference: %.1f dB" % snr)
    assert snr > 60.0, snr


# ---------------------------------------------------------------------------- edge cases
@pytest.mark.parametrize("nfft", [64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_all_fft_sizes_and_odd_hop(pvmod, nfft):
    from pypevoc_b200.pv import analyze_device
    rng = np.random.RandomState(nfft)
    n = nfft * 6 + 37
    t = np.arange(n)
    x = (0.1 * rng.randn(n) + np.sin(2 * np.pi * 0.031 * t) + 0.5 * np.sin(2 * np.pi * 0.12 * t)).astype(np.float32)
    for hop, npks in ((nfft // 4, 20), (nfft // 4 + 1, 7), (nfft, 3)):
        pv = pvmod.PV(x, 44100, nfft=nfft, hop=hop, npks=npks, progress=False)
        d = analyze_device(pv._xd, 44100, nfft, hop, npks, pv.peakthresh, pv._tb, spectra=True)
        got = {k: d[k][0].cpu().numpy() for k in ("f", "mag", "ph", "realph", "binno", "npk", "totalmag")}
        o = orc.analyze(np.zeros(1), 44100, nfft=nfft, hop=hop, npks=npks, fx_given=d["fx"][0].cpu().numpy())
        pu.compare_exact_on_spectrum(got, o)
        ref = orc.analyze(x, 44100, nfft=nfft, hop=hop, npks=npks, margins=True)
        pu.compare_analysis(got, ref, 44100, nfft, margin=ref["margin"])


@pytest.mark.parametrize("kind", ["zeros", "impulse", "impulse_train", "dc", "silence_then_noise", "white", "flat_th_neg"])
def test_degenerate_signals(pvmod, kind):
    from pypevoc_b200.pv import analyze_device
    rng = np.random.RandomState(11)
    nfft, hop, npks, th = 1024, 256, 30, 0.005
    x = np.zeros(nfft * 6, dtype=np.float32)
    if kind == "impulse":
        x[3000] = 1.0
    elif kind == "impulse_train":
        x[::64] = 1.0
    elif kind == "dc":
        x[:] = 1.0
    elif kind == "silence_then_noise":
        x[2500:] = 0.1 * rng.randn(len(x) - 2500)
    elif kind == "white":
        x[:] = rng.randn(len(x))
        npks = 5                      # far more candidates than npks: radix select path
    elif kind == "flat_th_neg":
        x[700] = 1.0
        x[701] = 1e-4
        th = 0.5                      # max*th < min: non-peaks qualify (PeakFinder.py:174-187)
    pv = pvmod.PV(x, 44100, nfft=nfft, hop=hop, npks=npks, pkthresh=th, progress=False)
    d = analyze_device(pv._xd, 44100, nfft, hop, npks, th, pv._tb, spectra=True)
    got = {k: d[k][0].cpu().numpy() for k in ("f", "mag", "ph", "realph", "binno", "npk", "totalmag")}
    o = orc.analyze(np.zeros(1), 44100, nfft=nfft, hop=hop, npks=npks, pkthresh=th, fx_given=d["fx"][0].cpu().numpy())
    pu.compare_exact_on_spectrum(got, o)
    if kind == "zeros":
        assert got["npk"].sum() == 0 and np.all(got["totalmag"] == 0)


def test_short_and_empty_inputs(pvmod):
    pv = pvmod.PV(np.zeros(100), 44100, nfft=1024, progress=False)
    pv.run_pv()
    assert pv.nframes == 0 and pv.f.shape == (0,) and pv.totalmag == [] and pv.t.shape == (0,)
    pv = pvmod.PV(np.zeros(1024), 44100, nfft=1024, progress=False)     # nsamp == nfft: frame excluded (:224)
    pv.run_pv()
    assert pv.nframes == 0
    pv = pvmod.PV(np.ones(1024 + 3 * 512), 44100, nfft=1024, hop=512, progress=False)
    pv.run_pv()
    assert pv.nframes == 3 and pv.f.shape == (3, 20) and pv.f.dtype == np.float64
    assert isinstance(pv.totalmag, list) and len(pv.totalmag) == 3


def test_api_errors(pvmod):
    with pytest.raises(ValueError):
        pvmod.PV(np.zeros(5000), 44100, nfft=1000)
    with pytest.raises(ValueError):
        pvmod.PV(np.zeros(5000), 44100, nfft=1024, npks=5000)
    from pypevoc_b200 import _lib
    L = _lib.lib()
    assert L.pvk_analyze_tables_bytes(1000) < 0
    st = L.pvk_analyze(None, 1, 0, 10, None, None, None, None, 1000, 1, 1, 0.0, 0.0, 0.0, 0, 1, 1, 0,
                       None, None, None, None, None, None, None, None, None)
    assert st != 0 and b"power of two" in L.pvk_last_error()


# ---------------------------------------------------------------------------- size-independent properties
def test_segment_sharding_is_bit_exact(pvmod):
    """Frames [j0, j1) analysed from samples [(j0-1)*hop, ...) with one warm-up frame equal the
    unsharded result bit for bit, at the 10-minute config's parameters."""
    from pypevoc_b200 import signals
    from pypevoc_b200.pv import analyze_device, host_tables
    sr, nfft, hop, npks = 44100, 4096, 512, 50
    x = signals.harm_torch(sr, sr * 20, 110, 150, 0.5, 0.01, 2, torch.device("cuda"))
    tb = host_tables(sr, nfft, hop)
    full = analyze_device(x, sr, nfft, hop, npks, 0.005, tb)
    F = full["f"].shape[1]
    for j0 in (1, F // 3, F - 5):
        seg = analyze_device(x[(j0 - 1) * hop:], sr, nfft, hop, npks, 0.005, tb, frame0=1, prev_zero=False)
        for k in ("f", "mag", "ph", "realph", "binno", "npk", "totalmag"):
            assert torch.equal(seg[k][0], full[k][0, j0:]), (j0, k)
    # run length must not change results
    for run in (1, 7, 1000):
        o = analyze_device(x, sr, nfft, hop, npks, 0.005, tb, run_frames=run)
        for k in ("f", "binno", "npk"):
            assert torch.equal(o[k], full[k]), (run, k)


def test_clip_batch_equals_single_clips(pvmod):
    from pypevoc_b200 import signals
    from pypevoc_b200.pv import analyze_device, host_tables, track_device
    sr, nfft, hop, npks = 16000, 512, 128, 20
    clips = np.stack([signals.speech_like_clip(1000 + i) for i in range(6)])
    xd = torch.from_numpy(clips).cuda()
    tb = host_tables(sr, nfft, hop)
    batch = analyze_device(xd, sr, nfft, hop, npks, 0.005, tb)
    trb = track_device(batch["f"], batch["mag"])
    for i in range(6):
        one = analyze_device(xd[i], sr, nfft, hop, npks, 0.005, tb)
        for k in ("f", "mag", "ph", "realph", "binno", "npk", "totalmag"):
            assert torch.equal(one[k][0], batch[k][i]), (i, k)
        tr1 = track_device(one["f"][0], one["mag"][0])
        assert torch.equal(tr1["tid"], trb["tid"][i])
        o = orc.track(one["f"][0].cpu().numpy(), one["mag"][0].cpu().numpy())
        assert np.array_equal(tr1["tid"].cpu().numpy(), o["tid"])


def test_full_size_pipeline_properties(pvmod):
    """60 s at the metric config (nfft 2048 / hop 512 / npks 50): structural invariants that do
    not need the (slow) oracle: rows zero padded and sorted by bin, tracks are runs of
    consecutive frames with unique ids per frame, resynthesis is deterministic and block-range
    rendering equals the full rendering."""
    from pypevoc_b200 import signals
    from pypevoc_b200.pv import resynth_device
    sr, nfft, hop, npks = 44100, 2048, 512, 50
    x = signals.harm_torch(sr, sr * 60, 220, 90, 0.5, 0.01, 1, torch.device("cuda"))
    pv = pvmod.PV(x, sr, nfft=nfft, hop=hop, npks=npks, progress=False)
    pv.run_pv()
    d = pv.device_tables
    npk = d["npk"].cpu().numpy()
    binno = pv.binno
    assert pv.nframes == orc.n_frames(len(x), nfft, hop)
    col = np.arange(npks)[None, :]
    assert np.all((binno > 0) == (col < npk[:, None]))
    assert np.all((np.diff(binno, axis=1) > 0) | (col[:, 1:] >= npk[:, None]))
    assert np.all((pv.f > 0) == (col < npk[:, None]))
    ss = pv.toSinSum()
    tid = ss.track_ids
    h = ss._host_tracks()
    assert np.array_equal(np.bincount(tid[tid >= 0], minlength=len(h["tlen"])), h["tlen"])
    fr, cc = np.nonzero(tid >= 0)
    ids = tid[fr, cc]
    assert np.array_equal(np.sort(np.unique(ids)), np.arange(len(h["tlen"])))
    # consecutive frames: last - first + 1 == count
    first = np.full(len(h["tlen"]), 1 << 30); last = np.zeros(len(h["tlen"]), dtype=np.int64)
    np.minimum.at(first, ids, fr); np.maximum.at(last, ids, fr)
    assert np.array_equal(last - first + 1, h["tlen"]) and np.array_equal(first, h["tstart"])
    # ids are numbered in creation order: start frames are non-decreasing
    assert np.all(np.diff(h["tstart"]) >= 0)
    w1 = ss.synth(sr, hop, to_host=False)
    w2 = ss.synth(sr, hop, to_host=False)
    assert torch.equal(w1, w2)
    tr, pk = ss._trk, ss._pk
    nblk = (len(w1) + hop - 1) // hop
    cut = nblk // 3
    pa = resynth_device(tr["tid"], pk, sr, hop, nfft, hop, block0=0, nblocks=cut)
    pb = resynth_device(tr["tid"], pk, sr, hop, nfft, hop, block0=cut, nblocks=nblk - cut)
    assert torch.equal(torch.cat([pa, pb]), w1)
    assert torch.isfinite(w1).all() and float(w1.abs().max()) > 0.05


def test_streamed_pipeline_equals_plain_calls(pvmod):
    """run_pv(hostbuf=...) / synth(hostbuf=...) (chunked launches, copies overlapped on side
    streams, pinned host results) give bit-identical results to the one-launch calls."""
    from pypevoc_b200 import signals
    x = signals.harm(44100, 3.0, 220, 90, 0.5, 0.01, 7)
    sr = 44100
    pv0 = pvmod.PV(x, sr, nfft=2048, hop=512, npks=50, progress=False)
    pv0.run_pv()
    w0 = pv0.toSinSum().synth(sr, 512)
    xh = torch.from_numpy(x).pin_memory()
    hb = {}
    for rep in range(2):                       # second pass reuses the pinned buffers
        pv1 = pvmod.PV(xh, sr, nfft=2048, hop=512, npks=50, progress=False)
        pv1.run_pv(hostbuf=hb, chunks=3)
        ss1 = pv1.toSinSum()
        w1 = ss1.synth(sr, 512, hostbuf=hb, chunks=3)
        for k in ("f", "mag", "ph", "realph", "binno"):
            assert np.array_equal(getattr(pv0, k), getattr(pv1, k)), k
        assert np.array_equal(np.asarray(pv0.totalmag), np.asarray(pv1.totalmag))
        assert w1.shape == w0.shape and np.array_equal(w0, w1)
        assert pv1.d2h_bytes == pv1.nframes * (5 * 50 + 1) * 8 and ss1.d2h_bytes == w1.nbytes


@pytest.mark.parametrize("world", [2, 5])
def test_sharded_pipeline_equals_unsharded(pvmod, world):
    """Every 'rank' (run one after the other on this GPU, the all_gathers replaced by their
    definition) analyses its window, links it locally and renders its own blocks from LOCAL
    tables: own rows, global track ids and the concatenated signal equal the unsharded run bit
    for bit (pypevoc_b200/dist.py)."""
    from pypevoc_b200 import signals, dist as D
    from pypevoc_b200 import pv as P
    sr, nfft, hop, npks = 44100, 2048, 512, 50
    x = signals.harm(sr, 4.0, 220, 90, 0.5, 0.02, 9)
    x[int(2.2 * sr):int(2.6 * sr)] = 0.0                      # a silent gap: partials end and restart
    pv0 = pvmod.PV(x, sr, nfft=nfft, hop=hop, npks=npks, progress=False)
    pv0.run_pv()
    ss0 = pv0.toSinSum()
    w0 = ss0.synth(sr, hop)
    tid0 = ss0.track_ids
    plans = D.plan_segments(len(x), nfft, hop, world)
    locs, summ, tabs = [], [], []
    for p in plans:
        xs = torch.from_numpy(x[p["sample0"]:p["sample0"] + p["nsamp"]]).cuda()
        a = P.analyze_device(xs, sr, nfft, hop, npks, 0.005, pv0._tb, frame0=p["frame0"], nframes=p["nframes"],
                             prev_zero=p["prev_zero"])
        tab = {k: a[k][0] for k in ("f", "mag", "ph", "realph")}
        own = slice(p["own0"], p["own0"] + p["nown"])
        for k in ("f", "mag", "ph", "realph"):
            assert np.array_equal(tab[k][own].cpu().numpy(), getattr(pv0, k)[p["j0"]:p["j1"]]), k
        tr = P.track_device(tab["f"], tab["mag"])
        locs.append((tab, tr))
        summ.append(D.local_summary(tr["tid"], p).cpu().numpy())
    summ = np.stack(summ)
    bases, gprevs, ntot, max_end = D.resolve_ids(summ, npks)
    assert ntot == len(ss0.st) and max_end == max(ss0.end)
    rows, sig = [], []
    for r, p in enumerate(plans):
        tab, tr = locs[r]
        rows.append(D.global_ids(tr["tid"], p, bases[r], gprevs[r], int(summ[r, 2 * npks]), int(summ[r, 2 * npks + 1])))
        pk = P.pack_device(tab["f"], tab["mag"], tab["ph"], tab["realph"], tr["tid"], None, int(tr["ntracks"][0].item()))
        sig.append(D.resynth_local(tr["tid"], pk, p, plans, max_end, sr, hop, nfft, hop).cpu().numpy())
        # the overlapped variant: block range from LOCAL knowledge, cut once max_end is known
        last = P.track_counts(tr)[2]
        ll = last + p["w0"] if last >= 0 else -1
        wl = D.resynth_local(tr["tid"], pk, p, plans, None, sr, hop, nfft, hop, local_last=ll)
        b0 = D.render_range_local(p, plans, ll, hop, nfft, hop)[0]
        n, s0 = D.trim_local(wl.numel(), b0, p, plans, max_end, hop, nfft, hop)
        assert np.array_equal(wl[:n].cpu().numpy(), sig[-1]), r
    table = torch.cat(rows)
    assert np.array_equal(table.cpu().numpy(), tid0)
    # the same numbering through the libpvk segment kernels (summary -> [all_gather] -> resolve -> rename)
    allv = torch.cat([D.segment_summary_device(locs[r][1]["tid"].contiguous(), p) for r, p in enumerate(plans)])
    for r, p in enumerate(plans):
        st = D.segment_rename_device(locs[r][1]["tid"].contiguous(), p, world, allv, max(q["own0"] for q in plans))
        assert st["ntracks"] == ntot and st["max_end"] == max_end
        assert torch.equal(st["tid_own"], rows[r]), r
    # ... and through the fused rename + gather kernel: every segment stores its renamed rows into
    # the tables of all ranks (here `world` local tensors stand in for local + peer memory)
    tables = [torch.full((tid0.shape[0], npks), -7, dtype=torch.int32, device="cuda") for _ in range(world)]
    for r, p in enumerate(plans):
        prm = D.segment_rename_push_device(locs[r][1]["tid"].contiguous(), p, world, allv, tables,
                                           max(q["own0"] for q in plans))
        assert prm.cpu().numpy()[3] == ntot
    for t in tables:
        assert np.array_equal(t.cpu().numpy(), tid0)
    tstart, tlen = P.spans_device(table.contiguous(), ntot)
    assert tstart.cpu().numpy().tolist() == ss0.st
    assert (tstart + tlen - 1).cpu().numpy().tolist() == ss0.end
    w = np.concatenate(sig)
    assert w.shape == w0.shape
    assert np.array_equal(w, w0)


def test_chunked_run_pv_under_a_device_budget_is_bit_identical(pvmod):
    """run_pv(device_budget=...) (signals larger than a device-memory budget; the reference reads
    the whole file, AudioInterface.py:15-37): frame-aligned chunks with one warm-up frame each,
    tables streamed into pinned host memory -- equal to the one-shot run bit for bit, and toSinSum
    (which uploads the host tables) gives the same partials and the same resynthesis."""
    from pypevoc_b200 import signals
    sr = 44100
    x = signals.harm(sr, 4.0, 220, 60, 0.5, 0.02, 21)
    x[int(1.5 * sr):int(1.7 * sr)] = 0.0
    pv0 = pvmod.PV(x, sr, nfft=2048, hop=512, npks=40, progress=False)
    pv0.run_pv(refine=True)
    ss0 = pv0.toSinSum()
    w0 = ss0.synth(sr, 512)
    xh = torch.from_numpy(x).pin_memory()
    for budget in (1 << 20, 3 << 20, 1 << 30):              # ~120-frame chunks ... one chunk
        pv = pvmod.PV(xh, sr, nfft=2048, hop=512, npks=40, progress=False)
        hb = {}
        pv.run_pv(hostbuf=hb, device_budget=budget, refine=True)
        assert pv.nframes == pv0.nframes and pv._devout is None
        if budget == 1 << 20:
            assert pv.chunk_frames * 2 < pv.nframes and hb["f"].shape[0] >= pv.nframes      # several chunks
        for k in ("f", "mag", "ph", "realph", "binno", "fine_pos", "fine_val", "t"):
            assert np.array_equal(getattr(pv, k), getattr(pv0, k)), (budget, k)
        assert np.array_equal(np.asarray(pv.totalmag), np.asarray(pv0.totalmag))
        ss = pv.toSinSum()
        assert np.array_equal(ss.track_ids, ss0.track_ids)
        assert np.array_equal(ss.synth(sr, 512), w0)
    with pytest.raises(ValueError):
        pvmod.PV(x, sr, nfft=2048, hop=512, npks=40, progress=False).run_pv(hostbuf={}, device_budget=1 << 20)


def test_tosinsum_honours_host_edited_tables(pvmod):
    """The reference tracks self.f / self.mag (PVAnalysis.py:320): edits made on the host arrays
    (in place or by assignment) change the partials here too."""
    from pypevoc_b200 import signals
    x, sr = signals.two_sines()
    pv = pvmod.PV(x, sr, nfft=2048, hop=512, npks=10, progress=False)
    pv.run_pv()
    n0 = len(pv.toSinSum().st)
    keep = pv.f < 800.0
    pv.mag[~keep] = 0.0                                      # in-place edit of a table that was read
    pv.f = np.where(keep, pv.f, 0.0)                         # assignment
    ss = pv.toSinSum()
    tr = orc.track(pv.f, pv.mag)
    assert np.array_equal(ss.track_ids, tr["tid"]) and len(ss.st) < n0


def test_streamed_tables_are_downloaded_lazily_and_bit_identical(pvmod):
    """run_pv(hostbuf, stream_tables=...): the named tables stream into pinned memory, the others are
    fetched from the device on first access; all equal the plain run bit for bit."""
    from pypevoc_b200 import signals
    sr = 44100
    x = signals.harm(sr, 3.0, 220, 60, 0.5, 0.02, 22)
    pv0 = pvmod.PV(x, sr, nfft=2048, hop=512, npks=40, progress=False)
    pv0.run_pv()
    pv = pvmod.PV(torch.from_numpy(x).pin_memory(), sr, nfft=2048, hop=512, npks=40, progress=False)
    hb = {}
    pv.run_pv(hostbuf=hb, stream_tables=("f", "mag", "ph"))
    assert sorted(hb) == ["f", "mag", "ph"] and pv.d2h_bytes == pv.nframes * 3 * 40 * 8
    for k in ("f", "mag", "ph", "realph", "binno"):
        assert np.array_equal(getattr(pv, k), getattr(pv0, k)), k
    assert np.array_equal(np.asarray(pv.totalmag), np.asarray(pv0.totalmag))
    with pytest.raises(ValueError):
        pv.run_pv(hostbuf=hb, stream_tables=("f", "nope"))


@pytest.mark.parametrize("name,sr,sec,nfft,hop,npks,f0,nharm,p,sigma,seed", [
    ("metric_60s", 44100, 60, 2048, 512, 50, 220.0, 90, 0.5, 0.01, 1),        # bench.py's workload signal, first 60 s
    ("cfg5_20s", 48000, 20, 8192, 1024, 400, 55.0, 420, 0.3, 0.001, 5),       # BASELINE configs[4] signal, first 20 s
])
def test_full_size_config_prefix_vs_oracle(pvmod, name, sr, sec, nfft, hop, npks, f0, nharm, p, sigma, seed):
    """Parity at the configs' own shapes (not the <= 1 s shape-alikes of the goldens): a 60 s / 20 s
    prefix of the bench signals through PV.run_pv -> toSinSum on the GPU against orc.analyze (fp64
    numpy, pinned to the reference) with per-frame decision margins, and orc.track on the GPU's own
    tables: bins bit-exact in every frame above the fp32 margin, values inside the north-star
    tolerances, track ids / st / end bit-exact.  The number of sub-margin frames is printed."""
    from pypevoc_b200 import signals
    x = signals.harm_torch(sr, sr * sec, f0, nharm, p, sigma, seed, torch.device("cuda"), scale=0.25)
    x.mul_(float(np.float32(0.9 / float(x.abs().max().item()))))
    xh = x.cpu().numpy()
    pv = pvmod.PV(x, sr, nfft=nfft, hop=hop, npks=npks, progress=False)
    pv.run_pv()
    o = orc.analyze(xh, sr, nfft=nfft, hop=hop, npks=npks, margins=True)
    assert pv.nframes == o["nframes"]
    got = dict(f=pv.f, mag=pv.mag, ph=pv.ph, realph=pv.realph, binno=pv.binno, totalmag=pv.totalmag)
    rep = pu.compare_analysis(got, o, sr, nfft, margin=o["margin"])
    sub = int((o["margin"] <= pu.MARGIN_FP32).sum())
    ss = pv.toSinSum()
    tr = orc.track(pv.f, pv.mag)
    assert np.array_equal(ss.track_ids, tr["tid"])
    assert ss.st == tr["st"].tolist() and ss.end == tr["end"].tolist()
    print("%s: %d frames, %d sub-margin frames (%d with different bins), %d partials, %s" % (
        name, pv.nframes, sub, rep["mismatched"], len(ss.st), rep))


def test_fused_back_half_equals_staged_calls(pvmod):
    """SinSum.synth on fresh partials queues link + pack + resynthesis with upper-bound sizes and ONE
    read-back at the end (pvk_track_pack_dev / pvk_resynth_dev); the result must equal the staged calls
    (counts read first, exact sizes) bit for bit -- also when the signal ends in silence (the rendered
    upper-bound tail is cut) and when there is no partial at all."""
    from pypevoc_b200 import signals
    from pypevoc_b200 import pv as P
    sr = 44100
    x = signals.harm(sr, 3.0, 220, 60, 0.5, 0.02, 23)
    x[int(2.2 * sr):] = 0.0
    for hop_s, edge, minframes in ((512, 1.0, 3), (384, 0.5, 2), (256, 1.0, 3)):
        pv = pvmod.PV(x, sr, nfft=2048, hop=512, npks=40, progress=False)
        pv.run_pv()
        ss1 = pv.toSinSum()
        w1 = ss1.synth(sr, hop_s, edge=edge, minframes=minframes)          # fused
        ss2 = pv.toSinSum()
        ss2._ensure_packed()                                                # staged: counts first
        w2 = ss2.synth(sr, hop_s, edge=edge, minframes=minframes)
        assert w1.shape == w2.shape and np.array_equal(w1, w2)
        assert np.array_equal(ss1.track_ids, ss2.track_ids) and ss1.st == ss2.st and ss1.end == ss2.end
        d1, d2 = ss1.device_tracks, ss2.device_tracks
        for k in ("toff", "pf", "pmag", "pph", "prealph"):
            assert torch.equal(d1[k], d2[k]), k
        assert np.array_equal(ss1.synth(sr, hop_s, edge=edge, minframes=minframes), w1)   # second call: staged path
        hb = {}
        ss3 = pv.toSinSum()
        w3 = np.array(ss3.synth(sr, hop_s, edge=edge, minframes=minframes, hostbuf=hb))     # fused + streamed download
        assert w3.shape == w1.shape and np.array_equal(w3, w1) and ss3.st == ss1.st
        assert np.array_equal(np.array(ss3.synth(sr, hop_s, edge=edge, minframes=minframes, hostbuf=hb)), w1)  # staged + streamed
    pz = pvmod.PV(np.zeros(8192, dtype=np.float32), sr, nfft=2048, hop=512, npks=10, progress=False)
    pz.run_pv()
    with pytest.raises(ValueError):
        pz.toSinSum().synth(sr, 512)
    # more partials than the speculative capacity: sized exactly afterwards, same result
    cap = P.PACK_SPEC_CAP
    try:
        P.PACK_SPEC_CAP = 16
        pv = pvmod.PV(x, sr, nfft=2048, hop=512, npks=40, progress=False)
        pv.run_pv()
        w3 = pv.toSinSum().synth(sr, 512)
    finally:
        P.PACK_SPEC_CAP = cap
    pv = pvmod.PV(x, sr, nfft=2048, hop=512, npks=40, progress=False)
    pv.run_pv()
    assert np.array_equal(w3, pv.toSinSum().synth(sr, 512))
