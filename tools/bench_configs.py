#!/usr/bin/env python
"""Per-config throughput of the hot path on one B200 (GPU box): every BASELINE.json config at
its full size, device resident, CUDA-event timed per stage (median of 3 after one warm-up).

    python tools/bench_configs.py [out.json] [--only name,name] [--cfg4-hours H]

Not the bench contract (bench.py is): this is the table DESIGN.md section 7 quotes for the
"named shapes" of BASELINE.json.  cfg3 (clip batch) has no resynthesis in its config.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from pypevoc_b200 import pv as P, signals  # noqa: E402


def ev():
    return torch.cuda.Event(enable_timing=True)


def med(fn, n=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = ev(), ev()
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), r


def run_long(name, sr, nsamp, nfft, hop, npks, gen, resynth=True):
    dev = torch.device("cuda")
    t0 = time.perf_counter()
    x = gen(dev)
    torch.cuda.synchronize()
    tgen = time.perf_counter() - t0
    tb = P.host_tables(sr, nfft, hop)
    out = P.analyze_device(x, sr, nfft, hop, npks, 0.005, tb)
    F = out["f"].shape[1]
    ms_an, _ = med(lambda: P.analyze_device(x, sr, nfft, hop, npks, 0.005, tb, out=out))
    f, mag, ph, realph = (out[k][0] for k in ("f", "mag", "ph", "realph"))
    ms_trk, tr = med(lambda: P.track_device(f, mag))
    nt, npts, last = P.track_counts(tr)
    res = dict(config=name, sr=sr, nsamp=int(nsamp), nfft=nfft, hop=hop, npks=npks, frames=int(F),
               peaks_per_frame=float(out["npk"].double().mean().item()), tracks=nt, points=npts,
               analysis_ms=ms_an, analysis_frames_per_s=F / (ms_an * 1e-3),
               analysis_alg_GBps=(4 * hop + 20 * npks + 8) * F / (ms_an * 1e-3) / 1e9,
               tracking_ms=ms_trk, gen_s=tgen)
    if resynth:
        ms_pack, pk = med(lambda: P.pack_device(f, mag, ph, realph, tr["tid"], None, nt, npts=npts))
        ms_syn, w = med(lambda: P.resynth_device(tr["tid"], pk, sr, hop, nfft, hop, max_end=last))
        tl = pk["tlen"].cpu().numpy().astype(np.int64)
        _, E = P.synth_geometry(last, hop, nfft, hop)
        ps = float((tl[tl >= 3] * hop + 2 * E).sum())
        res.update(pack_ms=ms_pack, resynth_ms=ms_syn, partial_samples=ps,
                   resynth_partial_samples_per_s=ps / (ms_syn * 1e-3), out_samples=int(w.numel()),
                   step_ms=ms_an + ms_trk + ms_pack + ms_syn,
                   step_frames_per_s=F / ((ms_an + ms_trk + ms_pack + ms_syn) * 1e-3))
    return res


def run_clips(name, sr, nclips, nsamp, nfft, hop, npks):
    """cfg3: clips generated on the host for 64 distinct seeds and tiled to the batch size
    (generation of 4096 distinct clips in numpy takes minutes; throughput does not depend on it)."""
    dev = torch.device("cuda")
    base = np.stack([signals.speech_like_clip(1000 + i, sr=sr, dur=nsamp / float(sr)) for i in range(64)])
    x = torch.from_numpy(base).to(dev).repeat((nclips + 63) // 64, 1)[:nclips].contiguous()
    tb = P.host_tables(sr, nfft, hop)
    out = P.analyze_device(x, sr, nfft, hop, npks, 0.005, tb)
    F = out["f"].shape[1]
    ms_an, _ = med(lambda: P.analyze_device(x, sr, nfft, hop, npks, 0.005, tb, out=out))
    ms_trk, tr = med(lambda: P.track_device(out["f"], out["mag"]))
    tot = nclips * F
    return dict(config=name, sr=sr, nclips=nclips, nsamp=int(nsamp), nfft=nfft, hop=hop, npks=npks,
                frames=int(tot), peaks_per_frame=float(out["npk"].double().mean().item()),
                tracks=int(tr["ntracks"].sum().item()), analysis_ms=ms_an,
                analysis_frames_per_s=tot / (ms_an * 1e-3),
                analysis_alg_GBps=(4 * hop + 20 * npks + 8) * tot / (ms_an * 1e-3) / 1e9,
                tracking_ms=ms_trk, step_ms=ms_an + ms_trk, step_frames_per_s=tot / ((ms_an + ms_trk) * 1e-3))


def main():
    args = sys.argv[1:]
    outp = args[0] if args and not args[0].startswith("--") else os.path.join(ROOT, "gpurun_out", "configs.json")
    only = None
    hours = 8.0
    for i, a in enumerate(args):
        if a == "--only":
            only = set(args[i + 1].split(","))
        if a == "--cfg4-hours":
            hours = float(args[i + 1])
    H = signals.harm_torch
    cases = [
        ("cfg1_readme", lambda: run_long("cfg1_readme", 44100, 44100, 2048, 1024, 3,
                                         lambda d: torch.from_numpy(signals.readme_vibrato()[0]).to(d))),
        ("metric_10min", lambda: run_long("metric_10min", 44100, 44100 * 600, 2048, 512, 50,
                                          lambda d: H(44100, 44100 * 600, 220.0, 90, 0.5, 0.01, 1, d, scale=0.25))),
        ("cfg2_10min", lambda: run_long("cfg2_10min", 44100, 44100 * 600, 4096, 512, 50,
                                        lambda d: H(44100, 44100 * 600, 110.0, 150, 0.5, 0.01, 2, d, scale=0.25))),
        ("cfg3_clips", lambda: run_clips("cfg3_clips", 16000, 4096, 48000, 512, 128, 20)),
        ("cfg4_8h", lambda: run_long("cfg4_%gh" % hours, 44100, int(44100 * 3600 * hours), 2048, 256, 100,
                                     lambda d: H(44100, int(44100 * 3600 * hours), 200.0, 100, 0.4, 0.01, 4000, d,
                                                 scale=0.25))),
        ("cfg5_60s", lambda: run_long("cfg5_60s", 48000, 48000 * 60, 8192, 1024, 400,
                                      lambda d: H(48000, 48000 * 60, 55.0, 420, 0.3, 0.001, 5, d, scale=0.25))),
    ]
    results = []
    for name, fn in cases:
        if only and name not in only:
            continue
        t0 = time.perf_counter()
        try:
            r = fn()
        except Exception as e:  # report, keep going
            r = dict(config=name, error=repr(e)[:400])
        r["wall_s"] = time.perf_counter() - t0
        results.append(r)
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(outp), exist_ok=True)
    with open(outp, "w") as fh:
        json.dump(results, fh, indent=1)


if __name__ == "__main__":
    main()
