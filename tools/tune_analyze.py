#!/usr/bin/env python
"""Tuning run (GPU box): time pvk_analyze alone for several build variants of libpvk.so.
Each variant is loaded in a fresh subprocess through PVK_LIB.  Usage: tune_analyze.py [variant.so ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [  # name, sr, seconds, nfft, hop, npks, f0, nharm, p, sigma, seed, nclips
    ("metric_10min", 44100, 600, 2048, 512, 50, 220.0, 90, 0.5, 0.01, 1),
    ("cfg2_10min", 44100, 600, 4096, 512, 50, 110.0, 150, 0.5, 0.01, 2),
    ("cfg4_20min", 44100, 1200, 2048, 256, 100, 200.0, 100, 0.4, 0.01, 4000),
    ("cfg5_60s", 48000, 60, 8192, 1024, 400, 55.0, 420, 0.3, 0.001, 5),
]


def child():
    sys.path.insert(0, ROOT)
    import torch
    from pypevoc_b200 import pv as P, signals
    dev = torch.device("cuda")
    res = {}
    runs = [int(r) for r in os.environ.get("PVK_RUNS", "0").split(",")]
    only = [c for c in os.environ.get("PVK_CASES", "").split(",") if c]
    for name, sr, sec, nfft, hop, npks, f0, nh, p, sg, seed in CASES:
        if only and name not in only:
            continue
        x = signals.harm_torch(sr, sr * sec, f0, nh, p, sg, seed, dev, scale=0.25)
        tb = P.host_tables(sr, nfft, hop)
        for run in runs:
            out = P.analyze_device(x, sr, nfft, hop, npks, 0.005, tb, run_frames=run)
            torch.cuda.synchronize()
            F = out["f"].shape[1]
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                P.analyze_device(x, sr, nfft, hop, npks, 0.005, tb, run_frames=run, out=out)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = sorted(ts)[len(ts) // 2]
            res["%s run=%d" % (name, run)] = dict(ms=round(ms, 4), mframes_s=round(F / ms / 1e3, 2),
                                                 npk=float(out["npk"].float().mean()))
        del x
    print(json.dumps(res))


if __name__ == "__main__":
    if os.environ.get("PVK_CHILD"):
        child()
    else:
        for so in sys.argv[1:] or [os.path.join(ROOT, "pypevoc_b200", "libpvk.so")]:
            env = dict(os.environ, PVK_CHILD="1", PVK_LIB=so)
            out = subprocess.run([sys.executable, __file__], env=env, stdout=subprocess.PIPE, text=True).stdout
            print(os.path.basename(so))
            try:
                for k, v in json.loads(out.strip().splitlines()[-1]).items():
                    print("   %-28s %8.3f ms %8.2f Mframes/s  (%.1f peaks/frame)" % (k, v["ms"], v["mframes_s"], v["npk"]))
            except Exception:
                print(out[-2000:])
