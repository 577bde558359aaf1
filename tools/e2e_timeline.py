#!/usr/bin/env python
"""Where does the end-to-end step go?  Host timestamps around the phases of the streamed
pipeline + raw PCIe copy bandwidths.  Run on the GPU box."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypevoc_b200 import PV, signals
import bench
c = bench.CFG
dev = torch.device("cuda", 0)
sr, nfft, hop, npks = c["sr"], c["nfft"], c["hop"], c["npks"]
n = sr * c["seconds"]
xd = signals.harm_torch(sr, n, c["f0"], c["nharm"], c["p"], c["sigma"], c["seed"], dev, scale=0.25)
xh = torch.empty(n, dtype=torch.float32).pin_memory(); xh.copy_(xd); torch.cuda.synchronize()
# raw copies
big = torch.empty(212 << 20, dtype=torch.uint8, device=dev); hbig = torch.empty(212 << 20, dtype=torch.uint8).pin_memory()
for name, fn in (("d2h", lambda: hbig.copy_(big, non_blocking=True)), ("h2d", lambda: big.copy_(hbig, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%s 212 MiB: %.2f ms = %.1f GB/s" % (name, dt * 1e3, (212 << 20) / dt / 1e9))
hb = {}
T = lambda: time.perf_counter()
for it in range(5):
    torch.cuda.synchronize(); t0 = T()
    pv = PV(xh, sr, nfft=nfft, hop=hop, npks=npks, progress=False, device=dev); t1 = T()
    pv.run_pv(hostbuf=hb, stream_tables=bench.E2E_TABLES); t2 = T()
    ss = pv.toSinSum(); ss._ensure_tracks(); t3 = T()
    ss._ensure_packed(); t4 = T()
    w = ss.synth(sr, hop, hostbuf=hb); t5 = T()
    _ = pv.f; t6 = T()
    torch.cuda.synchronize(); t7 = T()
    print("ctor %.2f run_pv(issue) %.2f track(+sync) %.2f pack(+sync) %.2f synth(+d2h sync) %.2f tables %.2f tail %.2f total %.2f ms" % tuple(
        1e3 * v for v in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t7 - t0)))

from pypevoc_b200 import pv as P
P.TRACE = []
torch.cuda.synchronize()
pv = PV(xh, sr, nfft=nfft, hop=hop, npks=npks, progress=False, device=dev)
pv.run_pv(hostbuf=hb, stream_tables=bench.E2E_TABLES)
ss = pv.toSinSum(); ss._ensure_tracks()
P._mark("track done", torch.cuda.current_stream())
w = ss.synth(sr, hop, hostbuf=hb)
torch.cuda.synchronize()
t0 = P.TRACE[0][1]
for label, ev in P.TRACE:
    print("%-16s %7.3f ms" % (label, t0.elapsed_time(ev)))
