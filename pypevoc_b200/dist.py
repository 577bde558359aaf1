"""Segment sharding of one long signal over the GPUs of a box.

One process per GPU (torchrun).  Rank g owns the contiguous, hop-aligned frame range
[j0, j1) and reads samples [(j0-2)*hop, (j1-1)*hop + nfft): the nfft-hop overlap with its
right neighbour plus two extra hops on the left, so that it can recompute
  * frame j0-2 as a pure warm-up (its spectrum feeds the phase difference of frame j0-1) and
  * frame j0-1 as an *overlap row*: the peak row its own first frame links into -- identical,
    bit for bit, to the last row of rank g-1 (tests/test_gpu_parity.py::test_segment_sharding).
Analysis and the frame-pair-local link step then need no communication at all.  Track ids are
made global by ONE all_gather (NCCL over NVLink) of each rank's peak/track tables: partials
born in the overlap row are renamed to the id they carry on the left neighbour (same column of
the same row), the others are offset by the number of partials born on earlier ranks, which
reproduces the reference's numbering (order of add_empty_partial calls, PVAnalysis.py:819-830).
Resynthesis is again embarrassingly parallel: every rank renders its own block range from the
gathered table.
"""
import numpy as np
import torch

from . import pv as P


def plan_segments(nsamp_total, nfft, hop, world):
    """Frame ranges and sample windows of every rank.  Keys: j0, j1 (global frames),
    sample0 / nsamp (window of the global signal the rank reads), frame0 (local index of the
    first emitted row), nframes (emitted rows, including the overlap row when has_overlap),
    prev_zero, has_overlap."""
    F = P.n_frames(nsamp_total, nfft, hop)
    plans = []
    shard = F >= 4 * world          # too short to shard: rank 0 takes everything
    for g in range(world):
        j0 = (F * g) // world if shard else (0 if g == 0 else F)
        j1 = (F * (g + 1)) // world if shard else F
        if j1 <= j0:
            plans.append(dict(rank=g, j0=j0, j1=j1, sample0=0, nsamp=0, frame0=0, nframes=0, prev_zero=True,
                              has_overlap=False, frames_total=F))
        elif g == 0:
            plans.append(dict(rank=g, j0=0, j1=j1, sample0=0, nsamp=(j1 - 1) * hop + nfft, frame0=0, nframes=j1,
                              prev_zero=True, has_overlap=False, frames_total=F))
        else:
            s0 = (j0 - 2) * hop
            plans.append(dict(rank=g, j0=j0, j1=j1, sample0=s0, nsamp=(j1 - 1) * hop + nfft - s0, frame0=1,
                              nframes=j1 - j0 + 1, prev_zero=False, has_overlap=True, frames_total=F))
    return plans


def stitch_ids(tids, ntracks, has_overlap):
    """Global track ids from per-segment local ids.

    tids[g]: int tensor [rows_g, K] of local ids (-1 = no point); row 0 is the overlap row
    when has_overlap[g].  Returns (list of global-id tensors for the OWN rows of every
    segment, total number of tracks).  Pure torch (runs on CPU tensors too: gloo tests).
    """
    out = []
    base = 0
    glast = None
    for g, tid in enumerate(tids):
        nt = int(ntracks[g])
        dev = tid.device
        if tid.shape[0] == 0:
            out.append(tid)
            continue
        gid = torch.empty((max(nt, 1),), dtype=torch.int64, device=dev)
        if has_overlap[g]:
            row0 = tid[0].long()
            cols = torch.nonzero(row0 >= 0).flatten()
            n0 = int(cols.numel())
            # partials born in the overlap row continue the left neighbour's partials
            gid[row0[cols]] = glast.to(dev)[cols]
            own = tid[1:]
        else:
            n0 = 0
            own = tid
        nnew = nt - n0
        if nnew > 0:
            gid[n0:nt] = base + torch.arange(nnew, device=dev)
        base += nnew
        ownl = own.long()
        gown = torch.where(ownl >= 0, gid[ownl.clamp(min=0)], torch.full_like(ownl, -1))
        out.append(gown.to(torch.int32))
        if gown.shape[0] > 0:
            glast = gown[-1].long()
        else:   # no own rows: the boundary row stays the overlap row's ids
            glast = torch.where(tid[0].long() >= 0, gid[tid[0].long().clamp(min=0)], torch.full_like(tid[0].long(), -1))
    return out, base


def _pack_bytes(tables, tid, ntracks, rows_max):
    """One contiguous uint8 buffer per rank: 4 float64 tables + int32 tid, padded to rows_max."""
    K = tid.shape[1]
    dev = tid.device
    rows = tid.shape[0]
    f64 = torch.zeros((4, rows_max, K), dtype=torch.float64, device=dev)
    for q, k in enumerate(("f", "mag", "ph", "realph")):
        f64[q, :rows] = tables[k]
    i32 = torch.full((rows_max + 1, K), -1, dtype=torch.int32, device=dev)
    i32[:rows] = tid
    i32[rows_max, 0] = int(ntracks)
    i32[rows_max, 1] = rows
    return torch.cat([f64.view(torch.uint8).flatten(), i32.view(torch.uint8).flatten()])


def _unpack_bytes(buf, rows_max, K):
    n64 = 4 * rows_max * K * 8
    f64 = buf[:n64].view(torch.float64).view(4, rows_max, K)
    i32 = buf[n64:].view(torch.int32).view(rows_max + 1, K)
    nt, rows = int(i32[rows_max, 0].item()), int(i32[rows_max, 1].item())
    return {k: f64[q, :rows] for q, k in enumerate(("f", "mag", "ph", "realph"))}, i32[:rows], nt


def gather_tables(tables, tid, ntracks, plans, group=None):
    """The single all_gather: every rank receives every rank's local tables and returns the
    global frame tables (own rows of all segments, overlap rows dropped) with global ids."""
    import torch.distributed as dist
    world = len(plans)
    K = tid.shape[1]
    rows_max = max(p["nframes"] for p in plans)
    mine = _pack_bytes(tables, tid, ntracks, rows_max)
    flat = torch.empty((world * mine.numel(),), dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(flat, mine, group=group)
    allb = flat.view(world, mine.numel())
    segs = [_unpack_bytes(allb[g], rows_max, K) for g in range(world)]
    gids, ntot = stitch_ids([s[1] for s in segs], [s[2] for s in segs], [p["has_overlap"] for p in plans])
    glob = {}
    for k in ("f", "mag", "ph", "realph"):
        glob[k] = torch.cat([segs[g][0][k][1:] if plans[g]["has_overlap"] else segs[g][0][k] for g in range(world)]).contiguous()
    glob["tid"] = torch.cat(gids).contiguous()
    glob["ntracks"] = ntot
    return glob


def track_segment(a, plan, world):
    """Link this rank's rows (first row = overlap row on ranks > 0), then -- when world > 1 --
    make ids global and assemble the global tables with one all_gather.

    ``a``: dict from analyze_device (nclips == 1).  Returns dict(f, mag, ph, realph [F*, K],
    tid, link, ntracks, block0, nblocks): F* = own frames (world == 1) or all frames.
    """
    tables = {k: a[k][0] for k in ("f", "mag", "ph", "realph")}
    tr = P.track_device(tables["f"], tables["mag"])
    nt = int(tr["ntracks"][0].item())
    if world == 1:
        out = dict(tables)
        out.update(tid=tr["tid"], link=tr["link"], ntracks=nt, block0=0, nblocks=-1)
        return out
    plans = plan["all"]
    glob = gather_tables(tables, tr["tid"], nt, plans)
    glob.update(link=None, block0=plan["j0"], nblocks=plan["j1"] - plan["j0"])
    return glob


class ShardedPV(object):
    """PV over one long signal sharded across the ranks of a torch.distributed job.

    Every rank constructs it with ITS window of the signal (``plan['sample0']`` ..
    ``+plan['nsamp']``; see :func:`plan_segments`), host or device.  ``run_pv`` analyses the
    rank's frames, ``toSinSum`` links them and gathers the global track table (one
    all_gather), ``synth`` renders the rank's own block range of the output signal.
    """

    def __init__(self, x_local, sr, nsamp_total, nfft=1024, hop=None, npks=20, pkthresh=0.005, wind=np.hanning,
                 rank=None, world=None, device=None):
        import torch.distributed as dist
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        hop = int(nfft / 2) if hop is None else int(hop)
        self.plans = plan_segments(nsamp_total, nfft, hop, self.world)
        self.plan = dict(self.plans[self.rank], all=self.plans)
        self.pv = P.PV(x_local, sr, nfft=nfft, hop=hop, npks=npks, pkthresh=pkthresh, wind=wind, progress=False,
                       device=device)
        if self.pv.nsamp != self.plan["nsamp"]:
            raise ValueError("rank %d expects %d samples starting at global sample %d, got %d" % (
                self.rank, self.plan["nsamp"], self.plan["sample0"], self.pv.nsamp))
        self.sr, self.nfft, self.hop = sr, nfft, hop

    def run_pv(self):
        """Analyse this rank's frames (rows: overlap row first on ranks > 0, then own frames)."""
        p, pv = self.plan, self.pv
        pv._devout = P.analyze_device(pv._xd, pv.sr, pv.nfft, pv.hop, pv.npeaks, pv.peakthresh, pv._tb,
                                      frame0=p["frame0"], nframes=p["nframes"], prev_zero=p["prev_zero"])
        pv.nframes = p["nframes"]
        pv._host = {}
        j_first = p["j0"] - (1 if p["has_overlap"] else 0)
        pv._host["t"] = ((np.arange(pv.nframes) + j_first) * pv.hop + pv.nfft / 2.0) / pv.sr

    def toSinSum(self):
        """Global SinSum (identical on every rank) + this rank's block range."""
        g = track_segment(self.pv._devout, self.plan, self.world)
        ss = P.SinSum(self.sr, nfft=self.nfft, hop=self.hop, device=self.pv._dev)
        ss._set_device_tables(g["f"], g["mag"], g["ph"], g["realph"])
        ss._set_device_tracks(g["tid"], g["link"], g["ntracks"])
        self.block0, self.nblocks = g["block0"], g["nblocks"]
        return ss

    def synth_local(self, ss, hop=None, edge=1.0, minframes=3):
        """Render this rank's own block range [j0, j1) (the last rank also renders the tail)."""
        hop = self.hop if hop is None else int(hop)
        pk = ss._ensure_packed()
        max_end = int((pk["tstart"] + pk["tlen"] - 1).max().item()) if ss._trk["ntracks"] else -1
        nout, _ = P.synth_geometry(max_end, hop, self.nfft, self.hop, edge)
        nblk = -(-nout // hop)
        b0 = self.block0
        nb = (nblk - b0) if self.rank == self.world - 1 else self.nblocks
        return P.resynth_device(ss._trk["tid"], pk, self.sr, hop, self.nfft, self.hop, edge=edge, minframes=minframes,
                                max_end=max_end, block0=b0, nblocks=nb)
