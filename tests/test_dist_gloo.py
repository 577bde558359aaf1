"""CPU, world_size 2/3 over gloo: the segment planner, the global numbering of locally linked
partials (tiny all_gather + sequential pass) and the all_gather of the track table
(pypevoc_b200/dist.py) reproduce the unsharded track numbering bit for bit.  Local linking is
done here by the oracle (the GPU kernel is checked in the -m gpu tests); what this test pins is
the host-side sharding logic and the collectives."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASE = dict(cfg3_clip=(512, 128), metric_1s=(2048, 512))


def _worker(rank, world, port, result_dir, name):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pv_oracle as orc
    from pypevoc_b200 import dist as D
    from golden_util import case_golden
    g = case_golden(name)                              # peak tables from the real reference
    F, K = g["f"].shape
    nfft, hop = CASE[name]
    nsamp = (F - 1) * hop + nfft + 1                   # sample counts are irrelevant for the stitch
    plans = D.plan_segments(nsamp, nfft, hop, world)
    assert plans[0]["frames_total"] == F
    p = plans[rank]
    f = np.ascontiguousarray(g["f"][p["w0"]:p["w1"]])
    mag = np.ascontiguousarray(g["mag"][p["w0"]:p["w1"]])
    assert f.shape[0] == p["nframes"]
    loc = orc.track(f, mag)                            # local ids over the window rows
    st = D.stitch(torch.from_numpy(loc["tid"]), p, plans)
    _, finish = D.gather_track_table(st["tid_own"], plans, async_op=False)
    table = finish()
    fglob = D.gather_rows(torch.from_numpy(f[p["own0"]:p["own0"] + p["nown"]]), plans)
    np.savez(os.path.join(result_dir, "r%d.npz" % rank), tid=table.numpy(), f=fglob.numpy(),
             ntracks=st["ntracks"], max_end=st["max_end"])
    dist.barrier()
    dist.destroy_process_group()


def _worker_short(rank, world, port, result_dir, name, nrows):
    """Too few frames to shard (F < 4 * world): rank 0 owns everything, the other ranks contribute
    empty [0, K] tables -- and still 2K + 4 integers each to the numbering all_gather."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pv_oracle as orc
    from pypevoc_b200 import dist as D
    from golden_util import case_golden
    g = case_golden(name)
    K = g["f"].shape[1]
    nfft, hop = CASE[name]
    plans = D.plan_segments((nrows - 1) * hop + nfft + 1, nfft, hop, world)
    p = plans[rank]
    assert plans[0]["nown"] == nrows and all(q["nown"] == 0 for q in plans[1:])
    f = np.ascontiguousarray(g["f"][p["w0"]:p["w1"]]).reshape(-1, K)
    mag = np.ascontiguousarray(g["mag"][p["w0"]:p["w1"]]).reshape(-1, K)
    tid = orc.track(f, mag)["tid"] if len(f) else np.zeros((0, K), dtype=np.int32)
    st = D.stitch(torch.from_numpy(tid), p, plans)
    _, finish = D.gather_track_table(st["tid_own"], plans, async_op=False)
    np.savez(os.path.join(result_dir, "r%d.npz" % rank), tid=finish().numpy(), ntracks=st["ntracks"], max_end=st["max_end"])
    dist.barrier()
    dist.destroy_process_group()


def test_stitch_unsharded_plan_with_empty_ranks(tmp_path):
    from oracle import pv_oracle as orc
    from golden_util import case_golden
    world, name, nrows = 3, "cfg3_clip", 9
    port = 29650 + (os.getpid() % 200)
    mp.spawn(_worker_short, args=(world, port, str(tmp_path), name, nrows), nprocs=world, join=True)
    g = case_golden(name)
    ref = orc.track(g["f"][:nrows], g["mag"][:nrows])
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        assert np.array_equal(z["tid"], ref["tid"]), r
        assert int(z["ntracks"]) == len(ref["st"]) and int(z["max_end"]) == int(np.max(ref["end"]))


@pytest.mark.parametrize("world,name", [(2, "cfg3_clip"), (3, "cfg3_clip"), (2, "metric_1s")])
def test_stitch_matches_unsharded_tracking(tmp_path, world, name):
    from golden_util import case_golden
    port = 29600 + world + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, str(tmp_path), name), nprocs=world, join=True)
    g = case_golden(name)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        assert np.array_equal(z["f"], g["f"])
        assert np.array_equal(z["tid"], g["tid"]), "rank %d: stitched ids differ from the reference numbering" % r
        assert int(z["ntracks"]) == len(g["st"])
        assert int(z["max_end"]) == int(np.max(g["end"]))


def test_resolve_ids_many_ranks_single_process():
    """The numbering pass alone, 5 'ranks' in one process (no collective)."""
    from oracle import pv_oracle as orc
    from pypevoc_b200 import dist as D
    from golden_util import case_golden
    g = case_golden("cfg3_clip")
    F, K = g["f"].shape
    world = 5
    plans = D.plan_segments((F - 1) * 128 + 512 + 1, 512, 128, world)
    locs, summ = [], []
    for p in plans:
        loc = orc.track(g["f"][p["w0"]:p["w1"]], g["mag"][p["w0"]:p["w1"]])
        locs.append(torch.from_numpy(loc["tid"]))
        summ.append(D.local_summary(locs[-1], p).numpy())
    summ = np.stack(summ)
    bases, gprevs, ntot, max_end = D.resolve_ids(summ, K)
    rows = [D.global_ids(locs[r], plans[r], bases[r], gprevs[r], int(summ[r, 2 * K]), int(summ[r, 2 * K + 1]))
            for r in range(world)]
    assert np.array_equal(torch.cat(rows).numpy(), g["tid"])
    assert ntot == len(g["st"])


def test_plan_covers_all_frames():
    from pypevoc_b200 import dist as D
    from pypevoc_b200.pv import n_frames
    for nsamp, nfft, hop, world in ((44100 * 60, 2048, 512, 8), (100000, 1024, 300, 4), (5000, 2048, 512, 2),
                                    (44100 * 600 * 8, 2048, 256, 8)):
        plans = D.plan_segments(nsamp, nfft, hop, world)
        F = n_frames(nsamp, nfft, hop)
        Lb, Lf = D.halos(nfft, hop)
        assert plans[0]["j0"] == 0 and plans[-1]["j1"] == F
        for a, b in zip(plans[:-1], plans[1:]):
            assert a["j1"] == b["j0"]
        for p in plans:
            if p["nframes"]:
                assert p["sample0"] + p["nsamp"] <= nsamp
                assert p["w0"] == max(0, p["j0"] - Lb) and p["w1"] == min(F, p["j1"] + Lf)
                assert p["sample0"] + p["frame0"] * hop == p["w0"] * hop
                assert (p["frame0"] + p["nframes"] - 1) * hop + nfft == p["nsamp"]
                assert p["own0"] == p["j0"] - p["w0"] and p["nown"] == p["j1"] - p["j0"]


def test_local_render_range_matches_global():
    """render_range_local() + trim_local() (block range from LOCAL knowledge, cut once the global
    last frame is known) give exactly render_range()'s sample range on every rank, whatever the
    position of the signal's last point."""
    from pypevoc_b200 import dist as D
    for nfft, hop in ((2048, 512), (512, 128), (1024, 300), (8192, 1024)):
        for world in (1, 2, 3, 8):
            F = 40 * world + 7
            nsamp = (F - 1) * hop + nfft + 1
            plans = D.plan_segments(nsamp, nfft, hop, world)
            assert plans[0]["frames_total"] == F
            for max_end in [-1, 0, 1, 5] + [p["j0"] + d for p in plans for d in (-3, -1, 0, 1, 2)] + [F - 2, F - 1]:
                if max_end < -1 or max_end >= F:
                    continue
                covered = 0
                for p in plans:
                    # what the rank sees: the last point inside its window, if any
                    ll = max_end if (p["nown"] and p["w0"] <= max_end < p["w1"]) else -1
                    if p["nown"] and max_end >= p["w1"]:
                        ll = p["w1"] - 1                      # some later point exists; only ranks before the last
                    b0, b1, bound = D.render_range_local(p, plans, ll, hop, nfft, hop)
                    n_r = max(min(b1 * hop, bound) - b0 * hop, 0)
                    g0, g1, nout = D.render_range(p, plans, max_end, hop, nfft, hop)
                    n_g = max(min(g1 * hop, nout) - g0 * hop, 0)
                    n, s0 = D.trim_local(n_r, b0, p, plans, max_end, hop, nfft, hop)
                    assert (n, s0) == (n_g, g0 * hop), (nfft, hop, world, max_end, p["rank"])
                    covered += n
                nout = D.P.synth_geometry(max_end, hop, nfft, hop)[0] if max_end >= 0 else 0
                assert covered == nout, (nfft, hop, world, max_end, covered, nout)


def test_clip_range_partitions_the_batch():
    from pypevoc_b200 import dist as D
    for nclips, world in ((4096, 8), (5, 3), (2, 4), (0, 2)):
        r = [D.clip_range(nclips, g, world) for g in range(world)]
        assert r[0][0] == 0 and r[-1][1] == nclips
        assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        assert max(c1 - c0 for c0, c1 in r) - min(c1 - c0 for c0, c1 in r) <= 1


def test_plan_random_sizes_keep_every_rank_busy():
    """Random (length, nfft, hop, world): contiguous cover, no empty rank once the signal is long
    enough to shard, all ranks but the last equal whenever that is possible (copy-free gather)."""
    from pypevoc_b200 import dist as D
    from pypevoc_b200.pv import n_frames
    rng = np.random.RandomState(5)
    for _ in range(400):
        nfft = int(2 ** rng.randint(6, 14))
        hop = int(rng.randint(1, nfft + 1))
        world = int(rng.randint(1, 9))
        nsamp = int(nfft + rng.randint(0, 600) * hop + rng.randint(0, hop + 1))
        F = n_frames(nsamp, nfft, hop)
        plans = D.plan_segments(nsamp, nfft, hop, world)
        assert plans[0]["j0"] == 0 and plans[-1]["j1"] == F
        assert all(a["j1"] == b["j0"] for a, b in zip(plans[:-1], plans[1:]))
        assert sum(p["nown"] for p in plans) == F
        if F >= 4 * world:
            assert all(p["nown"] > 0 for p in plans), (F, world)
            per = -(-F // world)
            if per * (world - 1) < F:
                assert all(p["nown"] == per for p in plans[:-1])
        else:
            assert plans[0]["nown"] == F and all(p["nown"] == 0 for p in plans[1:])
