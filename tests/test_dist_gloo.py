"""CPU, world_size 2 over gloo: the segment planner and the all_gather stitch of track ids
(pypevoc_b200/dist.py) reproduce the unsharded track numbering bit for bit.  Local linking is
done here by the oracle (the GPU kernel is checked in the -m gpu tests); what this test pins is
the host-side sharding logic and the collective."""
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


def _worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pv_oracle as orc
    from pypevoc_b200 import dist as D
    from golden_util import case_golden
    g = case_golden("cfg3_clip")                       # 371 frames x 20 peaks from the real reference
    F, K = g["f"].shape
    # a frame plan over `world` ranks (sample counts are irrelevant for the stitch)
    nfft, hop = 512, 128
    nsamp = (F - 1) * hop + nfft + 1
    plans = D.plan_segments(nsamp, nfft, hop, world)
    assert plans[0]["frames_total"] == F
    p = plans[rank]
    lo = p["j0"] - (1 if p["has_overlap"] else 0)
    tables = {k: torch.from_numpy(np.ascontiguousarray(g[k][lo:p["j1"]])) for k in ("f", "mag", "ph", "realph")}
    assert tables["f"].shape[0] == p["nframes"]
    loc = orc.track(tables["f"].numpy(), tables["mag"].numpy())     # local ids, overlap row first
    glob = D.gather_tables(tables, torch.from_numpy(loc["tid"]), len(loc["st"]), plans)
    np.savez(os.path.join(result_dir, "r%d.npz" % rank), tid=glob["tid"].numpy(), f=glob["f"].numpy(),
             ntracks=glob["ntracks"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_stitch_matches_unsharded_tracking(tmp_path, world):
    from golden_util import case_golden
    port = 29600 + world + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = case_golden("cfg3_clip")
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        assert np.array_equal(z["f"], g["f"])
        assert np.array_equal(z["tid"], g["tid"]), "rank %d: stitched ids differ from the reference numbering" % r
        assert int(z["ntracks"]) == len(g["st"])


def test_plan_covers_all_frames():
    from pypevoc_b200 import dist as D
    from pypevoc_b200.pv import n_frames
    for nsamp, nfft, hop, world in ((44100 * 60, 2048, 512, 8), (100000, 1024, 300, 4), (5000, 2048, 512, 2),
                                    (44100 * 600 * 8, 2048, 256, 8)):
        plans = D.plan_segments(nsamp, nfft, hop, world)
        F = n_frames(nsamp, nfft, hop)
        assert plans[0]["j0"] == 0 and plans[-1]["j1"] == F
        for a, b in zip(plans[:-1], plans[1:]):
            assert a["j1"] == b["j0"]
        for p in plans:
            if p["nframes"]:
                assert p["sample0"] + p["nsamp"] <= nsamp
                first = p["j0"] - (1 if p["has_overlap"] else 0)
                assert p["sample0"] + p["frame0"] * hop == first * hop
                assert (p["frame0"] + p["nframes"] - 1) * hop + nfft == p["nsamp"]
