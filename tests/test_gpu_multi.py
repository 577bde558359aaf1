"""-m gpu, needs >= 2 GPUs (skipped otherwise): ShardedPV over NCCL, one process per GPU.  Own
rows, the gathered global track table, spans and the concatenation of the locally rendered block
ranges equal the unsharded single-GPU run bit for bit (pypevoc_b200/dist.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _signal(short):
    from pypevoc_b200 import signals
    sr = 44100
    if short:                                              # 7 frames < 4 * world: rank 0 owns everything
        return signals.harm(sr, (2048 + 6 * 512 + 100) / float(sr), 220, 90, 0.5, 0.02, 9), sr
    x = signals.harm(sr, 6.0, 220, 90, 0.5, 0.02, 9)
    x[int(2.2 * sr):int(2.6 * sr)] = 0.0
    x[int(5.3 * sr):] = 0.0                               # the signal's last point lies before the last frames
    return x, sr


def _worker(rank, world, port, result_dir, streamed, peer=False, short=False):
    import torch
    import torch.distributed as dist
    # peer: fused rename + gather over peer memory (the default); otherwise the NCCL all_gather
    os.environ["PVK_PEER_GATHER"] = "1" if peer else "0"
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    from pypevoc_b200 import dist as D
    nfft, hop, npks = 2048, 512, 50
    x, sr = _signal(short)
    plans = D.plan_segments(len(x), nfft, hop, world)
    p = plans[rank]
    xl = torch.from_numpy(np.ascontiguousarray(x[p["sample0"]:p["sample0"] + p["nsamp"]]))
    hb = {} if streamed else None
    for rep in range(2):                                   # second pass reuses streams and pinned buffers
        spv = D.ShardedPV(xl.pin_memory() if streamed else xl.cuda(), sr, len(x), nfft=nfft, hop=hop, npks=npks,
                          rank=rank, world=world, device=torch.device("cuda", rank))
        spv.run_pv(hostbuf=hb)
        ss = spv.toSinSum()
        w, s0 = ss.synth_local(hostbuf=hb) if streamed else ss.synth_local(to_host=True)
        w = np.array(w)
        table = ss.track_ids
        st, end = ss.st, ss.end
        f_own = np.array(spv.pv.f[spv.own_rows])
    np.savez(os.path.join(result_dir, "r%d.npz" % rank), w=w, s0=s0, table=table, st=st, end=end, f=f_own,
             ntracks=ss.ntracks, max_end=ss.max_end, j0=p["j0"], j1=p["j1"], peer_used=bool(ss._h.peer_used))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("streamed,peer,short", [(False, False, False), (True, False, False), (False, True, False),
                                                 (False, False, True), (False, True, True)])
def test_sharded_pv_over_nccl(tmp_path, streamed, peer, short):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    from pypevoc_b200 import PV
    nfft, hop, npks = 2048, 512, 50
    x, sr = _signal(short)
    pv0 = PV(x, sr, nfft=nfft, hop=hop, npks=npks, progress=False)
    pv0.run_pv()
    ss0 = pv0.toSinSum()
    w0 = ss0.synth(sr, hop)
    port = 29700 + (os.getpid() % 200) + (50 if streamed else 0) + (25 if peer else 0) + (12 if short else 0)
    mp.spawn(_worker, args=(world, port, str(tmp_path), streamed, peer, short), nprocs=world, join=True)
    sig = np.zeros_like(w0)
    covered = 0
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        assert np.array_equal(z["table"], ss0.track_ids), r
        assert z["st"].tolist() == ss0.st and z["end"].tolist() == ss0.end
        assert int(z["ntracks"]) == len(ss0.st) and int(z["max_end"]) == max(ss0.end)
        if int(z["j1"]) > int(z["j0"]):
            assert np.array_equal(z["f"], pv0.f[int(z["j0"]):int(z["j1"])])
        else:                                              # a rank without frames: f.shape == (0,) as in the reference
            assert z["f"].size == 0
        s0, w = int(z["s0"]), z["w"]
        sig[s0:s0 + len(w)] = w
        covered += len(w)
    assert covered == len(w0)
    assert np.array_equal(sig, w0)
    if peer:                                               # falls back to NCCL (with a warning) where unavailable
        print("peer-memory gather used:", bool(z["peer_used"]))
