"""Multi-GPU arm (NCCL, world size 2) of the sharded paths; skipped with fewer than 2 GPUs.
With one GPU the same code path runs with world size 1 (the all-reduce is a no-op)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import cases, ref_numpy as rn  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, ret):
    import torch.distributed as dist

    from chord_detection_b200 import distributed as D

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, dev = D.init()
    x, fs = cases.make_input(dict(fn="s_poly_long", seed=3, fs=44100, n=n))
    f0, f1, s0, s1 = D.shard_frames(n, 2048, 512, rank, world)
    xl = torch.from_numpy(x[s0:s1]).to(dev)
    fused_ok = D.comm_init()
    tot = D.harmonic_energy_sharded(xl, fs, f1 - f0, 2048, hop=512, fused=False)  # NCCL all-reduce
    ret["fused_ok_%d" % rank] = bool(fused_ok)
    if fused_ok:
        # the in-kernel all-reduce over peer memory (CDB_FLAG_ALLREDUCE): several collectives in
        # a row (both mailbox parities), one of them with an empty shard on the last rank
        from chord_detection_b200 import _native as nat

        for it in range(5):
            t2 = D.harmonic_energy_sharded(xl, fs, f1 - f0, 2048, hop=512, fused=True)
            assert torch.allclose(t2, tot, rtol=1e-12, atol=0), (it, t2, tot)
        gathered = [torch.zeros_like(t2) for _ in range(world)]
        if world > 1:
            dist.all_gather(gathered, t2)
            assert all(torch.equal(g, gathered[0]) for g in gathered)  # bit-identical on all ranks
        empty = rank == world - 1 and world > 1
        t3 = D.harmonic_energy_sharded(xl, fs, 0 if empty else f1 - f0, 2048, hop=512, fused=True)
        solo = D.harmonic_energy_sharded(xl, fs, 0 if empty else f1 - f0, 2048, hop=512, fused=False)
        assert torch.allclose(t3, solo, rtol=1e-12, atol=0)
        assert nat.Handle.get(dev.index).comm_status() == 0
        # the end-to-end pipeline (pinned host buffer -> chunked H2D -> kernel) rides the exchange in
        # its LAST chunk (ACCUMULATE | ALLREDUCE): every rank gets the sum over the ranks' signals
        from chord_detection_b200 import ops

        xr, _ = cases.make_input(dict(fn="s_poly_long", seed=30 + rank, fs=44100, n=1500 * 512 + 11))
        host = torch.from_numpy(xr).pin_memory()
        pipe = ops.HostPipeline(dev, fs, 2048, hop=512, chunk_frames=400)
        local = torch.from_numpy(pipe.run(host)).to(dev)
        if world > 1:
            dist.all_reduce(local)
        fused = pipe.run(host, allreduce=True)
        assert np.allclose(fused, local.cpu().numpy(), rtol=1e-12, atol=0)
    clips = np.stack([cases.make_input(dict(fn="s_poly", seed=60 + i, fs=22050, n=9000))[0] for i in range(6)])
    c0, c1 = D.shard_range(6, rank, world)
    sums, _ = D.all_methods_sharded(torch.from_numpy(clips[c0:c1]).to(dev), 22050, methods=(2, 4))
    torch.cuda.synchronize()
    if rank == 0:
        ret["he"] = tot.cpu().numpy()
        ret["sums"] = sums.cpu().numpy()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_paths_sum_to_whole(world):
    import torch.multiprocessing as mp

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    n = 2000 * 512 + 3
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, ret), nprocs=world, join=True)
    x, fs = cases.make_input(dict(fn="s_poly_long", seed=3, fs=44100, n=n))
    want = rn.harmonic_energy_fast(x, fs, frame_size=2048, hop=512)
    assert np.max(np.abs(ret["he"] - want)) / np.max(want) < 1e-4
    # CUDA IPC between the ranks' processes must work on a B200 box: the fused path is the product
    assert all(ret["fused_ok_%d" % r] for r in range(world))
    clips = [cases.make_input(dict(fn="s_poly", seed=60 + i, fs=22050, n=9000))[0] for i in range(6)]
    w2 = sum(rn.harmonic_energy_fast(c, 22050) for c in clips)
    w4 = sum(rn.prime(c, 22050) for c in clips)
    assert np.max(np.abs(ret["sums"][0] - w2)) / np.max(w2) < 1e-4
    assert np.max(np.abs(ret["sums"][1] - w4)) / np.max(w4) < 1e-4
