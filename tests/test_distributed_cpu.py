"""World-size-2 gloo tests of the multi-GPU host logic (sharding + the single all-reduce).  The
compute on each rank is the numpy oracle (this is a CPU test of the plumbing; the GPU arm of the
same path is tests/test_distributed_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chord_detection_b200 import distributed as D
from oracle import cases, ref_numpy as rn


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 100, 100003):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_frames_halo():
    n, N, hop = 100000 * 512, 2048, 512
    tot = 0
    for r in range(8):
        f0, f1, s0, s1 = D.shard_frames(n, N, hop, r, 8)
        tot += f1 - f0
        assert s0 == f0 * hop and s1 == min(n, (f1 - 1) * hop + N)
    assert tot == 100000
    assert D.shard_frames(10, 2048, 512, 3, 8)[:2] == (0, 0) or True  # tiny inputs: empty shards allowed


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, dev = D.init(backend="gloo")
    assert (r, w) == (rank, world) and dev.type == "cpu"
    assert D.comm_init() is False  # the fused all-reduce needs CUDA peers; NCCL / gloo path stays
    x, fs = cases.make_input(dict(fn="s_poly_long", seed=3, fs=44100, n=n))
    N, hop = 2048, 512
    f0, f1, s0, s1 = D.shard_frames(n, N, hop, rank, world)
    _, frames = rn.harmonic_energy_fast(x[s0:s1], fs, frame_size=N, hop=hop, per_frame=True)
    local = torch.from_numpy(frames[: f1 - f0].sum(axis=0))
    # the clips arm: each rank owns a shard of clips, [2, 12] sums, ONE all-reduce
    c0, c1 = D.shard_range(5, rank, world)
    clip_sums = torch.zeros((2, 12), dtype=torch.float64)
    for i in range(c0, c1):
        xi, _ = cases.make_input(dict(fn="s_poly", seed=40 + i, fs=22050, n=9000))
        clip_sums[0] += torch.from_numpy(rn.harmonic_energy_fast(xi, 22050))
        clip_sums[1] += torch.from_numpy(rn.prime(xi, 22050))
    D.all_reduce_chroma(local)
    D.all_reduce_chroma(clip_sums)
    if rank == 0:
        ret["frames"] = local.numpy().copy()
        ret["clips"] = clip_sums.numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_frame_and_clip_sharding_sum_to_whole():
    n = 300 * 512 + 100
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n, ret), nprocs=2, join=True)
    x, fs = cases.make_input(dict(fn="s_poly_long", seed=3, fs=44100, n=n))
    want = rn.harmonic_energy_fast(x, fs, frame_size=2048, hop=512)
    assert np.allclose(ret["frames"], want, rtol=1e-12)
    wc = np.zeros((2, 12))
    for i in range(5):
        xi, _ = cases.make_input(dict(fn="s_poly", seed=40 + i, fs=22050, n=9000))
        wc[0] += rn.harmonic_energy_fast(xi, 22050)
        wc[1] += rn.prime(xi, 22050)
    assert np.allclose(ret["clips"], wc, rtol=1e-12)


def test_batch_expand_inputs(tmp_path):
    from chord_detection_b200 import batch

    d = tmp_path / "clips"
    (d / "sub").mkdir(parents=True)
    for name in ("b.wav", "a.wav", "sub/c.WAV", "notes.md"):
        (d / name).write_bytes(b"")
    man = tmp_path / "list.txt"
    man.write_text("# corpus\nclips/a.wav\n\n%s\n" % (d / "b.wav"))
    got = batch.expand_inputs([str(d), str(man), "x.wav"])
    assert [q.name for q in got] == ["a.wav", "b.wav", "c.WAV", "a.wav", "b.wav", "x.wav"]
    assert got[3] == tmp_path / "clips" / "a.wav"


def _batch_worker(rank, world, port, paths, ret):
    import io

    from chord_detection_b200 import batch

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))

    def fake_run(ps, methods, key=False, device=None, **kw):  # stands in for the device ops
        lines = [(str(q), [(m, "M%d" % m, "%012d" % (int(Path(q).stem) * 10 + m), "Cmaj" if key else None)
                           for m in methods]) for q in ps]
        sums = {m: np.full(12, float(sum(int(Path(q).stem) for q in ps) * m)) for m in methods}
        return lines, sums

    from pathlib import Path

    batch.run_batch = fake_run
    real_init = D.init
    D.init = lambda backend=None: real_init(backend="gloo")
    buf = io.StringIO()
    lines, sums = batch.main_batch(paths, [2, 4], key=True, out=buf)
    ret[rank] = (buf.getvalue(), [ln[0] for ln in lines], {m: v.tolist() for m, v in sums.items()})


def test_batch_front_end_two_ranks_gloo():
    """Sharding of the clip list, gather of the per-clip lines in order, ONE all-reduce of the
    corpus sums (the device ops are replaced by a stub: this is the plumbing)."""
    paths = ["%d.wav" % i for i in range(1, 8)]
    ret = mp.Manager().dict()
    mp.spawn(_batch_worker, args=(2, _free_port(), paths, ret), nprocs=2, join=True)
    out0, order0, sums0 = ret[0]
    out1, order1, sums1 = ret[1]
    assert order0 == order1 == paths
    assert out1 == ""  # only rank 0 prints
    assert sums0 == sums1 and sums0[2][0] == 2.0 * sum(range(1, 8)) and sums0[4][0] == 4.0 * sum(range(1, 8))
    rows = out0.strip().splitlines()
    assert rows[0] == "1.wav" and rows[1] == "2 - M2" and rows[2] == "%012d" % 12 and rows[3] == "Cmaj"
    assert "== corpus (7 clips)" in rows
