"""Multi-GPU: one process per GPU (torchrun), data-parallel over frames or clips, ONE all-reduce.

The reference has no parallelism of any kind.  Every method's frames (HE,
ESACF, prime) or clips (iterative F0 carries IIR state across a clip) are independent, so rank r
takes a contiguous shard, runs the same kernels, and the per-rank 12-bin sums are combined with a
single `all_reduce(SUM)` of 12 doubles per method (NCCL over NVLink on GPUs; gloo in CPU tests).
No data-path collective exists or is needed.
"""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK).
    Returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if cuda else "gloo"
        if backend == "nccl":
            dist.init_process_group(backend, device_id=device)
        else:
            dist.init_process_group(backend)
    return rank, world, device


def comm_init(device=None):
    """Set up the fused all-reduce of the harmonic-energy kernel (include/chordb200.h cdb_comm_*):
    every rank allocates its mailbox, the CUDA IPC handles travel through an all_gather_object on
    the existing process group, and every rank maps its peers' mailboxes.  Returns True when the
    in-kernel path is usable (ops.harmonic_energy(..., allreduce=True)); with one rank it is set up
    as a world of one.  Raises nothing on failure (e.g. CUDA IPC unavailable in a sandbox): returns
    False and callers keep using all_reduce_chroma (NCCL)."""
    from . import _native as nat

    if not torch.cuda.is_available():
        return False
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    h = nat.Handle.get(dev.index)
    if getattr(h, "comm_world", 0):
        return True
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0

    def exchange(b):
        if world == 1:
            return [b]
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    ok = True
    try:
        h.comm_init(rank, world, exchange)
    except (RuntimeError, ValueError):
        ok = False
    if world > 1:  # all ranks or none
        t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
        if not ok and getattr(h, "comm_world", 0):
            h.comm_destroy()
    return ok


def shard_range(n_items, rank, world):
    """Contiguous, balanced [begin, end) of n_items for this rank (sizes differ by at most 1)."""
    return (n_items * rank) // world, (n_items * (rank + 1)) // world


def shard_frames(n_samples, frame_size, hop, rank, world):
    """Shard the frames of ONE long signal.  Returns (f0, f1, s0, s1): this rank owns frames
    [f0, f1) and must hold samples [s0, s1) -- its frames plus the frame_size-hop halo; frames that
    run past n_samples are zero padded by the kernel (explicit frames_per_clip)."""
    hop = frame_size if not hop else hop
    n_frames = (n_samples + hop - 1) // hop if n_samples > 0 else 0
    f0, f1 = shard_range(n_frames, rank, world)
    if f1 <= f0:
        return f0, f1, 0, 0
    s0 = f0 * hop
    s1 = min(n_samples, (f1 - 1) * hop + frame_size)
    return f0, f1, s0, s1


def all_reduce_chroma(t):
    """Sum the per-rank chroma tensor(s) ([12] or [k, 12]) over all ranks, in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def harmonic_energy_sharded(x_local, fs, frames_local, frame_size, hop=None, fused=None, **kw):
    """HE over this rank's shard of a long signal (x_local = samples [s0, s1) of shard_frames) and
    the sum over ranks.  Returns the GLOBAL 12-bin sum on every rank (CUDA float64 tensor).
    fused=None: use the in-kernel all-reduce over peer memory when comm_init() succeeded and the
    frame size is 2048, else NCCL; True / False force one or the other."""
    from . import _native as nat, ops

    dev = x_local.device
    total = torch.zeros(12, dtype=torch.float64, device=dev)
    if fused is None:
        fused = (int(frame_size) == 2048 and x_local.is_cuda
                 and bool(getattr(nat.Handle.get(dev.index), "comm_world", 0)))
    if fused:
        # a collective: a rank with an empty shard still launches (one CTA that contributes zeros)
        xs = x_local if frames_local > 0 else x_local[:0]
        ops.harmonic_energy(xs, fs, frame_size=frame_size, hop=hop,
                            frames_per_clip=frames_local if frames_local > 0 else None,
                            out_total=total, allreduce=True, **kw)
        return total
    if frames_local > 0:
        ops.harmonic_energy(x_local, fs, frame_size=frame_size, hop=hop,
                            frames_per_clip=frames_local, out_total=total, **kw)
    return all_reduce_chroma(total)


_side = {}


def _side_streams(dev):
    """Streams (and their own library handles) for the methods that run next to each other."""
    key = (dev.type, dev.index)
    if key not in _side:
        lo, hi = torch.cuda.Stream.priority_range()
        # ESACF's fit kernel is one large persistent CTA per SM: give it priority so that its CTAs
        # are placed as soon as an SM has room, the small CTAs of the other kernels fill the rest
        _side[key] = {1: torch.cuda.Stream(dev, priority=hi), 3: torch.cuda.Stream(dev),
                      4: torch.cuda.Stream(dev)}
    return _side[key]


def all_methods_sharded(clips_local, fs, methods=(1, 2, 3, 4), reduce=True, concurrent=None):
    """Config C5: every rank runs the requested methods on its shard of clips ([n_local, clip_len]
    CUDA float32) and ONE all-reduce combines the [n_methods, 12] sums (reduce=False leaves the
    local sums for callers that loop over chunks and reduce once at the end).  Returns (global sums
    [n_methods, 12], dict of per-clip results that stay sharded).

    concurrent=True (or CDB_CONCURRENT_METHODS=1) runs every method on its own stream with its own
    library handle (a handle serves one stream at a time) and joins the streams before the results
    are used.  The kernels stress different resources (ESACF's fits are latency-bound with the FP64
    pipe < 10 % busy, prime's Goertzel sums are FP64-throughput-bound, the iterative-F0 spectrum is
    shared-memory-bound), but MEASURED on one B200 (8192 clips, scripts/time_c5.py, r02) the
    concurrent run is within 1 % of the sequential one (1003 vs 1012 ms; 1041 ms without matching
    shared-memory carve-outs, and forcing the maximal carve-out slows the sequential kernels by
    10 %): the 188-226 KB persistent fit CTAs leave no room for co-resident CTAs.  Off by default."""
    from . import _native as nat, ops

    dev = clips_local.device
    sums = torch.zeros((len(methods), 12), dtype=torch.float64, device=dev)
    per_clip = {}
    if concurrent is None:
        concurrent = clips_local.is_cuda and os.environ.get("CDB_CONCURRENT_METHODS") == "1"
    fns = {1: ops.esacf, 2: ops.harmonic_energy, 3: ops.iterative_f0, 4: ops.prime_multif0}
    for m in methods:
        if m not in fns:
            raise ValueError("valid methods: 1, 2, 3, 4")
    if clips_local.shape[0] == 0:
        return (all_reduce_chroma(sums) if reduce else sums), per_clip
    if not concurrent:
        for i, m in enumerate(methods):
            r = fns[m](clips_local, fs, per_clip=True)
            sums[i] = r.total
            per_clip[m] = r.clips
        return (all_reduce_chroma(sums) if reduce else sums), per_clip
    cur = torch.cuda.current_stream(dev)
    side = _side_streams(dev)
    ready = torch.cuda.Event()
    ready.record(cur)
    results, done = {}, []
    for m in methods:
        if m == 2:
            continue  # harmonic energy is ~0.1 % of the work: it stays on the caller's stream
        s = side[m]
        s.wait_event(ready)
        clips_local.record_stream(s)
        h = nat.Handle.get(dev.index, tag="method%d" % m)
        if m == 1 and not getattr(h, "_co_run", False):
            h.set_option("esacf_fit_warps", 5)  # leave shared memory for the co-running kernels
            h._co_run = True
        with torch.cuda.stream(s):
            results[m] = fns[m](clips_local, fs, per_clip=True, handle=h)
            e = torch.cuda.Event()
            e.record(s)
            done.append(e)
    if 2 in methods:
        results[2] = ops.harmonic_energy(clips_local, fs, per_clip=True)
    for e in done:
        cur.wait_event(e)
    for i, m in enumerate(methods):
        r = results[m]
        r.total.record_stream(cur)
        r.clips.record_stream(cur)
        sums[i] = r.total
        per_clip[m] = r.clips
    return (all_reduce_chroma(sums) if reduce else sums), per_clip
