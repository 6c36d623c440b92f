#!/usr/bin/env python
"""bench.py — harmonic-energy chromagram throughput (BASELINE.json metric, config C2) plus the
other BASELINE configs (C3 ESACF, C4 iterative F0, C5 all four methods) as a `secondary` block.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

Headline workload (configs[1]): 44.1 kHz mono float32, 2048-sample frames, hop 512, 100 000 frames
per GPU (51.2 M samples, 204.8 MB).  One step = one pass of the hot path over that batch = ONE
kernel: cdb_he_chroma; for N > 1 the sum of the per-GPU 12-bin vectors over NVLink is fused into
that kernel's last CTA (CDB_FLAG_ALLREDUCE: P2P stores into peer mailboxes), NCCL only if CUDA IPC
is unavailable.  Weak scaling: every rank owns its own 100 000 frames.  Prints ONE JSON line on
rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

FS, FRAME, HOP, FRAMES_PER_GPU = 44100, 2048, 512, 100_000
N_ROTATING = 4  # distinct 204.8 MB inputs rotated between steps: each step reads L2-cold data
ALG_BYTES_PER_FRAME = 4 * HOP  # every fp32 sample is read from HBM exactly once
ALG_FLOP_PER_FRAME = 6.0e4     # SURVEY.md 8d (window + 2048-pt real FFT + magnitudes + maxima)
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # CUDA-core fp32 at max clock (no measured value)
METRIC = "frames/sec (2048-pt STFT, hop 512) harmonic-energy chromagram"
PORT_NOTE = ("oracle/ref_numpy port of the reference's per-frame Python loop; it hoists the k' / "
             "window-bound computation (numpy.round per probe, harmonic_energy.py:51-55) out of the "
             "frame loop, so it is FASTER than the unmodified reference and the ratio is conservative")


def _peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic():
    try:
        with open(os.path.join(REPO, "profiles", "he_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, int(r)))
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        nv = self.nv
        names = {
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        mhz = sorted(s[0] for s in self.samples)
        bits = 0
        for s in self.samples:
            bits |= s[1]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": [k for k, v in names.items() if bits & v], "samples": len(mhz)}


# ------------------------------------------------------------------------------------------------
# CPU baseline = the oracle port of the reference (the ONLY place bench.py executes oracle/):
# timed on the box's host cores, on bounded samples of the same workloads
# ------------------------------------------------------------------------------------------------
REF_FRAMES_PER_CORE = 20000  # reference arm: ~2 s of the per-frame Python loop per core and step


def _cpu_worker(args):
    """One core's share of a CPU-baseline sample.  kind: he (C2 frames), esacf (C3 frames), iterf0
    (C4 clips of 8 frames), all4 (C5 clips: all four methods).  -> (seconds, units, checksum)"""
    kind, seed, size = args
    import warnings

    import numpy as np

    from chord_detection_b200 import synth
    from oracle import ref_numpy as rn

    warnings.simplefilter("ignore")
    if kind == "he":
        x = synth.noise(seed, (size - 1) * HOP + FRAME, sigma=0.1)
        t0 = time.perf_counter()
        c = rn.harmonic_energy(x, FS, frame_size=FRAME, hop=HOP)
        return time.perf_counter() - t0, size, float(np.sum(c))
    if kind == "esacf":
        x = synth.s_poly(seed, 44100, 2046 * size)
        t0 = time.perf_counter()
        c = rn.esacf(x, 44100)
        return time.perf_counter() - t0, size, float(np.sum(c))
    if kind == "iterf0":
        t, n = 0.0, 0
        for i in range(size):
            x = synth.s_poly(seed + 1000 * i, 22050, 65536)
            t0 = time.perf_counter()
            c = rn.iterf0(x, 22050)
            t += time.perf_counter() - t0
            n += 8
        return t, n, float(np.sum(c))
    if kind == "all4":
        t = 0.0
        for i in range(size):
            x = synth.s_poly(seed + 1000 * i, 22050, 44100)
            t0 = time.perf_counter()
            c = rn.esacf(x, 22050) + rn.harmonic_energy(x, 22050) + rn.iterf0(x, 22050) + rn.prime(x, 22050)
            t += time.perf_counter() - t0
        return t, size, float(np.sum(c))
    raise ValueError(kind)


def _cpu_pool(cores):
    import multiprocessing as mp

    pool = mp.get_context("spawn").Pool(cores)
    pool.map(_cpu_worker, [("he", 0, 8)] * cores)  # warm the workers (imports) outside any timing
    return pool


def _cpu_rate(pool, kind, size, cores):
    """-> (units/s on `cores` workers, seconds): every worker times its own loop (input synthesis
    excluded); the job time is the slowest worker's."""
    res = pool.map(_cpu_worker, [(kind, 100 + i, size) for i in range(cores)])
    dt = max(r[0] for r in res)
    return sum(r[1] for r in res) / dt, dt


def cpu_baseline(frames_per_core=40000, cores=None, pool=None, one_core=True):
    """Oracle port timed on all host cores (and on ONE core): each worker runs the
    reference's Python per-frame loop on its own shard of `frames_per_core` frames."""
    cores = cores or os.cpu_count() or 1
    own = pool is None
    if own:
        pool = _cpu_pool(cores)
    try:
        v, dt = _cpu_rate(pool, "he", frames_per_core, cores)
        v1 = None
        if one_core:
            v1, dt1 = _cpu_rate(pool, "he", frames_per_core, 1)
    finally:
        if own:
            pool.close()
            pool.join()
    total = frames_per_core * cores
    out = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
           "sample": "%d frames (%d per core x %d cores) of the 2048/512 @44.1 kHz workload, "
                     "oracle/ref_numpy.harmonic_energy (reference-style per-frame loop), %.1f s"
                     % (total, frames_per_core, cores, dt),
           "note": PORT_NOTE}
    if v1 is not None:
        out["one_core"] = {"value": v1, "unit": "frames/s", "cores": 1,
                           "sample": "%d frames on one core, %.1f s" % (frames_per_core, dt1)}
    return out, dt


def cpu_baseline_secondary(pool, cores):
    """Bounded CPU samples of C3 / C4 / C5 (all cores and one core)."""
    out = {}
    for key, kind, size, unit, what in (
            ("c3_esacf", "esacf", 150, "frames/s", "ESACF frames of 2046 samples @44.1 kHz"),
            ("c4_iterf0", "iterf0", 1, "frames/s", "iterative-F0 clips of 65 536 samples @22.05 kHz (8 frames each)"),
            ("c5_all4", "all4", 1, "clips/s", "clips of 44 100 samples @22.05 kHz through all four methods")):
        v, dt = _cpu_rate(pool, kind, size, cores)
        v1, dt1 = _cpu_rate(pool, kind, size, 1)
        out[key] = {"value": v, "unit": unit, "cores": cores, "kind": "port",
                    "sample": "%d per core x %d cores: %s (oracle/ref_numpy), %.1f s" % (size, cores, what, dt),
                    "one_core": {"value": v1, "unit": unit, "cores": 1, "sample": "%d on one core, %.1f s" % (size, dt1)}}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals, secs = [], []
    pool = _cpu_pool(cores)
    try:
        for _ in range(max(1, min(args.warmup, 2))):
            cpu_baseline(frames_per_core=200, cores=cores, pool=pool, one_core=False)
        steps = max(1, min(args.steps, 5))
        for _ in range(steps):
            cb, dt = cpu_baseline(frames_per_core=REF_FRAMES_PER_CORE, cores=cores, pool=pool,
                                  one_core=False)
            vals.append(cb["value"])
            secs.append(dt)
        v1, dt1 = _cpu_rate(pool, "he", REF_FRAMES_PER_CORE, 1)
        sec = cpu_baseline_secondary(pool, cores) if not args.no_secondary else None
    finally:
        pool.close()
        pool.join()
    v = sorted(vals)[len(vals) // 2]
    cb["value"] = v
    cb["one_core"] = {"value": v1, "unit": "frames/s", "cores": 1,
                      "sample": "%d frames on one core, %.1f s" % (REF_FRAMES_PER_CORE, dt1)}
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "harmonic-energy chromagram, 44.1 kHz mono, 2048-pt frames hop 512 "
                               "(BASELINE configs[1]); bounded sample per step on host cores",
                   "frames_per_step": REF_FRAMES_PER_CORE * cores},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if sec:
        line["secondary"] = sec
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# secondary configs (BASELINE.json configs[2..4]) on the CUDA path
# ------------------------------------------------------------------------------------------------
def _tiled(seed, fs, n, rows, dev, base_rows=8):
    """rows clips: base_rows distinct S-poly clips, tiled, with a per-sample amplitude modulation so
    no two clips are identical."""
    import numpy as np
    import torch

    from chord_detection_b200 import synth

    base = torch.from_numpy(np.stack([synth.s_poly(seed + i, fs, n) for i in range(base_rows)])).to(dev)
    x = base.repeat((rows + base_rows - 1) // base_rows, 1)[:rows].contiguous()
    g = torch.Generator(device=dev).manual_seed(seed)
    # in chunks: a [rows, n] rand tensor next to x would double the footprint
    for r0 in range(0, rows, 4096):
        xs = x[r0:r0 + 4096]
        xs.mul_(0.8 + 0.4 * torch.rand(xs.shape, device=dev, generator=g))
    return x


def _share(prof):
    tot = sum(prof.values()) or 1.0
    k = max(prof, key=prof.get)
    return {"dominant_kernel": k, "share_of_kernel_time": prof[k] / tot,
            "kernel_ms": {n: round(v, 3) for n, v in sorted(prof.items(), key=lambda kv: -kv[1])}}


def run_secondary(dev, rank, world, peak, reps=2):
    """C3 / C4 / C5 on the CUDA path.  Device-timed (CUDA events on the launching stream, max over
    ranks); inputs resident in HBM and larger than L2; per-kernel shares from the library's own
    event marks (cdb_profile_*) on an extra, untimed pass."""
    import torch
    import torch.distributed as dist

    from chord_detection_b200 import _native as nat, distributed as D, ops

    h = nat.Handle.get(dev.index)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        ts = []
        for _ in range(reps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([min(ts)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def profiled(fn):
        barrier()
        h.profile_start()
        fn()
        return h.profile_stop()

    out = {}
    # ---- C3: ESACF, 44.1 kHz, 1024 clips x 1 000 000 samples per GPU (489 frames of 2046 per clip)
    nc, n = 1024, 1_000_000
    x = _tiled(1 + 7 * rank, 44100, n, nc, dev)
    fn = lambda: ops.esacf(x, 44100)  # noqa: E731
    fn()
    ms = timed(fn, reps)
    frames = nc * 489
    prof = profiled(fn)
    ach = frames * 8184 / (ms * 1e-3) / 1e9
    out["c3_esacf"] = {
        "config": "ESACF two-channel SACF + enhance + peak fit, 44.1 kHz, 1024 clips x 1 000 000 samples "
                  "(500 736 frames of 2046) per GPU (BASELINE configs[2])",
        "value": world * frames / (ms * 1e-3), "unit": "frames/s", "ms": ms, "n_gpus": world, "scaling": "weak",
        "dtype": "f64", "input_bytes_per_gpu": int(x.numel() * 4),
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "algorithmic_bytes_per_frame": 8184,
                     "binding_roof": "FP64 issue / latency of the Levenberg-Marquardt fits (125-300 flop/B, "
                                     "SURVEY.md 8d), not HBM"},
        **_share(prof)}
    del x
    torch.cuda.empty_cache()
    # ---- C4: iterative F0, 65 536 frames = 8192 clips x 65 536 samples @22.05 kHz, clips sharded
    c0, c1 = D.shard_range(8192, rank, world)
    x = _tiled(3 + 7 * rank, 22050, 65536, c1 - c0, dev)
    fn = lambda: ops.iterative_f0(x, 22050)  # noqa: E731
    ops.iterative_f0(x[: min(c1 - c0, 2048)], 22050)  # workspace + plan tables
    fn()
    ms = timed(fn, reps)
    prof = profiled(fn)
    ach = (c1 - c0) * 8 * 32768 / (ms * 1e-3) / 1e9
    out["c4_iterf0"] = {
        "config": "iterative F0 (auditory filterbank + summary spectrum + periodicity), 22.05 kHz, 65 536 "
                  "frames of 8192 = 8192 clips x 65 536 samples sharded by clip over the GPUs (BASELINE configs[3])",
        "value": 65536 / (ms * 1e-3), "unit": "frames/s", "ms": ms, "n_gpus": world, "scaling": "strong",
        "dtype": "f64 filterbank / f32 spectrum / f64 periodicity", "clips_per_gpu": c1 - c0,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "algorithmic_bytes_per_frame": 32768, "per_gpu": True,
                     "binding_roof": "FP64 issue (17 IIR sections x 70 channels per sample) and the 70 "
                                     "16384-point FFTs per frame (1.4e8 flop/frame), not HBM"},
        **_share(prof)}
    del x
    torch.cuda.empty_cache()
    # ---- C5: all four methods over 100 000 clips (44 100 samples @22.05 kHz), clips sharded, ONE
    # all-reduce of the [4, 12] sums, batched pack + key per clip
    c0, c1 = D.shard_range(100_000, rank, world)
    x = _tiled(5 + 7 * rank, 22050, 44100, c1 - c0, dev, base_rows=64)
    chunk = 4096
    sums = torch.zeros((4, 12), dtype=torch.float64, device=dev)
    n_keys = [0]

    def c5(n_local):
        sums.zero_()
        n_keys[0] = 0
        for s in range(0, n_local, chunk):
            part, per_clip = D.all_methods_sharded(x[s:s + chunk], 22050, reduce=False)
            sums.add_(part)
            for m, pc in per_clip.items():
                digits, keys = ops.pack_and_key(pc, resolve=False)
                n_keys[0] += keys.numel()
        D.all_reduce_chroma(sums)

    c5(min(c1 - c0, chunk))  # warm-up: workspaces, plan tables, NCCL
    ms = timed(lambda: c5(c1 - c0), 1)
    keys_timed = n_keys[0]
    prof = profiled(lambda: c5(min(c1 - c0, 2 * chunk)))
    ach = (c1 - c0) * 44100 * 4 / (ms * 1e-3) / 1e9
    out["c5_all4"] = {
        "config": "all 4 methods over 100 000 synthetic 3-6-note polyphonic clips (44 100 samples @22.05 kHz), "
                  "clips sharded over the GPUs, one all-reduce of the [4, 12] chroma sums, batched digits + key "
                  "per clip and method (BASELINE configs[4])",
        "value": 100_000 / (ms * 1e-3), "unit": "clips/s", "ms": ms, "n_gpus": world, "scaling": "strong",
        "clips_per_gpu": c1 - c0, "keys_per_gpu": keys_timed,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "algorithmic_bytes_per_clip": 44100 * 4, "per_gpu": True,
                     "binding_roof": "FP64 issue (prime Goertzel, ESACF fits, IterF0 filterbank), not HBM"},
        **_share(prof)}
    del x
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-secondary", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"], help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist

    from chord_detection_b200 import distributed as D, ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime

        # a mismatched collective must fail loudly, not hang the driver
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    fused = world > 1 and args.collective == "fused" and D.comm_init(dev)

    nfr = args.frames
    n = nfr * HOP  # ceil(n/hop) = nfr frames, the last three zero-padded (dsp/frame.py tail rule)
    # synthetic polyphonic audio (S-poly): a 2^20-sample unique segment per buffer,
    # tiled and amplitude-modulated so no two frames are identical
    bufs = []
    for b in range(N_ROTATING):
        seg = torch.from_numpy(synth.s_poly_long(1000 * rank + b, FS, 1 << 20)).to(dev)
        reps = (n + seg.numel() - 1) // seg.numel()
        g = torch.Generator(device=dev).manual_seed(17 + b + 100 * rank)
        x = seg.repeat(reps)[:n] * (0.75 + 0.5 * torch.rand(n, device=dev, generator=g))
        bufs.append(x.contiguous())
    total = torch.zeros(12, dtype=torch.float64, device=dev)

    def step(i):
        # ONE kernel per step; for N > 1 the 12-double sum over the GPUs happens inside it
        ops.harmonic_energy(bufs[i % N_ROTATING], FS, frame_size=FRAME, hop=HOP, out_total=total,
                            allreduce=fused)
        if world > 1 and not fused:
            dist.all_reduce(total)  # NCCL fallback: one 12-double all-reduce per step, stream order

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()

    # ---- value: K steps, device time (CUDA events on the launching stream), max over ranks
    sampler = ClockSampler(local_rank)
    launches0 = ops.launch_count(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    prof = os.environ.get("CDB_PROFILE_RANGE") == "1"  # ncu --profile-from-start off
    if prof:
        torch.cuda.cudart().cudaProfilerStart()
    if world > 1:
        # one more untimed step after the host-side barrier: its in-kernel all-reduce ends on all
        # GPUs within an NVLink round trip, so the device-timed region below starts aligned on every
        # rank instead of charging the hosts' barrier-exit skew (tens of us) to the first timed step
        step(args.warmup)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    if prof:
        torch.cuda.cudart().cudaProfilerStop()
    ms_total = e0.elapsed_time(e1)
    launches = ops.launch_count(local_rank) - launches0 - (1 if world > 1 else 0)  # minus the aligning step
    if world > 1 and not fused:
        launches += args.steps  # the NCCL all-reduce kernel of every step
    clock_note = "sampled during the timed region"
    if len(sampler.samples) < 8:
        # the timed region was shorter than a few NVML polls: keep the same kernel running
        # (untimed) for ~0.3 s so that the clocks / throttle reasons are sampled under this load.
        # LOCAL work only: this branch and its trip count differ between ranks, so it must not
        # contain a collective (a mismatched all-reduce here deadlocked the N = 2 run of r01B).
        clock_note = "timed region too short for NVML polling; sampled under the same kernel loop right after it"
        t_end = time.perf_counter() + 0.3
        i = 0
        while time.perf_counter() < t_end:
            for _ in range(20):
                ops.harmonic_energy(bufs[i % N_ROTATING], FS, frame_size=FRAME, hop=HOP, out_total=total)
                i += 1
            torch.cuda.synchronize()
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    # ---- dominant kernel alone (roofline numerator): the same K launches back to back between two
    # events, no collective (identical to the timed region at N = 1)
    ksteps = max(10, args.steps)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(ksteps):
        ops.harmonic_energy(bufs[i % N_ROTATING], FS, frame_size=FRAME, hop=HOP, out_total=total)
    k1.record()
    k1.synchronize()
    kernel_ms = k0.elapsed_time(k1) / ksteps

    # ---- e2e: host (pinned) buffers through the public API, H2D + compute + D2H every step; for
    # N > 1 the cross-GPU sum rides in the last chunk's kernel
    host = bufs[0].cpu().pin_memory()
    pipe = ops.HostPipeline(dev, FS, FRAME, hop=HOP, chunk_frames=25600)
    for _ in range(2):
        pipe.run(host, allreduce=fused)
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = pipe.run(host, allreduce=fused)
        if world > 1 and not fused:
            t = torch.from_numpy(out).to(dev)
            dist.all_reduce(t)
            out = t.cpu().numpy()
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    barrier()

    # ---- the same end-to-end path with PCM16 on the wire (WAV payload as stored):
    # informational, the headline e2e above stays the float32 signal of BASELINE's config
    host16 = (bufs[0] * 32767.0).round().clamp_(-32768, 32767).to(torch.int16).cpu().pin_memory()
    pipe16 = ops.HostPipeline(dev, FS, FRAME, hop=HOP, chunk_frames=25600, dtype=torch.int16)
    for _ in range(2):
        pipe16.run(host16, allreduce=fused)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pipe16.run(host16, allreduce=fused)
    torch.cuda.synchronize()
    e2e16_ms = 1e3 * (time.perf_counter() - t0)
    barrier()

    times = torch.tensor([ms_total, e2e_ms, e2e16_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, e2e16_ms = float(times[0]), float(times[1]), float(times[2])
    h2d_bytes, h2d16_bytes = int(pipe.h2d_bytes), int(pipe16.h2d_bytes)
    del bufs, host, host16, pipe, pipe16
    torch.cuda.empty_cache()

    peak, peak_src = _peaks()
    secondary = None
    if not args.no_secondary:
        secondary = run_secondary(dev, rank, world, peak)

    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = world * nfr / (ms_per_step * 1e-3)
        achieved = nfr * ALG_BYTES_PER_FRAME / (kernel_ms * 1e-3) / 1e9
        fps_kernel = nfr / (kernel_ms * 1e-3)
        coll = ("none (1 GPU)" if world == 1 else
                "fused into the kernel: the last CTA of every rank stores its 12 doubles + a flag into every "
                "peer's mailbox over NVLink (P2P stores, CUDA IPC mappings), waits for the peers' flags and sums "
                "in rank order (CDB_FLAG_ALLREDUCE); no NCCL call in the timed region" if fused else
                "one 12-double NCCL all-reduce per step, in stream order (CUDA IPC unavailable or --collective nccl)")
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "harmonic-energy chromagram, 44.1 kHz mono fp32, 2048-pt frames, "
                                   "hop 512, %d frames per GPU (BASELINE configs[1])" % nfr,
                       "window": "hamming (reference harmonic_energy.py:42)",
                       "accumulate": "fp64", "parallelism": "frames sharded, dp%d" % world,
                       "collective": coll,
                       "l2": "%d rotating %.1f MB inputs (> 126 MB L2): every step reads cold data"
                             % (N_ROTATING, n * 4 / 1e6)},
            "gpu_launches": int(launches),
            "gpu_launches_note": "one kernel per step; the output is written by the kernel's last CTA, so there "
                                 "is no memset node (and for N > 1 no collective kernel) in the timed region",
            "clocks": dict(sampler.summary(), note=clock_note),
            "e2e": {"value": world * nfr * e2e_steps / (e2e_ms * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 96,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "path": "ops.HostPipeline: pinned host signal -> chunked H2D on a copy stream "
                            "overlapped with cdb_he_chroma on a compute stream -> 12 doubles D2H; wall clock, "
                            "max over ranks"},
            "e2e_pcm16": {"value": world * nfr * e2e_steps / (e2e16_ms * 1e-3), "unit": "frames/s",
                          "h2d_bytes_per_step": h2d16_bytes, "d2h_bytes_per_step": 96,
                          "ms_per_step": e2e16_ms / e2e_steps,
                          "note": "same pipeline, int16 PCM on the wire (CDB_FLAG_PCM16, decoded in the "
                                  "kernel, bit-identical chroma); informational, not the headline"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": _traffic(), "peak_source": peak_src,
                         "kernel": "he2048w_kernel<16,5,false>", "kernel_ms": kernel_ms,
                         "kernel_ms_how": "%d back-to-back launches between two CUDA events, no collective" % ksteps,
                         "algorithmic_bytes_per_launch": nfr * ALG_BYTES_PER_FRAME,
                         "binding_roof": "shared-memory wavefronts + FMA pipe, not HBM (29 flop/B, SURVEY.md "
                                         "8d; profiles/, DESIGN.md 3.1)",
                         "fp32_tflops": fps_kernel * ALG_FLOP_PER_FRAME / 1e12,
                         "fp32_frac_of_nominal": fps_kernel * ALG_FLOP_PER_FRAME / 1e12 / FP32_PEAK_TFLOPS},
        }
        if secondary:
            line["secondary"] = secondary
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            pool = _cpu_pool(cores)
            try:
                line["cpu_baseline"], _ = cpu_baseline(cores=cores, pool=pool)
                if secondary:
                    for k, v in cpu_baseline_secondary(pool, cores).items():
                        line["secondary"][k]["cpu_baseline"] = v
            finally:
                pool.close()
                pool.join()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
