"""Batch front end: many clips x any of the four methods in one run.

The reference CLI (/root/reference/chord_detection/chord_detect.py:45-63) handles one clip per
process and its classes one clip per object; config C5 (100 000 clips x 4 methods) needs the
opposite shape.  Here clips are loaded like the reference does (mono, 22 050 Hz, float32), grouped
by length (a batch is a [n_clips, clip_len] tensor: frames never span clips, and zero-padding a
clip to a common length would change the iterative-F0 result because its filterbank rings on past
the end), run through the batched device ops, and packed to the 12-digit strings / keys on the
device (cdb_pack_and_key).  Under torchrun every rank takes a contiguous shard of the clip list; the
per-clip lines are gathered on rank 0 and the per-method corpus sums are all-reduced once.
"""
import os
from collections import OrderedDict
from pathlib import Path

import numpy as np

AUDIO_EXT = (".wav", ".wave")


def expand_inputs(inputs):
    """paths / directories (recursive, *.wav) / manifests (*.txt or *.lst, one path per line,
    '#' comments, relative to the manifest) -> ordered list of clip paths."""
    out = []
    for item in inputs:
        p = Path(item)
        if p.is_dir():
            out.extend(sorted(q for q in p.rglob("*") if q.suffix.lower() in AUDIO_EXT))
        elif p.suffix.lower() in (".txt", ".lst"):
            for line in p.read_text().splitlines():
                line = line.strip()
                if line and not line.startswith("#"):
                    q = Path(line)
                    out.append(q if q.is_absolute() else p.parent / q)
        else:
            out.append(p)
    return out


def run_batch(paths, methods, key=False, device=None, max_batch_bytes=1 << 30, loader=None):
    """-> (lines per clip: list of (path, [(method_number, display_name, digits, key or None)]),
           corpus sums {method_number: float64[12]} over THIS rank's clips)."""
    import torch

    from . import METHODS, audio, ops

    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    # librosa.load on the device (decode / down-mix / resample); clips come back as
    # CUDA tensors and are stacked on the device
    load = loader or (lambda p: audio.load_device(p, dev))
    fn = {1: ops.esacf, 2: ops.harmonic_energy, 3: ops.iterative_f0, 4: ops.prime_multif0}
    for m in methods:
        if m not in fn:
            raise ValueError("valid methods: {0}".format(", ".join(str(k) for k in METHODS)))
    clips, groups = [], OrderedDict()
    for i, p in enumerate(paths):
        x, fs = load(p)
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        if x.dim() != 1:
            raise ValueError("Only 1D numpy ndarrays are supported")  # dsp/frame.py:6-7
        clips.append(x)
        groups.setdefault((int(fs), int(x.shape[0])), []).append(i)
    results = [[] for _ in paths]
    sums = {m: np.zeros(12) for m in methods}
    for (fs, n), idxs in groups.items():
        if n == 0:
            for i in idxs:
                for m in methods:
                    results[i].append((m, METHODS[m].display_name(), "0" * 12, None))
            continue
        per = max(1, int(max_batch_bytes // (4 * n)))
        for s in range(0, len(idxs), per):
            part = idxs[s:s + per]
            xd = torch.stack([clips[i].to(dev) for i in part])
            for m in methods:
                r = fn[m](xd, fs, per_clip=True)
                digits, keys = ops.pack_and_key(r.clips, resolve=key)
                sums[m] += r.total.cpu().numpy()
                dg = digits.cpu().numpy()
                for j, i in enumerate(part):
                    ks = keys[j] if key else None
                    results[i].append((m, METHODS[m].display_name(), "".join(str(int(d)) for d in dg[j]), ks))
    return [(str(p), results[i]) for i, p in enumerate(paths)], sums


def main_batch(inputs, methods, key=False, out=None):
    """Batch entry point of the CLI.  Single process or torchrun (RANK / WORLD_SIZE in the env)."""
    import sys

    from . import distributed as D

    out = out or sys.stdout
    paths = expand_inputs(inputs)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist

        rank, world, dev = D.init()
        lo, hi = D.shard_range(len(paths), rank, world)
        lines, sums = run_batch(paths[lo:hi], methods, key, dev)
        gathered = [None] * world
        dist.all_gather_object(gathered, lines)
        t = torch.tensor(np.stack([sums[m] for m in methods]), dtype=torch.float64, device=dev)
        D.all_reduce_chroma(t)  # the one collective on chroma data: [n_methods, 12] doubles
        sums = {m: t[i].cpu().numpy() for i, m in enumerate(methods)}
        lines = [ln for part in gathered for ln in part]
        if rank != 0:
            dist.destroy_process_group()
            return lines, sums
        dist.destroy_process_group()
    else:
        lines, sums = run_batch(paths, methods, key)
    for path, res in lines:
        out.write("{0}\n".format(path))
        for m, name, digits, ks in res:
            out.write("{0} - {1}\n{2}\n".format(m, name, digits))
            if ks is not None:
                out.write("{0}\n".format(ks))
    from . import METHODS
    from .chromagram import Chromagram

    out.write("== corpus ({0} clips)\n".format(len(lines)))
    for m in methods:
        c = Chromagram(sums[m])
        out.write("{0} - {1}\n{2}\n".format(m, METHODS[m].display_name(), c))
        if key:
            out.write("{0}\n".format(c.key()))
    return lines, sums
