"""GPU tests of the PCM16 ingestion path: cdb_pcm16_to_mono_f32 and the decode fused
into the frame-2048 harmonic-energy kernel (CDB_FLAG_PCM16).  Both must reproduce, bit for bit, the
float32 signal soundfile + librosa.to_mono hand the reference (s/32768, float32 mean over channels),
so every downstream parity statement carries over unchanged."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import cases, ref_numpy as rn  # noqa: E402


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("n,channels", [(0, 1), (1, 1), (7, 1), (8, 1), (100003, 1), (4096, 2), (999, 3), (50, 6)])
def test_pcm16_to_mono_is_exact(n, channels):
    from chord_detection_b200 import ops

    rng = np.random.default_rng(n + channels)
    pcm = rng.integers(-32768, 32768, size=(n, channels) if channels > 1 else (n,), dtype=np.int16)
    if n > 4:
        pcm.reshape(-1)[:4] = [-32768, 32767, 0, -1]
    got = ops.pcm16_to_mono(torch.from_numpy(pcm).to(_dev())).cpu().numpy()
    f = pcm.astype(np.float32) / np.float32(32768.0)       # soundfile PCM_16 -> float32
    want = f if channels == 1 else np.mean(f, axis=1)       # librosa.to_mono (float32 mean)
    assert got.dtype == np.float32 and got.shape == (n,)
    assert np.array_equal(got, want.astype(np.float32))


def test_pcm16_unaligned_views():
    from chord_detection_b200 import ops

    pcm = torch.from_numpy(np.arange(-500, 500, dtype=np.int16)).to(_dev())
    got = ops.pcm16_to_mono(pcm[3:]).cpu().numpy()          # 2-byte aligned only: scalar path
    assert np.array_equal(got, np.arange(-497, 500, dtype=np.float32) / np.float32(32768.0))


@pytest.mark.parametrize("hop", [512, 2048, 300])
def test_he_pcm16_equals_float_input_bitwise(hop):
    """The fused decode folds 1/32768 into the window constants (a power of two): per-frame chroma
    must be IDENTICAL to feeding the float32 samples."""
    from chord_detection_b200 import ops

    x, fs = cases.make_input(dict(fn="s_poly_long", seed=21, fs=44100, n=60000))
    q = np.clip(np.round(x.astype(np.float64) * 32768.0 * 1.7), -32768, 32767).astype(np.int16)
    xf = q.astype(np.float32) / np.float32(32768.0)
    a = ops.harmonic_energy(torch.from_numpy(q).to(_dev()), fs, frame_size=2048, hop=hop, per_frame=True)
    b = ops.harmonic_energy(torch.from_numpy(xf).to(_dev()), fs, frame_size=2048, hop=hop, per_frame=True)
    torch.cuda.synchronize()
    assert torch.equal(a.frames, b.frames)
    assert torch.equal(a.total, b.total)
    want = rn.harmonic_energy_fast(xf, fs, frame_size=2048, hop=hop)
    got = a.total.cpu().numpy()
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 1e-4
    assert rn.pack_chroma(got) == rn.pack_chroma(want)


def test_he_pcm16_batch_and_other_frame_sizes():
    from chord_detection_b200 import ops

    rows = np.stack([cases.make_input(dict(fn="s_poly", seed=70 + i, fs=22050, n=20000))[0] for i in range(5)])
    q = np.clip(np.round(rows.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)
    xf = q.astype(np.float32) / np.float32(32768.0)
    qd, fd = torch.from_numpy(q).to(_dev()), torch.from_numpy(xf).to(_dev())
    for kw in (dict(frame_size=2048, hop=512), dict(frame_size=8192), dict(frame_size=1024)):
        a = ops.harmonic_energy(qd, 22050, per_clip=True, **kw)
        b = ops.harmonic_energy(fd, 22050, per_clip=True, **kw)
        torch.cuda.synchronize()
        assert torch.equal(a.clips, b.clips), kw


def test_pcm16_golden_clips_through_int16_wire():
    """The reference's own test clips as soundfile would store them (PCM_16): feed the int16 samples,
    compare with the golden vectors of the unmodified reference for the float signal."""
    from chord_detection_b200 import ops, synth

    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "reference_golden.json")) as f:
        golden = json.load(f)["cases"]
    ids = sorted(k for k, v in golden.items() if k.startswith("clips_pcm16/") and v["method"] == 2)
    assert ids
    for cid in ids:
        g = golden[cid]
        x, fs = cases.make_input(g["input"])
        q = np.round(x.astype(np.float64) * 32768.0).astype(np.int16)
        assert np.array_equal(q.astype(np.float32) / np.float32(32768.0), x)
        xd = ops.pcm16_to_mono(torch.from_numpy(q).to(_dev()))
        got = ops.harmonic_energy(xd, fs, **g["kwargs"]).total.cpu().numpy()
        want = np.asarray(g["chroma"], dtype=np.float64)
        assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 1e-4, cid
        assert rn.pack_chroma(got) == g["digits"], cid


def test_host_pipeline_pcm16_wire():
    from chord_detection_b200 import ops

    x, fs = cases.make_input(dict(fn="s_poly_long", seed=9, fs=44100, n=3000 * 512))
    q = np.clip(np.round(x.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)
    xf = torch.from_numpy(q.astype(np.float32) / np.float32(32768.0)).to(_dev())
    want = ops.harmonic_energy(xf, fs, frame_size=2048, hop=512).total.cpu().numpy()
    pipe = ops.HostPipeline(_dev(), fs, 2048, hop=512, chunk_frames=1024, dtype=torch.int16)
    got = pipe.run(torch.from_numpy(q).pin_memory())
    assert pipe.h2d_bytes < 1.01 * q.nbytes + 16 * 4096
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 1e-9
    with pytest.raises(ValueError):
        pipe.run(torch.from_numpy(x))


def test_device_load_matches_host_load(tmp_path):
    """audio.load_device (int16 on the wire, mono down-mix and polyphase resampling on the GPU)
    against audio.load (scipy on the host) for mono / stereo clips at several sample rates."""
    import scipy.io.wavfile as wavfile

    from chord_detection_b200 import audio

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(8)
    for i, (fs, ch, n) in enumerate(((44100, 1, 30011), (48000, 2, 20000), (22050, 1, 5000),
                                     (16000, 2, 7001), (8000, 1, 999),
                                     # 12 801 taps (> 48 KB of shared memory) and 882 021 taps (L2 path)
                                     (32000, 1, 8000), (96000, 2, 9000), (44101, 1, 3000))):
        pcm = rng.integers(-20000, 20000, size=(n, ch) if ch > 1 else n).astype(np.int16)
        path = os.path.join(tmp_path, "c%d.wav" % i)
        wavfile.write(path, fs, pcm)
        want, fs_w = audio.load(path)
        got, fs_g = audio.load_device(path, dev)
        assert fs_w == fs_g == 22050
        got = got.cpu().numpy()
        assert got.shape == want.shape and got.dtype == np.float32
        assert np.max(np.abs(got - want)) <= 1e-6 * max(np.max(np.abs(want)), 1e-30), (fs, ch)
    with pytest.raises(ValueError):
        from chord_detection_b200 import ops
        ops.resample_poly(torch.zeros(4), 44100, 22050)  # CPU tensor: no fallback
