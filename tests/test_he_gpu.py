"""GPU parity tests for the harmonic-energy path (cdb_he_chroma through ctypes) against
(i) the golden vectors made by the unmodified reference and (ii) the numpy oracle.

Tolerance (BASELINE.json north_star): raw float chroma within 1e-4 relative; 12-digit string identical.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import cases, ref_numpy as rn  # noqa: E402

RTOL = 1e-4


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _he(x, fs, **kw):
    from chord_detection_b200 import ops

    xd = torch.from_numpy(np.ascontiguousarray(x)).to(_dev())
    res = ops.harmonic_energy(xd, fs, **kw)
    torch.cuda.synchronize()
    return res


def _assert_close(got, want, tol=RTOL):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = max(np.max(np.abs(want)), 1e-300)
    assert np.max(np.abs(got - want)) / scale <= tol, (got, want)
    nz = np.abs(want) > 1e-6 * scale
    assert np.all(np.abs(got[nz] - want[nz]) <= tol * np.abs(want[nz])), (got, want)


def _he_golden_ids():
    import json

    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "reference_golden.json")) as f:
        g = json.load(f)
    return sorted(k for k, v in g["cases"].items() if v["method"] == 2)


@pytest.mark.parametrize("cid", _he_golden_ids())
def test_he_matches_reference_golden(golden, cid):
    g = golden["cases"][cid]
    x, fs = cases.make_input(g["input"])
    res = _he(x, fs, **g["kwargs"])
    got = res.total.cpu().numpy()
    _assert_close(got, g["chroma"])
    assert rn.pack_chroma(got) == g["digits"]


def test_he_hop_matches_reference_offsets(golden):
    for cid, g in golden["hop_cases"].items():
        x, fs = cases.make_input(g["input"])
        got = _he(x, fs, **g["kwargs"]).total.cpu().numpy()
        _assert_close(got, g["chroma"])


@pytest.mark.parametrize("frame_size,hop", [(2048, 512), (2048, 2048), (2048, 300), (2048, 1),
                                            (8192, None), (1024, 256), (64, 64), (16384, 4096)])
def test_he_per_frame_matches_oracle(frame_size, hop):
    fs = 44100 if frame_size == 2048 else 22050
    n = 30000 if hop != 1 else 2300
    x, _ = cases.make_input(dict(fn="s_poly", seed=31, fs=fs, n=n))
    if frame_size == 64:
        kw = dict(num_harmonic=1, num_octave=1, num_bins=1)
        fs = 2000  # keeps the probed bins inside the 33 rfft bins
    else:
        kw = {}
    res = _he(x, fs, frame_size=frame_size, hop=hop, per_frame=True, **kw)
    want_total, want_frames = rn.harmonic_energy_fast(x, fs, frame_size=frame_size, hop=hop,
                                                      per_frame=True, **kw)
    _assert_close(res.total.cpu().numpy(), want_total)
    got_frames = res.frames.cpu().numpy()
    assert got_frames.shape == want_frames.shape
    _assert_close(got_frames, want_frames, tol=2e-4)


@pytest.mark.parametrize("frame_size,hop,fs", [(2048, 512, 44100), (8192, None, 22050),
                                               (8192, 2048, 22050), (8192, 1001, 44100)])
def test_he_fast_and_generic_kernels_agree(monkeypatch, frame_size, hop, fs):
    """The register-FFT kernels (he2048w_kernel, he8192_kernel) against the generic radix-2 kernel."""
    x, fs = cases.make_input(dict(fn="s_poly_long", seed=2, fs=fs, n=200 * 512 + 77))
    a = _he(x, fs, frame_size=frame_size, hop=hop, per_frame=True)
    monkeypatch.setenv("CDB_HE_FORCE_GENERIC", "1")
    b = _he(x, fs, frame_size=frame_size, hop=hop, per_frame=True)
    # (two fp32 FFTs of different structure: radix-2 in shared memory vs radix-32 / radix-64 in
    # registers; the parity bar against the oracle is 1e-4 and is tested separately)
    _assert_close(a.total.cpu().numpy(), b.total.cpu().numpy(), tol=1e-5 if frame_size == 2048 else 5e-5)
    _assert_close(a.frames.cpu().numpy(), b.frames.cpu().numpy(), tol=1e-4)
    if frame_size == 8192:
        # the frame-8192 kernels: scalar butterflies / packed butterflies / packed + the next frame
        # staged by bulk async copy / 64-thread teams with radix-64 x 64 register FFTs (ragged and
        # unaligned frames are read directly: hop 1001 and the clip tail exercise that path)
        monkeypatch.delenv("CDB_HE_FORCE_GENERIC")
        res = {}
        for mode in ("scalar", "packed", "staged", "team"):
            monkeypatch.setenv("CDB_HE8192", mode)
            res[mode] = _he(x, fs, frame_size=frame_size, hop=hop, per_frame=True).frames.cpu().numpy()
        # the radix-16^3 kernels share one butterfly structure; the radix-64^2 team kernel (the
        # default) rounds differently, like the generic kernel above
        for mode in ("scalar", "packed"):
            _assert_close(res[mode], res["staged"], tol=1e-5)
        _assert_close(res["team"], res["staged"], tol=1e-4)
        assert np.array_equal(res["team"], a.frames.cpu().numpy())  # team is the default
        monkeypatch.delenv("CDB_HE8192")
        xo = np.concatenate([np.zeros(1, dtype=np.float32), x])  # 4-byte-aligned view
        xd = torch.from_numpy(xo).to(_dev())[1:]
        from chord_detection_b200 import ops
        d = ops.harmonic_energy(xd, fs, frame_size=frame_size, hop=hop, per_frame=True)
        assert np.array_equal(d.frames.cpu().numpy(), a.frames.cpu().numpy())


def test_he_batch_of_clips_and_strided_rows():
    from chord_detection_b200 import ops

    rows = [cases.make_input(dict(fn="s_poly", seed=50 + i, fs=44100, n=10000))[0] for i in range(5)]
    want = np.stack([rn.harmonic_energy_fast(r, 44100, frame_size=2048, hop=512) for r in rows])
    big = torch.zeros((5, 10240), dtype=torch.float32, device=_dev())
    big[:, :10000] = torch.from_numpy(np.stack(rows)).to(_dev())
    res = ops.harmonic_energy(big[:, :10000], 44100, frame_size=2048, hop=512, per_clip=True)
    torch.cuda.synchronize()
    _assert_close(res.clips.cpu().numpy(), want)
    _assert_close(res.total.cpu().numpy(), want.sum(axis=0))
    # generic kernel, reference default frame size
    want8 = np.stack([rn.harmonic_energy_fast(r, 44100, frame_size=8192) for r in rows])
    res8 = ops.harmonic_energy(big[:, :10000], 44100, frame_size=8192, per_clip=True)
    _assert_close(res8.clips.cpu().numpy(), want8)


def test_he_sharded_frames_sum_to_whole():
    """shard a long signal by frame ranges with an N-hop halo; sums must add up."""
    from chord_detection_b200 import ops

    N, hop, nfr = 2048, 512, 1000
    x, fs = cases.make_input(dict(fn="s_poly_long", seed=5, fs=44100, n=nfr * hop))
    xd = torch.from_numpy(x).to(_dev())
    whole = ops.harmonic_energy(xd, fs, frame_size=N, hop=hop).total.cpu().numpy()
    acc = torch.zeros(12, dtype=torch.float64, device=_dev())
    for r in range(4):
        f0, f1 = r * nfr // 4, (r + 1) * nfr // 4
        seg = xd[f0 * hop : min(len(x), (f1 - 1) * hop + N)]
        ops.harmonic_energy(seg, fs, frame_size=N, hop=hop, frames_per_clip=f1 - f0,
                            out_total=acc, accumulate=True)
    torch.cuda.synchronize()
    _assert_close(acc.cpu().numpy(), whole, tol=1e-9)
    _assert_close(whole, rn.harmonic_energy_fast(x, fs, frame_size=N, hop=hop))


def test_he_full_size_properties():
    """Config C2 at full size (100 000 frames, 2048/512 @44.1 kHz): size-independent properties
    -- per-frame rows sum to the total, homogeneity chroma(a*x) = sqrt(a)*chroma(x), and a
    sampled subset of frames equals the oracle."""
    from chord_detection_b200 import ops

    N, hop, nfr = 2048, 512, 100000
    g = torch.Generator(device="cpu").manual_seed(0)
    seg = torch.from_numpy(cases.make_input(dict(fn="s_poly_long", seed=9, fs=44100, n=1 << 20))[0])
    reps = (nfr * hop + len(seg) - 1) // len(seg)
    x = (seg.repeat(reps)[: nfr * hop] * (0.5 + torch.rand(nfr * hop, generator=g))).contiguous()
    xd = x.to(_dev())
    res = ops.harmonic_energy(xd, 44100, frame_size=N, hop=hop, per_frame=True)
    torch.cuda.synchronize()
    assert res.frames.shape == (nfr, 12)
    tot = res.total.cpu().numpy()
    _assert_close(res.frames.double().sum(dim=0).cpu().numpy(), tot, tol=1e-6)
    res4 = ops.harmonic_energy(xd * 4.0, 44100, frame_size=N, hop=hop)
    _assert_close(res4.total.cpu().numpy(), 2.0 * tot, tol=1e-5)
    xs = x.numpy()
    for f in (0, 1, 777, 54321, nfr - 4, nfr - 1):
        w = rn.harmonic_energy_fast(xs[f * hop : f * hop + N], 44100, frame_size=N)
        _assert_close(res.frames[f].cpu().numpy(), w, tol=2e-4)


def test_he_class_api_and_errors():
    import chord_detection_b200 as cd
    from chord_detection_b200 import ops

    x, fs = cases.make_input(dict(fn="gen_test_clip", name="test_1_note_E4"))
    obj = cd.MultipitchHarmonicEnergy(x, fs=fs)
    c = obj.compute_pitches()
    assert isinstance(c, cd.Chromagram) and len(c) == 12
    assert repr(c) == "221111111111"  # SURVEY.md Appendix B / golden
    assert c.key() == "Cmin"
    assert cd.METHODS[2] is cd.MultipitchHarmonicEnergy
    with pytest.raises(ValueError):
        cd.MultipitchHarmonicEnergy(np.zeros((2, 100), dtype=np.float32), fs=fs)
    with pytest.raises(ValueError):  # not a power of two: device path refuses loudly
        ops.harmonic_energy(torch.zeros(5000, device=_dev()), fs, frame_size=3000)
    with pytest.raises(RuntimeError):  # CPU tensor: no CPU fallback
        ops.harmonic_energy(torch.zeros(5000), fs)


def test_he_host_pipeline_matches_device_path():
    from chord_detection_b200 import ops

    x, fs = cases.make_input(dict(fn="s_poly_long", seed=6, fs=44100, n=5000 * 512 + 77))
    xd = torch.from_numpy(x).to(_dev())
    want = ops.harmonic_energy(xd, fs, frame_size=2048, hop=512).total.cpu().numpy()
    pipe = ops.HostPipeline(_dev(), fs, 2048, hop=512, chunk_frames=1024)
    got = pipe.run(torch.from_numpy(x).pin_memory())
    _assert_close(got, want, tol=1e-9)
    assert pipe.h2d_bytes >= x.nbytes


@pytest.mark.parametrize("kw", [dict(), dict(num_bins=3), dict(num_bins=1, num_harmonic=3),
                                dict(num_bins=4, num_harmonic=1, num_octave=3), dict(num_bins=5),
                                dict(num_bins=5, num_octave=1), dict(num_bins=7, num_octave=1)])
def test_he2048_level_epilogue_is_bit_identical_to_the_gather(monkeypatch, kw):
    """CDB_HE_EPILOGUE=levels takes the frame-2048 kernel's window maxima from a sparse table of range
    maxima built with warp shuffles; the default (gather) scans every window bin by bin (measured
    6 % faster: fewer instructions).
    max() is exact in any order and the fp64 sums keep their order, so per-frame chroma, per-clip sums
    and totals must be bit-identical -- for power-of-two and other window widths (6, 10, 12, 14, 20,
    28), and for parameter sets that fall back to the gather (more than 64 windows, bins >= 192)."""
    from chord_detection_b200 import ops

    rows = np.stack([cases.make_input(dict(fn="s_poly", seed=700 + i, fs=44100, n=30 * 512 + 333))[0]
                     for i in range(4)])
    xd = torch.from_numpy(rows).to(_dev())
    b = ops.harmonic_energy(xd, 44100, frame_size=2048, hop=512, per_clip=True, per_frame=True, **kw)
    monkeypatch.setenv("CDB_HE_EPILOGUE", "levels")
    a = ops.harmonic_energy(xd, 44100, frame_size=2048, hop=512, per_clip=True, per_frame=True, **kw)
    torch.cuda.synchronize()
    assert torch.equal(a.frames, b.frames)
    # (per-clip sums are fp64 atomics from several warps: their order varies from run to run)
    _assert_close(a.clips.cpu().numpy(), b.clips.cpu().numpy(), tol=1e-13)
    want = np.stack([rn.harmonic_energy_fast(r, 44100, frame_size=2048, hop=512, **kw) for r in rows])
    _assert_close(a.clips.cpu().numpy(), want)
