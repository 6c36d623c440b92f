"""
TEST INFRASTRUCTURE — NOT PRODUCT CODE.

The parity case list: deterministic inputs (rebuilt from chord_detection_b200.synth
by spec, so no audio is committed) x methods x kwargs.  oracle/gen_golden.py runs
the unmodified reference on every case and writes tests/golden/reference_golden.json;
the tests rebuild the same inputs with make_input() on any box.
"""
import numpy as np

from chord_detection_b200 import synth


def make_input(spec):
    """spec: dict(fn=..., **args) -> (float32 array, fs)."""
    fn = spec["fn"]
    if fn == "gen_test_clip":
        return synth.gen_test_clips(pcm16=spec.get("pcm16", False))[spec["name"]], 22050
    if fn == "piano_like_cmaj":
        return synth.piano_like_cmaj(spec.get("fs", 22050), spec.get("n", 44100)), spec.get("fs", 22050)
    if fn == "s_poly":
        return synth.s_poly(spec["seed"], spec["fs"], spec["n"]), spec["fs"]
    if fn == "s_poly_long":
        return synth.s_poly_long(spec["seed"], spec["fs"], spec["n"]), spec["fs"]
    if fn == "noise":
        return synth.noise(spec["seed"], spec["n"], spec.get("sigma", 0.1)), spec["fs"]
    if fn == "silence":
        return np.zeros(spec["n"], dtype=np.float32), spec["fs"]
    if fn == "impulse":
        x = np.zeros(spec["n"], dtype=np.float32)
        x[spec.get("at", 0)] = 1.0
        return x, spec["fs"]
    raise ValueError(fn)


_CLIP_NAMES = [
    "test_1_note_Csharp3",
    "test_1_note_E4",
    "test_2_notes_E2_F3",
    "test_2_notes_G3_Asharp4",
    "test_3_notes_G2_B2_G#3",
]

ALL = (1, 2, 3, 4)


def case_list():
    """[(case_id, input_spec, method_number, kwargs)]"""
    cases = []

    def add(cid, spec, methods, **kw):
        for m in methods:
            cases.append(("%s/m%d%s" % (cid, m, "".join("_%s%s" % (k, v) for k, v in sorted(kw.items()))),
                          spec, m, dict(kw)))

    # C1 / reference test clips (tests/gen_test_clips.py), float and PCM16-clipped
    for nm in _CLIP_NAMES:
        add("clips/" + nm, dict(fn="gen_test_clip", name=nm), ALL)
        add("clips_pcm16/" + nm, dict(fn="gen_test_clip", name=nm, pcm16=True), ALL)
    add("piano_like", dict(fn="piano_like_cmaj"), ALL)
    # C5 shape: 3..6-note polyphonic clips, 22 050 Hz, 44 100 samples
    for seed in range(8):
        add("spoly22k/%d" % seed, dict(fn="s_poly", seed=seed, fs=22050, n=44100), ALL)
    for seed in range(8, 40):
        add("spoly22k/%d" % seed, dict(fn="s_poly", seed=seed, fs=22050, n=44100), (2,))
    for seed in range(8, 16):
        add("spoly22k/%d" % seed, dict(fn="s_poly", seed=seed, fs=22050, n=44100), (1, 4))
    # C2 shape: HE, 44.1 kHz, frame 2048 (hop = frame here; hop 512 is a derived case below)
    add("c2/long", dict(fn="s_poly_long", seed=0, fs=44100, n=256 * 512 + 1536), (2,), frame_size=2048)
    add("c2/noise", dict(fn="noise", seed=1, fs=44100, n=100000), (2,), frame_size=2048)
    add("he/fs44k_default", dict(fn="s_poly", seed=3, fs=44100, n=88200), (2,))
    add("he/params", dict(fn="s_poly", seed=4, fs=22050, n=44100), (2,),
        frame_size=4096, num_harmonic=3, num_octave=3, num_bins=1)
    add("he/params2", dict(fn="s_poly", seed=5, fs=22050, n=30001), (2,),
        frame_size=1024, num_harmonic=1, num_octave=1, num_bins=3)
    # C3 shape: ESACF @44.1 kHz (2046-sample frames)
    for seed in range(4):
        add("c3/%d" % seed, dict(fn="s_poly", seed=100 + seed, fs=44100, n=8 * 2046 + 100), (1,))
    # C4 shape: IterF0, 8 frames of 8192 @22 050
    for seed in range(2):
        add("c4/%d" % seed, dict(fn="s_poly", seed=200 + seed, fs=22050, n=65536), (3,))
    # edge cases: ragged / short / exact-multiple / silence / impulse
    add("edge/short", dict(fn="s_poly", seed=7, fs=22050, n=700), ALL)
    add("edge/ragged", dict(fn="s_poly", seed=8, fs=22050, n=10000), ALL)
    add("edge/exact", dict(fn="s_poly", seed=9, fs=22050, n=16384), (2, 3))
    add("edge/exact_esacf", dict(fn="s_poly", seed=9, fs=22050, n=1023 * 4), (1,))
    add("edge/silence", dict(fn="silence", fs=22050, n=9000), ALL)
    add("edge/impulse", dict(fn="impulse", fs=22050, n=9000, at=100), ALL)
    add("edge/noise", dict(fn="noise", seed=2, fs=22050, n=20000), ALL)
    return cases


def hop_case_list():
    """HE with hop < frame: oracle = reference on x[off:], summed."""
    return [
        ("c2hop/long", dict(fn="s_poly_long", seed=0, fs=44100, n=256 * 512), dict(frame_size=2048, hop=512)),
        ("c2hop/noise", dict(fn="noise", seed=3, fs=44100, n=50001), dict(frame_size=2048, hop=512)),
        ("hop/8192_2048", dict(fn="s_poly", seed=11, fs=22050, n=44100), dict(frame_size=8192, hop=2048)),
        ("hop/2048_1024", dict(fn="s_poly", seed=12, fs=22050, n=20000), dict(frame_size=2048, hop=1024)),
    ]
