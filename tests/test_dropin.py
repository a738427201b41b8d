"""The drop-in boundary: iac_b200/libiamf.so exports the reference's public API (include/IAMF_decoder.h) on top of the
CUDA engine.

CPU (`-m "not gpu"`): the library loads, exports every public symbol, setters behave like the reference's, and - with
no GPU - configure fails loudly instead of falling back to a CPU renderer.
GPU (`-m gpu`): driven through the public API exactly like the compiled reference was when tests/golden/*.npz were
made, it returns the same per-call sample counts and byte-identical PCM; the UNMODIFIED stock iamfplayer linked against
it writes the same WAV files as the one linked against the reference; the additive batch call steps many handles in one
launch with the same results.
"""
import ctypes as C
import glob
import importlib.util
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import iamfapi
import iamfgen as G
import refstreams
import scenarios as S

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBIAMF = os.path.join(ROOT, "iac_b200", "libiamf.so")
PLAYER_OURS = os.path.join(ROOT, "oracle", "_ref", "iamfplayer_b200")
PLAYER_REF = os.path.join(ROOT, "oracle", "_ref", "iamfplayer_ref")

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
MG = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MG)

PUBLIC = ["IAMF_decoder_open", "IAMF_decoder_close", "IAMF_decoder_configure", "IAMF_decoder_decode",
          "IAMF_decoder_set_mix_presentation_id", "IAMF_decoder_output_layout_set_sound_system",
          "IAMF_decoder_output_layout_set_binaural", "IAMF_layout_sound_system_channels_count",
          "IAMF_layout_binaural_channels_count", "IAMF_decoder_get_codec_capability",
          "IAMF_decoder_set_normalization_loudness", "IAMF_decoder_set_bit_depth", "IAMF_decoder_peak_limiter_enable",
          "IAMF_decoder_peak_limiter_set_threshold", "IAMF_decoder_peak_limiter_get_threshold",
          "IAMF_decoder_set_sampling_rate", "IAMF_decoder_get_stream_info", "IAMF_decoder_set_pts",
          "IAMF_decoder_get_last_metadata", "IAMF_decoder_decode_batch"]


def test_library_exports_the_public_api():
    L = C.CDLL(LIBIAMF, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    for name in PUBLIC:
        assert hasattr(L, name), name
    # nothing internal leaks out
    out = subprocess.run(["nm", "-D", "--defined-only", LIBIAMF], capture_output=True, text=True).stdout
    assert " ih_" not in out and "iamfb_" not in out


def test_setters_and_tables_match_the_reference_conventions():
    L = C.CDLL(LIBIAMF, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    L.IAMF_decoder_open.restype = C.c_void_p
    L.IAMF_decoder_peak_limiter_get_threshold.restype = C.c_float
    L.IAMF_decoder_get_codec_capability.restype = C.c_void_p
    vp = C.c_void_p
    for f in ("IAMF_decoder_close", "IAMF_decoder_output_layout_set_binaural"):
        getattr(L, f).argtypes = [vp]
    L.IAMF_decoder_set_sampling_rate.argtypes = [vp, C.c_uint32]
    L.IAMF_decoder_output_layout_set_sound_system.argtypes = [vp, C.c_int]
    L.IAMF_decoder_peak_limiter_get_threshold.argtypes = [vp]
    L.IAMF_decoder_decode.argtypes = [vp, C.c_char_p, C.c_int32, C.POINTER(C.c_uint32), vp]
    # IAMF_decoder.c:208-219,3998-4008
    assert [L.IAMF_layout_sound_system_channels_count(i) for i in range(13)] == [2, 6, 8, 10, 11, 12, 14, 24, 8, 12, 10, 6, 1]
    assert L.IAMF_layout_sound_system_channels_count(13) == -1 and L.IAMF_layout_binaural_channels_count() == 2
    h = L.IAMF_decoder_open()
    assert h
    assert L.IAMF_decoder_peak_limiter_get_threshold(h) == -1.0            # LIMITER_MaximumTruePeak
    assert L.IAMF_decoder_set_sampling_rate(h, 44100) == 0
    assert L.IAMF_decoder_set_sampling_rate(h, 12345) == -1                # IAMF_ERR_BAD_ARG
    assert L.IAMF_decoder_output_layout_set_sound_system(h, 99) == -1
    assert L.IAMF_decoder_output_layout_set_sound_system(h, 1) == 0
    buf = C.create_string_buffer(64)
    assert L.IAMF_decoder_decode(h, b"\x00", 1, None, buf) == -5           # IAMF_ERR_INVALID_STATE before configure
    cap = C.string_at(L.IAMF_decoder_get_codec_capability())
    assert b"iamf.001.001.ipcm" in cap
    L.IAMF_decoder_close(h)


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU behaviour")
def test_configure_fails_loudly_without_a_gpu():
    """no CPU rendering path: descriptors parse, then the engine cannot be created -> IAMF_ERR_INTERNAL"""
    L = C.CDLL(LIBIAMF, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    L.IAMF_decoder_open.restype = C.c_void_p
    L.IAMF_decoder_configure.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.IAMF_decoder_output_layout_set_sound_system.argtypes = [C.c_void_p, C.c_int]
    L.IAMF_decoder_set_bit_depth.argtypes = [C.c_void_p, C.c_uint32]
    L.IAMF_decoder_close.argtypes = [C.c_void_p]
    st = G.cfg_stereo()
    x = np.zeros((2, 960), np.int16)
    blob = st.descriptors() + st.temporal_unit([x])
    h = L.IAMF_decoder_open()
    L.IAMF_decoder_output_layout_set_sound_system(h, 0)
    L.IAMF_decoder_set_bit_depth(h, 16)
    r = C.c_uint32(0)
    assert L.IAMF_decoder_configure(h, blob, len(blob), C.byref(r)) == -3    # IAMF_ERR_INTERNAL
    assert r.value == len(st.descriptors())                                  # the descriptors themselves parsed
    L.IAMF_decoder_close(h)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("entry", MG.CASES, ids=[c[0] for c in MG.CASES])
def test_dropin_reproduces_reference_fixture(entry):
    name, case, kw, n_streams, n_frames, seed_in, seed_p = entry
    fx = np.load(os.path.join(HERE, "golden", name + ".npz"))
    sc, st, api_kw, unit_kw, inputs, P, _, _ = MG.build_case(case, kw, n_streams, n_frames, seed_in, seed_p)
    api = iamfapi.Api(LIBIAMF)
    for s in range(n_streams):
        units = refstreams.temporal_units(sc, st, inputs, P, unit_kw, s)
        pcm, counts = api.render(st.descriptors(), units, **api_kw)
        assert counts == [int(c) for c in fx[f"counts{s}"]], f"{name}: stream {s} per-call sample counts"
        assert pcm.tobytes() == fx[f"pcm{s}"].tobytes(), f"{name}: stream {s} PCM differs from the reference decoder"


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(PLAYER_OURS) and os.path.exists(PLAYER_REF)), reason="players not built (make -C oracle ref)")
@pytest.mark.parametrize("case,args", [("c1", ["-o2", "-s0"]), ("c2", ["-o2", "-s1"]), ("c4", ["-o2", "-sb"]),
                                       ("c5", ["-o2", "-s0", "-r", "48000", "-l", "-24"]), ("c3", ["-o2", "-s7", "-d", "24"])])
def test_stock_iamfplayer_writes_identical_wav(case, args):
    """the unmodified player (compiled against the reference's headers) linked against our libiamf.so vs the reference's"""
    sc, st, api_kw, unit_kw = refstreams.case(case)
    F = 12
    inputs = S.synth_inputs(sc, 1, F, seed=77)
    P, _, _ = S.synth_params(sc, 1, F, seed=78)
    refstreams.no_param_gaps(sc, P)
    blob = st.descriptors() + b"".join(refstreams.temporal_units(sc, st, inputs, P, unit_kw, 0))
    outs = []
    for player in (PLAYER_REF, PLAYER_OURS):
        d = tempfile.mkdtemp(prefix="iamfplayer_")
        try:
            with open(os.path.join(d, "in.iamf"), "wb") as f:
                f.write(blob)
            r = subprocess.run([player] + args + ["in.iamf"], cwd=d, capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stderr[-2000:]
            wavs = glob.glob(os.path.join(d, "*.wav"))
            assert len(wavs) == 1, (wavs, r.stdout[-500:])
            outs.append(open(wavs[0], "rb").read())
        finally:
            shutil.rmtree(d, ignore_errors=True)
    assert len(outs[0]) > 44 + 1000
    assert outs[0] == outs[1]


@pytest.mark.gpu
@pytest.mark.skipif(not (G.OpusEncoders.available() and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libiamf_ref.so"))),
                    reason="needs oracle/_ref (libopus_ref.so encoder + compiled reference)")
def test_opus_coded_stereo_matches_reference():
    """configuration 1 as named: simple profile, Opus-coded stereo -> sound system A; entropy decode happens in
    libopus on the host in both libraries, everything after it on the GPU in ours"""
    import refbind
    sc = S.c1_stereo(peak_db=(-3.0, 3.0))
    x = S.synth_inputs(sc, 1, 25, seed=5)[0][0]
    st = G.cfg_stereo(codec="opus")
    units = [st.temporal_unit([refstreams.to_i16(x[f])]) for f in range(x.shape[0])]
    ref_pcm, ref_counts = iamfapi.Api(refbind.REF_SO).render(st.descriptors(), units, sound_system=0)
    pcm, counts = iamfapi.Api(LIBIAMF).render(st.descriptors(), units, sound_system=0)
    assert counts == ref_counts
    assert pcm.tobytes() == ref_pcm.tobytes()


@pytest.mark.gpu
def test_decode_batch_matches_per_handle_decode():
    """IAMF_decoder_decode_batch: 9 handles of configuration 2 step together; same bytes as one handle at a time"""
    sc, st, api_kw, unit_kw = refstreams.case("c2")
    n, F = 9, 7
    inputs = S.synth_inputs(sc, n, F, seed=31)
    P, _, _ = S.synth_params(sc, n, F, seed=32)
    refstreams.no_param_gaps(sc, P)
    desc = st.descriptors()
    units = [refstreams.temporal_units(sc, st, inputs, P, unit_kw, s) for s in range(n)]
    api = iamfapi.Api(LIBIAMF)
    single = [api.render(desc, units[s], **api_kw) for s in range(n)]

    L = api.L
    vp = C.c_void_p
    L.IAMF_decoder_decode_batch.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int32),
                                            C.POINTER(C.c_uint32), C.POINTER(vp), C.POINTER(C.c_int)]
    hs = (vp * n)()
    ch = 6
    for s in range(n):
        h = L.IAMF_decoder_open()
        L.IAMF_decoder_peak_limiter_set_threshold(h, -1.0)
        L.IAMF_decoder_set_bit_depth(h, 16)
        L.IAMF_decoder_output_layout_set_sound_system(h, 1)
        blob = desc + units[s][0]
        r = C.c_uint32(0)
        assert L.IAMF_decoder_configure(h, blob, len(blob), C.byref(r)) == 0
        hs[s] = h
    bufs = [C.create_string_buffer(2 * 6144 * ch) for _ in range(n)]
    pcm = (vp * n)(*[C.cast(b, vp) for b in bufs])
    got = [[] for _ in range(n)]
    counts = [[] for _ in range(n)]
    for f in range(F + 1):
        data = (C.c_char_p * n)(*[(units[s][f] if f < F else None) for s in range(n)])
        size = (C.c_int32 * n)(*[(len(units[s][f]) if f < F else 0) for s in range(n)])
        rs = (C.c_uint32 * n)()
        ret = (C.c_int * n)()
        assert L.IAMF_decoder_decode_batch(hs, n, data, size, rs, pcm, ret) == 0
        for s in range(n):
            counts[s].append(ret[s])
            if ret[s] > 0:
                got[s].append(bufs[s].raw[: ret[s] * ch * 2])
    for s in range(n):
        assert counts[s] == single[s][1]
        assert b"".join(got[s]) == single[s][0].tobytes()
        L.IAMF_decoder_close(hs[s])


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,n,F,K", [("c2", 21, 9, 4), ("c1", 40, 12, 5), ("c5", 10, 7, 3), ("c4", 9, 6, 6)])
def test_decode_batch_units_matches_per_handle_decode(cfg, n, F, K):
    """IAMF_decoder_decode_batch_units: K temporal units per handle and call, int16 ingest (16-bit ipcm), the host part on
    the thread pool; buffers that end inside a temporal unit or hold fewer units than K; same bytes as one handle, one
    unit at a time"""
    sc, st, api_kw, unit_kw = refstreams.case(cfg)
    inputs = S.synth_inputs(sc, n, F, seed=41)
    P, _, _ = S.synth_params(sc, n, F, seed=42)
    refstreams.no_param_gaps(sc, P)
    desc = st.descriptors()
    units = [refstreams.temporal_units(sc, st, inputs, P, unit_kw, s) for s in range(n)]
    api = iamfapi.Api(LIBIAMF)
    single = [api.render(desc, units[s], **api_kw) for s in range(n)]
    L = api.L
    vp = C.c_void_p
    L.IAMF_decoder_decode_batch_units.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int32),
                                                  C.POINTER(C.c_uint32), C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
    hs = (vp * n)()
    ch = sc.out_channels
    for s in range(n):
        hs[s] = api.open_configured(desc + units[s][0], **api_kw)
    bufs = [C.create_string_buffer(2 * 6144 * ch * K) for _ in range(n)]
    pcm = (vp * n)(*[C.cast(b, vp) for b in bufs])
    # every handle's stream as one byte string; handle s is fed in pieces that do not line up with its temporal units
    blob = [b"".join(units[s]) for s in range(n)]
    pos = [0] * n
    got = [b"" for _ in range(n)]
    total = [0] * n
    rng = np.random.default_rng(7)
    for _ in range(10 * F):
        if all(pos[s] >= len(blob[s]) for s in range(n)):
            break
        piece = []
        for s in range(n):
            left = len(blob[s]) - pos[s]
            take = left if s % 3 == 0 else min(left, int(rng.integers(1, 3 * len(units[s][0]))))
            piece.append(blob[s][pos[s]: pos[s] + take])
        data = (C.c_char_p * n)(*piece)
        size = (C.c_int32 * n)(*[len(p) for p in piece])
        rs = (C.c_uint32 * n)()
        ret = (C.c_int * n)()
        done = (C.c_int * n)()
        assert L.IAMF_decoder_decode_batch_units(hs, n, data, size, rs, pcm, ret, K, done) == 0
        for s in range(n):
            assert ret[s] >= 0 and done[s] <= K
            pos[s] += rs[s]
            total[s] += ret[s]
            got[s] += bufs[s].raw[: ret[s] * ch * 2]
    # flush
    data = (C.c_char_p * n)(*[None] * n)
    size = (C.c_int32 * n)()
    ret = (C.c_int * n)()
    assert L.IAMF_decoder_decode_batch_units(hs, n, data, size, None, pcm, ret, K, None) == 0
    for s in range(n):
        got[s] += bufs[s].raw[: ret[s] * ch * 2]
        assert pos[s] == len(blob[s])
        assert total[s] + ret[s] == sum(c for c in single[s][1] if c > 0)
        assert got[s] == single[s][0].tobytes(), f"handle {s}"
        L.IAMF_decoder_close(hs[s])


@pytest.mark.gpu
def test_binauraliser_switch_through_the_public_api(monkeypatch):
    """IAMF_B200_BINAURALIZER=1 (the run-time counterpart of the reference's DISABLE_BINAURALIZER 0): configuration 4 through
    IAMF_decoder_output_layout_set_binaural renders the 7.1.4 element (headphones_rendering_mode 1) with the per-speaker
    HRIRs and the ambisonics element with the SH-domain ones.  Checked against the pipeline oracle with its HRTF renderer
    (self-oracle: the reference's binauraliser libraries are absent) - and the switch off still gives the
    as-built stereo-matrix bytes."""
    sc, st, api_kw, unit_kw = refstreams.case("c4h")
    st = G.cfg_714_foa(headphones_mode=1)
    n, F = 3, 6
    inputs = S.synth_inputs(sc, n, F, seed=51)
    P, _, _ = S.synth_params(sc, n, F, seed=52)
    desc = st.descriptors()
    api = iamfapi.Api(LIBIAMF)
    ref = S.run_oracle(sc, inputs, P)
    asbuilt_sc, _, _, _ = refstreams.case("c4")
    ref_asbuilt = S.run_oracle(asbuilt_sc, inputs, P)
    for s in range(n):
        units = refstreams.temporal_units(sc, st, inputs, P, unit_kw, s)
        monkeypatch.setenv("IAMF_B200_BINAURALIZER", "1")
        pcm, counts = api.render(desc, units, **api_kw)
        assert counts == ref[s][0]
        assert pcm.tobytes() == ref[s][1].tobytes(), f"stream {s}: HRTF rendering through the public API"
        monkeypatch.delenv("IAMF_B200_BINAURALIZER")
        pcm, counts = api.render(desc, units, **api_kw)
        assert pcm.tobytes() == ref_asbuilt[s][1].tobytes(), f"stream {s}: as built"


def _animated_units(st, sc, inputs, P, n_frames, seed, trim_first=0):
    """temporal units whose element / output mix gains are animated: step, linear and Bezier parameter blocks"""
    rng = np.random.default_rng(seed)
    units = []
    for f in range(n_frames):
        mg = {}
        for pid in [e.mixgain_pid() for e in st.elements] + [st.OUT_GAIN_PID]:
            kind = int(rng.integers(0, 4))
            q = [int(v) for v in rng.integers(-0x0600, 0x0200, 3)]
            if kind == 0:
                continue                                     # no block this frame: the default gain applies
            mg[pid] = ("step", q[0]) if kind == 1 else (("linear", q[0], q[1]) if kind == 2 else
                                                         ("bezier", q[0], q[1], q[2], int(rng.integers(1, 255))))
        pcm = [refstreams.to_i16(inputs[e][0, f]) for e in range(len(sc.elements))]
        units.append(st.temporal_unit(pcm, mix_gain=mg, trim_start=trim_first if f == 0 else 0))
    return units


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,trim", [("c1", 0), ("c1", 312), ("c4", 0), ("c5", 100)])
def test_animated_mix_gains_match_the_reference(cfg, trim):
    """element and output mix gains animated per frame (step / linear / quadratic Bezier parameter blocks,
    IAMF_decoder.c:639-664,857-982), a start-trimmed first frame (the gain time line is read at the pts AFTER trimming,
    :1379): the drop-in library against the compiled reference, byte for byte.  The ramps are evaluated ON THE DEVICE from
    the segment descriptions (k_gain_expand)."""
    import refbind
    if not refbind.have_ref():
        pytest.skip("compiled reference not available")
    sc, st, api_kw, unit_kw = refstreams.case(cfg)
    F = 9
    inputs = S.synth_inputs(sc, 1, F, seed=61)
    P, _, _ = S.synth_params(sc, 1, F, seed=62)
    units = _animated_units(st, sc, inputs, P, F, seed=63 + trim, trim_first=trim)
    desc = st.descriptors()
    ref_pcm, ref_counts = iamfapi.Api(refbind.REF_SO).render(desc, units, **api_kw)
    pcm, counts = iamfapi.Api(LIBIAMF).render(desc, units, **api_kw)
    assert counts == ref_counts
    assert pcm.tobytes() == ref_pcm.tobytes()


@pytest.mark.gpu
def test_animated_mix_gains_in_a_batch_step():
    """the same through IAMF_decoder_decode_batch_units: handles with and without animated gains in one group (frames
    without segments take their constant on the device), several temporal units per call"""
    import refbind
    if not refbind.have_ref():
        pytest.skip("compiled reference not available")
    sc, st, api_kw, unit_kw = refstreams.case("c1")
    n, F, K = 7, 8, 3
    inputs = S.synth_inputs(sc, n, F, seed=71)
    P, _, _ = S.synth_params(sc, n, F, seed=72)
    desc = st.descriptors()
    units = []
    for s in range(n):
        one = [x[s:s + 1] for x in inputs]
        if s % 2:
            units.append(_animated_units(st, sc, one, P, F, seed=80 + s))
        else:
            units.append(refstreams.temporal_units(sc, st, one, P[s:s + 1], unit_kw, 0))
    ref = [iamfapi.Api(refbind.REF_SO).render(desc, units[s], **api_kw) for s in range(n)]
    api = iamfapi.Api(LIBIAMF)
    L = api.L
    vp = C.c_void_p
    L.IAMF_decoder_decode_batch_units.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int32),
                                                  C.POINTER(C.c_uint32), C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
    hs = (vp * n)(*[api.open_configured(desc + units[s][0], **api_kw) for s in range(n)])
    ch = sc.out_channels
    bufs = [C.create_string_buffer(2 * 6144 * ch * K) for _ in range(n)]
    pcm = (vp * n)(*[C.cast(b, vp) for b in bufs])
    got = [b"" for _ in range(n)]
    for f0 in list(range(0, F, K)) + [None]:
        if f0 is None:
            piece = [None] * n
        else:
            piece = [b"".join(units[s][f0:f0 + K]) for s in range(n)]
        data = (C.c_char_p * n)(*piece)
        size = (C.c_int32 * n)(*[len(p) if p else 0 for p in piece])
        rs = (C.c_uint32 * n)()
        ret = (C.c_int * n)()
        assert L.IAMF_decoder_decode_batch_units(hs, n, data, size, rs, pcm, ret, K, None) == 0
        for s in range(n):
            assert ret[s] >= 0
            got[s] += bufs[s].raw[: ret[s] * ch * 2]
    for s in range(n):
        assert got[s] == ref[s][0].tobytes(), f"handle {s}"
        L.IAMF_decoder_close(hs[s])


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,bits", [("c1", 16), ("c2", 16), ("c4", 16), ("c1", 24)])
def test_flac_coded_streams_match_reference(cfg, bits):
    """FLAC core decode (libFLAC through its public API, one stream decoder per sub-stream, the channel count of the
    config's STREAMINFO patched for mono sub-streams - flac/flac_multistream_decoder.c) in front of the GPU path: FLAC-coded
    streams against the compiled reference, byte for byte; 16-bit FLAC also through the batch call (int16 hand-over)"""
    import dataclasses
    import refbind
    if not (refbind.have_ref() and G.FlacEncoder.available()):
        pytest.skip("compiled reference / libFLAC encoder not available")
    sc, st, api_kw, unit_kw = refstreams.case(cfg)
    st = dataclasses.replace(st, codec="flac", flac_bits=bits)
    F, n = 6, 3
    inputs = S.synth_inputs(sc, n, F, seed=81)
    P, _, _ = S.synth_params(sc, n, F, seed=82)
    refstreams.no_param_gaps(sc, P)
    desc = st.descriptors()

    def units_of(s):
        out = []
        for f in range(F):
            kw = dict(unit_kw(f, P, s) if unit_kw else {})
            pcm = [refstreams.to_i16(inputs[e][s, f]).astype(np.int32) * (1 << (bits - 16)) + (s if bits > 16 else 0)
                   for e in range(len(sc.elements))]
            out.append(st.temporal_unit(pcm, **kw))
        return out
    units = [units_of(s) for s in range(n)]
    ref = [iamfapi.Api(refbind.REF_SO).render(desc, units[s], **api_kw) for s in range(n)]
    api = iamfapi.Api(LIBIAMF)
    for s in range(n):
        pcm, counts = api.render(desc, units[s], **api_kw)
        assert counts == ref[s][1]
        assert pcm.tobytes() == ref[s][0].tobytes(), f"stream {s}"
    # the batch call (16-bit FLAC travels to the device as int16)
    L = api.L
    vp = C.c_void_p
    L.IAMF_decoder_decode_batch_units.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int32),
                                                  C.POINTER(C.c_uint32), C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
    hs = (vp * n)(*[api.open_configured(desc + units[s][0], **api_kw) for s in range(n)])
    ch, K = sc.out_channels, 4
    bufs = [C.create_string_buffer(2 * 6144 * ch * K) for _ in range(n)]
    pcm = (vp * n)(*[C.cast(b, vp) for b in bufs])
    got = [b"" for _ in range(n)]
    for f0 in list(range(0, F, K)) + [None]:
        piece = [None] * n if f0 is None else [b"".join(units[s][f0:f0 + K]) for s in range(n)]
        data = (C.c_char_p * n)(*piece)
        size = (C.c_int32 * n)(*[len(p) if p else 0 for p in piece])
        ret = (C.c_int * n)()
        assert L.IAMF_decoder_decode_batch_units(hs, n, data, size, None, pcm, ret, K, None) == 0
        for s in range(n):
            got[s] += bufs[s].raw[: ret[s] * ch * 2]
    for s in range(n):
        assert got[s] == ref[s][0].tobytes(), f"handle {s} (batch)"
        L.IAMF_decoder_close(hs[s])


@pytest.mark.gpu
def test_binauraliser_renders_a_scalable_stream_through_the_public_api(monkeypatch):
    """configuration 2's stream (2.0 -> 7.1.4 scalable layers, demixing + recon-gain parameter blocks every frame) on
    headphones with headphones_rendering_mode 1 and IAMF_B200_BINAURALIZER=1: the 7.1.4 layer is de-mixed in front of the
    HRTF renderer.  Against the pipeline oracle (pinned de-mixer + HRTF self-oracle)."""
    sc, st, api_kw, unit_kw = refstreams.case("c2")
    sc.target = S.TGT_BIN
    sc.elements[0].hrtf = True
    st.elements[0].headphones_mode = 1
    api_kw = dict(api_kw, binaural=True)
    api_kw.pop("sound_system", None)
    n, F = 2, 7
    inputs = S.synth_inputs(sc, n, F, seed=91)
    P, _, _ = S.synth_params(sc, n, F, seed=92)
    refstreams.no_param_gaps(sc, P)
    desc = st.descriptors()
    ref = S.run_oracle(sc, inputs, P)
    monkeypatch.setenv("IAMF_B200_BINAURALIZER", "1")
    api = iamfapi.Api(LIBIAMF)
    for s in range(n):
        units = refstreams.temporal_units(sc, st, inputs, P, unit_kw, s)
        pcm, counts = api.render(desc, units, **api_kw)
        assert counts == ref[s][0]
        assert pcm.tobytes() == ref[s][1].tobytes(), f"stream {s}"


def _batch_units_run(n, F, K, env):
    """n handles of configuration 2 through IAMF_decoder_decode_batch_units in a child process with `env` set; returns the
    concatenated PCM of every handle"""
    import json
    code = f"""
import ctypes as C, os, sys, json, base64
sys.path.insert(0, {os.path.join(ROOT, 'tests')!r}); sys.path.insert(0, {ROOT!r})
import numpy as np, scenarios as S, refstreams, iamfapi
sc, st, api_kw, unit_kw = refstreams.case("c2")
n, F, K = {n}, {F}, {K}
inputs = S.synth_inputs(sc, n, F, seed=41); P, _, _ = S.synth_params(sc, n, F, seed=42); refstreams.no_param_gaps(sc, P)
desc = st.descriptors()
units = [refstreams.temporal_units(sc, st, inputs, P, unit_kw, s) for s in range(n)]
api = iamfapi.Api({LIBIAMF!r}); L = api.L; vp = C.c_void_p
L.IAMF_decoder_open.restype = vp
L.IAMF_decoder_decode_batch_units.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
hs = (vp * n)()
for s in range(n):
    h = L.IAMF_decoder_open(); L.IAMF_decoder_peak_limiter_set_threshold(vp(h), C.c_float(-1.0)); L.IAMF_decoder_set_bit_depth(vp(h), 16)
    L.IAMF_decoder_output_layout_set_sound_system(vp(h), 1)
    blob = desc + units[s][0]; r = C.c_uint32(0)
    assert L.IAMF_decoder_configure(vp(h), blob, len(blob), C.byref(r)) == 0 and r.value == len(desc)
    hs[s] = h
bufs = [C.create_string_buffer(K * 960 * 6 * 2) for _ in range(n)]
pcm = (vp * n)(*[C.addressof(b) for b in bufs])
out = [b"" for _ in range(n)]
for f0 in list(range(0, F, K)) + [None]:
    blobs = [None if f0 is None else b"".join(units[s][f0:f0 + K]) for s in range(n)]
    data = (C.c_char_p * n)(*blobs); size = (C.c_int32 * n)(*[0 if b is None else len(b) for b in blobs])
    rs = (C.c_uint32 * n)(); ret = (C.c_int * n)(); ud = (C.c_int * n)()
    rc = L.IAMF_decoder_decode_batch_units(hs, n, data, size, rs, pcm, ret, K, ud)
    assert rc == 0, rc
    for s in range(n):
        if ret[s] > 0: out[s] += bufs[s].raw[: ret[s] * 12]
print("RESULT " + json.dumps([base64.b64encode(o).decode() for o in out]))
"""
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0, r.stderr[-3000:]
    import base64
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return [base64.b64decode(x) for x in json.loads(line[7:])]


@pytest.mark.gpu
def test_batch_call_split_over_a_device_list_matches_one_device():
    """IAMF_B200_DEVICES: the handles of a batch call are dealt round the device list (handle i -> entry i mod D), one host
    thread, context and group batch per entry; same bytes as on one device.  With a single GPU the list names it twice -
    the split, the threads and the scatter of the results are exercised all the same; with two or more GPUs visible the
    list is the first two devices."""
    import torch
    one = _batch_units_run(11, 9, 4, {"IAMF_B200_DEVICES": "0"})
    lst = "0,1" if torch.cuda.device_count() >= 2 else "0,0"
    two = _batch_units_run(11, 9, 4, {"IAMF_B200_DEVICES": lst})
    three = _batch_units_run(11, 9, 4, {"IAMF_B200_DEVICES": lst + ",0"})
    assert all(len(x) > 1000 for x in one)
    assert one == two and one == three
