"""GPU parity: the CUDA path behind the C ABI must be BIT-IDENTICAL to the oracle (which is itself pinned bit-exact to
the compiled reference) - integer PCM at 16/24/32 bit and the float debug output alike - on every BASELINE.json
configuration and on the edge cases (ragged submits, trims, ramps, both resampler paths, two elements, DMR, ...)."""
import numpy as np
import pytest

import scenarios as S

pytestmark = pytest.mark.gpu

ALL = [S.c1_stereo(), S.c2_714_to_B(), S.c3_toa_to_H(), S.c4_714_foa_binaural(), S.c5_resample()] + S.edge_cases()


def compare(sc, n_streams, F, splits, seed=0, s16=False, expect_path=None, edit_params=None):
    from gpu_harness import run_product
    inputs = S.synth_inputs(sc, n_streams, F, seed=0x1A3F + seed)
    P, ramps, oramp = S.synth_params(sc, n_streams, F, seed=0x77 + seed)
    if edit_params:
        edit_params(P)
    got, launches = run_product(sc, inputs, P, ramps, oramp, splits=splits, s16=s16, expect_path=expect_path)
    ref = S.run_oracle(sc, inputs, P, ramps, oramp)
    assert launches > 0
    for s in range(n_streams):
        assert got[s][0] == ref[s][0], f"{sc.name}: stream {s} per-call sample counts"
        a, b = got[s][1], ref[s][1]
        assert a.shape == b.shape, f"{sc.name}: stream {s} byte count {a.shape} vs {b.shape}"
        if not np.array_equal(a, b):
            bps = sc.bit_depth // 8 if sc.bit_depth else 4
            bad = np.nonzero(a != b)[0]
            first = bad[0] // (bps * sc.out_channels)
            raise AssertionError(f"{sc.name}: stream {s} differs in {len(bad)} bytes, first at output sample {first}")


@pytest.mark.parametrize("sc", ALL, ids=[s.name for s in ALL])
def test_bit_exact_single_submit(sc):
    compare(sc, 5, 6, [6])


@pytest.mark.parametrize("sc", ALL, ids=[s.name for s in ALL])
def test_bit_exact_ragged_submits(sc):
    compare(sc, 37, 12, [1, 4, 2, 5], seed=5)


def test_limiter_heavy_long():
    # every stream well above the threshold for 60 frames: exercises the serial gain recurrence across many tiles
    compare(S.c1_stereo(peak_db=(0.0, 6.0)), 40, 60, [20, 20, 20], seed=9)
    compare(S.c2_714_to_B(peak_db=(-1.5, 4.0)), 33, 30, [30], seed=10)


def test_fast_path_quiet_streams():
    compare(S.c1_stereo(peak_db=(-30.0, -20.0)), 64, 10, [10], seed=11)


@pytest.mark.parametrize("sc", [S.c2_714_to_B(), S.c4_714_foa_binaural(), S.c5_resample()], ids=["c2", "c4", "c5"])
def test_int16_upload_and_stream_groups(sc):
    # int16 input scaled on the device (IAMFB_IN_S16); 200 streams so that the host path pipelines groups of streams
    compare(sc, 200, 5, [2, 3], seed=21, s16=True)


def test_multi_kernel_path_still_bit_exact(monkeypatch):
    # the fused kernel is the default for these signatures; the multi-kernel path must stay correct behind IAMFB_FUSED=0
    monkeypatch.setenv("IAMFB_FUSED", "0")
    compare(S.c2_714_to_B(), 9, 6, [4, 2], seed=3)
    compare(S.c3_toa_to_H(), 5, 4, [4], seed=4)
    compare(S.c4_714_foa_binaural(), 7, 6, [1, 5], seed=5)


def test_sequential_fused_kernel_still_bit_exact(monkeypatch):
    # k_stream (register-resident, pipelined) is the default for the signatures it is instantiated for; k_fused must
    # stay correct for them behind IAMFB_STREAM=0 (it also renders their trimmed / flushed streams)
    monkeypatch.setenv("IAMFB_STREAM", "0")
    monkeypatch.setenv("IAMFB_PIPE", "0")
    compare(S.c2_714_to_B(), 9, 6, [4, 2], seed=3, expect_path=1)   # IAMFB_PATH_FUSED
    compare(S.c1_stereo(peak_db=(-3.0, 3.0)), 5, 6, [6], seed=4, expect_path=1)


def test_pipe_kernel_serves_the_named_configurations(monkeypatch):
    # k_pipe (double-buffered int16 / float32 staging) serves configs 1-4: channel-based, scene-based and two-element mixes.
    # Where k_stream exists too (16-bit channel-based single-element plans: configs 1-2) it keeps the float32 submits
    for sc in (S.c1_stereo(), S.c2_714_to_B(), S.c3_toa_to_H(), S.c4_714_foa_binaural()):
        both = len(sc.elements) == 1 and sc.elements[0].kind == "channel"
        compare(sc, 13, 8, [3, 5], seed=61, expect_path=2 if both else 3)
        compare(sc, 13, 8, [3, 5], seed=62, s16=True, expect_path=3)   # int16 staging, two stages
    monkeypatch.setenv("IAMFB_STREAM", "0")                            # k_pipe with its one float32 stage for configs 1-2 as well
    for sc in (S.c1_stereo(), S.c2_714_to_B()):
        compare(sc, 13, 8, [3, 5], seed=61, expect_path=3)
    monkeypatch.delenv("IAMFB_STREAM")
    compare(S.edge_cases()[9], 9, 6, [2, 4], seed=63, s16=True, expect_path=3)   # FOA + 7.1.4 -> sound system H


def test_pipe_kernel_other_output_depths():
    import dataclasses
    for bd in (24, 32, 0):
        compare(dataclasses.replace(S.c2_714_to_B(peak_db=(-3.0, 3.0)), bit_depth=bd, name=f"c2_{bd}bit"), 7, 6, [2, 4], seed=64, expect_path=3)
        compare(dataclasses.replace(S.c3_toa_to_H(), bit_depth=bd, name=f"c3_{bd}bit"), 5, 4, [4], seed=65, s16=True, expect_path=3)
    compare(dataclasses.replace(S.c4_714_foa_binaural(), limiter=False, name="c4_nolimiter"), 7, 6, [2, 4], seed=66, expect_path=3)
    compare(dataclasses.replace(S.c2_714_to_B(), limiter=False, bit_depth=0, name="c2_nolimiter_float"), 7, 6, [6], seed=67, s16=True, expect_path=3)


def test_pipe_rs_kernel_resampling_pipelines():
    # k_pipe_rs: render -> pre-resample ring -> FIR -> limiter on chip, for the regular streams of a resampling plan
    import dataclasses
    c5 = S.c5_resample()
    compare(c5, 9, 7, [7], seed=71, expect_path=3)
    compare(c5, 21, 9, [1, 3, 5], seed=72, s16=True, expect_path=3)
    compare(S.c5_resample(peak_db=(-2.0, 4.0)), 6, 12, [5, 7], seed=73, expect_path=3)                      # limiter busy
    down = S.Scenario("stereo_48k_to_44k1", [S.El("channel", S.LY_STEREO, [S.L2, S.R2])], S.TGT_A, in_rate=48000, out_rate=44100,
                      loudness_gain=0.7, peak_db=(-3.0, 3.0))
    compare(down, 7, 9, [2, 3, 4], seed=74, expect_path=3)
    compare(dataclasses.replace(down, bit_depth=24, limiter=False, name="stereo_48k_to_44k1_24bit_nolimiter"), 5, 6, [6], seed=75, s16=True, expect_path=3)
    c2rs = dataclasses.replace(S.c2_714_to_B(peak_db=(-3.0, 3.0)), in_rate=48000, out_rate=44100, name="c2_714_to_B_44k1")
    compare(c2rs, 5, 6, [2, 4], seed=76, expect_path=3)
    c2up = dataclasses.replace(S.c2_714_to_B(), frame_size=1024, in_rate=44100, out_rate=48000, name="c2_714_to_B_44k1_to_48k")
    compare(c2up, 5, 5, [1, 4], seed=77, s16=True, expect_path=3)


def test_pipe_rs_mixed_with_trimmed_streams():
    # trims on SOME streams: those take the multi-kernel path inside the same submit, the others k_pipe_rs; and a stream
    # changes path between submits (both keep their histories in the heads of the batch's time lines)
    def trims_on_every_third_stream(P):
        P["trim_start"][1::3, 2] = 96
        P["trim_end"][2::3, 5] = 300
    compare(S.c5_resample(peak_db=(-3.0, 3.0)), 20, 9, [3, 3, 3], seed=78, expect_path=3, edit_params=trims_on_every_third_stream)
    compare(S.c5_resample(peak_db=(-3.0, 3.0)), 20, 9, [3, 3, 3], seed=79, s16=True, expect_path=3, edit_params=trims_on_every_third_stream)


@pytest.mark.parametrize("mode", ["0", "1"], ids=["one_kernel", "split"])
def test_resampling_pipeline_forms_bit_exact(monkeypatch, mode):
    # the two forms of a resampling pipeline: k_pipe_rs alone (IAMFB_RS_SPLIT=0, test hook) and the default split form -
    # k_pipe_prerender, k_resample_ls (one stream per lane), k_pipe_rs<PRE> (the limiter half).  Same PCM.
    import dataclasses
    monkeypatch.setenv("IAMFB_RS_SPLIT", mode)
    test_pipe_rs_kernel_resampling_pipelines()
    test_pipe_rs_mixed_with_trimmed_streams()
    # several groups of 32 streams, the last one ragged; small work items; streams at different resampler phases inside a
    # group (a start trim in the first submit shifts every third stream)
    monkeypatch.setenv("IAMFB_LS_CHUNK", "16")
    def shift_every_third(P):
        P["trim_start"][1::3, 0] = 100
        P["trim_start"][2::5, 1] = 37
    compare(S.c5_resample(peak_db=(-3.0, 3.0)), 77, 8, [2, 3, 3], seed=81, expect_path=3, edit_params=shift_every_third)
    c2up = dataclasses.replace(S.c2_714_to_B(), frame_size=1024, in_rate=44100, out_rate=48000, name="c2_714_to_B_44k1_to_48k")
    compare(c2up, 40, 4, [1, 3], seed=82, expect_path=3)


def test_stream_kernel_still_bit_exact(monkeypatch):
    # k_stream (single float32 stage) stays selectable behind IAMFB_PIPE=0 until k_pipe has replaced it everywhere
    monkeypatch.setenv("IAMFB_PIPE", "0")
    compare(S.c2_714_to_B(), 9, 6, [4, 2], seed=3, expect_path=2)
    compare(S.c2_714_to_B(), 9, 6, [4, 2], seed=3, s16=True, expect_path=2)


def test_stream_kernel_mixed_with_trimmed_submits():
    # first submit untrimmed (k_stream), later submits carry trimmed frames (k_fused takes those streams): the limiter
    # history and state must hand over between the two kernels
    compare(S.c1_stereo(trims={5: (0, 100), 9: (200, 0)}, peak_db=(-2.0, 3.0)), 12, 12, [4, 4, 4], seed=31, expect_path=2)
    compare(S.c2_714_to_B(trims={7: (0, 480)}, peak_db=(-3.0, 3.0)), 10, 10, [3, 3, 4], seed=32, expect_path=2)
    compare(S.c2_714_to_B(trims={7: (0, 480)}, peak_db=(-3.0, 3.0)), 10, 10, [3, 3, 4], seed=32, s16=True, expect_path=3)   # k_pipe <-> k_fused


def test_stream_and_fused_kernels_side_by_side_in_one_submit():
    # trims on SOME streams only: within one submit those go to k_fused while the others stay on k_stream, and the two
    # launches run beside each other (programmatic dependent launch) - disjoint streams, shared buffers
    def trims_on_every_third_stream(P):
        P["trim_start"][1::3, 2] = 96
        P["trim_end"][2::3, 5] = 300
        P["trim_start"][2::3, 7] = 959
    compare(S.c2_714_to_B(peak_db=(-3.0, 3.0)), 23, 9, [3, 3, 3], seed=51, expect_path=2, edit_params=trims_on_every_third_stream)
    compare(S.c1_stereo(peak_db=(-2.0, 3.0)), 40, 8, [4, 4], seed=52, expect_path=2, edit_params=trims_on_every_third_stream)
    compare(S.c1_stereo(peak_db=(-2.0, 3.0)), 40, 8, [4, 4], seed=52, s16=True, expect_path=3, edit_params=trims_on_every_third_stream)


def test_stream_kernel_clipping_quantiser():
    # limiter threshold above full scale: samples beyond +-1.0 reach the quantiser and must saturate like
    # FLOAT2INT16 (IAMF_decoder.c:100-103) does
    compare(S.c1_stereo(threshold_db=6.0, peak_db=(0.0, 8.0)), 9, 6, [6], seed=41, expect_path=2)
    compare(S.c2_714_to_B(threshold_db=9.0, peak_db=(-3.0, 3.0)), 7, 5, [2, 3], seed=42, expect_path=2)
    compare(S.c2_714_to_B(threshold_db=9.0, peak_db=(-3.0, 3.0)), 7, 5, [2, 3], seed=42, s16=True, expect_path=3)


STREAM_CASES = S.stream_kernel_cases()


@pytest.mark.parametrize("sc", STREAM_CASES, ids=[s.name for s in STREAM_CASES])
def test_stream_kernel_signatures(sc):
    # the other (layout, target) pairs k_stream is instantiated for, layered (de-mixing + recon gain) where the layout allows
    compare(sc, 11, 9, [4, 5], seed=51, expect_path=2)             # IAMFB_PATH_STREAM: float32 submits
    compare(sc, 11, 9, [4, 5], seed=51, s16=True, expect_path=3)   # IAMFB_PATH_PIPE: int16 submits


@pytest.mark.parametrize("thr_db", [-1.0, 0.0, -6.0, -0.1, -20.0, 3.0])
def test_scan_quotient_exhaustive(thr_db):
    """k_stream's limiter scan divides thr by the look-ahead peak with a branch-free sequence (so that the division is
    scheduled into the gain chain); it must equal the IEEE division for every float of the range it serves."""
    import ctypes as C
    from iac_b200 import binding
    L = binding.lib()
    ctx = C.c_void_p()
    assert L.iamfb_ctx_create(0, C.byref(ctx)) == 0
    try:
        thr = np.float32(10.0 ** (thr_db / 20.0))       # audio_effect_peak_limiter.c:73-92 (double pow, then float)
        bad = C.c_uint64(12345)
        assert L.iamfb_selftest_quotient(ctx, C.c_float(float(thr)), C.byref(bad)) == 0
        assert bad.value == 0, f"{bad.value} quotients differ from thr / w at thr = {thr}"
    finally:
        L.iamfb_ctx_destroy(ctx)


@pytest.mark.parametrize("s16", [False, True], ids=["f32", "s16"])
@pytest.mark.parametrize("mk", [S.c2_714_to_B, S.c4_714_foa_binaural, S.c5_resample], ids=["c2", "c4", "c5"])
def test_streams_without_frames_in_some_submits(mk, s16):
    # handles of a group step through IAMF_decoder_decode_batch_units at their own pace: a stream may have no frame in the
    # tail of a submit, or in a whole submit (also the last one before the flush) - its state must stay untouched
    from gpu_harness import run_product
    sc = mk(peak_db=(-3.0, 3.0))
    n, F = 12, 9
    inputs = S.synth_inputs(sc, n, F, seed=0x1A3F + 91)
    P, ramps, oramp = S.synth_params(sc, n, F, seed=0x77 + 91)
    P["trim_start"][1::3, 6:] = 0xFFFF          # nothing in the last submit
    P["trim_start"][2::3, 3:6] = 0xFFFF         # nothing in the middle submit
    P["trim_start"][0::4, 2] = 0xFFFF           # the tail of the first submit
    P["trim_start"][0::4, 8] = 0xFFFF           # ... and of the last
    got, _ = run_product(sc, inputs, P, ramps, oramp, splits=[3, 3, 3], s16=s16)
    for s in range(n):
        keep = [f for f in range(F) if P["trim_start"][s, f] != 0xFFFF]
        ref = S.run_oracle(sc, [x[s:s + 1][:, keep] for x in inputs], P[s:s + 1][:, keep])[0]
        cnt = [c for f, c in enumerate(got[s][0][:F]) if f in keep] + got[s][0][F:]
        assert cnt == ref[0], f"stream {s}: counts {got[s][0]} vs {ref[0]}"
        assert np.array_equal(got[s][1], ref[1]), f"stream {s}: PCM differs"


def test_parametric_downmixer_on_the_fused_kernel(monkeypatch):
    # DMRenderer plans (downmix_renderer.c: element with demixing info toward a layout with fewer surrounds / tops) run in
    # k_fused's down-mixer variant - one kernel per submit instead of the multi-kernel path, which stays correct behind
    # IAMFB_FUSED=0
    dmr = [sc for sc in S.edge_cases() if "dmr" in sc.name]
    assert len(dmr) == 2
    more = S.Scenario("714_scalable_dmr_to_514", [S.El("channel", S.LY_714, [S.L2, S.R2, S.L5, S.R5, S.SL7, S.SR7, S.HFL, S.HFR, S.HBL, S.HBR, S.CC, S.LFE],
                                                       out_gain=[(S.L2, 1.1), (S.R2, 1.1)], demix=(1, 0), first_layer_layout=S.LY_STEREO,
                                                       selected_layer=1, recon_flags=0x780, dmr_out_layout=S.LY_514)], S.TGT_D,
                      peak_db=(-6.0, 3.0), trims={0: (100, 0), 7: (0, 333)})
    for sc in dmr + [more]:
        compare(sc, 9, 8, [3, 5], seed=101, expect_path=1)
        compare(sc, 9, 8, [3, 5], seed=101, s16=True, expect_path=1)
    monkeypatch.setenv("IAMFB_FUSED", "0")
    compare(more, 5, 6, [6], seed=102, expect_path=0)


def test_animated_gains_inside_the_pipelined_kernel():
    # per-sample element / output mix gains (animated mix gain) are applied by k_pipe itself; only streams with trimmed /
    # missing frames of such a submit are left to k_fused
    import dataclasses
    from gpu_harness import run_product
    for sc in (S.c1_stereo(ramp=True, out_gain=1.2, peak_db=(-3.0, 3.0)), S.c2_714_to_B(ramp=True),
               S.c4_714_foa_binaural(ramp=True), dataclasses.replace(S.c3_toa_to_H(ramp=True), bit_depth=24)):
        sc.name += "_ramps"
        n, F = 9, 8
        inputs = S.synth_inputs(sc, n, F, seed=0x1A3F + 111)
        P, ramps, oramp = S.synth_params(sc, n, F, seed=0x77 + 111)
        P["trim_start"][2::4, 3] = 200                      # some streams irregular in the second submit
        for s16 in (False, True):
            ran = {}
            got, _ = run_product(sc, inputs, P, ramps, oramp, splits=[3, 5], s16=s16, kernels=ran)
            assert ran.get("k_pipe", 0) >= 2 and "k_stream" not in ran, f"{sc.name}: kernels {ran}"
            ref = S.run_oracle(sc, inputs, P, ramps, oramp)
            for s in range(n):
                assert got[s][0] == ref[s][0], f"{sc.name}: stream {s} counts"
                assert np.array_equal(got[s][1], ref[s][1]), f"{sc.name}: stream {s} PCM differs (s16={s16})"


def _pcm_values(sc, raw):
    """interleaved PCM bytes -> float64 values in units of one LSB (16 / 24 / 32 bit) or of full scale (float output)"""
    if sc.bit_depth == 0:
        return raw.view(np.float32).astype(np.float64)
    if sc.bit_depth == 16:
        return raw.view(np.int16).astype(np.float64)
    if sc.bit_depth == 32:
        return raw.view(np.int32).astype(np.float64)
    b = raw.reshape(-1, 3).astype(np.int32)
    v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
    return np.where(v >= 1 << 23, v - (1 << 24), v).astype(np.float64)


def test_fma_arithmetic_stays_within_the_stated_tolerance():
    # IAMFB_ARITH_FMA (include/iamf_b200.h): the HOA matrix fuses multiply and add - no longer bit for bit, but within
    # BASELINE.json's tolerance against the (exact) oracle: +-1 LSB at 16 bit, 1e-5 of full scale as float (measured: 2e-7);
    # at 24 bit one LSB is two float32 ulps near full scale, so the bound there is +-2 LSB.
    # Signatures without a fused variant must stay bit-identical under the switch.
    import dataclasses
    from gpu_harness import run_product
    cases = [(S.c3_toa_to_H(), 1.0), (dataclasses.replace(S.c3_toa_to_H(), bit_depth=24), 2.0),
             (dataclasses.replace(S.c3_toa_to_H(), bit_depth=0), 1e-6), (dataclasses.replace(S.c3_toa_to_H(), limiter=False), 1.0),
             (S.c5_resample(), 0.0), (S.c2_714_to_B(), 0.0)]
    for k, (sc, tol) in enumerate(cases):
        sc = dataclasses.replace(sc, arithmetic=1, name=sc.name + "_fma")
        n, F = 21, 10
        inputs = S.synth_inputs(sc, n, F, seed=0x1A3F + 300 + k)
        P, ramps, oramp = S.synth_params(sc, n, F, seed=0x77 + 300 + k)
        ref = S.run_oracle(sc, inputs, P, ramps, oramp)
        for s16 in (False, True):
            from iac_b200 import Engine
            eng = Engine(S.plan_desc(sc), 1, 1)
            assert eng.arithmetic == (1 if tol else 0), f"{sc.name}: arithmetic {eng.arithmetic}"
            eng.close()
            got, _ = run_product(sc, inputs, P, ramps, oramp, splits=[4, 6], s16=s16)
            worst, differing = 0.0, 0
            for s in range(n):
                assert got[s][0] == ref[s][0], f"{sc.name}: stream {s} counts"
                a, b = _pcm_values(sc, got[s][1]), _pcm_values(sc, ref[s][1])
                assert a.shape == b.shape
                d = np.abs(a - b)
                worst = max(worst, float(d.max()))
                differing += int((d != 0).sum())
            assert worst <= tol, f"{sc.name}: differs from the oracle by {worst} (allowed {tol}), s16={s16}"
            if tol:
                assert differing > 0, f"{sc.name}: identical to the exact kernels - the fused variant did not run?"
