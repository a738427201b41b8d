"""GPU parity of the binaural HRTF renderer (tcgen05 int8 contraction, iac_b200/csrc/iamfb_hrtf.cuh) through the C ABI.

PARITY UNPINNED BY THE REFERENCE: m2b_rdr.c / h2b_rdr.c call closed libraries that are absent from the reference tree
(SURVEY 8c), so the checker is the self-oracle oracle/oracle_hrtf.c (exact integer FIR, rounded once) inside the pinned
pipeline oracle.  The renderer is exact integer arithmetic, so the bar is still BIT-EXACT - float output and 16/24-bit
PCM alike (north_star allows 1e-5 of full scale / +-1 LSB)."""
import numpy as np
import pytest

import scenarios as S
from test_gpu_parity import compare

pytestmark = pytest.mark.gpu

CASES = S.hrtf_cases()


@pytest.mark.parametrize("sc", CASES, ids=[s.name for s in CASES])
def test_hrtf_single_submit(sc):
    compare(sc, 5, 6, [6], seed=31)


@pytest.mark.parametrize("sc", CASES, ids=[s.name for s in CASES])
def test_hrtf_ragged_submits(sc):
    # the filter state (255 instants per channel) crosses submit boundaries of every length, incl. a 1-frame submit
    compare(sc, 21, 12, [1, 4, 2, 5], seed=32)


@pytest.mark.parametrize("sc", [c for c in CASES if all(e.hrtf for e in c.elements)], ids=[c.name for c in CASES if all(e.hrtf for e in c.elements)])
def test_hrtf_ragged_submits_int16(sc):
    # the same with int16 hand-over: elements whose PCM reaches the renderer untouched take the contraction kernel's RAW
    # variant (limb rows made in the kernel, history in / out by its converter warps), the others the two-limb prep pass;
    # 14 frames in one submit = two tiles per stream
    compare(sc, 21, 12, [1, 4, 2, 5], seed=37, s16=True)
    compare(sc, 5, 14, [14], seed=38, s16=True)


def test_hrtf_int16_upload_two_limbs():
    # 16-bit decoded PCM travels as two limbs (Q15) instead of three (Q20): same bits out
    compare(S.c4_hrtf(), 70, 5, [2, 3], seed=33, s16=True)
    compare(S.hrtf_cases()[2], 9, 4, [4], seed=34, s16=True)


@pytest.mark.parametrize("s16", [False, True], ids=["f32", "s16"])
def test_hrtf_missing_frames_keep_the_filter_state(s16):
    # streams that have no frame in some steps (trim_start 0xFFFF): the renderer sees their present frames as one signal
    sc = S.c4_hrtf()

    def edit(P):
        P["trim_start"][1::3, 2] = 0xFFFF
        P["trim_start"][2::5, 0:3] = 0xFFFF
    from gpu_harness import run_product
    n, F = 11, 8
    inputs = S.synth_inputs(sc, n, F, seed=0x1A3F + 35)
    P, ramps, oramp = S.synth_params(sc, n, F, seed=0x77 + 35)
    edit(P)
    P["trim_start"][4::5, 3:] = 0xFFFF           # nothing at all in the second submit
    got, _ = run_product(sc, inputs, P, ramps, oramp, splits=[3, 5], s16=s16)
    # the oracle is fed the present frames only (a stream without a frame in a step is simply not called)
    for s in range(n):
        keep = [f for f in range(F) if P["trim_start"][s, f] != 0xFFFF]
        ref = S.run_oracle(sc, [x[s:s + 1][:, keep] for x in inputs], P[s:s + 1][:, keep])[0]
        cnt = [c for f, c in enumerate(got[s][0][:F]) if f in keep] + got[s][0][F:]
        assert cnt == ref[0], f"stream {s}: counts"
        assert np.array_equal(got[s][1], ref[1]), f"stream {s}: PCM differs"


def test_hrtf_full_size_config4():
    # BASELINE.json configs[3] at its stream count: 16 distinct streams replicated over 2048 must give the same bytes
    # wherever they sit, equal to the oracle
    from gpu_harness import run_product
    sc = S.c4_hrtf()
    n, F, reps = 16, 4, 128
    inputs = S.synth_inputs(sc, n, F, seed=0x1A3F + 36)
    P, _, _ = S.synth_params(sc, n, F, seed=0x77 + 36)
    big = [np.tile(x, (reps, 1, 1, 1)) for x in inputs]
    got, _ = run_product(sc, big, np.tile(P, (reps, 1)), splits=[F], s16=True)
    ref = S.run_oracle(sc, inputs, P)
    for s in range(n * reps):
        assert got[s][0] == ref[s % n][0]
        assert np.array_equal(got[s][1], ref[s % n][1]), f"stream {s}"


def test_hrtf_scalable_element_int16_and_missing_frames():
    # a scalable element (its layout channels derived by the de-mixer, recon gain, output gain) in front of the renderer, with
    # int16 hand-over (widened for the de-mixer) and streams without frames in some steps: the de-mixer state of the front
    # end and the filter state stay in step
    from gpu_harness import run_product
    sc = S.hrtf_cases()[5]
    assert sc.name == "hrtf_714_scalable_recon"
    compare(sc, 9, 7, [3, 4], seed=41, s16=True)
    n, F = 8, 9
    inputs = S.synth_inputs(sc, n, F, seed=0x1A3F + 42)
    P, ramps, oramp = S.synth_params(sc, n, F, seed=0x77 + 42)
    P["trim_start"][1::3, 2] = 0xFFFF
    P["trim_start"][2::3, 4:] = 0xFFFF
    got, _ = run_product(sc, inputs, P, ramps, oramp, splits=[4, 5])
    for s in range(n):
        keep = [f for f in range(F) if P["trim_start"][s, f] != 0xFFFF]
        ref = S.run_oracle(sc, [x[s:s + 1][:, keep] for x in inputs], P[s:s + 1][:, keep])[0]
        assert np.array_equal(got[s][1], ref[1]), f"stream {s}: PCM differs"


def test_hrtf_refuses_what_it_does_not_render():
    from iac_b200 import Engine
    sc = S.c4_hrtf()
    sc.in_rate = sc.out_rate = 44100     # the HRIR set is sampled at 48 kHz
    with pytest.raises(Exception):
        Engine(S.plan_desc(sc), 4, 2)
