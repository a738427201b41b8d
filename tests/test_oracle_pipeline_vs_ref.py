"""Pins the oracle's whole-path driver (oracle_pipeline.c) and the scenario/parameter conventions against the
compiled reference driven END TO END through its public API (IAMF_decoder_open/configure/decode) on synthetic ipcm
bitstreams: same sample counts per call and bit-identical PCM.  Skipped where oracle/_ref is not built."""
import numpy as np
import pytest

import iamfapi
import iamfgen as G
import orcbind
import refbind
import scenarios as S

pytestmark = pytest.mark.skipif(not refbind.have_ref(), reason="oracle/_ref not built (make -C oracle ref)")


def lin(q78):
    L = orcbind.lib()
    return float(L.orc_db2lin(L.orc_q_to_float(q78, 8)))


def to_i16(x):
    return np.rint(x * 32768.0).astype(np.int16)


def run_case(sc, stream, F, api_kw, unit_kw=None, seed=11):
    """sc: scenario (1 stream), stream: iamfgen.Stream describing the same thing"""
    inputs = S.synth_inputs(sc, 1, F, seed=seed)
    P, ramps, oramp = S.synth_params(sc, 1, F, seed=seed)
    # a temporal unit without its parameter block de-synchronises the reference's parameter time line for good
    # (IAMF_decoder.c:791-803,1089-1126); that is host-side behaviour, exercised by the drop-in tests, not here
    for e in range(len(sc.elements)):
        if sc.elements[e].recon_flags:
            P[f"has_recon{e}"] = 1
    units = []
    for f in range(F):
        kw = dict(unit_kw(f, P) if unit_kw else {})
        pcm = [to_i16(inputs[e][0, f]) for e in range(len(sc.elements))]
        units.append(stream.temporal_unit(pcm, trim_start=int(P["trim_start"][0, f]), trim_end=int(P["trim_end"][0, f]), **kw))
    api = iamfapi.Api(refbind.REF_SO)
    ref_pcm, ref_counts = api.render(stream.descriptors(), units, **api_kw)
    res = S.run_oracle(sc, inputs, P, ramps, oramp)
    counts, raw = res[0]
    assert counts == ref_counts
    assert raw.tobytes() == ref_pcm.tobytes()


def test_c1_stereo():
    run_case(S.c1_stereo(), G.cfg_stereo(), 12, dict(sound_system=0))


def test_c1_stereo_loud_peaks_24bit():
    run_case(S.c1_stereo(bit_depth=24, peak_db=(1.0, 3.0)), G.cfg_stereo(), 12, dict(sound_system=0, bit_depth=24))


def test_c1_stereo_no_limiter_32bit():
    run_case(S.c1_stereo(bit_depth=32, limiter=False, peak_db=(-1.0, 0.0)), G.cfg_stereo(), 6,
             dict(sound_system=0, bit_depth=32, limiter=False))


def test_c2_714_scalable_to_B():
    sc = S.c2_714_to_B()
    sc.elements[0].out_gain = [(S.R2, lin(0x0100)), (S.L2, lin(0x0100))]
    st = G.cfg_714_scalable()

    def unit_kw(f, P):
        rg = None
        if P["has_recon0"][0, f]:
            rg = {0: [None, (int(P["recon_flags0"][0, f]), [int(v) for v in P["recon_gain0"][0, f][:4]])]}
        return dict(demix_mode={0: int(P["dmx_mode0"][0, f])}, recon=rg)
    run_case(sc, st, 14, dict(sound_system=1), unit_kw)


def test_c3_toa_to_H():
    run_case(S.c3_toa_to_H(), G.cfg_toa(), 6, dict(sound_system=7))


def test_c4_714_foa_binaural():
    sc = S.c4_714_foa_binaural()
    sc.elements[0].mix_gain = lin(-0x0300)
    sc.elements[1].mix_gain = lin(-0x0300)
    run_case(sc, G.cfg_714_foa(), 6, dict(binaural=True))


def test_c5_resample_loudness():
    sc = S.c5_resample()
    sc.loudness_gain = float(orcbind.lib().orc_db2lin(-24.0 - (-16.0)))
    run_case(sc, G.cfg_stereo(rate=44100, frame_size=1024, loud_q78=-16 * 256), 9, dict(sound_system=0, rate=48000, loudness=-24.0))


def test_trims():
    sc = S.c1_stereo(trims={0: (312, 0), 2: (960, 0), 5: (0, 100), 7: (0, 960)}, peak_db=(0.0, 2.0))
    run_case(sc, G.cfg_stereo(), 9, dict(sound_system=0))
