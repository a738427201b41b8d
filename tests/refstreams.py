"""Maps the BASELINE.json scenarios onto synthetic IAMF bitstreams (tests/iamfgen.py) so that the compiled reference
can be driven end to end through its public API on exactly the decoded PCM / parameters a Scenario describes.
Shared by tests/test_oracle_pipeline_vs_ref.py, tests/test_dropin_vs_ref.py and bench.py's CPU reference leg."""
import numpy as np

import iamfgen as G
import scenarios as S


def lin_q78(q78):
    """db2lin(q_to_float(q, 8)) with the reference's float arithmetic (fixedp11_5.c:45-47,72)"""
    db = np.float32(q78) * np.float32(2.0 ** -8)
    return float(np.float32(np.power(np.float32(10.0), np.float32(0.05) * db, dtype=np.float32)))


def to_i16(x):
    return np.rint(np.asarray(x, np.float64) * 32768.0).astype(np.int16)


def case(name, **kw):
    """returns (scenario, iamfgen.Stream, api kwargs, unit_kw(f, P, s) -> temporal_unit kwargs)"""
    import orcbind
    L = orcbind.lib()

    def lin(q):
        return float(L.orc_db2lin(L.orc_q_to_float(q, 8)))

    unit_kw = None
    if name == "c1":
        sc, st, api = S.c1_stereo(**kw), G.cfg_stereo(), dict(sound_system=0)
    elif name == "c2":
        sc = S.c2_714_to_B(**kw)
        sc.elements[0].out_gain = [(S.R2, lin(0x0100)), (S.L2, lin(0x0100))]
        st, api = G.cfg_714_scalable(), dict(sound_system=1)

        def unit_kw(f, P, s=0):
            rg = None
            if P["has_recon0"][s, f]:
                rg = {0: [None, (int(P["recon_flags0"][s, f]), [int(v) for v in P["recon_gain0"][s, f][:4]])]}
            return dict(demix_mode={0: int(P["dmx_mode0"][s, f])}, recon=rg)
    elif name == "c3":
        sc, st, api = S.c3_toa_to_H(**kw), G.cfg_toa(), dict(sound_system=7)
    elif name in ("c4", "c4h"):
        # c4h: the same mix through the HRTF renderer (the reference as built cannot run it: bench.py's CPU leg for it is the port)
        sc = S.c4_hrtf(**kw) if name == "c4h" else S.c4_714_foa_binaural(**kw)
        sc.elements[0].mix_gain = lin(-0x0300)
        sc.elements[1].mix_gain = lin(-0x0300)
        st, api = G.cfg_714_foa(), dict(binaural=True)
    elif name == "c5":
        sc = S.c5_resample(**kw)
        sc.loudness_gain = float(L.orc_db2lin(-24.0 - (-16.0)))
        st = G.cfg_stereo(rate=44100, frame_size=1024, loud_q78=-16 * 256)
        api = dict(sound_system=0, rate=48000, loudness=-24.0)
    else:
        raise KeyError(name)
    api["bit_depth"] = sc.bit_depth
    api["limiter"] = sc.limiter
    api["threshold_db"] = sc.threshold_db
    return sc, st, api, unit_kw


def no_param_gaps(sc, P):
    """a temporal unit without its parameter block de-synchronises the reference's parameter time line for good
    (IAMF_decoder.c:791-803,1089-1126); that host-side behaviour is exercised by the drop-in tests only"""
    for e in range(len(sc.elements)):
        if sc.elements[e].recon_flags:
            P[f"has_recon{e}"] = 1
    return P


def temporal_units(sc, st, inputs, P, unit_kw, s=0):
    F = P.shape[1]
    units = []
    for f in range(F):
        kw = dict(unit_kw(f, P, s) if unit_kw else {})
        pcm = [to_i16(inputs[e][s, f]) for e in range(len(sc.elements))]
        units.append(st.temporal_unit(pcm, trim_start=int(P["trim_start"][s, f]), trim_end=int(P["trim_end"][s, f]), **kw))
    return units
