"""Pins the oracle: our C restatement (oracle/liboracle.so) must be BIT-IDENTICAL to the compiled, unmodified
reference (oracle/_ref/libiamf_ref.so) stage by stage on seeded random inputs.  Skipped where _ref is not built."""
import numpy as np
import pytest

import orcbind
import refbind

pytestmark = pytest.mark.skipif(not refbind.have_ref(), reason="oracle/_ref not built (make -C oracle ref)")

# channel ids (IAMF_types.h:61-90)
L7, R7, C_, LFE, SL7, SR7, BL7, BR7, HFL, HFR, HBL, HBR, MONO, L2, R2, TL, TR, L3, R3, SL5, SR5, HL, HR = range(1, 24)
L5, R5 = L7, R7
MONO_, STEREO, L510, L512, L514, L710, L712, L714, L312, BIN = range(10)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same(a, b):
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


def test_scalars():
    R, O = refbind.lib(), orcbind.lib()
    for q in range(-32768, 32768, 97):
        assert R.q_to_float(q, 8) == O.orc_q_to_float(q, 8)
        assert R.q_to_float(q, 15) == O.orc_q_to_float(q, 15)
        assert R.db2lin(R.q_to_float(q, 8)) == O.orc_db2lin(O.orc_q_to_float(q, 8))
    for q in range(256):
        assert R.qf_to_float(q, 8) == O.orc_qf_to_float(q, 8)


DEMIX_CASES = [
    # (layout, transmission order, gain channels, recon channels+flags)
    ("mono->stereo", STEREO, [MONO, L2], [MONO], [R2], 0x5),
    ("2.0->7.1.4 (2 layers)", L714, [L2, R2, L7, R7, SL7, SR7, HFL, HFR, TL, TR, C_, LFE][:0] or
     [L2, R2, L5, R5, SL7, SR7, HFL, HFR, TL, TR, C_, LFE], [L2, R2], [BL7, BR7, HBL, HBR], 0x780),
    ("2.0->3.1.2->5.1.2->7.1.4", L714, [L2, R2, TL, TR, C_, LFE, L5, R5, SL7, SR7, HFL, HFR], [L2, R2],
     [BL7, BR7, HBL, HBR], 0x780),
    ("2.0->5.1", L510, [L2, R2, L5, R5, C_, LFE], [], [SL5, SR5], 0x18),
    ("5.1->7.1", L710, [L5, R5, SL5, SR5, C_, LFE, SL7, SR7], [SL5, SR5], [BL7, BR7], 0x180),
    ("3.1.2->5.1.2", L512, [L3, R3, TL, TR, C_, LFE, L5, R5], [], [SL5, SR5, HL, HR], 0x78),
    ("5.1.2->5.1.4", L514, [L5, R5, SL5, SR5, HL, HR, C_, LFE, HFL, HFR], [], [HBL, HBR], 0x600),
    ("7.1.4 single layer", L714, [L7, R7, SL7, SR7, BL7, BR7, HFL, HFR, HBL, HBR, C_, LFE], [], [], 0),
    ("2.0->3.1.2", L312, [L2, R2, TL, TR, C_, LFE], [L2], [L3, R3], 0x5),
]


@pytest.mark.parametrize("case", DEMIX_CASES, ids=[c[0] for c in DEMIX_CASES])
@pytest.mark.parametrize("fs", [960, 1024, 128])
def test_demixer(case, fs):
    _, layout, chs_in, gain_chs, recon_chs, flags = case
    rng = np.random.default_rng(1234 + fs)
    gains = [1.25 + 0.1 * i for i in range(len(gain_chs))]
    r = refbind.RefDemixer(fs, layout, chs_in, gain_chs, gains, mode=1, w_idx=0)
    o = orcbind.OrcDemixer(fs, layout, chs_in, gain_chs, gains, mode=1, w_idx=0)
    r.set_offset(0)
    o.set_offset(0)
    modes = [1, 2, 4, 5, 6, 0, 4, 4, 4, 4, 4, 6, 5, 0, 1, 1]
    for f in range(24):
        x = (rng.standard_normal((len(chs_in), fs)) * 0.3).astype(np.float32)
        if recon_chs:
            g = (rng.integers(0, 256, len(recon_chs)) / np.float32(255)).astype(np.float32)
            r.set_recon(recon_chs, g, flags)
            o.set_recon(recon_chs, g, flags)
        if f % 3 != 2:
            m = modes[f % len(modes)]
            assert r.set_mode(m) == o.set_mode(m)
        rr, ro = r.demix(x)
        orr, oo = o.demix(x)
        assert (rr == 0) == (orr == 0)
        assert same(ro, oo), f"frame {f}"
    r.close()
    o.close()


def test_demixer_frame_offset():
    fs = 960
    chs = [L2, R2, L5, R5, SL7, SR7, HFL, HFR, TL, TR, C_, LFE]
    rng = np.random.default_rng(7)
    for off in (0, 312, 900, 959):
        r = refbind.RefDemixer(fs, L714, chs, [], [], mode=2, w_idx=3)
        o = orcbind.OrcDemixer(fs, L714, chs, [], [], mode=2, w_idx=3)
        r.set_offset(off)
        o.set_offset(off)
        for f in range(6):
            x = (rng.standard_normal((12, fs)) * 0.3).astype(np.float32)
            g = rng.random(4).astype(np.float32)
            r.set_recon([BL7, BR7, HBL, HBR], g, 0x780)
            o.set_recon([BL7, BR7, HBL, HBR], g, 0x780)
            m = [4, 1, 6, 0, 5, 2][f]
            r.set_mode(m)
            o.set_mode(m)
            assert same(r.demix(x)[1], o.demix(x)[1])
        r.close()
        o.close()


def test_dmr_all_pairs():
    rng = np.random.default_rng(99)
    counts = [1, 2, 6, 8, 10, 8, 10, 12, 6, 2]
    n_valid = 0
    for lin in range(10):
        for lout in range(10):
            r, o = refbind.RefDmr(lin, lout), orcbind.OrcDmr(lin, lout)
            assert r.ok() == o.ok(), (lin, lout)
            if not r.ok():
                continue
            n_valid += 1
            r.set_mode_weight(1, 2)
            o.set_mode_weight(1, 2)
            for f, m in enumerate([1, 4, 5, 6, 0, 2, 4, 4]):
                x = (rng.standard_normal((counts[lin], 256)) * 0.4).astype(np.float32)
                if f:
                    assert r.set_mode_weight(m, -1) == o.set_mode_weight(m, -1)
                assert same(r.downmix(x, counts[lout]), o.downmix(x, counts[lout])), (lin, lout, f)
            r.close()
            o.close()
    assert n_valid == 15  # the 15 pairs listed in SURVEY 9.4-5


def test_m2m_all_pairs():
    rng = np.random.default_rng(5)
    n = 0
    for lin in refbind.LAYER_IDS:
        for name, out_id in refbind.SS_IDS.items():
            mm = refbind.m2m_matrix(lin, out_id)
            if mm is None:
                continue
            m, nn, mat = mm
            x = (rng.standard_normal((m, 200)) * 0.5).astype(np.float32)
            assert same(refbind.render_m2m(lin, out_id, x), orcbind.render_m2m(mat, x)), (hex(lin), name)
            n += 1
    assert n == 140


def test_h2m_all_pairs():
    rng = np.random.default_rng(6)
    chans = {"A": 2, "B": 6, "C": 8, "D": 10, "E": 11, "F": 12, "G": 14, "H": 24, "I": 8, "J": 12, "712": 10,
             "312": 6, "MONO": 1, "BINAURAL": 2}
    n = 0
    for order in range(4):
        for name, out_id in refbind.SS_IDS.items():
            hm = refbind.h2m_matrix(order, out_id)
            assert hm is not None
            m, nn, l1, l2, mat = hm
            x = (rng.standard_normal((m, 200)) * 0.5).astype(np.float32)
            a = refbind.render_h2m(order, out_id, x, chans[name])
            b = orcbind.render_h2m(mat, l1, l2, x, chans[name])
            assert same(a, b), (order, name)
            n += 1
    assert n == 56


@pytest.mark.parametrize("ch", [1, 2, 6, 24])
def test_limiter(ch):
    rng = np.random.default_rng(ch)
    r, o = refbind.RefLimiter(-1.0, 48000, ch), orcbind.OrcLimiter(-1.0, 48000, ch)
    for f in range(60):
        fs = [960, 960, 100, 140, 960, 1024, 17][f % 7]
        amp = [0.2, 0.95, 1.6, 0.5, 0.05, 1.1][(f // 3) % 6]
        x = (rng.standard_normal((ch, fs)) * 0.35 * amp).astype(np.float32)
        if f == 20:
            x[:] = 0
        a, b = r.process(x), o.process(x)
        assert same(a, b), f"frame {f}"
    r.close()
    o.close()


@pytest.mark.parametrize("rates", [(44100, 48000), (48000, 44100), (48000, 96000), (96000, 48000), (48000, 16000),
                                   (32000, 48000), (22050, 48000)])
def test_resampler(rates):
    rng = np.random.default_rng(rates[0] % 1000)
    r, o = refbind.RefResampler(2, *rates), orcbind.OrcResampler(2, *rates)
    for f in range(12):
        fs = [1024, 960, 1024, 333][f % 4]
        x = (rng.standard_normal((2, fs)) * 0.4).astype(np.float32)
        if f == 5:
            x *= 4  # exercise the +-1 clamp
        a, b = r.process(x), o.process(x)
        assert same(a, b), f"frame {f}: {a.shape} {b.shape}"
    assert same(r.flush(), o.flush())
    r.close()
    o.close()
