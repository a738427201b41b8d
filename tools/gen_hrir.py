#!/usr/bin/env python3
"""Generates the in-repo SYNTHETIC head-related impulse responses of the binaural (HRTF) renderer:

    iac_b200/csrc/iamfb_hrir.inc

The reference's binaural path hands the planar frames to two closed libraries (BEAR: per-loudspeaker HRIRs of its
default.tf; Resonance Audio: spherical-harmonic-domain HRIRs) that are absent from the tree (SURVEY 8c, m2b_rdr.c:103-121,
h2b_rdr.c:109-131), so no reference data exists to carry.  This set stands in for them: a rigid-sphere head model
(Woodworth inter-aural delay, first-order head shadow), three pinna echoes and a short decaying diffuse tail, 256 taps
at 48 kHz, written as Q15 int16 (what 16-bit SOFA / WAV HRIR sets hold): the renderer is an exact integer contraction on
the int8 tensor cores (iac_b200/csrc/iamfb_hrtf.cuh), so the product and the test oracle agree bit for bit.

  speaker set  (M2B): one (left ear, right ear) pair per IAChannel id 1..23 (IAMF_types.h:61-90)
  ambisonic set (H2B): one pair per ACN channel 0..15 (SN3D), = the speaker-domain responses of a 14-direction virtual
                       array (cube faces + corners) weighted with the real spherical harmonics of each direction

Deterministic (seeded); run:  python tools/gen_hrir.py
"""
import hashlib
import math
import os

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OUT = os.path.join(ROOT, "iac_b200", "csrc", "iamfb_hrir.inc")
TAPS, FS = 256, 48000.0
HEAD_R, C_SOUND = 0.0875, 343.0

# IAChannel id -> (azimuth deg, + = left; elevation deg)
SPEAKERS = {
    1: (30, 0), 2: (-30, 0), 3: (0, 0), 4: (0, -15), 5: (90, 0), 6: (-90, 0), 7: (135, 0), 8: (-135, 0),
    9: (45, 45), 10: (-45, 45), 11: (135, 45), 12: (-135, 45), 13: (0, 0), 14: (30, 0), 15: (-30, 0),
    16: (90, 45), 17: (-90, 45), 18: (30, 0), 19: (-30, 0), 20: (110, 0), 21: (-110, 0), 22: (90, 45), 23: (-90, 45),
}


def direction(az, el):
    a, e = math.radians(az), math.radians(el)
    return np.array([math.cos(e) * math.cos(a), math.cos(e) * math.sin(a), math.sin(e)])   # x front, y left, z up


def frac_delay(delay, gain):
    """windowed-sinc fractional delay of `delay` samples"""
    n = np.arange(TAPS, dtype=np.float64)
    x = n - delay
    w = np.where(np.abs(x) < 16, 0.5 * (1 + np.cos(np.pi * x / 16)), 0.0)
    return gain * np.sinc(x) * w


def hrir(az, el, ear, rng):
    """ear = +1 left, -1 right"""
    d = direction(az, el)
    ear_axis = np.array([0.0, float(ear), 0.0])
    cos_inc = float(np.dot(d, ear_axis))                      # +1: source on this ear's side
    theta = math.acos(max(-1.0, min(1.0, cos_inc)))            # angle between source and ear axis
    # Woodworth: extra path around the sphere for the far ear
    extra = HEAD_R * (theta - math.pi / 2 + (1 - math.cos(theta - math.pi / 2))) if theta > math.pi / 2 else HEAD_R * (1 - math.sin(theta)) * 0.0
    delay = 24.0 + extra / C_SOUND * FS
    shadow = 0.5 * (1 + cos_inc)                               # 1 near ear ... 0 far ear
    gain = 0.35 + 0.65 * shadow
    h = frac_delay(delay, gain)
    # head shadow: one-pole low-pass, stronger for the far ear
    a = 0.15 + 0.6 * (1 - shadow)
    y = np.zeros(TAPS)
    acc = 0.0
    for i in range(TAPS):
        acc = (1 - a) * h[i] + a * acc
        y[i] = acc
    # pinna echoes (elevation dependent) and a decaying diffuse tail
    for k, (dl, g) in enumerate([(7.3, 0.28), (11.6, -0.17), (17.2, 0.11)]):
        y += frac_delay(delay + dl * (1 + 0.25 * math.sin(math.radians(el)) * (k + 1) / 3), g * gain * (0.6 + 0.4 * shadow))
    tail = rng.standard_normal(TAPS) * np.exp(-np.arange(TAPS) / 40.0) * 0.02 * gain
    tail[: int(delay) + 20] = 0.0
    y += tail
    return y.astype(np.float32)


def real_sh(order_max, d):
    """real spherical harmonics, ACN order, SN3D normalisation, up to order 3"""
    x, y, z = d
    sh = [1.0, y, z, x]
    if order_max >= 2:
        sh += [math.sqrt(3) * x * y, math.sqrt(3) * y * z, 0.5 * (3 * z * z - 1), math.sqrt(3) * x * z, math.sqrt(3) / 2 * (x * x - y * y)]
    if order_max >= 3:
        sh += [math.sqrt(5 / 8) * y * (3 * x * x - y * y), math.sqrt(15) * x * y * z, math.sqrt(3 / 8) * y * (5 * z * z - 1),
               0.5 * z * (5 * z * z - 3), math.sqrt(3 / 8) * x * (5 * z * z - 1), math.sqrt(15) / 2 * z * (x * x - y * y),
               math.sqrt(5 / 8) * x * (x * x - 3 * y * y)]
    return np.array(sh)


def main():
    rng = np.random.default_rng(0x1A3F)
    spk = np.zeros((24, 2, TAPS), np.float32)
    for ch, (az, el) in sorted(SPEAKERS.items()):
        for e, ear in enumerate((+1, -1)):
            spk[ch, e] = hrir(az, el, ear, rng)
    # virtual array for the ambisonic set: 6 cube faces + 8 corners
    dirs = [(0, 0), (180, 0), (90, 0), (-90, 0), (0, 90), (0, -90)] + [(a, e) for a in (45, 135, -45, -135) for e in (35.26, -35.26)]
    amb = np.zeros((16, 2, TAPS), np.float64)
    for az, el in dirs:
        y = real_sh(3, direction(az, el))
        for e, ear in enumerate((+1, -1)):
            h = hrir(az, el, ear, rng).astype(np.float64)
            for m in range(16):
                amb[m, e] += (2.0 / len(dirs)) * y[m] * h
    amb = amb.astype(np.float32)
    allf = np.concatenate([spk.reshape(-1), amb.reshape(-1)]).astype(np.float64)
    peak = float(np.abs(allf).max())
    assert peak < 1.0, peak
    pool = np.clip(np.rint(allf * 32768.0), -32767, 32767).astype(np.int16)
    sha = hashlib.sha256(pool.tobytes()).hexdigest()
    with open(OUT, "w") as f:
        f.write("// GENERATED by tools/gen_hrir.py - do not edit.  Synthetic HRIR set of the binaural renderer (see the generator).\n")
        f.write(f"// {TAPS} taps at 48 kHz, Q15 int16; speaker set [24 IAChannel ids][2 ears][taps], then the ambisonic set\n")
        f.write(f"// [16 ACN channels][2 ears][taps].  peak {peak:.6f}  sha256 {sha}\n")
        f.write(f"static constexpr int k_hrir_taps = {TAPS};\n")
        f.write(f"static constexpr int k_hrir_amb_off = {24 * 2 * TAPS};\n")
        f.write("static const int16_t k_hrir_q15[] = {\n")
        for i in range(0, len(pool), 16):
            f.write("  " + ", ".join(str(int(v)) for v in pool[i:i + 16]) + ",\n")
        f.write("};\n")
    print(OUT, len(pool), "taps", "peak", peak, sha)


if __name__ == "__main__":
    main()
