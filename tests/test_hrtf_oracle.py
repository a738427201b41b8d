"""CPU checks around the binaural HRTF renderer: the self-oracle (oracle/oracle_hrtf.c) against an independent numpy
statement of its definition, the pinned HRIR set, and the tensor-core instructions in the shipped library."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

import orcbind
from iac_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _taps(kind, idx):
    return np.stack([binding.get_hrir(kind, i) for i in idx])


def test_hrir_set_is_pinned():
    spk = _taps(0, range(1, 24))
    amb = _taps(1, range(16))
    assert spk.shape == (23, 2, 256) and amb.shape == (16, 2, 256)
    h = hashlib.sha256(np.concatenate([np.zeros(512, np.int16), spk.reshape(-1), amb.reshape(-1)]).tobytes()).hexdigest()
    line = [l for l in open(os.path.join(ROOT, "iac_b200", "csrc", "iamfb_hrir.inc")) if "sha256" in l][0]
    assert h in line                                        # the generated table is what the accessor serves
    assert h == "a4864a8158a6942d43f7751d3e474f704ccfcc9b760842e7903c203d6c44afb8"
    # a head model: the ear on the source's side is louder and earlier
    left, right = spk[5 - 1, 0].astype(np.float64), spk[5 - 1, 1].astype(np.float64)      # SL7 at +90 degrees
    assert (left ** 2).sum() > 4 * (right ** 2).sum()
    assert np.argmax(np.abs(left)) < np.argmax(np.abs(right))


def test_oracle_hrtf_matches_its_definition():
    L = orcbind.lib()
    L.orc_hrtf_open.restype = C.c_void_p
    L.orc_hrtf_open.argtypes = [C.c_int, C.POINTER(C.c_int16)]
    L.orc_hrtf_render.argtypes = [C.c_void_p, orcbind.f32p, orcbind.f32p, C.c_int]
    L.orc_hrtf_close.argtypes = [C.c_void_p]
    rng = np.random.default_rng(5)
    nch, n, calls = 5, 300, 4
    taps = np.ascontiguousarray(_taps(0, [1, 2, 3, 7, 12]), np.int16)
    x = (rng.uniform(-1.2, 1.2, (nch, n * calls))).astype(np.float32)
    x[0, :50] = rng.integers(-32768, 32768, 50) / np.float32(32768)       # 16-bit content is carried exactly
    x[1, 5] = 9.0                                                         # clamped to the Q20 range
    h = L.orc_hrtf_open(nch, taps.ctypes.data_as(C.POINTER(C.c_int16)))
    out = np.zeros((2, n * calls), np.float32)
    for k in range(calls):                                                # the state is carried from call to call
        xi = np.ascontiguousarray(x[:, k * n:(k + 1) * n])
        o = np.zeros((2, n), np.float32)
        L.orc_hrtf_render(h, xi.ctypes.data_as(orcbind.f32p), o.ctypes.data_as(orcbind.f32p), n)
        out[:, k * n:(k + 1) * n] = o
    L.orc_hrtf_close(h)
    xq = np.rint(np.clip(x.astype(np.float32) * np.float32(1048576.0), -8388607.0, 8388607.0)).astype(np.int64)
    ref = np.zeros((2, n * calls), np.int64)
    for c in range(nch):
        for ear in range(2):
            ref[ear] += np.convolve(xq[c], taps[c, ear].astype(np.int64))[: n * calls]
    want = (ref.astype(np.float64) * 2.0 ** -35).astype(np.float32)       # int64 -> float32: one rounding
    want = np.array([[np.float32(int(v)) for v in row] for row in ref], np.float32) * np.float32(2.0 ** -35)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


def test_library_carries_tensor_core_instructions():
    """the HRTF contraction is tcgen05 code: UTCIMMA (int8 tensor MMA into TMEM) and LDTM (TMEM loads) in the SASS"""
    so = os.path.join(ROOT, "iac_b200", "libiamf_b200.so")
    # (the translation unit's cubin alone: dumping the whole library takes a quarter of a minute)
    lst = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    name = [l.split(":")[1].strip() for l in lst.splitlines() if "iamfb_hrtf" in l]
    assert name, lst
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", name[0], so], cwd=tmp, check=True, capture_output=True)
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(tmp, name[0])], capture_output=True, text=True).stdout
    assert "UTCIMMA" in sass and "LDTM" in sass and "UBLKCP" in sass
