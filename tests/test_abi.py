"""CPU-side checks of the native product: the C-ABI library loads, exports every symbol include/iamf_b200.h declares,
refuses to run without a device (no CPU fallback), and carries the reference's rendering matrices unchanged."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

import iac_b200
from iac_b200 import binding

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "iamf_b200.h")).read()
    return sorted(set(re.findall(r"\b(iamfb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = iac_b200.lib()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/iamf_b200.h but not exported"


def test_frame_params_layout():
    assert C.sizeof(binding.FrameParams) == 48
    assert binding.FRAME_PARAMS_DTYPE.itemsize == 48


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = iac_b200.lib()
    ctx = C.c_void_p()
    r = L.iamfb_ctx_create(0, C.byref(ctx))
    assert r == -100  # IAMFB_ERR_NO_DEVICE
    assert b"no CPU path" in L.iamfb_last_error()


def test_matrix_table_pinned():
    inc = os.path.join(ROOT, "iac_b200", "csrc", "iamfb_matrices.inc")
    assert hashlib.sha256(open(inc, "rb").read()).hexdigest() == \
        "5f28d753d607602c7ee6011aed58e096da621a7df8142bb82ffb63e92fecfddd"


def test_matrices_equal_reference():
    import refbind
    if not refbind.have_ref():
        pytest.skip("oracle/_ref not built")
    out_ids = [0x020, 0x050, 0x250, 0x450, 0x451, 0x370, 0x490, 0x9A3, 0x070, 0x470, 0x712, 0x312, 0x100, 0x1020]
    n = 0
    for li, lin in enumerate(refbind.LAYER_IDS):
        for oi, out in enumerate(out_ids):
            r = refbind.m2m_matrix(lin, out)
            mine = binding.get_m2m_matrix(li, oi)
            assert (r is None) == (mine is None)
            if r is not None:
                assert np.array_equal(r[2].view(np.uint32), mine.view(np.uint32))
                n += 1
    for order in range(4):
        for oi, out in enumerate(out_ids):
            r = refbind.h2m_matrix(order, out)
            mine = binding.get_h2m_matrix(order, oi)
            assert r is not None and mine is not None
            assert np.array_equal(r[4].view(np.uint32), mine[0].view(np.uint32)) and (r[2], r[3]) == mine[1:]
            n += 1
    assert n == 196


def test_target_channel_counts():
    L = iac_b200.lib()
    assert [L.iamfb_target_channels(t) for t in range(14)] == [2, 6, 8, 10, 11, 12, 14, 24, 8, 12, 10, 6, 1, 2]
