"""iac_b200 - B200 (sm_100a) implementation of libiamf's post-decode rendering path.

The product is native: iac_b200/libiamf_b200.so (CUDA kernels + the C ABI of include/iamf_b200.h) and, on top of it,
the drop-in libiamf.so exporting the reference's public IAMF_decoder.h API.  This Python package only binds the C ABI
(ctypes) for tests and bench.py; there is no Python or CPU compute path.
"""
from .binding import (Engine, ElementDesc, FrameParams, PlanDesc, IamfB200Error, lib, lib_path,  # noqa: F401
                      channel_element, scene_element, frame_params_array)
