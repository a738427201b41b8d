"""ctypes binding of include/iamf_b200.h.  Fails loudly when the CUDA library is missing - no fallback."""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
MAXL, MAXS, MAXE = 12, 16, 2


class IamfB200Error(RuntimeError):
    pass


def lib_path():
    return os.path.join(PKG, "libiamf_b200.so")


class ElementDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_in", C.c_int32), ("layout", C.c_int32), ("chs_in", C.c_int32 * MAXL),
                ("n_out_gain", C.c_int32), ("out_gain_ch", C.c_int32 * MAXL), ("out_gain", C.c_float * MAXL),
                ("has_demix_info", C.c_int32), ("default_mode", C.c_int32), ("default_w_idx", C.c_int32),
                ("first_layer_layout", C.c_int32), ("selected_layer", C.c_int32), ("recon_present", C.c_int32),
                ("use_dmr", C.c_int32), ("dmr_out_layout", C.c_int32),
                ("ambi_mode", C.c_int32), ("ambi_channels", C.c_int32), ("ambi_map", C.c_uint8 * MAXS),
                ("ambi_cols", C.c_int32), ("ambi_matrix", C.c_float * (MAXS * MAXS)), ("binaural_hrtf", C.c_int32)]


class PlanDesc(C.Structure):
    _fields_ = [("frame_size", C.c_int32), ("in_rate", C.c_int32), ("out_rate", C.c_int32), ("n_elements", C.c_int32),
                ("el", ElementDesc * MAXE), ("target", C.c_int32), ("loudness_gain", C.c_float),
                ("limiter", C.c_int32), ("limiter_threshold_db", C.c_float), ("bit_depth", C.c_int32),
                ("arithmetic", C.c_int32)]    # 0 exact (bit-identical to the reference), 1 IAMFB_ARITH_FMA (tolerance mode)


class _ElParams(C.Structure):
    _fields_ = [("dmx_mode", C.c_int8), ("has_recon", C.c_uint8), ("recon_flags", C.c_uint16),
                ("recon_gain", C.c_uint8 * 12), ("mix_gain", C.c_float)]


class FrameParams(C.Structure):
    _fields_ = [("el", _ElParams * MAXE), ("out_gain", C.c_float), ("trim_start", C.c_uint16),
                ("trim_end", C.c_uint16)]


FRAME_PARAMS_DTYPE = np.dtype({
    "names": ["dmx_mode0", "has_recon0", "recon_flags0", "recon_gain0", "mix_gain0",
              "dmx_mode1", "has_recon1", "recon_flags1", "recon_gain1", "mix_gain1",
              "out_gain", "trim_start", "trim_end"],
    "formats": ["i1", "u1", "u2", ("u1", 12), "f4", "i1", "u1", "u2", ("u1", 12), "f4", "f4", "u2", "u2"],
    "offsets": [0, 1, 2, 4, 16, 20, 21, 22, 24, 36, 40, 44, 46],
    "itemsize": 48})
assert C.sizeof(FrameParams) == 48


def frame_params_array(n_streams, n_frames):
    """numpy structured array [S][F] laid out as iamfb_frame_params, with neutral defaults"""
    a = np.zeros((n_streams, n_frames), FRAME_PARAMS_DTYPE)
    a["dmx_mode0"] = -1
    a["dmx_mode1"] = -1
    a["mix_gain0"] = 1.0
    a["mix_gain1"] = 1.0
    a["out_gain"] = 1.0
    return a


class Io(C.Structure):
    _fields_ = [("in_", C.c_void_p * MAXE), ("params", C.c_void_p), ("gain_ramp", C.c_void_p * MAXE),
                ("out_gain_ramp", C.c_void_p), ("pcm", C.c_void_p), ("out_counts", C.c_void_p), ("in_format", C.c_int32),
                ("gain_segs", C.c_void_p * MAXE), ("out_gain_segs", C.c_void_p)]


_lib = None


def lib():
    """loads iac_b200/libiamf_b200.so; raises if it was not built (python -m iac_b200.build)"""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise IamfB200Error(f"{path} is missing: build it with `python -m iac_b200.build` (nvcc, sm_100a). "
                            "There is no CPU fallback.")
    L = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    vp = C.c_void_p
    L.iamfb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.iamfb_ctx_set_stream.argtypes = [vp, vp]
    L.iamfb_ctx_synchronize.argtypes = [vp]
    L.iamfb_ctx_destroy.argtypes = [vp]
    L.iamfb_ctx_destroy.restype = None
    L.iamfb_last_error.restype = C.c_char_p
    L.iamfb_version.restype = C.c_char_p
    L.iamfb_plan_create.argtypes = [vp, C.POINTER(PlanDesc), C.POINTER(vp)]
    L.iamfb_plan_destroy.argtypes = [vp]
    L.iamfb_plan_destroy.restype = None
    L.iamfb_plan_out_channels.argtypes = [vp]
    L.iamfb_selftest_quotient.argtypes = [vp, C.c_float, C.POINTER(C.c_uint64)]
    L.iamfb_plan_kernel_path.argtypes = [vp]
    L.iamfb_plan_kernel_path.restype = C.c_int
    L.iamfb_plan_arithmetic.argtypes = [vp]
    L.iamfb_plan_arithmetic.restype = C.c_int
    L.iamfb_plan_kernel_path_fmt.argtypes = [vp, C.c_int]
    L.iamfb_plan_kernel_path_fmt.restype = C.c_int
    L.iamfb_plan_max_out_samples.argtypes = [vp, C.c_int]
    L.iamfb_plan_out_stride_bytes.argtypes = [vp, C.c_int]
    L.iamfb_plan_out_stride_bytes.restype = C.c_size_t
    L.iamfb_batch_create.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.iamfb_batch_destroy.argtypes = [vp]
    L.iamfb_batch_destroy.restype = None
    L.iamfb_batch_reset.argtypes = [vp]
    L.iamfb_batch_submit_device.argtypes = [vp, C.POINTER(Io), C.c_int]
    L.iamfb_batch_submit_host.argtypes = [vp, C.POINTER(Io), C.c_int]
    L.iamfb_batch_flush_device.argtypes = [vp, vp, vp]
    L.iamfb_batch_flush_host.argtypes = [vp, vp, vp]
    L.iamfb_host_alloc.argtypes = [C.c_size_t]
    L.iamfb_host_alloc.restype = vp
    L.iamfb_host_free.argtypes = [vp]
    L.iamfb_host_free.restype = None
    L.iamfb_ctx_launch_count.argtypes = [vp]
    L.iamfb_ctx_launch_count.restype = C.c_uint64
    L.iamfb_ctx_set_timing.argtypes = [vp, C.c_int]
    L.iamfb_ctx_get_timing.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double),
                                       C.POINTER(C.c_uint64)]
    L.iamfb_get_hrir.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int16)]
    L.iamfb_hrir_taps.restype = C.c_int
    L.iamfb_ctx_get_timing_median.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.iamfb_target_channels.argtypes = [C.c_int]
    L.iamfb_layout_channels.argtypes = [C.c_int, C.POINTER(C.c_int32)]
    L.iamfb_get_m2m_matrix.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_float)]
    L.iamfb_get_h2m_matrix.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
    _lib = L
    return L


def _check(r, what):
    if r != 0:
        raise IamfB200Error(f"{what} failed ({r}): {lib().iamfb_last_error().decode()}")


def get_m2m_matrix(layout, target):
    m, n = C.c_int32(), C.c_int32()
    buf = (C.c_float * (24 * 16))()
    if lib().iamfb_get_m2m_matrix(layout, target, C.byref(m), C.byref(n), buf) != 0:
        return None
    return np.array(buf[: m.value * n.value], np.float32).reshape(m.value, n.value)


def get_h2m_matrix(order, target):
    m, n, l1, l2 = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    buf = (C.c_float * (24 * 16))()
    if lib().iamfb_get_h2m_matrix(order, target, C.byref(m), C.byref(n), C.byref(l1), C.byref(l2), buf) != 0:
        return None
    return np.array(buf[: m.value * n.value], np.float32).reshape(n.value, m.value), l1.value, l2.value


def layout_channels(layout):
    """IAChannel ids of a layout in rendering order (IAMF_utils.c:117-133)"""
    chs = (C.c_int32 * MAXL)()
    n = lib().iamfb_layout_channels(layout, chs)
    return [int(chs[i]) for i in range(n)]


def get_hrir(kind, index):
    """[2 ears][taps] Q15 int16 of IAChannel `index` (kind 0) or ambisonics channel `index` (kind 1)"""
    n = lib().iamfb_hrir_taps()
    buf = (C.c_int16 * (2 * n))()
    if lib().iamfb_get_hrir(kind, index, buf) != 0:
        return None
    return np.array(buf[:], np.int16).reshape(2, n)


def channel_element(layout, chs_in, out_gain=(), demix=None, first_layer_layout=None, selected_layer=0,
                    recon_present=False, dmr_out_layout=None):
    """iamfb_element_desc for a channel-based element. out_gain: [(IAChannel, linear gain)]; demix: (mode, w_idx)"""
    e = ElementDesc()
    e.kind = 0
    e.n_in = len(chs_in)
    e.layout = layout
    for i, c in enumerate(chs_in):
        e.chs_in[i] = c
    e.n_out_gain = len(out_gain)
    for i, (c, g) in enumerate(out_gain):
        e.out_gain_ch[i] = c
        e.out_gain[i] = g
    if demix is not None:
        e.has_demix_info, e.default_mode, e.default_w_idx = 1, demix[0], demix[1]
    else:
        e.default_mode, e.default_w_idx = -1, -1
    e.first_layer_layout = layout if first_layer_layout is None else first_layer_layout
    e.selected_layer = selected_layer
    e.recon_present = 1 if recon_present else 0
    if dmr_out_layout is not None:
        e.use_dmr, e.dmr_out_layout = 1, dmr_out_layout
    return e


def scene_element(channels, n_in=None, mapping=None, projection=None):
    """projection: float32 [cols][rows] (rows = channels)"""
    e = ElementDesc()
    e.kind = 1
    e.ambi_channels = channels
    if projection is None:
        mapping = list(range(channels)) if mapping is None else list(mapping)
        e.ambi_mode = 0
        e.n_in = (max(mapping) + 1) if n_in is None else n_in
        for i, m in enumerate(mapping):
            e.ambi_map[i] = m
    else:
        projection = np.ascontiguousarray(projection, np.float32)
        e.ambi_mode = 1
        e.ambi_cols = projection.shape[0]
        e.n_in = projection.shape[0] if n_in is None else n_in
        flat = projection.reshape(-1)
        for i, v in enumerate(flat):
            e.ambi_matrix[i] = v
    return e


class Engine:
    """One context + plan + batch.  Host-resident submits take numpy arrays; device-resident submits take raw device
    pointers (ints), e.g. torch.Tensor.data_ptr()."""

    def __init__(self, desc: PlanDesc, n_streams: int, max_frames: int, device: int = 0, cuda_stream=None):
        L = lib()
        self.L, self.desc, self.S, self.Fmax = L, desc, n_streams, max_frames
        self.ctx, self.plan, self.batch = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(L.iamfb_ctx_create(device, C.byref(self.ctx)), "iamfb_ctx_create")
        if cuda_stream is not None:
            _check(L.iamfb_ctx_set_stream(self.ctx, C.c_void_p(cuda_stream)), "iamfb_ctx_set_stream")
        _check(L.iamfb_plan_create(self.ctx, C.byref(desc), C.byref(self.plan)), "iamfb_plan_create")
        _check(L.iamfb_batch_create(self.plan, n_streams, max_frames, C.byref(self.batch)), "iamfb_batch_create")
        self.out_channels = L.iamfb_plan_out_channels(self.plan)
        self.kernel_path = L.iamfb_plan_kernel_path(self.plan)   # 0 multi-kernel, 1 k_fused, 2 k_stream, 3 k_pipe (float32 submits)
        self.kernel_path_s16 = L.iamfb_plan_kernel_path_fmt(self.plan, 1)   # the same for int16 submits
        self.arithmetic = L.iamfb_plan_arithmetic(self.plan)                # 1: an IAMFB_ARITH_FMA kernel variant serves the plan
        self.bytes_per_sample = desc.bit_depth // 8 if desc.bit_depth else 4

    def close(self):
        if self.batch:
            self.L.iamfb_batch_destroy(self.batch)
            self.L.iamfb_plan_destroy(self.plan)
            self.L.iamfb_ctx_destroy(self.ctx)
            self.batch = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def out_stride_bytes(self, n_frames):
        return self.L.iamfb_plan_out_stride_bytes(self.plan, n_frames)

    def max_out_samples(self, n_frames):
        return self.L.iamfb_plan_max_out_samples(self.plan, n_frames)

    def launch_count(self):
        return self.L.iamfb_ctx_launch_count(self.ctx)

    def set_timing(self, enable):
        _check(self.L.iamfb_ctx_set_timing(self.ctx, 1 if enable else 0), "iamfb_ctx_set_timing")

    def get_timing(self):
        """{kernel name: (total ms, launches, median ms per launch)} measured with CUDA events on the launching stream"""
        out, i = {}, 0
        while True:
            name, ms, n, med = C.c_char_p(), C.c_double(), C.c_uint64(), C.c_double()
            if self.L.iamfb_ctx_get_timing(self.ctx, i, C.byref(name), C.byref(ms), C.byref(n)) != 0:
                return out
            self.L.iamfb_ctx_get_timing_median(self.ctx, i, C.byref(med))
            out[name.value.decode()] = (ms.value, n.value, med.value)
            i += 1

    def synchronize(self):
        _check(self.L.iamfb_ctx_synchronize(self.ctx), "iamfb_ctx_synchronize")

    def reset(self):
        _check(self.L.iamfb_batch_reset(self.batch), "iamfb_batch_reset")

    # ---- host-resident ----
    def submit_host(self, inputs, params, gain_ramps=None, out_gain_ramp=None):
        """inputs[e]: float32 [S][F][n_in][N]; params: frame_params_array [S][F].
        returns (pcm bytes ndarray [S][stride], counts int32 [S][F])"""
        F = params.shape[1]
        io = Io()
        keep = []
        s16 = all(np.asarray(x).dtype == np.int16 for x in inputs)     # int16 in -> scaled by 1/32768 on the device
        io.in_format = 1 if s16 else 0
        for e, x in enumerate(inputs):
            x = np.ascontiguousarray(x, np.int16 if s16 else np.float32)
            keep.append(x)
            io.in_[e] = x.ctypes.data
        if gain_ramps:
            for e, g in enumerate(gain_ramps):
                if g is not None:
                    g = np.ascontiguousarray(g, np.float32)
                    keep.append(g)
                    io.gain_ramp[e] = g.ctypes.data
        if out_gain_ramp is not None:
            g = np.ascontiguousarray(out_gain_ramp, np.float32)
            keep.append(g)
            io.out_gain_ramp = g.ctypes.data
        params = np.ascontiguousarray(params)
        io.params = params.ctypes.data
        stride = self.out_stride_bytes(F)
        pcm = np.zeros((self.S, stride), np.uint8)
        counts = np.zeros((self.S, F), np.int32)
        io.pcm = pcm.ctypes.data
        io.out_counts = counts.ctypes.data
        _check(self.L.iamfb_batch_submit_host(self.batch, C.byref(io), F), "iamfb_batch_submit_host")
        return pcm, counts

    def flush_host(self):
        stride = self.out_stride_bytes(1)
        pcm = np.zeros((self.S, stride), np.uint8)
        counts = np.zeros((self.S,), np.int32)
        _check(self.L.iamfb_batch_flush_host(self.batch, pcm.ctypes.data, counts.ctypes.data), "iamfb_batch_flush_host")
        return pcm, counts

    def decode_pcm(self, pcm_row, n_samples):
        """view one stream's bytes as [n_samples][channels] of the plan's sample type"""
        co, bd = self.out_channels, self.desc.bit_depth
        nbytes = n_samples * co * self.bytes_per_sample
        raw = pcm_row[:nbytes]
        if bd == 16:
            return raw.view(np.int16).reshape(n_samples, co)
        if bd == 32:
            return raw.view(np.int32).reshape(n_samples, co)
        if bd == 24:
            return raw.reshape(n_samples, co, 3)
        return raw.view(np.float32).reshape(n_samples, co)

    # ---- device-resident (raw pointers) ----
    def submit_device(self, in_ptrs, params_ptr, pcm_ptr, counts_ptr, n_frames, gain_ramp_ptrs=None,
                      out_gain_ramp_ptr=None, in_format=0):
        io = Io()
        io.in_format = in_format
        for e, p in enumerate(in_ptrs):
            io.in_[e] = p
        if gain_ramp_ptrs:
            for e, p in enumerate(gain_ramp_ptrs):
                io.gain_ramp[e] = p
        io.out_gain_ramp = out_gain_ramp_ptr
        io.params = params_ptr
        io.pcm = pcm_ptr
        io.out_counts = counts_ptr
        _check(self.L.iamfb_batch_submit_device(self.batch, C.byref(io), n_frames), "iamfb_batch_submit_device")

    def flush_device(self, pcm_ptr, counts_ptr):
        _check(self.L.iamfb_batch_flush_device(self.batch, pcm_ptr, counts_ptr), "iamfb_batch_flush_device")
