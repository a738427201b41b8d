"""ctypes bindings of oracle/liboracle.so (our plain-C restatement of the reference path).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(HERE, "..", "oracle")
ORC_SO = os.path.join(ORACLE_DIR, "liboracle.so")
f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)

MAXL = 12


class ElementCfg(C.Structure):
    _fields_ = [("type", C.c_int), ("n_in", C.c_int), ("layout", C.c_int), ("chs_in", C.c_int * MAXL),
                ("n_out_gain", C.c_int), ("out_gain_ch", C.c_int * MAXL), ("out_gain", C.c_float * MAXL),
                ("has_demix_info", C.c_int), ("default_mode", C.c_int), ("default_w_idx", C.c_int),
                ("first_layer_layout", C.c_int), ("selected_layer", C.c_int),
                ("use_dmr", C.c_int), ("dmr_out_layout", C.c_int),
                ("ambi_mode", C.c_int), ("ambi_map", C.c_uint8 * 16), ("ambi_matrix", f32p), ("ambi_cols", C.c_int),
                ("mat", f32p), ("mat_in", C.c_int), ("mat_out", C.c_int), ("lfe1", C.c_int), ("lfe2", C.c_int),
                ("hrtf_taps", C.POINTER(C.c_int16))]


class StreamCfg(C.Structure):
    _fields_ = [("frame_size", C.c_int), ("in_rate", C.c_int), ("out_rate", C.c_int), ("n_elements", C.c_int),
                ("el", ElementCfg * 2), ("out_channels", C.c_int), ("loudness_gain", C.c_float),
                ("limiter", C.c_int), ("limiter_threshold_db", C.c_float), ("bit_depth", C.c_int)]


class FrameParams(C.Structure):
    _fields_ = [("dmx_mode", C.c_int), ("n_recon", C.c_int), ("recon_ch", C.c_int * MAXL),
                ("recon_gain", C.c_float * MAXL), ("recon_flags", C.c_uint32), ("has_recon", C.c_int),
                ("gain_const", C.c_float), ("gain_ramp", f32p)]


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORC_SO):
            build()
        L = C.CDLL(ORC_SO, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        vp = C.c_void_p
        L.orc_demixer_open.restype = vp
        L.orc_demixer_open.argtypes = [C.c_int]
        L.orc_demixer_close.argtypes = [vp]
        L.orc_demixer_set_frame_offset.argtypes = [vp, C.c_uint32]
        L.orc_demixer_set_layout.argtypes = [vp, C.c_int]
        L.orc_demixer_set_channels_order.argtypes = [vp, i32p, C.c_int]
        L.orc_demixer_set_output_gain.argtypes = [vp, i32p, f32p, C.c_int]
        L.orc_demixer_set_demixing_info.argtypes = [vp, C.c_int, C.c_int]
        L.orc_demixer_set_recon_gain.argtypes = [vp, C.c_int, i32p, f32p, C.c_uint32]
        L.orc_demixer_demix.argtypes = [vp, f32p, f32p, C.c_uint32]
        L.orc_dmr_open.restype = vp
        L.orc_dmr_open.argtypes = [C.c_int, C.c_int]
        L.orc_dmr_close.argtypes = [vp]
        L.orc_dmr_set_mode_weight.argtypes = [vp, C.c_int, C.c_int]
        L.orc_dmr_downmix.argtypes = [vp, f32p, f32p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_render_m2m.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p, C.c_int]
        L.orc_render_h2m.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, f32p, C.c_int]
        L.orc_ambisonics_projection.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p, C.c_int]
        L.orc_gain_linear.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, C.c_uint32, f32p]
        L.orc_gain_bezier.argtypes = [C.c_float, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, C.c_uint32, f32p]
        L.orc_plane2stride.argtypes = [vp, f32p, C.c_int, C.c_int, C.c_uint32, C.c_uint32]
        L.orc_resampler_open.restype = vp
        L.orc_resampler_open.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_resampler_close.argtypes = [vp]
        L.orc_resample.argtypes = [vp, f32p, f32p, C.c_int]
        L.orc_limiter_new.restype = vp
        L.orc_limiter_new.argtypes = [C.c_float, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
        L.orc_limiter_free.argtypes = [vp]
        L.orc_limiter_process.argtypes = [vp, f32p, f32p, C.c_int]
        L.orc_stream_open.restype = vp
        L.orc_stream_open.argtypes = [C.POINTER(StreamCfg)]
        L.orc_stream_close.argtypes = [vp]
        L.orc_stream_decode.argtypes = [vp, C.POINTER(f32p), C.POINTER(FrameParams), C.c_float, f32p, C.c_int,
                                        C.c_int, vp]
        L.orc_stream_flush.argtypes = [vp, vp]
        L.orc_db2lin.restype = C.c_float
        L.orc_db2lin.argtypes = [C.c_float]
        L.orc_qf_to_float.restype = C.c_float
        L.orc_qf_to_float.argtypes = [C.c_uint8, C.c_int]
        L.orc_q_to_float.restype = C.c_float
        L.orc_q_to_float.argtypes = [C.c_int16, C.c_int]
        L.orc_layout_channels.argtypes = [C.c_int, i32p]
        L.orc_layer_channels.argtypes = [C.c_int, i32p]
        L.orc_new_channels.argtypes = [C.c_int, C.c_int, i32p]
        L.orc_recon_flags.restype = C.c_uint32
        L.orc_recon_order.argtypes = [C.c_int, C.c_uint32, i32p]
        _lib = L
    return _lib


def fp(a):
    return a.ctypes.data_as(f32p)


def ip(a):
    return a.ctypes.data_as(i32p)


class OrcDemixer:
    def __init__(self, frame_size, layout, chs_in, gain_chs=(), gains=(), mode=None, w_idx=None):
        L = lib()
        self.L, self.n, self.nch = L, frame_size, len(chs_in)
        self.h = L.orc_demixer_open(frame_size)
        L.orc_demixer_set_layout(self.h, layout)
        a = np.asarray(chs_in, np.int32)
        L.orc_demixer_set_channels_order(self.h, ip(a), len(a))
        g = np.asarray(gain_chs, np.int32)
        gv = np.asarray(gains, np.float32)
        L.orc_demixer_set_output_gain(self.h, ip(g), fp(gv), len(g))
        if mode is not None:
            L.orc_demixer_set_demixing_info(self.h, mode, w_idx)

    def set_recon(self, chs, gains, flags):
        c = np.asarray(chs, np.int32)
        g = np.asarray(gains, np.float32)
        self.L.orc_demixer_set_recon_gain(self.h, len(c), ip(c), fp(g), flags)

    def set_mode(self, mode, w_idx=-1):
        return self.L.orc_demixer_set_demixing_info(self.h, mode, w_idx)

    def set_offset(self, off):
        self.L.orc_demixer_set_frame_offset(self.h, off)

    def demix(self, x):
        x = np.ascontiguousarray(x, np.float32).copy()
        out = np.zeros((self.nch, self.n), np.float32)
        r = self.L.orc_demixer_demix(self.h, fp(out), fp(x), self.n)
        return r, out

    def close(self):
        self.L.orc_demixer_close(self.h)


class OrcDmr:
    def __init__(self, lin, lout):
        self.L = lib()
        self.h = self.L.orc_dmr_open(lin, lout)

    def ok(self):
        return bool(self.h)

    def set_mode_weight(self, mode, w):
        return self.L.orc_dmr_set_mode_weight(self.h, mode, w)

    def downmix(self, x, nout, s=0, dur=None):
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[1]
        out = np.zeros((nout, n), np.float32)
        self.L.orc_dmr_downmix(self.h, fp(x), fp(out), s, n if dur is None else dur, n)
        return out

    def close(self):
        if self.h:
            self.L.orc_dmr_close(self.h)


def render_m2m(mat, x):
    mat = np.ascontiguousarray(mat, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    m, n = mat.shape
    out = np.zeros((n, x.shape[1]), np.float32)
    lib().orc_render_m2m(fp(mat), m, n, fp(x), fp(out), x.shape[1])
    return out


def render_h2m(mat, lfe1, lfe2, x, out_channels):
    mat = np.ascontiguousarray(mat, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    n, m = mat.shape
    out = np.zeros((out_channels, x.shape[1]), np.float32)
    lib().orc_render_h2m(fp(mat), m, n, lfe1, lfe2, fp(x), fp(out), x.shape[1])
    return out


class OrcLimiter:
    def __init__(self, thr_db, rate, ch, atk=0.001, rel=0.2, delay=240):
        self.L = lib()
        self.ch = ch
        self.h = self.L.orc_limiter_new(thr_db, rate, ch, atk, rel, delay)

    def process(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros_like(x)
        r = self.L.orc_limiter_process(self.h, fp(x), fp(out), x.shape[1])
        return out.reshape(-1)[: r * self.ch].reshape(self.ch, r).copy()

    def close(self):
        self.L.orc_limiter_free(self.h)


class OrcResampler:
    def __init__(self, ch, in_rate, out_rate, quality=4):
        self.L = lib()
        self.ch, self.in_rate, self.out_rate = ch, in_rate, out_rate
        self.h = self.L.orc_resampler_open(ch, in_rate, out_rate, quality)

    def process(self, x):
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[1]
        cap = n * (self.out_rate // self.in_rate + 1)
        out = np.zeros(self.ch * cap, np.float32)
        r = self.L.orc_resample(self.h, fp(x), fp(out), n)
        return out[: r * self.ch].reshape(self.ch, r).copy()

    def flush(self):
        out = np.zeros(self.ch * 4096, np.float32)
        r = self.L.orc_resample(self.h, None, fp(out), -1)
        return out[: r * self.ch].reshape(self.ch, r).copy()

    def close(self):
        self.L.orc_resampler_close(self.h)


def plane2stride(x, bit_depth):
    """x [ch][n] -> interleaved ints (16: int16 [n][ch]; 24: uint8 [n][ch][3]; 32: int32 [n][ch])"""
    x = np.ascontiguousarray(x, np.float32)
    ch, n = x.shape
    if bit_depth == 16:
        out = np.zeros((n, ch), np.int16)
    elif bit_depth == 24:
        out = np.zeros((n, ch, 3), np.uint8)
    elif bit_depth == 32:
        out = np.zeros((n, ch), np.int32)
    else:
        out = np.zeros((n, ch), np.float32)
    lib().orc_plane2stride(out.ctypes.data_as(C.c_void_p), fp(x), n, ch, bit_depth, ch)
    return out
