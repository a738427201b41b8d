"""ctypes bindings of the stage functions exported by the compiled, unmodified reference (oracle/_ref/libiamf_ref.so).

Test infrastructure only.  The library is built by `make -C oracle ref` in the authoring container (where
/root/reference exists) and travels to the GPU box as a built artefact; when it is absent every user of this module
skips.  Struct layouts mirror src/iamf_dec/ae_rdr.h:98-151 and audio_effect_peak_limiter.h:48-73.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "..", "oracle", "_ref", "libiamf_ref.so")

f32p = C.POINTER(C.c_float)


def have_ref():
    return os.path.exists(REF_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(REF_SO, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        _proto(_lib)
    return _lib


class LfeFilter(C.Structure):
    _fields_ = [("init", C.c_int), ("c", C.c_float), ("a1", C.c_float), ("a2", C.c_float), ("a3", C.c_float),
                ("b1", C.c_float), ("b2", C.c_float), ("ih", C.c_float * 2), ("oh", C.c_float * 2)]


class PredefSp(C.Structure):
    _fields_ = [("system", C.c_int), ("lfe1", C.c_int), ("lfe2", C.c_int)]


class SpLayout(C.Structure):
    _fields_ = [("sp_type", C.c_int), ("predefined_sp", C.POINTER(PredefSp)), ("lfe_f", LfeFilter)]


class HoaLayout(C.Structure):
    _fields_ = [("order", C.c_int), ("lfe_on", C.c_int)]


class M2M(C.Structure):
    _fields_ = [("in_", C.c_int), ("out", C.c_int), ("mat", f32p), ("m", C.c_int), ("n", C.c_int)]


class H2M(C.Structure):
    _fields_ = [("in_", C.c_int), ("out", C.c_int), ("channels", C.c_int), ("lfe1", C.c_int), ("lfe2", C.c_int),
                ("mat", f32p), ("m", C.c_int), ("n", C.c_int)]


# IAMF_SOUND_SYSTEM ids, ae_rdr.h:40-61
SS_IDS = {"A": 0x020, "B": 0x050, "C": 0x250, "D": 0x450, "E": 0x451, "F": 0x370, "G": 0x490, "H": 0x9A3,
          "I": 0x070, "J": 0x470, "712": 0x712, "312": 0x312, "MONO": 0x100, "BINAURAL": 0x1020}
LAYER_IDS = [0x100, 0x200, 0x510, 0x512, 0x514, 0x710, 0x712, 0x714, 0x312, 0x1020]  # IAMF_decoder.c:255-260


def _proto(L):
    vp = C.c_void_p
    L.demixer_open.restype = vp
    L.demixer_open.argtypes = [C.c_uint32]
    L.demixer_close.argtypes = [vp]
    L.demixer_set_channel_layout.argtypes = [vp, C.c_int]
    L.demixer_set_channels_order.argtypes = [vp, C.POINTER(C.c_int), C.c_int]
    L.demixer_set_output_gain.argtypes = [vp, C.POINTER(C.c_int), f32p, C.c_int]
    L.demixer_set_demixing_info.argtypes = [vp, C.c_int, C.c_int]
    L.demixer_set_recon_gain.argtypes = [vp, C.c_int, C.POINTER(C.c_int), f32p, C.c_uint32]
    L.demixer_set_frame_offset.argtypes = [vp, C.c_uint32]
    L.demixer_demixing.argtypes = [vp, f32p, f32p, C.c_uint32]
    L.DMRenderer_open.restype = vp
    L.DMRenderer_open.argtypes = [C.c_int, C.c_int]
    L.DMRenderer_close.argtypes = [vp]
    L.DMRenderer_set_mode_weight.argtypes = [vp, C.c_int, C.c_int]
    L.DMRenderer_downmix.argtypes = [vp, f32p, f32p, C.c_uint32, C.c_uint32, C.c_uint32]
    L.IAMF_element_renderer_get_M2M_matrix.argtypes = [C.POINTER(SpLayout), C.POINTER(SpLayout), C.POINTER(M2M)]
    L.IAMF_element_renderer_render_M2M.argtypes = [C.POINTER(M2M), C.POINTER(f32p), C.POINTER(f32p), C.c_int]
    L.IAMF_element_renderer_get_H2M_matrix.argtypes = [C.POINTER(HoaLayout), C.POINTER(PredefSp), C.POINTER(H2M)]
    L.IAMF_element_renderer_render_H2M.argtypes = [C.POINTER(H2M), C.POINTER(f32p), C.POINTER(f32p), C.c_int, vp]
    L.audio_effect_peak_limiter_create.restype = vp
    L.audio_effect_peak_limiter_init.argtypes = [vp, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
    L.audio_effect_peak_limiter_process_block.argtypes = [vp, f32p, f32p, C.c_int]
    L.audio_effect_peak_limiter_destroy.argtypes = [vp]
    L.speex_resampler_init.restype = vp
    L.speex_resampler_init.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_int)]
    L.speex_resampler_skip_zeros.argtypes = [vp]
    L.speex_resampler_destroy.argtypes = [vp]
    L.speex_resampler_get_output_latency.argtypes = [vp]
    L.speex_resampler_get_input_latency.argtypes = [vp]
    L.speex_resampler_process_interleaved_float.argtypes = [vp, f32p, C.POINTER(C.c_uint32), f32p,
                                                            C.POINTER(C.c_uint32)]
    L.db2lin.restype = C.c_float
    L.db2lin.argtypes = [C.c_float]
    L.qf_to_float.restype = C.c_float
    L.qf_to_float.argtypes = [C.c_uint8, C.c_int]
    L.q_to_float.restype = C.c_float
    L.q_to_float.argtypes = [C.c_int16, C.c_int]


def fp(a):
    return a.ctypes.data_as(f32p)


def ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def m2m_matrix(in_id, out_id):
    """(m, n, mat[m][n]) of the reference's static table, or None.  m2m_rdr.c:1786-1804"""
    L = lib()
    pi, po = PredefSp(in_id, 0, 0), PredefSp(out_id, 0, 0)
    li, lo = SpLayout(0, C.pointer(pi)), SpLayout(0, C.pointer(po))
    m = M2M()
    if L.IAMF_element_renderer_get_M2M_matrix(C.byref(li), C.byref(lo), C.byref(m)) != 0:
        return None
    mat = np.ctypeslib.as_array(m.mat, shape=(m.m * m.n,)).copy().reshape(m.m, m.n)
    return m.m, m.n, mat


def h2m_matrix(order, out_id):
    """(m_in, n_out, lfe1, lfe2, mat[n][m]) or None.  h2m_rdr.c:1070-1081"""
    L = lib()
    h = HoaLayout(order, 0)
    po = PredefSp(out_id, 0, 0)
    m = H2M()
    if L.IAMF_element_renderer_get_H2M_matrix(C.byref(h), C.byref(po), C.byref(m)) != 0:
        return None
    mat = np.ctypeslib.as_array(m.mat, shape=(m.m * m.n,)).copy().reshape(m.n, m.m)
    return m.m, m.n, m.lfe1, m.lfe2, mat


def render_m2m(in_id, out_id, x):
    """x: [m][ns] float32 -> [n][ns]"""
    L = lib()
    pi, po = PredefSp(in_id, 0, 0), PredefSp(out_id, 0, 0)
    li, lo = SpLayout(0, C.pointer(pi)), SpLayout(0, C.pointer(po))
    m = M2M()
    assert L.IAMF_element_renderer_get_M2M_matrix(C.byref(li), C.byref(lo), C.byref(m)) == 0
    ns = x.shape[1]
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros((m.n, ns), np.float32)
    pin = (f32p * m.m)(*[fp(x[i]) for i in range(m.m)])
    pout = (f32p * m.n)(*[fp(out[i]) for i in range(m.n)])
    L.IAMF_element_renderer_render_M2M(C.byref(m), pin, pout, ns)
    return out


def render_h2m(order, out_id, x, out_channels):
    L = lib()
    h = HoaLayout(order, 0)
    po = PredefSp(out_id, 0, 0)
    m = H2M()
    assert L.IAMF_element_renderer_get_H2M_matrix(C.byref(h), C.byref(po), C.byref(m)) == 0
    ns = x.shape[1]
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros((out_channels, ns), np.float32)
    pin = (f32p * m.m)(*[fp(x[i]) for i in range(m.m)])
    pout = (f32p * out_channels)(*[fp(out[i]) for i in range(out_channels)])
    L.IAMF_element_renderer_render_H2M(C.byref(m), pin, pout, ns, None)
    return out


class RefDemixer:
    def __init__(self, frame_size, layout, chs_in, gain_chs=(), gains=(), mode=None, w_idx=None):
        L = lib()
        self.L, self.n, self.nch = L, frame_size, len(chs_in)
        self.h = L.demixer_open(frame_size)
        L.demixer_set_channel_layout(self.h, layout)
        a = np.asarray(chs_in, np.int32)
        L.demixer_set_channels_order(self.h, ip(a), len(a))
        g = np.asarray(gain_chs, np.int32)
        gv = np.asarray(gains, np.float32)
        L.demixer_set_output_gain(self.h, ip(g), fp(gv), len(g))
        if mode is not None:
            L.demixer_set_demixing_info(self.h, mode, w_idx)

    def set_recon(self, chs, gains, flags):
        c = np.asarray(chs, np.int32)
        g = np.asarray(gains, np.float32)
        self.L.demixer_set_recon_gain(self.h, len(c), ip(c), fp(g), flags)

    def set_mode(self, mode, w_idx=-1):
        return self.L.demixer_set_demixing_info(self.h, mode, w_idx)

    def set_offset(self, off):
        self.L.demixer_set_frame_offset(self.h, off)

    def demix(self, x):
        x = np.ascontiguousarray(x, np.float32).copy()
        out = np.zeros((self.nch, self.n), np.float32)
        r = self.L.demixer_demixing(self.h, fp(out), fp(x), self.n)
        return r, out

    def close(self):
        self.L.demixer_close(self.h)


class RefDmr:
    def __init__(self, lin, lout):
        self.L = lib()
        self.h = self.L.DMRenderer_open(lin, lout)

    def ok(self):
        return bool(self.h)

    def set_mode_weight(self, mode, w):
        return self.L.DMRenderer_set_mode_weight(self.h, mode, w)

    def downmix(self, x, nout, s=0, dur=None):
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[1]
        out = np.zeros((nout, n), np.float32)
        self.L.DMRenderer_downmix(self.h, fp(x), fp(out), s, n if dur is None else dur, n)
        return out

    def close(self):
        if self.h:
            self.L.DMRenderer_close(self.h)


class RefLimiter:
    def __init__(self, thr_db, rate, ch, atk=0.001, rel=0.2, delay=240):
        self.L = lib()
        self.h = self.L.audio_effect_peak_limiter_create()
        self.ch = ch
        self.L.audio_effect_peak_limiter_init(self.h, thr_db, rate, ch, atk, rel, delay)

    def process(self, x):
        """x [ch][n] -> [ch][ret] (planar, compacted like the reference)"""
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[1]
        out = np.zeros_like(x)
        r = self.L.audio_effect_peak_limiter_process_block(self.h, fp(x), fp(out), n)
        return out.reshape(-1)[: r * self.ch].reshape(self.ch, r).copy()

    def close(self):
        self.L.audio_effect_peak_limiter_destroy(self.h)


class RefResampler:
    """speex resampler driven exactly as iamf_resample does (IAMF_decoder.c:3223-3248)."""

    def __init__(self, ch, in_rate, out_rate, quality=4):
        self.L = lib()
        err = C.c_int(0)
        self.h = self.L.speex_resampler_init(ch, in_rate, out_rate, quality, C.byref(err))
        assert err.value == 0
        self.L.speex_resampler_skip_zeros(self.h)
        self.ch, self.in_rate, self.out_rate = ch, in_rate, out_rate

    def process(self, x):
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[1]
        inter = np.ascontiguousarray(x.T)
        cap = n * (self.out_rate // self.in_rate + 1)
        out = np.zeros((cap, self.ch), np.float32)
        il, ol = C.c_uint32(n), C.c_uint32(cap)
        self.L.speex_resampler_process_interleaved_float(self.h, fp(inter), C.byref(il), fp(out), C.byref(ol))
        return np.ascontiguousarray(out[: ol.value].T)

    def flush(self):
        ol = C.c_uint32(self.L.speex_resampler_get_output_latency(self.h))
        il = C.c_uint32(self.L.speex_resampler_get_input_latency(self.h))
        out = np.zeros((ol.value, self.ch), np.float32)
        self.L.speex_resampler_process_interleaved_float(self.h, None, C.byref(il), fp(out), C.byref(ol))
        return np.ascontiguousarray(out[: ol.value].T)

    def close(self):
        self.L.speex_resampler_destroy(self.h)
