"""Drives a library exporting include/IAMF_decoder.h (the compiled reference, or our drop-in libiamf.so) through the
public API exactly like test/tools/iamfplayer/player/iamfplayer.c:380-650 does.  Test infrastructure."""
import ctypes as C
import os

import numpy as np


class Api:
    def __init__(self, so_path):
        L = C.CDLL(so_path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        vp, u8p = C.c_void_p, C.POINTER(C.c_uint8)
        L.IAMF_decoder_open.restype = vp
        L.IAMF_decoder_close.argtypes = [vp]
        L.IAMF_decoder_configure.argtypes = [vp, C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.IAMF_decoder_decode.argtypes = [vp, C.c_char_p, C.c_int32, C.POINTER(C.c_uint32), vp]
        L.IAMF_decoder_output_layout_set_sound_system.argtypes = [vp, C.c_int]
        L.IAMF_decoder_output_layout_set_binaural.argtypes = [vp]
        L.IAMF_layout_sound_system_channels_count.argtypes = [C.c_int]
        L.IAMF_decoder_set_normalization_loudness.argtypes = [vp, C.c_float]
        L.IAMF_decoder_set_bit_depth.argtypes = [vp, C.c_uint32]
        L.IAMF_decoder_peak_limiter_enable.argtypes = [vp, C.c_uint32]
        L.IAMF_decoder_peak_limiter_set_threshold.argtypes = [vp, C.c_float]
        L.IAMF_decoder_set_sampling_rate.argtypes = [vp, C.c_uint32]
        L.IAMF_decoder_get_stream_info.restype = C.POINTER(C.c_uint32)
        L.IAMF_decoder_get_stream_info.argtypes = [vp]
        L.IAMF_decoder_set_pts.argtypes = [vp, C.c_int64, C.c_uint32]
        self.L = L

    def open_configured(self, blob, sound_system=0, binaural=False, bit_depth=16, rate=0, loudness=0.0, limiter=True,
                        threshold_db=-1.0):
        """a handle taken through the player's set-up calls and IAMF_decoder_configure (blob = descriptors + first unit)"""
        L = self.L
        h = L.IAMF_decoder_open()
        assert h
        L.IAMF_decoder_peak_limiter_set_threshold(h, threshold_db)
        L.IAMF_decoder_set_normalization_loudness(h, loudness)
        L.IAMF_decoder_set_bit_depth(h, bit_depth)
        if not limiter:
            L.IAMF_decoder_peak_limiter_enable(h, 0)
        if rate:
            L.IAMF_decoder_set_sampling_rate(h, rate)
        if binaural:
            L.IAMF_decoder_output_layout_set_binaural(h)
        else:
            L.IAMF_decoder_output_layout_set_sound_system(h, sound_system)
        rsize = C.c_uint32(0)
        L.IAMF_decoder_set_pts(h, 0, 90000)
        ret = L.IAMF_decoder_configure(h, blob, len(blob), C.byref(rsize))
        assert ret == 0, f"configure -> {ret}"
        return h

    def render(self, descriptors, units, sound_system=0, binaural=False, bit_depth=16, rate=0, loudness=0.0,
               limiter=True, threshold_db=-1.0, flush=True):
        """returns (pcm ndarray [samples][channels], per-call sample counts)"""
        L = self.L
        h = L.IAMF_decoder_open()
        assert h
        # iamfplayer.c:402-428
        L.IAMF_decoder_peak_limiter_set_threshold(h, threshold_db)
        L.IAMF_decoder_set_normalization_loudness(h, loudness)
        L.IAMF_decoder_set_bit_depth(h, bit_depth)
        if not limiter:
            L.IAMF_decoder_peak_limiter_enable(h, 0)
        if rate:
            L.IAMF_decoder_set_sampling_rate(h, rate)
        if binaural:
            L.IAMF_decoder_output_layout_set_binaural(h)
            channels = 2
        else:
            L.IAMF_decoder_output_layout_set_sound_system(h, sound_system)
            channels = L.IAMF_layout_sound_system_channels_count(sound_system)
        rsize = C.c_uint32(0)
        L.IAMF_decoder_set_pts(h, 0, 90000)
        # like iamfplayer.c:574: the block handed to configure runs past the descriptors (the first non-descriptor
        # OBU is what tells the decoder that the descriptor set is complete, IAMF_decoder.c:2829-2833)
        blob = descriptors + (units[0] if units else b"")
        ret = L.IAMF_decoder_configure(h, blob, len(blob), C.byref(rsize))
        assert ret == 0, f"configure -> {ret}"
        assert rsize.value == len(descriptors), (rsize.value, len(descriptors))
        info = L.IAMF_decoder_get_stream_info(h)
        max_frame = info[0]
        bps = bit_depth // 8
        buf = C.create_string_buffer(bps * max_frame * channels * 2)
        chunks, counts = [], []
        for u in units:
            ret = L.IAMF_decoder_decode(h, u, len(u), C.byref(rsize), buf)
            counts.append(ret)
            if ret > 0:
                chunks.append(buf.raw[: ret * channels * bps])
        if flush:
            ret = L.IAMF_decoder_decode(h, None, 0, C.byref(rsize), buf)
            counts.append(ret)
            if ret > 0:
                chunks.append(buf.raw[: ret * channels * bps])
        L.IAMF_decoder_close(h)
        raw = b"".join(chunks)
        if bit_depth == 16:
            pcm = np.frombuffer(raw, np.int16).reshape(-1, channels)
        elif bit_depth == 32:
            pcm = np.frombuffer(raw, np.int32).reshape(-1, channels)
        else:
            pcm = np.frombuffer(raw, np.uint8).reshape(-1, channels, 3)
        return pcm, counts
