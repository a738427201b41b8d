"""Synthetic IAMF bitstream writer (test infrastructure).

Writes the early (pre-v1.0) OBU syntax the reference parser accepts (src/iamf_dec/IAMF_OBU.c:79-138,
260-297, 300-356, 391-607, 641-932, 990-1215, 1227-1254).  ipcm-coded (16-bit little-endian) so that entropy decode
is trivial and bit-exact.  Used to drive BOTH the compiled reference and our drop-in libiamf.so through the public
IAMF_decoder.h API with identical bytes.
"""
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

OBU_CODEC_CONFIG, OBU_AUDIO_ELEMENT, OBU_MIX_PRESENTATION, OBU_PARAMETER_BLOCK = 0, 1, 2, 3
OBU_TEMPORAL_DELIMITER, OBU_AUDIO_FRAME, OBU_AUDIO_FRAME_ID0, OBU_SEQUENCE_HEADER = 4, 5, 6, 31
PARAM_MIX_GAIN, PARAM_DEMIXING, PARAM_RECON_GAIN = 0, 1, 2

# IAChannelLayoutType
MONO, STEREO, L510, L512, L514, L710, L712, L714, L312, BINAURAL = range(10)
LAYOUT_CH = [1, 2, 6, 8, 10, 8, 10, 12, 6, 2]
LAYOUT_S = [1, 2, 5, 5, 5, 7, 7, 7, 3, 2]
LAYOUT_T = [0, 0, 0, 2, 4, 0, 2, 4, 2, 0]


def leb128(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def obu(obu_type: int, payload: bytes, trim_start: int = 0, trim_end: int = 0, redundant: int = 0) -> bytes:
    trimming = 1 if (trim_start or trim_end) else 0
    head = bytes([(obu_type << 3) | (redundant << 2) | (trimming << 1)])
    body = b""
    if trimming:
        body += leb128(trim_end) + leb128(trim_start)
    body += payload
    return head + leb128(len(body)) + body


def s16be(v: int) -> bytes:
    return struct.pack(">h", v)


def param_base(pid: int, rate: int, duration: int) -> bytes:
    """param definition with mode 0: one constant segment spanning `duration` per parameter block"""
    return leb128(pid) + leb128(rate) + bytes([0 << 7]) + leb128(duration) + leb128(duration)


@dataclass
class Layer:
    layout: int
    n_substreams: int
    n_coupled: int
    out_gain_flags: int = 0       # 6-bit flags (L R LS RS LTF RTF from MSB); 0 => no output gain
    out_gain_q78: int = 0
    recon_gain: bool = False


@dataclass
class Element:
    eid: int
    kind: str                      # "channel" | "scene"
    layers: List[Layer] = field(default_factory=list)
    demix: Optional[tuple] = None  # (default_mode, default_w) -> demixing param definition
    # scene
    ambi_mode: int = 0             # 0 mono, 1 projection
    ambi_channels: int = 0
    ambi_mapping: Optional[list] = None       # mono: channel map
    ambi_coupled: int = 0
    ambi_matrix_q15: Optional[np.ndarray] = None  # projection: int16 [cols][rows] (column-major as on the wire)
    mix_gain_q78: int = 0
    headphones_mode: int = 0
    substream_base: int = 0

    @property
    def n_substreams(self):
        if self.kind == "channel":
            return sum(l.n_substreams for l in self.layers)
        if self.ambi_mode == 0:
            return self.ambi_channels if self.ambi_mapping is None else len(set(self.ambi_mapping))
        return self._proj_substreams

    @property
    def n_channels_tx(self):
        if self.kind == "channel":
            return sum(l.n_substreams + l.n_coupled for l in self.layers)
        if self.ambi_mode == 0:
            return self.n_substreams
        return self._proj_substreams + self.ambi_coupled

    _proj_substreams: int = 0

    def demix_pid(self):
        return 100 + self.eid

    def recon_pid(self):
        return 200 + self.eid

    def mixgain_pid(self):
        return 300 + self.eid


class OpusEncoders:
    """libopus encoders (one per sub-stream) from oracle/_ref/libopus_ref.so - the reference tree's own prebuilt libopus"""
    import os as _os
    SO = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "oracle", "_ref", "libopus_ref.so")

    @classmethod
    def available(cls):
        import os
        return os.path.exists(cls.SO)

    def __init__(self, rate, bitrate=128000):
        import ctypes as C
        self.C = C
        self.L = C.CDLL(self.SO)
        self.L.opus_encoder_create.restype = C.c_void_p
        self.L.opus_encoder_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        self.L.opus_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        self.L.opus_encoder_ctl.argtypes = [C.c_void_p, C.c_int, C.c_int]
        self.rate, self.bitrate, self.enc = rate, bitrate, {}

    def encode(self, key, x):
        """x: int16 [channels][n] -> packet bytes"""
        C = self.C
        ch = x.shape[0]
        if key not in self.enc:
            err = C.c_int(0)
            e = self.L.opus_encoder_create(self.rate, ch, 2049, C.byref(err))     # OPUS_APPLICATION_AUDIO
            assert e and err.value == 0
            self.L.opus_encoder_ctl(e, 4002, self.bitrate * ch // 2)              # OPUS_SET_BITRATE
            self.enc[key] = e
        inter = np.ascontiguousarray(x.T).astype(np.int16)
        buf = C.create_string_buffer(4000)
        n = self.L.opus_encode(self.enc[key], inter.ctypes.data, x.shape[1], buf, 4000)
        assert n > 0, n
        return buf.raw[:n]


class FlacEncoder:
    """libFLAC stream encoder from oracle/_ref/libFLAC_ref.so (the reference tree's own prebuilt libFLAC): one FLAC frame
    per call - a fresh encoder per frame, frames are self-contained"""
    import os as _os
    SO = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "oracle", "_ref", "libFLAC_ref.so")

    @classmethod
    def available(cls):
        import os
        return os.path.exists(cls.SO)

    def __init__(self, rate, bits=16):
        import ctypes as C
        self.C, self.rate, self.bits = C, rate, bits
        L = C.CDLL(self.SO)
        L.FLAC__stream_encoder_new.restype = C.c_void_p
        for f in ("set_channels", "set_bits_per_sample", "set_sample_rate", "set_blocksize", "set_compression_level"):
            getattr(L, "FLAC__stream_encoder_" + f).argtypes = [C.c_void_p, C.c_uint32]
        self.WCB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_ubyte), C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p)
        L.FLAC__stream_encoder_init_stream.argtypes = [C.c_void_p, self.WCB, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.FLAC__stream_encoder_process_interleaved.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.FLAC__stream_encoder_finish.argtypes = [C.c_void_p]
        L.FLAC__stream_encoder_delete.argtypes = [C.c_void_p]
        self.L = L

    def encode(self, key, x):
        """x: int [channels][n] -> the bytes of ONE FLAC frame"""
        C, L = self.C, self.L
        ch, n = x.shape
        frames = []

        def cb(enc, buf, nbytes, samples, cur, client):
            if samples:                                   # (samples == 0: stream marker / metadata)
                frames.append(bytes(bytearray(buf[:nbytes])))
            return 0
        wcb = self.WCB(cb)
        e = L.FLAC__stream_encoder_new()
        L.FLAC__stream_encoder_set_channels(e, ch)
        L.FLAC__stream_encoder_set_bits_per_sample(e, self.bits)
        L.FLAC__stream_encoder_set_sample_rate(e, self.rate)
        L.FLAC__stream_encoder_set_blocksize(e, n)
        L.FLAC__stream_encoder_set_compression_level(e, 5)
        assert L.FLAC__stream_encoder_init_stream(e, wcb, None, None, None, None) == 0
        inter = np.ascontiguousarray(np.asarray(x, np.int32).T)
        assert L.FLAC__stream_encoder_process_interleaved(e, inter.ctypes.data, n)
        L.FLAC__stream_encoder_finish(e)
        L.FLAC__stream_encoder_delete(e)
        assert len(frames) == 1, len(frames)
        return frames[0]


@dataclass
class Stream:
    elements: List[Element]
    frame_size: int = 960
    rate: int = 48000
    profile: int = 1
    codec: str = "ipcm"            # "ipcm" (16-bit little-endian) | "opus" | "flac" (need oracle/_ref/lib{opus,FLAC}_ref.so)
    _opus: Optional[object] = None
    _flac: Optional[object] = None
    flac_bits: int = 16
    out_gain_q78: int = 0
    layouts: List[tuple] = field(default_factory=lambda: [("ss", 0, 0)])  # ("ss", sound_system, loudness_q78) | ("bin", 0, q78)
    OUT_GAIN_PID = 400

    # ---------------------------------------------------------------- descriptors
    def descriptors(self) -> bytes:
        out = obu(OBU_SEQUENCE_HEADER, b"iamf" + bytes([self.profile, self.profile]))
        if self.codec == "opus":
            # OpusHead-like decoder config, big endian (opus/IAMF_opus_decoder.c:56-99): version, channels(2), pre-skip,
            # input rate, output gain, mapping family 0
            cc = leb128(0) + b"Opus" + leb128(self.frame_size) + s16be(-4) + bytes([1, 2]) + struct.pack(">HIhB", 0, self.rate, 0, 0)
        elif self.codec == "flac":
            # decoder config = the METADATA_BLOCKs without the "fLaC" marker: one STREAMINFO (last block), stereo - the
            # decoder patches the channel count for mono sub-streams (flac_multistream_decoder.c:124-150,206-217)
            si = struct.pack(">HH", self.frame_size, self.frame_size) + b"\0" * 6
            v = (self.rate << 44) | (1 << 41) | ((self.flac_bits - 1) << 36)
            si += v.to_bytes(8, "big") + b"\0" * 16
            cc = leb128(0) + b"fLaC" + leb128(self.frame_size) + s16be(0) + bytes([0x80, 0, 0, 34]) + si
        else:
            # codec config: id 0, ipcm, 16-bit LE
            cc = leb128(0) + b"ipcm" + leb128(self.frame_size) + s16be(0) + bytes([1, 16]) + struct.pack(">I", self.rate)
        out += obu(OBU_CODEC_CONFIG, cc)
        for e in self.elements:
            out += obu(OBU_AUDIO_ELEMENT, self._element(e))
        out += obu(OBU_MIX_PRESENTATION, self._mix_presentation())
        return out

    def _element(self, e: Element) -> bytes:
        p = leb128(e.eid) + bytes([(0 if e.kind == "channel" else 1) << 5]) + leb128(0)
        p += leb128(e.n_substreams)
        for i in range(e.n_substreams):
            p += leb128(e.substream_base + i)
        params = []
        if e.kind == "channel" and e.demix is not None:
            params.append(leb128(PARAM_DEMIXING) + param_base(e.demix_pid(), self.rate, self.frame_size) +
                          bytes([e.demix[0] << 5, e.demix[1] << 4]))
        if e.kind == "channel" and any(l.recon_gain for l in e.layers):
            params.append(leb128(PARAM_RECON_GAIN) + param_base(e.recon_pid(), self.rate, self.frame_size))
        p += leb128(len(params)) + b"".join(params)
        if e.kind == "channel":
            p += bytes([len(e.layers) << 5])
            for l in e.layers:
                og = 1 if l.out_gain_flags else 0
                p += bytes([(l.layout << 4) | (og << 3) | ((1 if l.recon_gain else 0) << 2), l.n_substreams, l.n_coupled])
                if og:
                    p += bytes([l.out_gain_flags << 2]) + s16be(l.out_gain_q78)
        else:
            p += leb128(e.ambi_mode)
            if e.ambi_mode == 0:
                mapping = e.ambi_mapping if e.ambi_mapping is not None else list(range(e.ambi_channels))
                p += bytes([e.ambi_channels, e.n_substreams]) + bytes(mapping)
            else:
                p += bytes([e.ambi_channels, e._proj_substreams, e.ambi_coupled])
                p += np.ascontiguousarray(e.ambi_matrix_q15, dtype=">i2").tobytes()
        return p

    def _mix_presentation(self) -> bytes:
        p = leb128(42) + leb128(0) + leb128(1) + leb128(len(self.elements))
        for e in self.elements:
            p += leb128(e.eid) + bytes([e.headphones_mode << 6]) + leb128(0)
            p += param_base(e.mixgain_pid(), self.rate, self.frame_size) + s16be(e.mix_gain_q78)
        p += param_base(self.OUT_GAIN_PID, self.rate, self.frame_size) + s16be(self.out_gain_q78)
        p += leb128(len(self.layouts))
        for kind, ss, loud in self.layouts:
            if kind == "ss":
                p += bytes([(2 << 6) | (ss << 2)])
            else:
                p += bytes([3 << 6])
            p += bytes([0]) + s16be(loud) + s16be(0)
        return p

    # ---------------------------------------------------------------- temporal units
    def temporal_unit(self, pcm: List[np.ndarray], demix_mode: Optional[dict] = None, recon: Optional[dict] = None,
                      mix_gain: Optional[dict] = None, trim_start: int = 0, trim_end: int = 0) -> bytes:
        """pcm[e]: int16 [n_channels_tx][frame_size] in transmission order.
        demix_mode: {eid: mode}; recon: {eid: [(flags, [u8 gains]) or None per layer]};
        mix_gain: {pid: ("step", q) | ("linear", q0, q1) | ("bezier", q0, q1, qc, t_u8)}"""
        out = obu(OBU_TEMPORAL_DELIMITER, b"")
        for e in self.elements:
            if demix_mode and e.eid in demix_mode and e.demix is not None:
                out += obu(OBU_PARAMETER_BLOCK, leb128(e.demix_pid()) + bytes([demix_mode[e.eid] << 5]))
            if recon and e.eid in recon:
                p = leb128(e.recon_pid())
                for l, item in zip(e.layers, recon[e.eid]):
                    if not l.recon_gain:
                        continue
                    flags, gains = item
                    p += leb128(flags) + bytes(gains)
                out += obu(OBU_PARAMETER_BLOCK, p)
        for pid, g in (mix_gain or {}).items():
            p = leb128(pid)
            if g[0] == "step":
                p += leb128(0) + s16be(g[1])
            elif g[0] == "linear":
                p += leb128(1) + s16be(g[1]) + s16be(g[2])
            else:
                p += leb128(2) + s16be(g[1]) + s16be(g[2]) + s16be(g[3]) + bytes([g[4]])
            out += obu(OBU_PARAMETER_BLOCK, p)
        for e, x in zip(self.elements, pcm):
            x = np.asarray(x, np.int32 if (self.codec == "flac" and self.flac_bits > 16) else np.int16)
            ch = 0
            sid = e.substream_base
            groups = [(l.n_substreams, l.n_coupled) for l in e.layers] if e.kind == "channel" else \
                [(e.n_substreams, e.ambi_coupled if e.ambi_mode else 0)]
            for nsub, ncoupled in groups:
                for k in range(nsub):
                    nch = 2 if k < ncoupled else 1
                    if self.codec == "opus":
                        if self._opus is None:
                            self._opus = OpusEncoders(self.rate)
                        payload = self._opus.encode(sid, x[ch:ch + nch])
                    elif self.codec == "flac":
                        if self._flac is None:
                            self._flac = FlacEncoder(self.rate, self.flac_bits)
                        payload = self._flac.encode(sid, x[ch:ch + nch])
                    elif nch == 2:
                        payload = np.ascontiguousarray(x[ch:ch + 2].T).astype("<i2").tobytes()
                    else:
                        payload = x[ch].astype("<i2").tobytes()
                    ch += nch
                    if sid < 18:
                        out += obu(OBU_AUDIO_FRAME_ID0 + sid, payload, trim_start, trim_end)
                    else:
                        out += obu(OBU_AUDIO_FRAME, leb128(sid) + payload, trim_start, trim_end)
                    sid += 1
            assert ch == x.shape[0], (ch, x.shape)
        return out


# -------------------------------------------------------------------- canned configurations (BASELINE.json configs)
def cfg_stereo(rate=48000, frame_size=960, loud_q78=0, codec="ipcm"):
    e = Element(0, "channel", [Layer(STEREO, 1, 1)])
    return Stream([e], frame_size=frame_size, rate=rate, profile=0, layouts=[("ss", 0, loud_q78)], codec=codec)


def cfg_714_scalable(two_layer=True):
    """C2: base profile, [2.0 -> 7.1.4] with demixing + recon gain + output gain on layer 0"""
    if two_layer:
        # layer 2 adds [L5 R5][SL7 SR7][HFL HFR][HBL HBR] C LFE (IAMF_decoder.c:450-531): 4 coupled + 2 mono
        layers = [Layer(STEREO, 1, 1, out_gain_flags=0b110000, out_gain_q78=0x0100),
                  Layer(L714, 6, 4, recon_gain=True)]
    else:
        layers = [Layer(L714, 7, 5)]
    e = Element(0, "channel", layers, demix=(1, 0))
    return Stream([e], layouts=[("ss", 1, 0)])


def cfg_toa():
    e = Element(0, "scene", ambi_mode=0, ambi_channels=16)
    return Stream([e], layouts=[("ss", 7, 0)])


def cfg_714_foa(binaural=True, headphones_mode=0):
    e0 = Element(0, "channel", [Layer(L714, 7, 5)], mix_gain_q78=-0x0300, headphones_mode=headphones_mode)
    e1 = Element(1, "scene", ambi_mode=0, ambi_channels=4, mix_gain_q78=-0x0300, substream_base=7)
    return Stream([e0, e1], layouts=[("bin", 0, 0)] if binaural else [("ss", 9, 0)])
