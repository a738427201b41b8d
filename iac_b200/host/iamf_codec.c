/*
 * iamf_codec.c - core (codec) decode of the drop-in host layer.  Entropy decoding is OUTSIDE the accelerated path and
 * stays in the reference codec libraries (README.md:118-130): linear PCM ('ipcm') is decoded here; Opus is decoded by
 * libopus and FLAC by libFLAC through their public APIs when the libraries were available at build time (-DIH_HAVE_OPUS /
 * -DIH_HAVE_FLAC, linked from the reference tree's dep_codecs/lib archives, never copied into this repository); AAC streams
 * are refused with IAMF_ERR_UNIMPLEMENTED at configure time rather than rendered wrongly (fdk-aac is missing upstream for
 * Linux).
 * Output contract (pcm/IAMF_pcm_decoder.c:52-151): planar float, integer sample / 2^(bits-1), coupled substreams
 * de-interleaved, substreams concatenated in transmission order.
 */
#include <stdlib.h>
#include <string.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "iamf_host.h"

#ifdef IH_HAVE_OPUS
/* public libopus API (opus.h) - declared here so that no third-party header is needed at build time */
typedef struct OpusDecoder OpusDecoder;
OpusDecoder *opus_decoder_create(int Fs, int channels, int *error);
int opus_decode(OpusDecoder *st, const unsigned char *data, int len, short *pcm, int frame_size, int decode_fec);
void opus_decoder_destroy(OpusDecoder *st);
#endif

#ifdef IH_HAVE_FLAC
/* public libFLAC stream-decoder API (FLAC/stream_decoder.h) - declared here so that no third-party header is needed at
 * build time; enums travel as int */
typedef struct FLAC__StreamDecoder FLAC__StreamDecoder;
typedef struct { uint32_t blocksize, sample_rate, channels; /* ... FLAC__FrameHeader continues */ } ih_flac_frame_head;
typedef int (*ih_flac_read_cb)(const FLAC__StreamDecoder *, uint8_t buffer[], size_t *bytes, void *client);
typedef int (*ih_flac_write_cb)(const FLAC__StreamDecoder *, const void *frame, const int32_t *const buffer[], void *client);
typedef void (*ih_flac_meta_cb)(const FLAC__StreamDecoder *, const void *metadata, void *client);
typedef void (*ih_flac_error_cb)(const FLAC__StreamDecoder *, int status, void *client);
FLAC__StreamDecoder *FLAC__stream_decoder_new(void);
void FLAC__stream_decoder_delete(FLAC__StreamDecoder *);
int FLAC__stream_decoder_set_md5_checking(FLAC__StreamDecoder *, int);
int FLAC__stream_decoder_init_stream(FLAC__StreamDecoder *, ih_flac_read_cb, void *seek, void *tell, void *length, void *eof,
                                     ih_flac_write_cb, ih_flac_meta_cb, ih_flac_error_cb, void *client);
int FLAC__stream_decoder_process_until_end_of_metadata(FLAC__StreamDecoder *);
int FLAC__stream_decoder_process_single(FLAC__StreamDecoder *);
int FLAC__stream_decoder_flush(FLAC__StreamDecoder *);
#endif

int ih_codec_supported(int codec) {
#ifdef IH_HAVE_OPUS
  if (codec == IAMF_CODEC_OPUS) return 1;
#endif
#ifdef IH_HAVE_FLAC
  if (codec == IAMF_CODEC_FLAC) return 1;
#endif
  return codec == IAMF_CODEC_PCM;
}

/* STREAMINFO of a FLAC decoder config (the metadata blocks without the "fLaC" marker): byte offset of its 34-byte body,
 * or -1 (iamf_codec_conf_get_sampling_rate IAMF_decoder.c:733-751, flac_header_set_channels flac_multistream_decoder.c:124-150) */
static int flac_streaminfo(const uint8_t *conf, int size) {
  int off = 0;
  while (off + 4 <= size) {
    const int last = conf[off] >> 7, type = conf[off] & 0x7f;
    const int len = conf[off + 1] << 16 | conf[off + 2] << 8 | conf[off + 3];
    off += 4;
    if (!type) return off + 34 <= size ? off : -1;
    off += len;
    if (last) break;
  }
  return -1;
}
int ih_flac_rate(const uint8_t *conf, int size) {
  const int o = flac_streaminfo(conf, size);
  if (o < 0) return -1;
  return conf[o + 10] << 12 | conf[o + 11] << 4 | conf[o + 12] >> 4;
}
int ih_flac_bits(const uint8_t *conf, int size) {
  const int o = flac_streaminfo(conf, size);
  if (o < 0) return -1;
  return (((conf[o + 12] & 1) << 4) | (conf[o + 13] >> 4)) + 1;
}

static int rd16le(const uint8_t *p) { return (int16_t)(p[0] | p[1] << 8); }
static int rd16be(const uint8_t *p) { return (int16_t)(p[0] << 8 | p[1]); }
static int rd24le(const uint8_t *p) { return ((int)((uint32_t)(p[0] | p[1] << 8 | p[2] << 16) << 8)) >> 8; }
/* the reference builds its big-endian 24-bit reader from the little-endian 16-bit one (bitstream.c:204-208);
   replicated so that 24-bit BE streams decode to the same values */
static int rd24be_ref(const uint8_t *p) { return ((int)((uint32_t)((p[0] | p[1] << 8) << 8 | p[2]) << 8)) >> 8; }
static int rd32le(const uint8_t *p) { return (int)((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24); }
static int rd32be(const uint8_t *p) { return (int)((uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | (uint32_t)p[3]); }

/* 16-bit little-endian linear PCM handed on as it is: planar int16 (the engine scales by 1/32768, IAMFB_IN_S16).  Mono
 * sub-streams are one memcpy, coupled ones a de-interleave */
static int pcm_decode_s16(const ih_codec *cc, uint8_t *const *pkt, const uint32_t *pkt_size, int n_sub, int n_coupled,
                          int16_t *out, int frame_size) {
  const int le = cc->conf[0] != 0;
  if (n_sub <= 0 || cc->conf[1] != 16) return IAMF_ERR_BAD_ARG;
  const int samples = n_coupled ? (int)(pkt_size[0] / 4) : (int)(pkt_size[0] / 2);
  for (int c = 0; c < n_sub; ++c)
    if ((c < n_coupled ? (int)(pkt_size[c] / 4) : (int)(pkt_size[c] / 2)) != samples) return IAMF_ERR_INTERNAL;
  if (samples > frame_size) return IAMF_ERR_INTERNAL;
  int ch = 0;
  for (int c = 0; c < n_sub; ++c) {
    const uint8_t *p = pkt[c];
    if (c < n_coupled) {
      int16_t *l = out + (size_t)samples * ch, *r = l + samples;
      if (le) {
        int s = 0;
#if defined(__SSE2__)
        /* 8 stereo samples per step: L = the low, R = the high half of every 32-bit pair (the values are int16, so the
         * saturating pack is exact) */
        for (; s + 8 <= samples; s += 8) {
          const __m128i a = _mm_loadu_si128((const __m128i *)(p + 4 * s)), b = _mm_loadu_si128((const __m128i *)(p + 4 * s + 16));
          _mm_storeu_si128((__m128i *)(l + s), _mm_packs_epi32(_mm_srai_epi32(_mm_slli_epi32(a, 16), 16), _mm_srai_epi32(_mm_slli_epi32(b, 16), 16)));
          _mm_storeu_si128((__m128i *)(r + s), _mm_packs_epi32(_mm_srai_epi32(a, 16), _mm_srai_epi32(b, 16)));
        }
#endif
        for (; s < samples; ++s) {
          int16_t v[2];
          memcpy(v, p + 4 * s, 4);
          l[s] = v[0];
          r[s] = v[1];
        }
      } else {
        for (int s = 0; s < samples; ++s) {
          l[s] = (int16_t)(p[4 * s] << 8 | p[4 * s + 1]);
          r[s] = (int16_t)(p[4 * s + 2] << 8 | p[4 * s + 3]);
        }
      }
      ch += 2;
    } else {
      int16_t *d = out + (size_t)samples * ch;
      if (le) memcpy(d, p, (size_t)samples * 2);
      else
        for (int s = 0; s < samples; ++s) d[s] = (int16_t)(p[2 * s] << 8 | p[2 * s + 1]);
      ch += 1;
    }
  }
  return samples;
}

static int pcm_decode(const ih_codec *cc, uint8_t *const *pkt, const uint32_t *pkt_size, int n_sub, int n_coupled,
                      float *out, int frame_size) {
  const int le = cc->conf[0] != 0, bits = cc->conf[1];
  const int bytes = bits / 8;
  int (*rd)(const uint8_t *) = rd16le;
  float scale = 1 << 15;
  if (bits == 16) { if (!le) rd = rd16be; }
  else if (bits == 24) { scale = 1 << 23; rd = le ? rd24le : rd24be_ref; }
  else if (bits == 32) { scale = 1U << 31; rd = le ? rd32le : rd32be; }
  if (n_sub <= 0 || bytes <= 0) return IAMF_ERR_BAD_ARG;
  int samples = n_coupled ? (int)(pkt_size[0] / 2 / bytes) : (int)(pkt_size[0] / bytes);
  for (int c = 0; c < n_sub; ++c) {
    int n = c < n_coupled ? (int)(pkt_size[c] / 2 / bytes) : (int)(pkt_size[c] / bytes);
    if (n != samples) return IAMF_ERR_INTERNAL;
  }
  if (samples > frame_size) return IAMF_ERR_INTERNAL;
  int ch = 0;
  for (int c = 0; c < n_sub; ++c) {
    if (c < n_coupled) {
      for (int s = 0; s < samples; ++s) {
        out[(size_t)samples * ch + s] = rd(pkt[c] + (s * 2) * bytes) / scale;
        out[(size_t)samples * (ch + 1) + s] = rd(pkt[c] + (s * 2 + 1) * bytes) / scale;
      }
      ch += 2;
    } else {
      for (int s = 0; s < samples; ++s) out[(size_t)samples * ch + s] = rd(pkt[c] + s * bytes) / scale;
      ch += 1;
    }
  }
  return samples;
}

#ifdef IH_HAVE_OPUS
/* one OpusDecoder per sub-stream (stereo for coupled ones), int16 output scaled by 1/32768 like
 * opus/IAMF_opus_decoder.c:119-138; coupled sub-streams first, planar output (opus_multistream2_decoder.c:125-165) */
static int opus_decode_group(ih_stream *st, int first_sub, const ih_codec *cc, uint8_t *const *pkt, const uint32_t *pkt_size,
                             int n_sub, int n_coupled, float *out, int16_t *out16, int frame_size) {
  short buf[2 * 5760];
  int ch = 0, samples = 0;
  if (frame_size > 5760) return IAMF_ERR_BAD_ARG;
  for (int c = 0; c < n_sub; ++c) {
    const int nch = c < n_coupled ? 2 : 1;
    OpusDecoder *dec = (OpusDecoder *)st->codec_state[first_sub + c];
    if (!dec) {
      int err = 0;
      dec = opus_decoder_create(cc->rate, nch, &err);
      if (!dec) return IAMF_ERR_INVALID_STATE;
      st->codec_state[first_sub + c] = dec;
    }
    int n = opus_decode(dec, pkt[c], (int)pkt_size[c], buf, frame_size, 0);
    if (n < 0) return IAMF_ERR_INTERNAL;
    if (c && n != samples) return IAMF_ERR_INTERNAL;
    samples = n;
    for (int k = 0; k < nch; ++k) {
      if (out16)
        for (int s = 0; s < n; ++s) out16[(size_t)n * (ch + k) + s] = buf[s * nch + k];
      else
        for (int s = 0; s < n; ++s) out[(size_t)n * (ch + k) + s] = buf[s * nch + k] / 32768.f;
    }
    ch += nch;
  }
  return samples;
}
#endif

#ifdef IH_HAVE_FLAC
/* one FLAC stream decoder per sub-stream, fed one frame per call through the read callback and writing its channels
 * straight into the element's planar frame (flac/flac_multistream_decoder.c:64-122,152-183; IAMF_flac_decoder.c:101-122:
 * integer sample / 2^(bits-1)) */
typedef struct {
  FLAC__StreamDecoder *dec;
  const uint8_t *packet;
  uint32_t packet_size;
  int nch, bits, fs, cap;
  float *out;            /* planar destination of this sub-stream's channels (row pitch = blocksize) */
  int16_t *out16;
} ih_flac_handle;

static int flac_read(const FLAC__StreamDecoder *d, uint8_t buffer[], size_t *bytes, void *client) {
  ih_flac_handle *h = (ih_flac_handle *)client;
  (void)d;
  if (!h->packet || *bytes < h->packet_size) { *bytes = 0; return 2; }   /* ..._READ_STATUS_ABORT */
  memcpy(buffer, h->packet, h->packet_size);
  *bytes = h->packet_size;
  h->packet = 0;
  return 0;                                                              /* ..._READ_STATUS_CONTINUE */
}
static int flac_write(const FLAC__StreamDecoder *d, const void *frame, const int32_t *const buffer[], void *client) {
  ih_flac_handle *h = (ih_flac_handle *)client;
  const int n = (int)((const ih_flac_frame_head *)frame)->blocksize;
  (void)d;
  h->fs = n;
  if (n > h->cap) return 1;                                              /* ..._WRITE_STATUS_ABORT */
  const float scale = (float)(1u << (h->bits - 1));
  for (int c = 0; c < h->nch; ++c) {
    if (h->out16)
      for (int s = 0; s < n; ++s) h->out16[(size_t)c * n + s] = (int16_t)buffer[c][s];
    else
      for (int s = 0; s < n; ++s) h->out[(size_t)c * n + s] = buffer[c][s] / scale;
  }
  return 0;
}
static void flac_meta(const FLAC__StreamDecoder *d, const void *m, void *client) { (void)d; (void)m; (void)client; }
static void flac_error(const FLAC__StreamDecoder *d, int status, void *client) { (void)d; (void)status; (void)client; }

static ih_flac_handle *flac_open(const ih_codec *cc, int nch) {
  uint8_t head[4 + sizeof(cc->conf)];
  const int o = flac_streaminfo(cc->conf, cc->conf_size);
  if (o < 0) return 0;
  ih_flac_handle *h = (ih_flac_handle *)calloc(1, sizeof(*h));
  if (!h) return 0;
  memcpy(head, "fLaC", 4);
  memcpy(head + 4, cc->conf, (size_t)cc->conf_size);
  /* the config describes the coupled (stereo) sub-streams; mono ones get their channel count patched in */
  head[4 + o + 12] = (uint8_t)((head[4 + o + 12] & ~(0x7 << 1)) | (((nch - 1) & 0x7) << 1));
  h->nch = nch;
  h->bits = ih_flac_bits(cc->conf, cc->conf_size);
  h->dec = FLAC__stream_decoder_new();
  if (!h->dec || h->bits < 4 || h->bits > 32) { if (h->dec) FLAC__stream_decoder_delete(h->dec); free(h); return 0; }
  FLAC__stream_decoder_set_md5_checking(h->dec, 0);
  h->packet = head;
  h->packet_size = (uint32_t)(4 + cc->conf_size);
  if (FLAC__stream_decoder_init_stream(h->dec, flac_read, 0, 0, 0, 0, flac_write, flac_meta, flac_error, h) != 0 ||
      !FLAC__stream_decoder_process_until_end_of_metadata(h->dec)) {
    FLAC__stream_decoder_delete(h->dec);
    free(h);
    return 0;
  }
  h->packet = 0;
  return h;
}

static int flac_decode_group(ih_stream *st, int first_sub, const ih_codec *cc, uint8_t *const *pkt, const uint32_t *pkt_size,
                             int n_sub, int n_coupled, float *out, int16_t *out16, int frame_size) {
  int ch = 0, samples = 0;
  for (int c = 0; c < n_sub; ++c) {
    const int nch = c < n_coupled ? 2 : 1;
    ih_flac_handle *h = (ih_flac_handle *)st->codec_state[first_sub + c];
    if (!h) {
      h = flac_open(cc, nch);
      if (!h) return IAMF_ERR_INVALID_STATE;
      st->codec_state[first_sub + c] = h;
    }
    h->packet = pkt[c];
    h->packet_size = pkt_size[c];
    h->cap = frame_size;
    h->fs = 0;
    /* the rows of this sub-stream follow the earlier ones at the pitch of the frame size the codec config promises */
    h->out = out ? out + (size_t)frame_size * ch : 0;
    h->out16 = out16 ? out16 + (size_t)frame_size * ch : 0;
    if (!FLAC__stream_decoder_process_single(h->dec)) return IAMF_ERR_INTERNAL;
    if (h->fs != frame_size) return IAMF_ERR_INTERNAL;
    samples = h->fs;
    ch += nch;
  }
  return samples;
}
#endif

void ih_codec_close(ih_stream *st) {
#ifdef IH_HAVE_FLAC
  for (int i = 0; i < IH_MAX_SUBSTREAMS; ++i)
    if (st->codec_state[i] && st->cc && st->cc->codec == IAMF_CODEC_FLAC) {
      ih_flac_handle *h = (ih_flac_handle *)st->codec_state[i];
      FLAC__stream_decoder_flush(h->dec);
      FLAC__stream_decoder_delete(h->dec);
      free(h);
      st->codec_state[i] = 0;
    }
#endif
  for (int i = 0; i < IH_MAX_SUBSTREAMS; ++i) {
#ifdef IH_HAVE_OPUS
    if (st->codec_state[i] && st->cc && st->cc->codec == IAMF_CODEC_OPUS) opus_decoder_destroy((OpusDecoder *)st->codec_state[i]);
#endif
    st->codec_state[i] = 0;
  }
}

int ih_codec_decode(ih_stream *st, int first_sub, const ih_codec *cc, uint8_t *const *pkt, const uint32_t *pkt_size, int n_sub,
                    int n_coupled, float *out, int frame_size) {
  (void)st; (void)first_sub;
  if (cc->codec == IAMF_CODEC_PCM) return pcm_decode(cc, pkt, pkt_size, n_sub, n_coupled, out, frame_size);
#ifdef IH_HAVE_OPUS
  if (cc->codec == IAMF_CODEC_OPUS) return opus_decode_group(st, first_sub, cc, pkt, pkt_size, n_sub, n_coupled, out, 0, frame_size);
#endif
#ifdef IH_HAVE_FLAC
  if (cc->codec == IAMF_CODEC_FLAC) return flac_decode_group(st, first_sub, cc, pkt, pkt_size, n_sub, n_coupled, out, 0, frame_size);
#endif
  return IAMF_ERR_UNIMPLEMENTED;
}

/* core decode produces 16-bit samples for this codec configuration (Opus; 16-bit linear PCM): they can be handed to the
 * engine as int16 (IAMFB_IN_S16) instead of the float scaling of opus/IAMF_opus_decoder.c:133-135, pcm/IAMF_pcm_decoder.c */
int ih_codec_is_s16(const ih_codec *cc) {
#ifdef IH_HAVE_OPUS
  if (cc->codec == IAMF_CODEC_OPUS) return 1;
#endif
#ifdef IH_HAVE_FLAC
  if (cc->codec == IAMF_CODEC_FLAC) return ih_flac_bits(cc->conf, cc->conf_size) == 16;
#endif
  return cc->codec == IAMF_CODEC_PCM && cc->conf[1] == 16;
}

int ih_codec_decode_s16(ih_stream *st, int first_sub, const ih_codec *cc, uint8_t *const *pkt, const uint32_t *pkt_size, int n_sub,
                        int n_coupled, int16_t *out, int frame_size) {
  (void)st; (void)first_sub;
  if (cc->codec == IAMF_CODEC_PCM) return pcm_decode_s16(cc, pkt, pkt_size, n_sub, n_coupled, out, frame_size);
#ifdef IH_HAVE_OPUS
  if (cc->codec == IAMF_CODEC_OPUS) return opus_decode_group(st, first_sub, cc, pkt, pkt_size, n_sub, n_coupled, 0, out, frame_size);
#endif
#ifdef IH_HAVE_FLAC
  if (cc->codec == IAMF_CODEC_FLAC) return flac_decode_group(st, first_sub, cc, pkt, pkt_size, n_sub, n_coupled, 0, out, frame_size);
#endif
  return IAMF_ERR_UNIMPLEMENTED;
}
