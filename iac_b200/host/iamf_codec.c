/*
 * iamf_codec.c - core (codec) decode of the drop-in host layer.  Entropy decoding is OUTSIDE the accelerated path:
 * linear PCM ('ipcm') is decoded here; Opus / AAC / FLAC belong to the reference codec libraries (libopus, fdk-aac,
 * libFLAC, README.md:118-130) which are not vendored into this repository, so streams using them are refused with
 * IAMF_ERR_UNIMPLEMENTED at configure time rather than rendered wrongly.
 * Output contract (pcm/IAMF_pcm_decoder.c:52-151): planar float, integer sample / 2^(bits-1), coupled substreams
 * de-interleaved, substreams concatenated in transmission order.
 */
#include <string.h>

#include "iamf_host.h"

int ih_codec_supported(int codec) { return codec == IAMF_CODEC_PCM; }

static int rd16le(const uint8_t *p) { return (int16_t)(p[0] | p[1] << 8); }
static int rd16be(const uint8_t *p) { return (int16_t)(p[0] << 8 | p[1]); }
static int rd24le(const uint8_t *p) { return ((int)((uint32_t)(p[0] | p[1] << 8 | p[2] << 16) << 8)) >> 8; }
/* the reference builds its big-endian 24-bit reader from the little-endian 16-bit one (bitstream.c:204-208);
   replicated so that 24-bit BE streams decode to the same values */
static int rd24be_ref(const uint8_t *p) { return ((int)((uint32_t)((p[0] | p[1] << 8) << 8 | p[2]) << 8)) >> 8; }
static int rd32le(const uint8_t *p) { return (int)((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24); }
static int rd32be(const uint8_t *p) { return (int)((uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | (uint32_t)p[3]); }

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

int ih_codec_decode(const ih_codec *cc, uint8_t *const *pkt, const uint32_t *pkt_size, int n_sub, int n_coupled,
                    float *out, int frame_size) {
  if (cc->codec == IAMF_CODEC_PCM) return pcm_decode(cc, pkt, pkt_size, n_sub, n_coupled, out, frame_size);
  return IAMF_ERR_UNIMPLEMENTED;
}
