/*
 * oracle_render.c - TEST INFRASTRUCTURE ONLY (see iamf_oracle.h).
 * CPU restatement of the matrix renderers, ambisonics channel conversion, mix gains, mixer, trimming, loudness and
 * PCM quantisation of the reference decoder.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "iamf_oracle.h"

/* src/iamf_dec/m2m_rdr.c:1820-1840 : sample-outer, input-middle, output-inner; accumulation over inputs ascending */
void orc_render_m2m(const float *mat, int m_in, int n_out, const float *in, float *out, int ns) {
  for (int i = 0; i < ns; ++i)
    for (int m = 0; m < m_in; ++m)
      for (int n = 0; n < n_out; ++n) {
        if (m == 0) out[n * ns + i] = 0;
        out[n * ns + i] += mat[m * n_out + n] * in[m * ns + i];
      }
}

/* src/iamf_dec/h2m_rdr.c:1088-1150 with DISABLE_LFE_HOA == 1 (ae_rdr.h:63-65) */
void orc_render_h2m(const float *mat, int m_in, int n_out, int lfe1, int lfe2, const float *in, float *out, int ns) {
  int map[24] = {0};
  for (int i = 0; i < ns; ++i)
    for (int n = 0; n < n_out; ++n)
      for (int m = 0; m < m_in; ++m) {
        if (m == 0) out[n * ns + i] = 0;
        out[n * ns + i] += mat[n * m_in + m] * in[m * ns + i];
      }
  if (lfe1 >= 0 || lfe2 >= 0) {
    int n = 0;
    for (int i = 0; i < n_out; ++i) {
      if (lfe1 == i) n++;
      if (lfe2 == i) n++;
      map[i] = n;
      n++;
    }
    for (int i = n_out - 1; i >= 0; --i)
      if (map[i] != i)
        for (int j = 0; j < ns; ++j) out[map[i] * ns + j] = out[i * ns + j];
    if (lfe1 >= 0)
      for (int j = 0; j < ns; ++j) out[lfe1 * ns + j] = 0;
    if (lfe2 >= 0)
      for (int j = 0; j < ns; ++j) out[lfe2 * ns + j] = 0;
  }
}

/* src/iamf_dec/IAMF_core_decoder.c:105-115 */
void orc_ambisonics_mono(const uint8_t *map, int channels, const float *in, float *out, int fs) {
  memset(out, 0, sizeof(float) * fs * channels);
  for (int i = 0; i < channels; ++i) memcpy(&out[fs * i], &in[fs * map[i]], fs * sizeof(float));
}

/* src/iamf_dec/IAMF_core_decoder.c:116-130 */
void orc_ambisonics_projection(const float *matrix, int rows, int cols, const float *in, float *out, int fs) {
  for (int s = 0; s < fs; ++s)
    for (int r = 0; r < rows; ++r) {
      out[r * fs + s] = .0f;
      for (int l = 0; l < cols; ++l) out[r * fs + s] += in[l * fs + s] * matrix[l * rows + r];
    }
}

/* src/iamf_dec/IAMF_decoder.c:639-645 */
void orc_gain_linear(float s, float e, int d, int o, uint32_t l, float *g) {
  int oe = o + l;
  for (int i = o, k = 0; i < oe; ++i, ++k) g[k] = s + (e - s) * i / d;
}

/* src/iamf_dec/IAMF_decoder.c:647-664 */
void orc_gain_bezier(float s, float e, int d, float c, int ct, int o, uint32_t l, float *g) {
  int oe = o + l;
  int64_t alpha = d - 2 * ct;
  float a = 1.0f;
  for (int i = o, k = 0; i < oe; ++i, ++k) {
    if (alpha) {
      a = (sqrt(pow(ct, 2) + alpha * i) - ct) / alpha;
    } else {
      a = i;
      a /= (2 * ct);
    }
    g[k] = (s + e - 2 * c) * pow(a, 2) + 2 * a * (c - s) + s;
  }
}

/* src/iamf_dec/IAMF_decoder.c:1392-1398 */
void orc_frame_gain_const(float *data, int samples, int channels, float gain) {
  if (gain != 1.f && gain > 0.f) {
    int count = samples * channels;
    for (int i = 0; i < count; ++i) data[i] *= gain;
  }
}
/* src/iamf_dec/IAMF_decoder.c:1401-1405 */
void orc_frame_gain_ramp(float *data, int samples, int channels, const float *gains) {
  for (int c = 0; c < channels; ++c)
    for (int i = 0; i < samples; ++i) data[c * samples + i] *= gains[i];
}

/* src/iamf_dec/IAMF_decoder.c:1361-1381 */
int orc_frame_trim(float *data, int samples, int channels, int start, int end, int start_ext) {
  int s = start + start_ext;
  int ret = samples - s - end;
  if (start < 0 || end < 0 || start_ext < 0 || ret < 0) return -1;
  if (ret > 0 && ret != samples)
    for (int c = 0; c < channels; ++c) memmove(&data[c * ret], &data[c * samples + s], ret * sizeof(float));
  return ret;
}

/* src/iamf_dec/IAMF_decoder.c:2719-2730 */
void orc_mix(float *dst, const float *const *elems, int n_elems, int samples, int channels) {
  memset(dst, 0, sizeof(float) * samples * channels);
  for (int e = 0; e < n_elems; ++e)
    for (int c = 0; c < channels; ++c)
      for (int i = 0; i < samples; ++i) dst[c * samples + i] += elems[e][c * samples + i];
}

/* src/iamf_dec/IAMF_decoder.c:3206-3221 */
void orc_loudness(float *block, int frame_size, int channels, float gain) {
  if (!frame_size || gain == 1.0f) return;
  for (int c = 0; c < channels; ++c)
    for (int i = 0; i < frame_size; ++i) block[c * frame_size + i] *= gain;
}

/* src/iamf_dec/IAMF_decoder.c:100-119 */
static int16_t f2i16(float x) {
  x = x * 32768.f;
  x = x > -32768.f ? x : -32768.f;
  x = x < 32767.f ? x : 32767.f;
  return (int16_t)lrintf(x);
}
static int32_t f2i24(float x) {
  x = x * 8388608.f;
  x = x > -8388608.f ? x : -8388608.f;
  x = x < 8388607.f ? x : 8388607.f;
  return (int32_t)lrintf(x);
}
static int32_t f2i32(float x) {
  x = x * 2147483648.f;
  x = x > -2147483648.f ? x : -2147483648.f;
  x = x < 2147483647.f ? x : 2147483647.f;
  return (int32_t)lrintf(x);
}

/* src/iamf_dec/IAMF_decoder.c:121-167 */
void orc_plane2stride(void *dst, const float *src, int fs, int channels, uint32_t bit_depth, uint32_t stride) {
  if (bit_depth == 16) {
    int16_t *d = (int16_t *)dst;
    memset(d, 0, 2 * fs * stride);
    for (int c = 0; c < channels; ++c)
      for (int i = 0; i < fs; ++i) d[i * stride + c] = src ? f2i16(src[fs * c + i]) : 0;
  } else if (bit_depth == 24) {
    uint8_t *d = (uint8_t *)dst;
    memset(d, 0, 3 * fs * stride);
    for (int c = 0; c < channels; ++c)
      for (int i = 0; i < fs; ++i) {
        int32_t t = src ? f2i24(src[fs * c + i]) : 0;
        d[(i * stride + c) * 3] = t & 0xff;
        d[(i * stride + c) * 3 + 1] = (t >> 8) & 0xff;
        d[(i * stride + c) * 3 + 2] = ((t >> 16) & 0x7f) | ((t >> 24) & 0x80);
      }
  } else if (bit_depth == 32) {
    int32_t *d = (int32_t *)dst;
    memset(d, 0, 4 * fs * stride);
    for (int c = 0; c < channels; ++c)
      for (int i = 0; i < fs; ++i) d[i * stride + c] = src ? f2i32(src[fs * c + i]) : 0;
  } else if (bit_depth == 0) {
    /* oracle-only debug format: float, interleaved with the same stride (the reference writes nothing for depth 0) */
    float *d = (float *)dst;
    for (int c = 0; c < channels; ++c)
      for (int i = 0; i < fs; ++i) d[i * stride + c] = src ? src[fs * c + i] : 0.f;
  }
}
