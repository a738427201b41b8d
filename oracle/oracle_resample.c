/*
 * oracle_resample.c - TEST INFRASTRUCTURE ONLY (see iamf_oracle.h).
 * CPU restatement of the Speex resampler (float build, quality <= 8 single-precision paths) exactly as the reference
 * decoder drives it: src/iamf_dec/resample.c and src/iamf_dec/IAMF_decoder.c:1892-1909,3223-3248.
 * Deliberately keeps the reference's *streaming* formulation (filter memory, 160-sample input chunks, per-channel
 * last_sample / samp_frac_num counters) so that it is an independent check of the closed-form indexing the CUDA
 * kernel uses.  "Magic samples" (filter-length changes while running) never occur on this path and are left out.
 */
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "iamf_oracle.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* Kaiser-window look-up tables of the Speex resampler (resample.c:105-153); only Q0..Q8 are restated. */
static const double k_kaiser10[36] = {
    0.99537781, 1.00000000, 0.99537781, 0.98162644, 0.95908712, 0.92831446, 0.89005583, 0.84522401, 0.79486424,
    0.74011713, 0.68217934, 0.62226347, 0.56155915, 0.50119680, 0.44221549, 0.38553619, 0.33194107, 0.28205962,
    0.23636152, 0.19515633, 0.15859932, 0.12670280, 0.09935205, 0.07632451, 0.05731132, 0.04193980, 0.02979584,
    0.02044510, 0.01345224, 0.00839739, 0.00488951, 0.00257636, 0.00115101, 0.00035515, 0.00000000, 0.00000000};
static const double k_kaiser8[36] = {
    0.99635258, 1.00000000, 0.99635258, 0.98548012, 0.96759014, 0.94302200, 0.91223751, 0.87580811, 0.83439927,
    0.78875245, 0.73966538, 0.68797126, 0.63451750, 0.58014482, 0.52566725, 0.47185369, 0.41941150, 0.36897272,
    0.32108304, 0.27619388, 0.23465776, 0.19672670, 0.16255380, 0.13219758, 0.10562887, 0.08273982, 0.06335451,
    0.04724088, 0.03412321, 0.02369490, 0.01563093, 0.00959968, 0.00527363, 0.00233883, 0.00050000, 0.00000000};
static const double k_kaiser6[36] = {
    0.99733006, 1.00000000, 0.99733006, 0.98935595, 0.97618418, 0.95799003, 0.93501423, 0.90755855, 0.87598009,
    0.84068475, 0.80211977, 0.76076565, 0.71712752, 0.67172623, 0.62508937, 0.57774224, 0.53019925, 0.48295561,
    0.43647969, 0.39120616, 0.34752997, 0.30580127, 0.26632152, 0.22934058, 0.19505503, 0.16360756, 0.13508755,
    0.10953262, 0.08693120, 0.06722600, 0.05031820, 0.03607231, 0.02432151, 0.01487334, 0.00752000, 0.00000000};

typedef struct { const double *table; int oversample; } WinFunc;
static const WinFunc W6 = {k_kaiser6, 32}, W8 = {k_kaiser8, 32}, W10 = {k_kaiser10, 32};
/* resample.c:187-207 (Q0..Q8) */
static const struct { int base_length, oversample; float down_bw, up_bw; const WinFunc *win; } k_quality[9] = {
    {8, 4, 0.830f, 0.860f, &W6},   {16, 4, 0.850f, 0.880f, &W6},  {32, 4, 0.882f, 0.910f, &W6},
    {48, 8, 0.895f, 0.917f, &W8},  {64, 8, 0.921f, 0.940f, &W8},  {80, 16, 0.922f, 0.940f, &W10},
    {96, 16, 0.940f, 0.945f, &W10}, {128, 16, 0.950f, 0.950f, &W10}, {160, 16, 0.960f, 0.960f, &W10}};

/* resample.c:210-229 */
static double window_value(float x, const WinFunc *f) {
  float y, frac;
  double interp[4];
  int ind;
  y = x * f->oversample;
  ind = (int)floor(y);
  frac = (y - ind);
  interp[3] = -0.1666666667 * frac + 0.1666666667 * (frac * frac * frac);
  interp[2] = frac + 0.5 * (frac * frac) - 0.5 * (frac * frac * frac);
  interp[0] = -0.3333333333 * frac + 0.5 * (frac * frac) - 0.1666666667 * (frac * frac * frac);
  interp[1] = 1.f - interp[3] - interp[2] - interp[0];
  return interp[0] * f->table[ind] + interp[1] * f->table[ind + 1] + interp[2] * f->table[ind + 2] +
         interp[3] * f->table[ind + 3];
}

/* resample.c:233-244 */
static float sinc_value(float cutoff, float x, int N, const WinFunc *f) {
  float xx = x * cutoff;
  if (fabs(x) < 1e-6)
    return cutoff;
  else if (fabs(x) > .5 * N)
    return 0;
  return cutoff * sin(M_PI * xx) / (M_PI * xx) * window_value(fabs(2. * x / N), f);
}

/* resample.c:246-256 */
static void cubic(float frac, float interp[4]) {
  interp[0] = -0.16667f * frac + 0.16667f * frac * frac * frac;
  interp[1] = frac + 0.5f * frac * frac - 0.5f * frac * frac * frac;
  interp[3] = -0.33333f * frac + 0.5f * frac * frac - 0.16667f * frac * frac * frac;
  interp[2] = 1. - interp[0] - interp[1] - interp[3];
}

static uint32_t gcd_u32(uint32_t a, uint32_t b) {
  while (b) { uint32_t t = a; a = b; b = t % b; }
  return a;
}

/* resample.c:527-593,628-640 (initial build only) */
static int build_filter(OrcResampler *st) {
  st->int_advance = st->num_rate / st->den_rate;
  st->frac_advance = st->num_rate % st->den_rate;
  st->oversample = k_quality[st->quality].oversample;
  st->filt_len = k_quality[st->quality].base_length;
  if (st->num_rate > st->den_rate) {
    st->cutoff = k_quality[st->quality].down_bw * st->den_rate / st->num_rate;
    { /* multiply_frac :512-525 */
      uint32_t major = st->filt_len / st->den_rate, remain = st->filt_len % st->den_rate;
      st->filt_len = remain * st->num_rate / st->den_rate + major * st->num_rate;
    }
    st->filt_len = ((st->filt_len - 1) & (~0x7U)) + 8;
    if (2 * st->den_rate < st->num_rate) st->oversample >>= 1;
    if (4 * st->den_rate < st->num_rate) st->oversample >>= 1;
    if (8 * st->den_rate < st->num_rate) st->oversample >>= 1;
    if (16 * st->den_rate < st->num_rate) st->oversample >>= 1;
    if (st->oversample < 1) st->oversample = 1;
  } else {
    st->cutoff = k_quality[st->quality].up_bw;
  }
  st->use_direct = st->filt_len * st->den_rate <= st->filt_len * st->oversample + 8 &&
                   INT_MAX / sizeof(float) / st->den_rate >= st->filt_len;
  st->sinc_table_length = st->use_direct ? st->filt_len * st->den_rate : st->filt_len * st->oversample + 8;
  st->sinc_table = (float *)malloc(st->sinc_table_length * sizeof(float));
  if (st->use_direct) {
    for (uint32_t i = 0; i < st->den_rate; ++i)
      for (int32_t j = 0; j < (int32_t)st->filt_len; ++j)
        st->sinc_table[i * st->filt_len + j] =
            sinc_value(st->cutoff, ((j - (int32_t)st->filt_len / 2 + 1) - ((float)i) / st->den_rate), st->filt_len,
                       k_quality[st->quality].win);
  } else {
    for (int32_t i = -4; i < (int32_t)(st->oversample * st->filt_len + 4); ++i)
      st->sinc_table[i + 4] = sinc_value(st->cutoff, (i / (float)st->oversample - st->filt_len / 2), st->filt_len,
                                         k_quality[st->quality].win);
  }
  st->mem_alloc_size = st->filt_len - 1 + st->buffer_size;
  st->mem = (float *)calloc(st->nb_channels * st->mem_alloc_size, sizeof(float));
  return 0;
}

OrcResampler *orc_resampler_open(uint32_t channels, uint32_t in_rate, uint32_t out_rate, int quality) {
  if (!channels || !in_rate || !out_rate || quality < 0 || quality > 8) return 0;
  OrcResampler *st = (OrcResampler *)calloc(1, sizeof(*st));
  uint32_t f = gcd_u32(in_rate, out_rate);
  st->nb_channels = channels;
  st->buffer_size = 160; /* :741 */
  st->quality = quality;
  st->in_rate = in_rate; st->out_rate = out_rate;
  st->num_rate = in_rate / f; st->den_rate = out_rate / f;
  st->last_sample = (int32_t *)calloc(channels, sizeof(int32_t));
  st->samp_frac_num = (uint32_t *)calloc(channels, sizeof(uint32_t));
  build_filter(st);
  for (uint32_t i = 0; i < channels; ++i) st->last_sample[i] = st->filt_len / 2; /* skip_zeros :1115-1119 */
  return st;
}

void orc_resampler_close(OrcResampler *st) {
  if (!st) return;
  free(st->last_sample); free(st->samp_frac_num); free(st->mem); free(st->sinc_table); free(st);
}

int orc_resampler_output_latency(const OrcResampler *st) {
  return ((st->filt_len / 2) * st->den_rate + (st->num_rate >> 1)) / st->num_rate;
}

/* resample.c:258-313 / 357-418 */
static int kernel_run(OrcResampler *st, uint32_t ch, const float *in, uint32_t *in_len, float *out, uint32_t *out_len) {
  const int N = st->filt_len;
  int out_sample = 0;
  int last_sample = st->last_sample[ch];
  uint32_t frac_num = st->samp_frac_num[ch];
  while (!(last_sample >= (int32_t)*in_len || out_sample >= (int32_t)*out_len)) {
    const float *iptr = &in[last_sample];
    float sum;
    if (st->use_direct) {
      const float *sinct = &st->sinc_table[frac_num * N];
      sum = 0;
      for (int j = 0; j < N; ++j) sum += ((float)(sinct[j]) * (float)(iptr[j]));
    } else {
      const int offset = frac_num * st->oversample / st->den_rate;
      const float frac = ((float)((frac_num * st->oversample) % st->den_rate)) / st->den_rate;
      float interp[4];
      float accum[4] = {0, 0, 0, 0};
      for (int j = 0; j < N; ++j) {
        const float x = iptr[j];
        accum[0] += ((float)(x) * (float)(st->sinc_table[4 + (j + 1) * st->oversample - offset - 2]));
        accum[1] += ((float)(x) * (float)(st->sinc_table[4 + (j + 1) * st->oversample - offset - 1]));
        accum[2] += ((float)(x) * (float)(st->sinc_table[4 + (j + 1) * st->oversample - offset]));
        accum[3] += ((float)(x) * (float)(st->sinc_table[4 + (j + 1) * st->oversample - offset + 1]));
      }
      cubic(frac, interp);
      sum = ((interp[0]) * (accum[0])) + ((interp[1]) * (accum[1])) + ((interp[2]) * (accum[2])) +
            ((interp[3]) * (accum[3]));
    }
    out[out_sample++] = sum;
    last_sample += st->int_advance;
    frac_num += st->frac_advance;
    if (frac_num >= st->den_rate) { frac_num -= st->den_rate; last_sample++; }
  }
  st->last_sample[ch] = last_sample;
  st->samp_frac_num[ch] = frac_num;
  return out_sample;
}

/* resample.c:786-811 + 917-972 for one channel, planar in/out (strides 1) */
static void process_channel(OrcResampler *st, uint32_t ch, const float *in, uint32_t *in_len, float *out, uint32_t *out_len) {
  uint32_t ilen = *in_len, olen = *out_len;
  float *x = st->mem + ch * st->mem_alloc_size;
  const uint32_t xlen = st->mem_alloc_size - (st->filt_len - 1);
  float ystack[1024];
  while (ilen && olen) {
    uint32_t ichunk = ilen > xlen ? xlen : ilen;
    uint32_t ochunk = olen > 1024 ? 1024 : olen;
    if (in) for (uint32_t j = 0; j < ichunk; ++j) x[j + st->filt_len - 1] = in[j];
    else for (uint32_t j = 0; j < ichunk; ++j) x[j + st->filt_len - 1] = 0;
    { /* process_native */
      int produced = kernel_run(st, ch, x, &ichunk, ystack, &ochunk);
      if (st->last_sample[ch] < (int32_t)ichunk) ichunk = st->last_sample[ch];
      ochunk = produced;
      st->last_sample[ch] -= ichunk;
      for (uint32_t j = 0; j < st->filt_len - 1; ++j) x[j] = x[j + ichunk];
    }
    for (uint32_t j = 0; j < ochunk; ++j) {
      float y = ystack[j];
      out[j] = ((y) < -1.0 ? -1.0 : ((y) > 1.0 ? 1.0 : (y))); /* FLTADJUST :84,959 */
    }
    ilen -= ichunk; olen -= ochunk; out += ochunk;
    if (in) in += ichunk;
  }
  *in_len -= ilen;
  *out_len -= olen;
}

/* IAMF_decoder.c:3223-3248: planar->interleaved->process->planar collapses to per-channel planar processing */
int orc_resample(OrcResampler *st, const float *in, float *out, int frame_size) {
  uint32_t out_cap, in_len, produced = 0;
  int flush = frame_size < 0;
  if (flush) {
    st->rest_flag = 2;
    out_cap = orc_resampler_output_latency(st);
    in_len = st->filt_len / 2;
  } else {
    out_cap = frame_size * (st->out_rate / st->in_rate + 1);
    in_len = frame_size;
  }
  /* each channel writes with planar stride = its own produced count; they are all equal. First pass finds it. */
  float *tmp = (float *)malloc(sizeof(float) * (out_cap ? out_cap : 1));
  for (uint32_t c = 0; c < st->nb_channels; ++c) {
    uint32_t il = in_len, ol = out_cap;
    process_channel(st, c, flush ? 0 : in + (size_t)c * frame_size, &il, tmp, &ol);
    produced = ol;
    memcpy(out + (size_t)c * ol, tmp, sizeof(float) * ol);
  }
  free(tmp);
  if (!st->rest_flag) st->rest_flag = 1;
  return (int)produced;
}
