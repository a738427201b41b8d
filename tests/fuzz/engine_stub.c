/* engine_stub.c - stands in for libiamf_b200.so in the host-layer fuzzer (tests/fuzz/run.sh ... stub).  TEST INFRASTRUCTURE
 * ONLY: it renders nothing.  What it does is hold the host layer to the buffer contract of include/iamf_b200.h under
 * AddressSanitizer: every submit READS every byte the contract says the caller provides (in[e] [S][F][n_in][N] float32 or
 * int16, params [S][F], ramps / segments when given) and WRITES every byte it may write (pcm [S][stride(F)], out_counts
 * [S][F]) - a host buffer that is too small is reported at once.  Sample counts follow the real engine's rule for streams
 * without trims (frame_size, or its resampled length); sizes follow iamfb_plan_max_out_samples / _out_stride_bytes. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "iamf_b200.h"

struct iamfb_ctx { int device; };
struct iamfb_plan { iamfb_plan_desc d; int co; };
struct iamfb_batch { iamfb_plan *p; int S, Fmax; };

static const int k_target_channels[IAMFB_TARGET_COUNT] = {2, 6, 8, 10, 11, 12, 14, 24, 8, 12, 10, 6, 1, 2};
static volatile uint32_t g_sink;

int iamfb_device_count(void) { return 1; }
int iamfb_ctx_create(int device, iamfb_ctx **ctx) {
  if (!ctx || device != 0) return IAMFB_ERR_BAD_ARG;
  *ctx = (iamfb_ctx *)calloc(1, sizeof(iamfb_ctx));
  return *ctx ? IAMFB_OK : IAMFB_ERR_ALLOC_FAIL;
}
void iamfb_ctx_destroy(iamfb_ctx *ctx) { free(ctx); }
const char *iamfb_last_error(void) { return "stub"; }

int iamfb_plan_create(iamfb_ctx *ctx, const iamfb_plan_desc *d, iamfb_plan **plan) {
  if (!ctx || !d || !plan) return IAMFB_ERR_BAD_ARG;
  /* the checks of the real iamfb_plan_create that the buffer arithmetic below relies on */
  if (d->frame_size <= 0 || d->frame_size > 1 << 16 || d->in_rate <= 0 || d->out_rate <= 0) return IAMFB_ERR_BAD_ARG;
  if (d->n_elements < 1 || d->n_elements > IAMFB_MAX_ELEMENTS) return IAMFB_ERR_BAD_ARG;
  if (d->target < 0 || d->target >= IAMFB_TARGET_COUNT) return IAMFB_ERR_BAD_ARG;
  if (d->bit_depth != 0 && d->bit_depth != 16 && d->bit_depth != 24 && d->bit_depth != 32) return IAMFB_ERR_BAD_ARG;
  for (int e = 0; e < d->n_elements; ++e) {
    const iamfb_element_desc *el = &d->el[e];
    const int max_in = el->kind == IAMFB_EL_SCENE ? IAMFB_MAX_SCENE_CH : IAMFB_MAX_LAYOUT_CH;
    if (el->n_in < 1 || el->n_in > max_in) return IAMFB_ERR_BAD_ARG;
  }
  iamfb_plan *p = (iamfb_plan *)calloc(1, sizeof(*p));
  if (!p) return IAMFB_ERR_ALLOC_FAIL;
  p->d = *d;
  p->co = k_target_channels[d->target];
  *plan = p;
  return IAMFB_OK;
}
void iamfb_plan_destroy(iamfb_plan *p) { free(p); }

static long long out_len(const iamfb_plan *p, long long in) {
  if (p->d.in_rate == p->d.out_rate) return in;
  return (in * p->d.out_rate + p->d.in_rate - 1) / p->d.in_rate + 2;
}
int iamfb_plan_max_out_samples(const iamfb_plan *p, int n_frames) {
  if (!p) return 0;
  const long long out = out_len(p, (long long)n_frames * p->d.frame_size);
  const long long flush = 240 + 64 + (p->d.in_rate != p->d.out_rate ? 256ll * p->d.out_rate / p->d.in_rate : 0);
  return (int)(out > flush ? out : flush);
}
size_t iamfb_plan_out_stride_bytes(const iamfb_plan *p, int n_frames) {
  if (!p) return 0;
  const size_t bps = p->d.bit_depth ? (size_t)p->d.bit_depth / 8 : 4;
  const size_t b = (size_t)iamfb_plan_max_out_samples(p, n_frames) * (size_t)p->co * bps;
  return (b + 15) & ~(size_t)15;
}

int iamfb_batch_create(iamfb_plan *p, int n_streams, int max_frames, iamfb_batch **out) {
  if (!p || !out || n_streams <= 0 || max_frames <= 0) return IAMFB_ERR_BAD_ARG;
  iamfb_batch *b = (iamfb_batch *)calloc(1, sizeof(*b));
  if (!b) return IAMFB_ERR_ALLOC_FAIL;
  b->p = p; b->S = n_streams; b->Fmax = max_frames;
  *out = b;
  return IAMFB_OK;
}
void iamfb_batch_destroy(iamfb_batch *b) { free(b); }

static void touch(const void *p, size_t n) {       /* reads every byte */
  const uint8_t *q = (const uint8_t *)p;
  uint32_t x = 0;
  for (size_t i = 0; i < n; ++i) x += q[i];
  g_sink += x;
}

static int submit_range(iamfb_batch *b, const iamfb_io *io, int F, int s_lo, int s_cnt) {
  const iamfb_plan *p = b->p;
  const size_t N = (size_t)p->d.frame_size, esz = io->in_format == IAMFB_IN_S16 ? 2 : 4;
  const size_t stride = iamfb_plan_out_stride_bytes(p, F);
  for (int e = 0; e < p->d.n_elements; ++e) {
    const size_t per = (size_t)F * (size_t)p->d.el[e].n_in * N;
    if (!io->in[e]) return IAMFB_ERR_BAD_ARG;
    touch((const uint8_t *)io->in[e] + (size_t)s_lo * per * esz, (size_t)s_cnt * per * esz);
    if (io->gain_ramp[e]) touch(io->gain_ramp[e] + (size_t)s_lo * F * N, (size_t)s_cnt * F * N * 4);
    if (io->gain_segs[e]) touch(io->gain_segs[e] + (size_t)s_lo * F, (size_t)s_cnt * F * sizeof(iamfb_gain_ramp));
  }
  if (io->out_gain_ramp) touch(io->out_gain_ramp + (size_t)s_lo * F * N, (size_t)s_cnt * F * N * 4);
  if (io->out_gain_segs) touch(io->out_gain_segs + (size_t)s_lo * F, (size_t)s_cnt * F * sizeof(iamfb_gain_ramp));
  if (!io->params || !io->pcm) return IAMFB_ERR_BAD_ARG;
  touch(io->params + (size_t)s_lo * F, (size_t)s_cnt * F * sizeof(iamfb_frame_params));
  memset((uint8_t *)io->pcm + (size_t)s_lo * stride, 0x5a, (size_t)s_cnt * stride);
  if (io->out_counts) {
    for (int s = s_lo; s < s_lo + s_cnt; ++s)
      for (int f = 0; f < F; ++f) {
        const iamfb_frame_params *fp = &io->params[(size_t)s * F + f];
        long long n = 0;
        if (fp->trim_start != 0xFFFF) {
          n = (long long)N - fp->trim_start - fp->trim_end;
          if (n < 0) n = 0;
          n = p->d.in_rate == p->d.out_rate ? n : n * p->d.out_rate / p->d.in_rate;
        }
        io->out_counts[(size_t)s * F + f] = (int32_t)n;
      }
  }
  return IAMFB_OK;
}

int iamfb_batch_submit_host(iamfb_batch *b, const iamfb_io *io, int F) {
  if (!b || !io || F <= 0 || F > b->Fmax) return IAMFB_ERR_BAD_ARG;
  return submit_range(b, io, F, 0, b->S);
}
int iamfb_batch_submit_host_hooks(iamfb_batch *b, const iamfb_io *io, int F, const iamfb_chunk_hooks *hooks) {
  if (!b || !io || !hooks || F <= 0 || F > b->Fmax) return IAMFB_ERR_BAD_ARG;
  iamfb_io lio = *io;
  for (int s = 0; s < b->S; s += 8) {
    const int cnt = b->S - s < 8 ? b->S - s : 8;
    if (hooks->fill) hooks->fill(hooks->user, s, cnt, &lio);
    int r = submit_range(b, &lio, F, s, cnt);
    if (r) return r;
    if (hooks->drain) hooks->drain(hooks->user, s, cnt);
  }
  return IAMFB_OK;
}
int iamfb_batch_flush_host(iamfb_batch *b, void *pcm, int32_t *out_counts) {
  if (!b || !pcm) return IAMFB_ERR_BAD_ARG;
  memset(pcm, 0x5a, (size_t)b->S * iamfb_plan_out_stride_bytes(b->p, 1));
  if (out_counts)
    for (int s = 0; s < b->S; ++s) out_counts[s] = b->p->d.limiter ? 240 : 0;
  return IAMFB_OK;
}
void *iamfb_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void iamfb_host_free(void *p) { free(p); }
