/*
 * oracle_hrtf.c - TEST INFRASTRUCTURE ONLY (see iamf_oracle.h).
 *
 * Binaural (HRTF) rendering of one audio element.  PARITY UNPINNED BY THE REFERENCE: the reference hands the planar frame
 * to two closed libraries that are absent from its tree and compiled out by default (ae_rdr.h:67-69) -
 *   IAMF_element_renderer_render_M2B  m2b_rdr.c:103-121   SetBearDirectSpeakerChannel / GetBearRenderedAudio
 *   IAMF_element_renderer_render_H2B  h2b_rdr.c:109-131   SetPlanarBufferFloat / FillPlanarOutputBufferFloat
 * so there is no reference output to compare with.  What is restated is the call CONTRACT (planar [C][n] in, planar [2][n]
 * out, the filter state carried from call to call inside the renderer, one renderer per element) around a definition of
 * our own, stated here independently of the CUDA implementation:
 *
 *     xq[c][t]    = lrintf(clamp(in[c][t] * 2^20, +-(2^23 - 1)))                      (Q20; exact for 16-bit PCM / 32768)
 *     out[ear][t] = (float)( sum_c sum_{k=0..255} (int64) taps[c][ear][k] * xq[c][t - k] ) * 2^-35       (taps: Q15 int16)
 *
 * i.e. a 256-tap FIR per channel and ear in exact integer arithmetic, rounded once.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "iamf_oracle.h"

struct OrcHrtf {
  int n_ch;
  const int16_t *taps; /* [n_ch][2][ORC_HRTF_TAPS] */
  int32_t *hist;       /* [n_ch][ORC_HRTF_TAPS - 1] most recent last */
};

OrcHrtf *orc_hrtf_open(int n_ch, const int16_t *taps) {
  OrcHrtf *h = (OrcHrtf *)calloc(1, sizeof(*h));
  h->n_ch = n_ch;
  h->taps = taps;
  h->hist = (int32_t *)calloc((size_t)n_ch * (ORC_HRTF_TAPS - 1), sizeof(int32_t));
  return h;
}

void orc_hrtf_close(OrcHrtf *h) {
  if (!h) return;
  free(h->hist);
  free(h);
}

int32_t orc_hrtf_quantise(float x) {
  float v = x * 1048576.0f;
  v = fminf(fmaxf(v, -8388607.0f), 8388607.0f);
  return (int32_t)lrintf(v);
}

void orc_hrtf_render(OrcHrtf *h, const float *in, float *out, int n) {
  const int H = ORC_HRTF_TAPS - 1;
  int64_t *acc = (int64_t *)calloc((size_t)2 * n, sizeof(int64_t));
  int32_t *x = (int32_t *)malloc(sizeof(int32_t) * (size_t)(H + n));
  for (int c = 0; c < h->n_ch; ++c) {
    memcpy(x, h->hist + (size_t)c * H, sizeof(int32_t) * H);
    for (int t = 0; t < n; ++t) x[H + t] = orc_hrtf_quantise(in[(size_t)c * n + t]);
    for (int ear = 0; ear < 2; ++ear) {
      const int16_t *tp = h->taps + ((size_t)c * 2 + ear) * ORC_HRTF_TAPS;
      for (int t = 0; t < n; ++t) {
        int64_t a = 0;
        const int32_t *xs = x + H + t;
        for (int k = 0; k < ORC_HRTF_TAPS; ++k) a += (int64_t)tp[k] * xs[-k];
        acc[(size_t)ear * n + t] += a;
      }
    }
    memcpy(h->hist + (size_t)c * H, x + n, sizeof(int32_t) * H);
  }
  for (int i = 0; i < 2 * n; ++i) out[i] = (float)acc[i] * 2.9103830456733704e-11f; /* 2^-35 */
  free(acc);
  free(x);
}
