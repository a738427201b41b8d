/*
 * oracle_limiter.c - TEST INFRASTRUCTURE ONLY (see iamf_oracle.h).
 * CPU restatement of the look-ahead peak limiter, src/iamf_dec/audio_effect_peak_limiter.c (USE_TRUEPEAK off,
 * OLD_CODE off), kept in the reference's own formulation (ring buffers + cached arg-max) so that it independently
 * checks the sliding-maximum / integer-time-index decomposition used on the GPU.
 */
#include <math.h>
#include <string.h>

#include "iamf_oracle.h"

/* :211-235 + :73-92 */
void orc_limiter_init(OrcLimiter *l, float threshold_db, int sample_rate, int channels, float atk, float rel, int delay) {
  memset(l, 0, sizeof(*l));
  l->current_gain = 1.0;
  l->target_start = -1.0;
  l->target_end = -1.0;
  l->current_tc = -1.0;
  l->peak_pos = -1;
  l->threshold = pow(10, threshold_db / 20);
  l->attack_sec = atk;
  l->release_sec = rel;
  l->inc_tc = (float)1 / (float)sample_rate;
  l->num_channels = channels;
  l->delay_size = delay;
  l->padsize = delay;
}

/* :267-271 */
static float accel(float x) {
  if (1.0 < x) return 1.0f;
  if (x < 0) return 0.0f;
  return 1.0f - powf(x - 1, 2.0);
}

/* :237-265 */
static float target_gain(OrcLimiter *l, float peak) {
  float r = 0, gain = 0;
  if (l->current_tc != -1 && l->current_tc < l->attack_sec) {
    l->current_tc += l->inc_tc;
    r = accel(l->current_tc / l->attack_sec);
    gain = l->target_start - r * (l->target_start - l->target_end);
    l->current_gain = gain;
  } else if (l->current_tc != -1 && l->current_tc < l->release_sec + l->attack_sec) {
    l->current_tc += l->inc_tc;
    r = accel((l->current_tc - l->attack_sec) / l->release_sec);
    gain = l->target_end + r * (1.0f - l->target_end);
    l->current_gain = gain;
  } else {
    l->current_gain = 1.0;
  }
  if (peak * l->current_gain > l->threshold) {
    l->target_start = l->current_gain;
    l->target_end = l->threshold / peak;
    l->current_tc = 0.0f;
  }
  return l->current_gain;
}

/* :94-204 */
int orc_limiter_process(OrcLimiter *l, const float *in, float *out, int frame_size) {
  if (!in) return 0;
#define RING(i) ((i) % l->delay_size)
  for (int k = 0; k < frame_size; ++k) {
    float peak = 0.0f, cp, gain, pmax;
    int idx = k + l->entry;
    if (l->peak_pos < 0) {
      for (int i = 0; i < l->delay_size; ++i) {
        cp = l->peak[RING(i + idx)];
        if (cp > peak) { peak = cp; l->peak_pos = RING(i + idx); }
      }
    } else {
      peak = l->peak[l->peak_pos];
    }
    gain = target_gain(l, peak);
    pmax = 0;
    for (int c = 0; c < l->num_channels; ++c) {
      int pos = c * frame_size;
      if (l->delay_size > 0) {
        float o = l->delay[c][RING(idx)] * gain;
        l->delay[c][RING(idx)] = in[pos + k];
        out[pos + k] = o;
        cp = fabs(l->delay[c][RING(idx)]);
      } else {
        out[pos + k] = in[pos + k] * gain;
        cp = fabs(l->delay[c][0]); /* the reference indexes DB_IDX with delaySize 0 here (UB); unused on this path */
      }
      if (cp > pmax) pmax = cp;
    }
    if (l->peak_pos == RING(idx)) l->peak_pos = -1;
    else if (l->peak_pos < 0 || l->peak[l->peak_pos] < pmax) l->peak_pos = RING(idx);
    l->peak[RING(idx)] = pmax;
  }
  if (l->delay_size > 0) l->entry = RING(l->entry + frame_size);
#undef RING
  if (!l->init) {
    if (l->padsize >= frame_size) {
      l->padsize -= frame_size;
      frame_size = 0;
    } else {
      int i = 0;
      for (int c = 0; c < l->num_channels; ++c) {
        int pos = c * frame_size;
        for (int k = l->padsize; k < frame_size; ++k) out[i++] = out[pos + k];
      }
      frame_size -= l->padsize;
      l->padsize = 0;
      l->init = 1;
    }
  }
  return frame_size;
}

/* heap helpers for the ctypes bindings in tests/ */
#include <stdlib.h>
OrcLimiter *orc_limiter_new(float threshold_db, int sample_rate, int channels, float atk, float rel, int delay) {
  OrcLimiter *l = (OrcLimiter *)malloc(sizeof(OrcLimiter));
  orc_limiter_init(l, threshold_db, sample_rate, channels, atk, rel, delay);
  return l;
}
void orc_limiter_free(OrcLimiter *l) { free(l); }
