/*
 * iamf_timeline.c - parameter time lines of the drop-in host layer: which demixing mode / recon-gain list / mix gain
 * applies to a frame.  Behaviour follows the reference's descriptor "database" (IAMF_decoder.c:760-1126), including
 * its time-stamp window test and the way segments are consumed as time elapses.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "iamf_host.h"

/* fixedp11_5.c:45-47, 53-55, 72: same libm calls so that the scalars are bit-identical */
float ih_q_to_float(int16_t q, int frac) { return ((float)q) * powf(2.0f, (float)-frac); }
float ih_qf_to_float(uint8_t q) { return ((float)q / (pow(2.0f, (float)8) - 1.0)); }
float ih_db2lin(float db) { return powf(10.0f, 0.05f * db); }

/* IAMF_decoder.c:91-95 */
int64_t ih_time_transform(int64_t t, int s1, int s2) {
  if (s1 == s2) return t;
  double r = (double)(t * s2);
  return (int64_t)(r / s1 + 0.5f);
}

ih_param_item *ih_param_find(struct IAMF_Decoder *d, uint64_t pid) {
  for (int i = 0; i < d->n_params; ++i)
    if (d->params[i].id == pid) return &d->params[i];
  return 0;
}

/* iamf_database_parameter_add_item, IAMF_decoder.c:985-1043 */
ih_param_item *ih_param_add(struct IAMF_Decoder *d, const ih_param_def *def, uint64_t parent, int rate) {
  ih_param_item *pi = ih_param_find(d, def->id);
  if (pi) return pi;
  if (d->n_params >= IH_MAX_PARAMS) return 0;
  pi = &d->params[d->n_params++];
  memset(pi, 0, sizeof(*pi));
  pi->id = def->id;
  pi->type = def->type;
  pi->parent = parent;
  pi->def = def;
  pi->rate = rate;
  if (def->type == IAMF_PARAMETER_TYPE_MIX_GAIN) pi->use_default = 1;
  return pi;
}

/* iamf_database_parameter_add, IAMF_decoder.c:1045-1075 */
void ih_param_push(ih_param_item *pi, ih_segment *segs) {
  if (pi->type == IAMF_PARAMETER_TYPE_MIX_GAIN) pi->use_default = 0;
  while (segs) {
    ih_segment *n = segs->next;
    segs->next = 0;
    if (pi->tail) pi->tail->next = segs; else pi->head = segs;
    pi->tail = segs;
    pi->duration += segs->interval;
    segs = n;
  }
}

void ih_param_clear(ih_param_item *pi) {
  while (pi->head) {
    ih_segment *n = pi->head->next;
    free(pi->head);
    pi->head = n;
  }
  pi->tail = 0;
}

/* iamf_database_parameter_get_segment, IAMF_decoder.c:791-830: pts must fall in (timestamp, timestamp+duration] */
const ih_segment *ih_param_segment_at(const ih_param_item *pi, uint64_t pts) {
  if (!pi) return 0;
  if (!(pts > pi->timestamp && pts <= pi->timestamp + pi->duration)) return 0;
  uint64_t start = pts - pi->timestamp;
  for (const ih_segment *s = pi->head; s; s = s->next) {
    if (start < s->interval) return s;
    start -= s->interval;
  }
  return 0;
}

/* iamf_database_parameters_time_elapse, IAMF_decoder.c:1089-1126 */
void ih_params_elapse(struct IAMF_Decoder *d, uint64_t duration, uint32_t rate) {
  for (int i = 0; i < d->n_params; ++i) {
    ih_param_item *pi = &d->params[i];
    pi->elapse += (uint64_t)ih_time_transform((int64_t)duration, (int)rate, (int)pi->def->rate);
    while (pi->head && pi->head->interval <= pi->elapse) {
      ih_segment *s = pi->head;
      pi->timestamp += s->interval;
      pi->duration -= s->interval;
      pi->elapse -= s->interval;
      pi->head = s->next;
      if (!pi->head) pi->tail = 0;
      free(s);
    }
  }
}

/* iamf_database_parameter_get_mix_gain_unit, IAMF_decoder.c:857-982.
 * returns 0 when there is no unit (or the unit covers fewer samples than the frame: iamf_frame_gain then applies
 * nothing, :1385-1390), 1 for a constant in *gain, 2 for an animated gain: *segs then lists the parameter segments that
 * cover the frame's samples - which segment, from which position inside it, for how many samples - and the per-sample
 * gains (mix_gain_bezier_linear / _quad, :639-664) are evaluated on the device (k_gain_expand);
 * -1 when the frame spans more segments than iamfb_gain_ramp holds. */
static int seg_push(iamfb_gain_ramp *r, int type, int count, int offset, int interval, int ct, float s, float e, float c) {
  if (count <= 0) return 0;
  if (r->n_segs >= IAMFB_MAX_GAIN_SEGS) return -1;
  iamfb_gain_seg *g = &r->seg[r->n_segs++];
  g->type = type; g->count = count; g->offset = offset; g->interval = interval; g->ct = ct;
  g->start = s; g->end = e; g->control = c;
  return 0;
}

int ih_mix_gain_unit(const ih_param_item *pi, uint64_t pt, int duration, int rate, float *gain, iamfb_gain_ramp *segs) {
  if (!pi) return 0;
  uint64_t start = 0;
  int use_default = 0;
  if (pt < pi->timestamp) use_default = 1;
  else start = pt - pi->timestamp;
  if (pi->use_default || use_default) {
    *gain = pi->default_gain;
    return 1;
  }
  float ratio = 1.f;
  if ((uint64_t)rate != pi->def->rate) ratio = (rate + 0.1f) / pi->def->rate;
  int64_t sgd = 0;
  int left = duration, count = 0, have_ramp = 0, overflow = 0;
  float constant = 0.f;
  segs->n_segs = 0;
  for (const ih_segment *seg = pi->head; seg; seg = seg->next) {
    int64_t minterval = (int64_t)(seg->interval * ratio);
    sgd += minterval;
    if ((int64_t)start < sgd) {
      if (seg->anim == ANIMATION_TYPE_STEP) {
        if (!count && (int64_t)(start + duration) <= sgd) {
          constant = seg->g_start;
          count = duration;
        } else if (!count) {
          have_ramp = 1;
          count = (int)(sgd - (int64_t)start);
          overflow |= seg_push(segs, 0, count, 0, 0, 0, seg->g_start, 0.f, 0.f);
          start = (uint64_t)sgd;
        } else {
          int e = count + (int)minterval;
          if (e >= duration) e = duration;
          else start = (uint64_t)sgd;
          overflow |= seg_push(segs, 0, e - count, 0, 0, 0, seg->g_start, 0.f, 0.f);
          count = e;
        }
      } else {
        int ss = (int)(sgd - minterval);
        int off = (int)start - ss;
        int dd;
        have_ramp = 1;
        if ((int64_t)(start + left) <= sgd) {
          dd = left;
        } else {
          dd = (int)(sgd - (int64_t)start);
          start = (uint64_t)sgd;
          left -= dd;
        }
        if (seg->anim == ANIMATION_TYPE_LINEAR)
          overflow |= seg_push(segs, 1, dd, off, (int)minterval, 0, seg->g_start, seg->g_end, 0.f);
        else
          overflow |= seg_push(segs, 2, dd, off, (int)minterval, (int)(seg->g_ctime * (minterval + .1f)), seg->g_start, seg->g_end,
                               seg->g_control);
        count += dd;
      }
    }
    if (count == duration) break;
  }
  if (duration > count) { segs->n_segs = 0; return 0; }
  if (!have_ramp) {
    segs->n_segs = 0;
    *gain = constant;
    return 1;
  }
  return overflow ? -1 : 2;
}
