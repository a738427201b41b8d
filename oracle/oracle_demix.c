/*
 * oracle_demix.c - TEST INFRASTRUCTURE ONLY (see iamf_oracle.h).
 * CPU restatement of the scalable-channel demixer (src/iamf_dec/demixer.c), the parametric down-mix renderer
 * (src/iamf_dec/downmix_renderer.c), the channel-layout tables (src/iamf_dec/IAMF_utils.c) and the small
 * fixed-point helpers (src/common/fixedp11_5.c).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "iamf_oracle.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ layout tables (IAMF_utils.c:111-196) ---- */
static const int k_count[ORC_LAYOUT_COUNT] = {1, 2, 6, 8, 10, 8, 10, 12, 6, 2};
static const int k_surround[ORC_LAYOUT_COUNT] = {1, 2, 5, 5, 5, 7, 7, 7, 3, 2};
static const int k_top[ORC_LAYOUT_COUNT] = {0, 0, 0, 2, 4, 0, 2, 4, 2, 0};

enum { L7 = ORC_CH_L7, R7 = ORC_CH_R7, C = ORC_CH_C, LFE = ORC_CH_LFE, SL7 = ORC_CH_SL7, SR7 = ORC_CH_SR7,
       BL7 = ORC_CH_BL7, BR7 = ORC_CH_BR7, HFL = ORC_CH_HFL, HFR = ORC_CH_HFR, HBL = ORC_CH_HBL, HBR = ORC_CH_HBR,
       MONO = ORC_CH_MONO, L2 = ORC_CH_L2, R2 = ORC_CH_R2, TL = ORC_CH_TL, TR = ORC_CH_TR, L3 = ORC_CH_L3,
       R3 = ORC_CH_R3, SL5 = ORC_CH_SL5, SR5 = ORC_CH_SR5, HL = ORC_CH_HL, HR = ORC_CH_HR, L5 = L7, R5 = R7 };

static const int k_render_order[ORC_LAYOUT_COUNT][ORC_MAX_LAYOUT_CH] = {
    {MONO},
    {L2, R2},
    {L5, R5, C, LFE, SL5, SR5},
    {L5, R5, C, LFE, SL5, SR5, HL, HR},
    {L5, R5, C, LFE, SL5, SR5, HFL, HFR, HBL, HBR},
    {L7, R7, C, LFE, SL7, SR7, BL7, BR7},
    {L7, R7, C, LFE, SL7, SR7, BL7, BR7, HL, HR},
    {L7, R7, C, LFE, SL7, SR7, BL7, BR7, HFL, HFR, HBL, HBR},
    {L3, R3, C, LFE, TL, TR},
    {L2, R2}};

static const int k_layer_order[ORC_LAYOUT_COUNT][ORC_MAX_LAYOUT_CH] = {
    {MONO},
    {L2, R2},
    {L5, R5, SL5, SR5, C, LFE},
    {L5, R5, SL5, SR5, HL, HR, C, LFE},
    {L5, R5, SL5, SR5, HFL, HFR, HBL, HBR, C, LFE},
    {L7, R7, SL7, SR7, BL7, BR7, C, LFE},
    {L7, R7, SL7, SR7, BL7, BR7, HL, HR, C, LFE},
    {L7, R7, SL7, SR7, BL7, BR7, HFL, HFR, HBL, HBR, C, LFE},
    {L3, R3, TL, TR, C, LFE},
    {L2, R2}};

static int layout_ok(int l) { return l >= 0 && l < ORC_LAYOUT_COUNT; }
int orc_layout_channel_count(int l) { return layout_ok(l) ? k_count[l] : 0; }
int orc_layout_surround(int l) { return layout_ok(l) ? k_surround[l] : 0; }
int orc_layout_top(int l) { return layout_ok(l) ? k_top[l] : 0; }
int orc_layout_channels(int l, int *chs) {
  if (!layout_ok(l)) return 0;
  for (int i = 0; i < k_count[l]; ++i) chs[i] = k_render_order[l][i];
  return k_count[l];
}
int orc_layer_channels(int l, int *chs) {
  if (!layout_ok(l)) return 0;
  for (int i = 0; i < k_count[l]; ++i) chs[i] = k_layer_order[l][i];
  return k_count[l];
}

/* IAMF_decoder.c:450-531 */
int orc_new_channels(int last, int cur, int *chs) {
  int n = 0;
  if (last < 0) return orc_layer_channels(cur, chs);
  int s1 = orc_layout_surround(last), s2 = orc_layout_surround(cur);
  int t1 = orc_layout_top(last), t2 = orc_layout_top(cur);
  if (s1 < 5 && 5 <= s2) { chs[n++] = L5; chs[n++] = R5; }
  if (s1 < 7 && 7 <= s2) { chs[n++] = SL7; chs[n++] = SR7; }
  if (t2 != t1 && t2 == 4) { chs[n++] = HFL; chs[n++] = HFR; }
  if (t2 - t1 == 4) { chs[n++] = HBL; chs[n++] = HBR; }
  else if (!t1 && t2 - t1 == 2) {
    if (s2 < 5) { chs[n++] = TL; chs[n++] = TR; }
    else { chs[n++] = HL; chs[n++] = HR; }
  }
  if (s1 < 3 && 3 <= s2) { chs[n++] = C; chs[n++] = LFE; }
  if (s1 < 2 && 2 <= s2) { chs[n++] = L2; }
  return n;
}

/* IAMF_decoder.c:371-407; bit order IAMF_types.h:38-59 */
enum { RE_L, RE_C, RE_R, RE_LS, RE_RS, RE_LTF, RE_RTF, RE_LB, RE_RB, RE_LTB, RE_RTB, RE_LFE, RE_COUNT };
uint32_t orc_recon_flags(int l1, int l2) {
  if (l1 == l2) return 0;
  int s1 = orc_layout_surround(l1), s2 = orc_layout_surround(l2);
  int t1 = orc_layout_top(l1), t2 = orc_layout_top(l2);
  uint32_t f = 0;
  if (s1 != s2) {
    if (s2 <= 3) f |= (1u << RE_L) | (1u << RE_R);
    else if (s2 == 5) f |= (1u << RE_LS) | (1u << RE_RS);
    else if (s2 == 7) f |= (1u << RE_LB) | (1u << RE_RB);
  }
  if (t2 != t1 && t2 == 4) f |= (1u << RE_LTB) | (1u << RE_RTB);
  if (s2 == 5 && t1 && t2 == t1) f |= (1u << RE_LTF) | (1u << RE_RTF);
  return f;
}

/* IAMF_decoder.c:409-448 */
int orc_recon_order(int layout, uint32_t flags, int *chs) {
  static const int map[ORC_LAYOUT_COUNT - 1][RE_COUNT] = {
      {MONO, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
      {L2, 0, R2, 0, 0, 0, 0, 0, 0, 0, 0, 0},
      {L5, C, R5, SL5, SR5, 0, 0, 0, 0, 0, 0, LFE},
      {L5, C, R5, SL5, SR5, HL, HR, 0, 0, 0, 0, LFE},
      {L5, C, R5, SL5, SR5, HFL, HFR, 0, 0, HBL, HBR, LFE},
      {L7, C, R7, SL7, SR7, 0, 0, BL7, BR7, 0, 0, LFE},
      {L7, C, R7, SL7, SR7, HL, HR, BL7, BR7, 0, 0, LFE},
      {L7, C, R7, SL7, SR7, HFL, HFR, BL7, BR7, HBL, HBR, LFE},
      {L3, C, R3, 0, 0, TL, TR, 0, 0, 0, 0, LFE}};
  int n = 0;
  if (layout < 0 || layout >= ORC_LAYOUT_COUNT - 1) return 0;
  for (int c = 0; c < RE_COUNT; ++c)
    if (flags & (1u << c)) chs[n++] = map[layout][c];
  return n;
}

/* IAMF_decoder.c:533-602; gain bit order IAMF_defines.h (L R LS RS LTF RTF) */
int orc_output_gain_channel(int layout, int g) {
  int s = orc_layout_surround(layout);
  switch (g) {
    case 0: return layout == ORC_LAYOUT_MONO ? MONO : layout == ORC_LAYOUT_STEREO ? L2 : layout == ORC_LAYOUT_312 ? L3 : 0;
    case 1: return layout == ORC_LAYOUT_STEREO ? R2 : layout == ORC_LAYOUT_312 ? R3 : 0;
    case 2: return s == 5 ? SL5 : 0;
    case 3: return s == 5 ? SR5 : 0;
    case 4: return s < 5 ? TL : HL;
    case 5: return s < 5 ? TR : HR;
    default: return 0;
  }
}

/* ------------------------------------------------------------------ scalars (fixedp11_5.c) ---- */
float orc_q_to_float(int16_t q, int frac) { return ((float)q) * powf(2.0f, (float)-frac); }                /* :45-47 */
float orc_qf_to_float(uint8_t qf, int frac) { return ((float)qf / (pow(2.0f, (float)frac) - 1.0)); }        /* :53-55 */
float orc_db2lin(float db) { return powf(10.0f, 0.05f * db); }                                               /* :72 */
static const float k_w[11] = {0.0, 0.0179, 0.0391, 0.0658, 0.1038, 0.25, 0.3962, 0.4342, 0.4609, 0.4821, 0.5}; /* :81-82 */
float orc_get_w(int i) { return k_w[i < 0 ? 0 : i > 10 ? 10 : i]; }                                          /* :92-99 */
int orc_calc_w_idx(int off, int prev) {                                                                      /* :83-90 */
  if (off > 0) return prev + 1 < 10 ? prev + 1 : 10;
  return prev - 1 > 0 ? prev - 1 : 0;
}

/* mode -> alpha beta gamma delta w_off   (demixer.c:62-72 == IAMF_utils.c:236-240) */
static const struct { float a, b, g, d; int w; } k_mix[8] = {
    {1.0, 1.0, 0.707, 0.707, -1},   {0.707, 0.707, 0.707, 0.707, -1}, {1.0, 0.866, 0.866, 0.866, -1}, {0, 0, 0, 0, 0},
    {1.0, 1.0, 0.707, 0.707, 1},    {0.707, 0.707, 0.707, 0.707, 1},  {1.0, 0.866, 0.866, 0.866, 1},  {0, 0, 0, 0, 0}};

/* ------------------------------------------------------------------ demixer ---- */
enum { SLOT_S_L, SLOT_S_R, SLOT_S5_L, SLOT_S5_R, SLOT_T_L, SLOT_T_R, SLOT_COUNT }; /* demixer.c:114-122 */

OrcDemixer *orc_demixer_open(int frame_size) {
  OrcDemixer *d = (OrcDemixer *)calloc(1, sizeof(*d));
  int wl = frame_size / 8;
  d->frame_size = frame_size;
  d->layout = -1;
  d->hann = (float *)malloc(sizeof(float) * (wl > 0 ? wl : 1));
  d->start_win = (float *)malloc(sizeof(float) * frame_size);
  d->stop_win = (float *)malloc(sizeof(float) * frame_size);
  d->scratch = (float *)malloc(sizeof(float) * SLOT_COUNT * frame_size);
  for (int i = 0; i < wl; ++i) d->hann[i] = (0.5 * (1.0 - cos(2.0 * M_PI * (double)i / (double)(wl - 1))));
  for (int i = 0; i < frame_size; ++i) { d->start_win[i] = 1; d->stop_win[i] = 0; }
  for (int i = 0; i < ORC_CH_COUNT; ++i) { d->last_sf[i] = 1.0; d->last_sfavg[i] = 1.0; }
  return d;
}

void orc_demixer_close(OrcDemixer *d) {
  if (!d) return;
  free(d->hann); free(d->start_win); free(d->stop_win); free(d->scratch); free(d);
}

int orc_demixer_set_frame_offset(OrcDemixer *d, uint32_t offset) {
  int wl = d->frame_size / 8, ol = wl / 2;
  int pre = offset % d->frame_size;
  d->skip = pre;
  if (pre + ol > d->frame_size) return 0;
  for (int i = 0; i < pre; ++i) { d->start_win[i] = 0; d->stop_win[i] = 1; }
  for (int i = pre, j = 0; j < ol; ++i, ++j) { d->start_win[i] = d->hann[j]; d->stop_win[i] = d->hann[j + ol]; }
  for (int i = pre + ol; i < d->frame_size; ++i) { d->start_win[i] = 1; d->stop_win[i] = 0; }
  return 0;
}

int orc_demixer_set_layout(OrcDemixer *d, int layout) {
  if (orc_layout_channels(layout, d->chs_out) > 0) { d->layout = layout; return 0; }
  return -1;
}

void orc_demixer_set_channels_order(OrcDemixer *d, const int *chs, int count) {
  memcpy(d->chs_in, chs, sizeof(int) * count);
  d->chs_count = count;
}

void orc_demixer_set_output_gain(OrcDemixer *d, const int *chs, const float *g, int count) {
  for (int i = 0; i < count; ++i) { d->gain_ch[i] = chs[i]; d->gain[i] = g[i]; }
  d->n_gain = count;
}

int orc_demixer_set_demixing_info(OrcDemixer *d, int mode, int w_idx) {
  if (mode < 0 || mode == 3 || mode > 6) return -1;
  if (w_idx < 0 || w_idx > 10) {
    d->last_mode = d->mode;
    d->mode = mode;
    d->last_w_idx = d->w_idx;
    d->w_idx = orc_calc_w_idx(k_mix[mode].w, d->last_w_idx);
  } else {
    if (mode != d->mode) d->last_mode = d->mode = mode;
    if (d->w_idx != w_idx) d->last_w_idx = d->w_idx = w_idx;
  }
  return 0;
}

void orc_demixer_set_recon_gain(OrcDemixer *d, int count, const int *chs, const float *g, uint32_t flags) {
  if (flags && (flags ^ d->recon_flags)) {
    for (int i = 0; i < count; ++i) d->recon_ch[i] = chs[i];
    d->n_recon = count;
    d->recon_flags = flags;
  }
  for (int i = 0; i < count; ++i) d->recon_gain[i] = g[i];
}

/* Recursive "make channel available" following dmx_s2..dmx_h4 (:127-378).  p[] mirrors ths->ch_data. */
typedef struct { OrcDemixer *d; float *p[ORC_CH_COUNT]; } DmxRun;

static int need_s2(DmxRun *r) {
  OrcDemixer *d = r->d; int n = d->frame_size;
  if (!r->p[L2]) return -1;
  if (r->p[R2]) return 0;
  if (!r->p[MONO]) return -1;
  float *o = d->scratch + n * SLOT_S_R;
  for (int i = 0; i < n; ++i) o[i] = 2 * r->p[MONO][i] - r->p[L2][i];
  r->p[R2] = o;
  return 0;
}
static int need_s3(DmxRun *r) {
  OrcDemixer *d = r->d; int n = d->frame_size;
  if (r->p[R3]) return 0;
  if (need_s2(r)) return -1;
  if (!r->p[C]) return -1;
  float *l = d->scratch + n * SLOT_S_L, *rr = d->scratch + n * SLOT_S_R;
  for (int i = 0; i < n; ++i) {
    l[i] = r->p[L2][i] - 0.707 * r->p[C][i]; /* double expression, demixer.c:166-167 */
    rr[i] = r->p[R2][i] - 0.707 * r->p[C][i];
  }
  r->p[L3] = l; r->p[R3] = rr;
  return 0;
}
static int need_s5(DmxRun *r) {
  OrcDemixer *d = r->d; int n = d->frame_size, i = 0;
  if (r->p[SR5]) return 0;
  if (need_s3(r)) return -1;
  if (!r->p[L5] || !r->p[R5]) return -1;
  float *l = d->scratch + n * SLOT_S5_L, *rr = d->scratch + n * SLOT_S5_R;
  for (; i < d->skip; ++i) {
    l[i] = (r->p[L3][i] - r->p[L5][i]) / k_mix[d->last_mode].d;
    rr[i] = (r->p[R3][i] - r->p[R5][i]) / k_mix[d->last_mode].d;
  }
  for (; i < n; ++i) {
    l[i] = (r->p[L3][i] - r->p[L5][i]) / k_mix[d->mode].d;
    rr[i] = (r->p[R3][i] - r->p[R5][i]) / k_mix[d->mode].d;
  }
  r->p[SL5] = l; r->p[SR5] = rr;
  return 0;
}
static int need_s7(DmxRun *r) {
  OrcDemixer *d = r->d; int n = d->frame_size, i = 0;
  if (r->p[BR7]) return 0;
  if (need_s5(r) < 0) return -1;
  if (!r->p[SL7] || !r->p[SR7]) return -1;
  float *l = d->scratch + n * SLOT_S_L, *rr = d->scratch + n * SLOT_S_R;
  for (; i < d->skip; ++i) {
    l[i] = (r->p[SL5][i] - r->p[SL7][i] * k_mix[d->last_mode].a) / k_mix[d->last_mode].b;
    rr[i] = (r->p[SR5][i] - r->p[SR7][i] * k_mix[d->last_mode].a) / k_mix[d->last_mode].b;
  }
  for (; i < n; ++i) {
    l[i] = (r->p[SL5][i] - r->p[SL7][i] * k_mix[d->mode].a) / k_mix[d->mode].b;
    rr[i] = (r->p[SR5][i] - r->p[SR7][i] * k_mix[d->mode].a) / k_mix[d->mode].b;
  }
  r->p[BL7] = l; r->p[BR7] = rr;
  return 0;
}
static int need_h2(DmxRun *r) {
  OrcDemixer *d = r->d; int n = d->frame_size, i = 0;
  if (r->p[HR]) return 0;
  if (!r->p[TL] || !r->p[TR]) return -1;
  if (need_s5(r)) return -1;
  float w = orc_get_w(d->w_idx), lw = orc_get_w(d->last_w_idx);
  float *l = d->scratch + n * SLOT_T_L, *rr = d->scratch + n * SLOT_T_R;
  for (; i < d->skip; ++i) {
    l[i] = r->p[TL][i] - k_mix[d->last_mode].d * lw * r->p[SL5][i];
    rr[i] = r->p[TR][i] - k_mix[d->last_mode].d * lw * r->p[SR5][i];
  }
  for (; i < n; ++i) {
    l[i] = r->p[TL][i] - k_mix[d->mode].d * w * r->p[SL5][i];
    rr[i] = r->p[TR][i] - k_mix[d->mode].d * w * r->p[SR5][i];
  }
  r->p[HL] = l; r->p[HR] = rr;
  return 0;
}
static int need_h4(DmxRun *r) {
  OrcDemixer *d = r->d; int n = d->frame_size, i = 0;
  if (r->p[HBR]) return 0;
  if (need_h2(r)) return -1;
  if (!r->p[HFR] || !r->p[HFL]) return -1;
  float *l = d->scratch + n * SLOT_T_L, *rr = d->scratch + n * SLOT_T_R;
  for (; i < d->skip; ++i) {
    l[i] = (r->p[HL][i] - r->p[HFL][i]) / k_mix[d->last_mode].g;
    rr[i] = (r->p[HR][i] - r->p[HFR][i]) / k_mix[d->last_mode].g;
  }
  for (; i < n; ++i) {
    l[i] = (r->p[HL][i] - r->p[HFL][i]) / k_mix[d->mode].g;
    rr[i] = (r->p[HR][i] - r->p[HFR][i]) / k_mix[d->mode].g;
  }
  r->p[HBL] = l; r->p[HBR] = rr;
  return 0;
}
static int need_channel(DmxRun *r, int ch) { /* dmx_channel :380-419 */
  if (r->p[ch]) return 0;
  switch (ch) {
    case R2: return need_s2(r);
    case L3: case R3: return need_s3(r);
    case SL5: case SR5: return need_s5(r);
    case BL7: case BR7: return need_s7(r);
    case HL: case HR: return need_h2(r);
    case HBL: case HBR: return need_h4(r);
    default: return -1;
  }
}

int orc_demixer_demix(OrcDemixer *d, float *dst, float *src, uint32_t size) {
  DmxRun r;
  int n = d->frame_size;
  if ((int)size != n) return -1;
  if (orc_layout_channel_count(d->layout) != d->chs_count) return -2;
  memset(&r, 0, sizeof(r));
  r.d = d;
  for (int c = 0; c < d->chs_count; ++c) r.p[d->chs_in[c]] = src + size * c;

  /* dmx_gainup :421-430 (in place on the input) */
  for (int c = 0; c < d->n_gain; ++c)
    for (int i = 0; i < n; ++i)
      if (r.p[d->gain_ch[c]]) r.p[d->gain_ch[c]][i] *= d->gain[c];

  /* dmx_demix :432-441 */
  for (int c = 0; c < orc_layout_channel_count(d->layout); ++c)
    if (need_channel(&r, d->chs_out[c]) < 0) return -2;

  /* dmx_rms :443-475 */
  {
    float N = 7;
    for (int c = 0; c < d->n_recon; ++c) {
      int ch = d->recon_ch[c];
      float sf = d->recon_gain[c], sfavg, f;
      float *out = r.p[ch];
      if (!out) continue; /* the reference would dereference NULL here; such inputs are not generated */
      sfavg = (2 / (N + 1)) * sf + (1 - 2 / (N + 1)) * d->last_sfavg[ch];
      for (int i = 0; i < n; ++i) {
        f = d->last_sfavg[ch] * d->stop_win[i] + sfavg * d->start_win[i];
        out[i] *= f;
      }
      d->last_sf[ch] = sf;
      d->last_sfavg[ch] = sfavg;
    }
  }

  for (int c = 0; c < d->chs_count; ++c) {
    int ch = d->chs_out[c];
    if (!r.p[ch]) continue;
    memcpy(&dst[c * size], r.p[ch], sizeof(float) * size);
  }
  return 0;
}

/* ------------------------------------------------------------------ parametric down-mix renderer ---- */
static int valid_downmix(int in, int out) { /* downmix_renderer.c:77-91 */
  int s1 = orc_layout_surround(in), s2 = orc_layout_surround(out);
  int t1 = orc_layout_top(in), t2 = orc_layout_top(out);
  if (t1 && !t2) return 0;
  return !(s1 < s2 || t1 < t2);
}

OrcDownmixer *orc_dmr_open(int in, int out) {
  if (in == out || !layout_ok(in) || !layout_ok(out) || in == ORC_LAYOUT_BINAURAL || out == ORC_LAYOUT_BINAURAL) return 0;
  if (!valid_downmix(in, out)) return 0;
  OrcDownmixer *d = (OrcDownmixer *)calloc(1, sizeof(*d));
  d->n_in = orc_layout_channels(in, d->chs_in);
  d->n_out = orc_layout_channels(out, d->chs_out);
  d->mode = -1;
  d->w_idx = -1;
  for (int i = 0; i < d->n_in; ++i) d->is_input[d->chs_in[i]] = 1;
  return d;
}
void orc_dmr_close(OrcDownmixer *d) { free(d); }

int orc_dmr_set_mode_weight(OrcDownmixer *d, int mode, int w_idx) {
  if (!d || mode < 0 || mode == 3 || mode >= 7) return -1;
  if (d->mode != mode) {
    d->mode = mode;
    d->alpha = k_mix[mode].a; d->beta = k_mix[mode].b; d->gamma = k_mix[mode].g; d->delta = k_mix[mode].d;
    d->w_off = k_mix[mode].w;
  }
  int tl_derived = !d->is_input[TL] && !d->is_input[TR];
  if (w_idx < 0 || w_idx > 10) {
    int nw = orc_calc_w_idx(d->w_off, d->w_idx);
    d->w_idx = nw;
    if (tl_derived) d->tl_scale = d->gamma * orc_get_w(nw);
  } else if (d->w_idx != w_idx) {
    d->w_idx = w_idx;
    if (tl_derived) d->tl_scale = d->gamma * orc_get_w(w_idx);
  }
  return 0;
}

/* _downmix_channel_data :115-129 with the dependency tables :65-75 written out as a switch */
static float dmr_value(const OrcDownmixer *d, float *const *p, int c, int i) {
  float sum = 0.f;
  if (p[c]) return p[c][i];
  switch (c) {
    case MONO: sum += dmr_value(d, p, R2, i) * 0.5f; sum += dmr_value(d, p, L2, i) * (float)0.5; break;
    case L2: sum += dmr_value(d, p, L3, i) * 1.f; sum += dmr_value(d, p, C, i) * (float)0.707; break;
    case R2: sum += dmr_value(d, p, R3, i) * 1.f; sum += dmr_value(d, p, C, i) * (float)0.707; break;
    case TL: sum += dmr_value(d, p, HL, i) * 1.f; sum += dmr_value(d, p, SL5, i) * d->tl_scale; break;
    case TR: sum += dmr_value(d, p, HR, i) * 1.f; sum += dmr_value(d, p, SR5, i) * d->tl_scale; break;
    case L3: sum += dmr_value(d, p, L5, i) * 1.f; sum += dmr_value(d, p, SL5, i) * d->delta; break;
    case R3: sum += dmr_value(d, p, R5, i) * 1.f; sum += dmr_value(d, p, SR5, i) * d->delta; break;
    case SL5: sum += dmr_value(d, p, SL7, i) * d->alpha; sum += dmr_value(d, p, BL7, i) * d->beta; break;
    case SR5: sum += dmr_value(d, p, SR7, i) * d->alpha; sum += dmr_value(d, p, BR7, i) * d->beta; break;
    case HL: sum += dmr_value(d, p, HFL, i) * 1.f; sum += dmr_value(d, p, HBL, i) * d->gamma; break;
    case HR: sum += dmr_value(d, p, HFR, i) * 1.f; sum += dmr_value(d, p, HBR, i) * d->gamma; break;
    default: return 0.f;
  }
  return sum;
}

int orc_dmr_downmix(OrcDownmixer *d, const float *in, float *out, uint32_t s, uint32_t duration, uint32_t size) {
  float *p[ORC_CH_COUNT];
  if (!d || !in || !out || !size || s >= size) return -1;
  memset(p, 0, sizeof(p));
  for (int i = 0; i < d->n_in; ++i) p[d->chs_in[i]] = (float *)in + size * i;
  uint32_t e = s + duration;
  if (e > size) e = size;
  for (int i = 0; i < d->n_out; ++i)
    for (uint32_t j = s; j < e; ++j) out[size * i + j] = dmr_value(d, p, d->chs_out[i], (int)j);
  return 0;
}
