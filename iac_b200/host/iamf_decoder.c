/*
 * iamf_decoder.c - the public IAMF_decoder.h API of the drop-in libiamf.so (plain C host layer).
 *
 * What stays on the CPU is what the reference keeps outside its sample loops: OBU parsing, the descriptor database,
 * mix-presentation / layer selection, parameter time lines and core (codec) decode.  Every sample after core decode
 * is handed to the CUDA engine through include/iamf_b200.h: one iamfb_batch_submit_host per temporal unit (or one per
 * GROUP of handles through IAMF_decoder_decode_batch).  There is no CPU rendering path: when the engine cannot be
 * created (no sm_100-class GPU) configure fails with IAMF_ERR_INTERNAL.
 *
 * Behaviour follows the reference's orchestration (Samsung/iac src/iamf_dec/IAMF_decoder.c, cited per function);
 * written from scratch against that behaviour.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "iamf_host.h"

#define IH_LIMITER_DEFAULT_DB (-1.0f) /* LIMITER_MaximumTruePeak, audio_defines.h:38 */
#define IH_OUTPUT_RATE 48000          /* OUTPUT_SAMPLERATE */

/* ------------------------------------------------------------------ layout tables ---- */
/* IAMF_utils.c:111-196 */
static const int k_layout_channels[10] = {1, 2, 6, 8, 10, 8, 10, 12, 6, 2};
static const int k_layout_s[10] = {1, 2, 5, 5, 5, 7, 7, 7, 3, 2};
static const int k_layout_t[10] = {0, 0, 0, 2, 4, 0, 2, 4, 2, 0};
/* transmission order of a first layer (ia_audio_layer_get_channels, IAMF_utils.c:166-196) */
static const unsigned char k_layer0_order[10][12] = {
    {IAMFB_CH_MONO},
    {IAMFB_CH_L2, IAMFB_CH_R2},
    {IAMFB_CH_L5, IAMFB_CH_R5, IAMFB_CH_SL5, IAMFB_CH_SR5, IAMFB_CH_C, IAMFB_CH_LFE},
    {IAMFB_CH_L5, IAMFB_CH_R5, IAMFB_CH_SL5, IAMFB_CH_SR5, IAMFB_CH_HL, IAMFB_CH_HR, IAMFB_CH_C, IAMFB_CH_LFE},
    {IAMFB_CH_L5, IAMFB_CH_R5, IAMFB_CH_SL5, IAMFB_CH_SR5, IAMFB_CH_HFL, IAMFB_CH_HFR, IAMFB_CH_HBL, IAMFB_CH_HBR,
     IAMFB_CH_C, IAMFB_CH_LFE},
    {IAMFB_CH_L7, IAMFB_CH_R7, IAMFB_CH_SL7, IAMFB_CH_SR7, IAMFB_CH_BL7, IAMFB_CH_BR7, IAMFB_CH_C, IAMFB_CH_LFE},
    {IAMFB_CH_L7, IAMFB_CH_R7, IAMFB_CH_SL7, IAMFB_CH_SR7, IAMFB_CH_BL7, IAMFB_CH_BR7, IAMFB_CH_HL, IAMFB_CH_HR,
     IAMFB_CH_C, IAMFB_CH_LFE},
    {IAMFB_CH_L7, IAMFB_CH_R7, IAMFB_CH_SL7, IAMFB_CH_SR7, IAMFB_CH_BL7, IAMFB_CH_BR7, IAMFB_CH_HFL, IAMFB_CH_HFR,
     IAMFB_CH_HBL, IAMFB_CH_HBR, IAMFB_CH_C, IAMFB_CH_LFE},
    {IAMFB_CH_L3, IAMFB_CH_R3, IAMFB_CH_TL, IAMFB_CH_TR, IAMFB_CH_C, IAMFB_CH_LFE},
    {IAMFB_CH_L2, IAMFB_CH_R2}};
/* IAMF_decoder.c:208-219 */
static const int k_ss_channels_no_lfe[13] = {2, 5, 7, 9, 10, 10, 13, 22, 7, 11, 9, 5, 1};
/* iamf_sound_system_get_channel_layout, IAMF_decoder.c:236-247 */
static const int k_ss_layout[13] = {IA_CHANNEL_LAYOUT_STEREO, IA_CHANNEL_LAYOUT_510, IA_CHANNEL_LAYOUT_512,
                                    IA_CHANNEL_LAYOUT_514, -1, -1, -1, -1, IA_CHANNEL_LAYOUT_710, IA_CHANNEL_LAYOUT_714,
                                    IA_CHANNEL_LAYOUT_712, IA_CHANNEL_LAYOUT_312, IA_CHANNEL_LAYOUT_MONO};
/* iamf_layer_layout_convert_sound_system, IAMF_decoder.c:276-283 */
static const int k_layout_ss[9] = {SOUND_SYSTEM_MONO, SOUND_SYSTEM_A, SOUND_SYSTEM_B, SOUND_SYSTEM_C, SOUND_SYSTEM_D,
                                   SOUND_SYSTEM_I, SOUND_SYSTEM_EXT_712, SOUND_SYSTEM_J, SOUND_SYSTEM_EXT_312};

static int ss_valid(int ss) { return ss > SOUND_SYSTEM_INVALID && ss < SOUND_SYSTEM_END; }
static int ss_lfe1(int ss) { return ss != SOUND_SYSTEM_A && ss != SOUND_SYSTEM_MONO; }
static int ss_lfe2(int ss) { return ss == SOUND_SYSTEM_F || ss == SOUND_SYSTEM_H; }

int IAMF_layout_sound_system_channels_count(IAMF_SoundSystem ss) {
  if (!ss_valid(ss)) return IAMF_ERR_BAD_ARG;
  return k_ss_channels_no_lfe[ss] + ss_lfe1(ss) + ss_lfe2(ss);
}
int IAMF_layout_binaural_channels_count(void) { return 2; }

/* new channels a higher layer adds, iamf_channel_layout_get_new_channels IAMF_decoder.c:450-531 */
static int layer_new_channels(int last, int cur, int *out, int room) {
  int n = 0;
  if (last < 0) {
    for (int i = 0; i < k_layout_channels[cur] && n < room; ++i) out[n++] = k_layer0_order[cur][i];
    return k_layout_channels[cur] > room ? -1 : n;
  }
  int tmp[12];
  const int s1 = k_layout_s[last], s2 = k_layout_s[cur], t1 = k_layout_t[last], t2 = k_layout_t[cur];
  if (s1 < 5 && 5 <= s2) { tmp[n++] = IAMFB_CH_L5; tmp[n++] = IAMFB_CH_R5; }
  if (s1 < 7 && 7 <= s2) { tmp[n++] = IAMFB_CH_SL7; tmp[n++] = IAMFB_CH_SR7; }
  if (t2 != t1 && t2 == 4) { tmp[n++] = IAMFB_CH_HFL; tmp[n++] = IAMFB_CH_HFR; }
  if (t2 - t1 == 4) { tmp[n++] = IAMFB_CH_HBL; tmp[n++] = IAMFB_CH_HBR; }
  else if (!t1 && t2 - t1 == 2) {
    if (s2 < 5) { tmp[n++] = IAMFB_CH_TL; tmp[n++] = IAMFB_CH_TR; }
    else { tmp[n++] = IAMFB_CH_HL; tmp[n++] = IAMFB_CH_HR; }
  }
  if (s1 < 3 && 3 <= s2) { tmp[n++] = IAMFB_CH_C; tmp[n++] = IAMFB_CH_LFE; }
  if (s1 < 2 && 2 <= s2) tmp[n++] = IAMFB_CH_L2;
  if (n > room) return -1;
  for (int i = 0; i < n; ++i) out[i] = tmp[i];
  return n;
}

/* iamf_output_gain_channel_map, IAMF_decoder.c:533-602; gch: 0 L, 1 R, 2 LS, 3 RS, 4 LTF, 5 RTF */
static int output_gain_channel(int layout, int gch) {
  const int s = k_layout_s[layout];
  switch (gch) {
    case 0: return layout == IA_CHANNEL_LAYOUT_MONO ? IAMFB_CH_MONO : layout == IA_CHANNEL_LAYOUT_STEREO ? IAMFB_CH_L2
                 : layout == IA_CHANNEL_LAYOUT_312 ? IAMFB_CH_L3 : 0;
    case 1: return layout == IA_CHANNEL_LAYOUT_STEREO ? IAMFB_CH_R2 : layout == IA_CHANNEL_LAYOUT_312 ? IAMFB_CH_R3 : 0;
    case 2: return s == 5 ? IAMFB_CH_SL5 : 0;
    case 3: return s == 5 ? IAMFB_CH_SR5 : 0;
    case 4: return s < 5 ? IAMFB_CH_TL : IAMFB_CH_HL;
    case 5: return s < 5 ? IAMFB_CH_TR : IAMFB_CH_HR;
    default: return 0;
  }
}

/* DMRenderer_open validity, downmix_renderer.c:77-91,131-139 */
static int dmr_pair_valid(int in, int out) {
  if (in == out || in < 0 || in > 8 || out < 0 || out > 8) return 0;
  const int s1 = k_layout_s[in], s2 = k_layout_s[out], t1 = k_layout_t[in], t2 = k_layout_t[out];
  if (t1 && !t2) return 0;
  return !(s1 < s2 || t1 < t2);
}

/* ------------------------------------------------------------------ database ---- */
static ih_codec *db_codec(IAMF_DecoderHandle h, uint64_t id) {
  for (int i = 0; i < h->n_codecs; ++i)
    if (h->codecs[i].id == id) return &h->codecs[i];
  return 0;
}
static ih_element *db_element(IAMF_DecoderHandle h, uint64_t id) {
  for (int i = 0; i < h->n_elements; ++i)
    if (h->elements[i].id == id) return &h->elements[i];
  return 0;
}
static ih_mix *db_mix(IAMF_DecoderHandle h, uint64_t id) {
  for (int i = 0; i < h->n_mixes; ++i)
    if (h->mixes[i].id == id) return &h->mixes[i];
  return 0;
}

/* ---- engine objects shared by the handles of a process: one iamfb context per device and one plan per pipeline signature
 * (the reference needs neither; a context per handle meant ~70 CUDA events, three streams and every constant table once
 * per handle).  Calls into the engine through a shared entry are serialised by its mutex: distinct handles may still be
 * used from distinct threads (SURVEY 8b "threading"), the device part of their decode calls then runs one at a time. */
#include <pthread.h>
typedef struct ih_shared {
  struct ih_shared *next;
  int device, refs;
  iamfb_plan_desc desc;
  iamfb_ctx *ctx;
  iamfb_plan *plan;
  pthread_mutex_t mu;
} ih_shared;
static ih_shared *g_shared = 0;
static pthread_mutex_t g_shared_mu = PTHREAD_MUTEX_INITIALIZER;

static ih_shared *shared_acquire(int device, const iamfb_plan_desc *d) {
  pthread_mutex_lock(&g_shared_mu);
  ih_shared *e = g_shared;
  for (; e; e = e->next)
    if (e->device == device && memcmp(&e->desc, d, sizeof(*d)) == 0) break;
  if (!e) {
    e = (ih_shared *)calloc(1, sizeof(*e));
    if (e) {
      e->device = device;
      e->desc = *d;
      pthread_mutex_init(&e->mu, 0);
      if (iamfb_ctx_create(device, &e->ctx) != IAMFB_OK || iamfb_plan_create(e->ctx, d, &e->plan) != IAMFB_OK) {
        if (e->ctx) iamfb_ctx_destroy(e->ctx);
        free(e);
        e = 0;
      } else {
        e->next = g_shared;
        g_shared = e;
      }
    }
  }
  if (e) ++e->refs;
  pthread_mutex_unlock(&g_shared_mu);
  return e;
}

static void shared_release(ih_shared *e) {
  if (!e) return;
  pthread_mutex_lock(&g_shared_mu);
  if (--e->refs == 0) {
    ih_shared **pp = &g_shared;
    while (*pp && *pp != e) pp = &(*pp)->next;
    if (*pp) *pp = e->next;
    iamfb_plan_destroy(e->plan);
    iamfb_ctx_destroy(e->ctx);
    pthread_mutex_destroy(&e->mu);
    free(e);
  }
  pthread_mutex_unlock(&g_shared_mu);
}

static void engine_release(IAMF_DecoderHandle h) {
  if (h->group_owner) {
    if (h->batch) {
      if (h->shared) pthread_mutex_lock(&((ih_shared *)h->shared)->mu);
      iamfb_batch_destroy(h->batch);
      if (h->shared) pthread_mutex_unlock(&((ih_shared *)h->shared)->mu);
    }
    for (int e = 0; e < IAMFB_MAX_ELEMENTS; ++e) {
      iamfb_host_free(h->in[e]);
      iamfb_host_free(h->ramp[e]);
    }
    iamfb_host_free(h->out_ramp);
    iamfb_host_free(h->pcm_stage);
    iamfb_host_free(h->fp_stage);
    iamfb_host_free(h->counts_stage);
  }
  shared_release((ih_shared *)h->shared);
  h->shared = 0;
  h->batch = 0; h->plan = 0; h->ctx = 0;
  for (int e = 0; e < IAMFB_MAX_ELEMENTS; ++e) { h->in[e] = 0; h->ramp[e] = 0; }
  h->out_ramp = 0; h->pcm_stage = 0; h->fp_stage = 0; h->counts_stage = 0;
  h->group_owner = 1; h->group_size = 1; h->group_index = 0; h->leader = 0;
  h->group_units = 0; h->group_s16 = 0;
}

/* ---- the GPUs of this process (streams shard over the GPUs of a box, no exchange between them).
 * IAMF_B200_DEVICES = "all" | a count | a comma-separated list of device ordinals; without it the one device
 * IAMF_B200_DEVICE names (default 0).  Handles are dealt round the list as they are configured; a batch call splits its
 * handles over the list - handle i to entry i mod D - and runs one host thread per device (batch_units below). */
#define IH_MAX_DEVICES 16
static int g_devs[IH_MAX_DEVICES], g_ndev = 0;
static unsigned g_next_dev = 0;
static pthread_once_t g_devs_once = PTHREAD_ONCE_INIT;
static void devices_parse(void) {
  const char *list = getenv("IAMF_B200_DEVICES");
  const char *one = getenv("IAMF_B200_DEVICE");
  g_ndev = 0;
  if (list && *list) {
    if (!strcmp(list, "all") || !strchr(list, ',')) {
      int have = iamfb_device_count(), want = strcmp(list, "all") ? atoi(list) : have;
      if (want > have && have > 0) want = have;
      for (int i = 0; i < want && i < IH_MAX_DEVICES; ++i) g_devs[g_ndev++] = i;
    } else {
      for (const char *p = list; *p && g_ndev < IH_MAX_DEVICES;) {
        g_devs[g_ndev++] = atoi(p);
        p = strchr(p, ',');
        if (!p) break;
        ++p;
      }
    }
  }
  if (g_ndev <= 0) { g_devs[0] = one ? atoi(one) : 0; g_ndev = 1; }
}
static void devices_init(void) { pthread_once(&g_devs_once, devices_parse); }

/* (re)binds a configured handle to `device`: its (context, plan) entry of that device */
static int engine_bind(IAMF_DecoderHandle h, int device) {
  if (h->shared && ((ih_shared *)h->shared)->device == device) return IAMF_OK;
  engine_release(h);
  ih_shared *sh = shared_acquire(device, &h->desc);
  if (!sh) return IAMF_ERR_INTERNAL;
  h->shared = sh;
  h->ctx = sh->ctx;
  h->plan = sh->plan;
  h->pcm_stage_size = iamfb_plan_out_stride_bytes(h->plan, 1);
  return IAMF_OK;
}

static void pkt_drop(ih_stream *st, int k);

static void db_reset(IAMF_DecoderHandle h) {
  for (int i = 0; i < h->n_params; ++i) ih_param_clear(&h->params[i]);
  for (int s = 0; s < IAMFB_MAX_ELEMENTS; ++s) {
    ih_codec_close(&h->streams[s]);
    for (int k = 0; k < IH_MAX_SUBSTREAMS; ++k) pkt_drop(&h->streams[s], k);
  }
  h->n_codecs = h->n_elements = h->n_mixes = h->n_params = 0;
  h->have_header = 0;
  h->mix = 0;
  h->n_streams = 0;
  free(h->metadata.loudness_layout); free(h->metadata.loudness); free(h->metadata.param);
  memset(&h->metadata, 0, sizeof(h->metadata));
  h->metadata.output_sound_mode = IAMF_SOUND_MODE_NONE;
}

/* iamf_database_element_add IAMF_decoder.c:1127-1166: register the element's parameter definitions */
static void db_register_element_params(IAMF_DecoderHandle h, ih_element *e) {
  ih_codec *cc = db_codec(h, e->codec_id);
  const int rate = cc ? cc->rate : -1;
  for (int i = 0; i < e->n_params; ++i) ih_param_add(h, &e->params[i], e->id, rate);
}

static int db_add_descriptor(IAMF_DecoderHandle h, const ih_obu *o) {
  int rc = IAMF_OK;
  switch (o->type) {
    case IH_OBU_SEQUENCE_HEADER:
      if (o->payload_size < 6 || memcmp(o->payload, "iamf", 4)) return IAMF_ERR_INVALID_PACKET;
      h->have_header = 1;
      break;
    case IH_OBU_CODEC_CONFIG: {
      ih_codec c;
      if ((rc = ih_parse_codec(o, &c)) != IAMF_OK) return rc;
      if (db_codec(h, c.id)) break;
      if (h->n_codecs >= IH_MAX_CODECS) return IAMF_ERR_ALLOC_FAIL;
      h->codecs[h->n_codecs++] = c;
    } break;
    case IH_OBU_AUDIO_ELEMENT: {
      ih_element *e = (ih_element *)malloc(sizeof(*e));
      if (!e) return IAMF_ERR_ALLOC_FAIL;
      rc = ih_parse_element(o, e);
      if (rc == IAMF_OK && !db_element(h, e->id)) {
        if (h->n_elements >= IH_MAX_ELEMENTS) rc = IAMF_ERR_ALLOC_FAIL;
        else {
          h->elements[h->n_elements] = *e;
          db_register_element_params(h, &h->elements[h->n_elements]);
          ++h->n_elements;
        }
      }
      free(e);
    } break;
    case IH_OBU_MIX_PRESENTATION: {
      ih_mix m;
      if ((rc = ih_parse_mix(o, &m)) != IAMF_OK) return rc;
      if (db_mix(h, m.id)) break;
      if (h->n_mixes >= IH_MAX_MIXES) return IAMF_ERR_ALLOC_FAIL;
      h->mixes[h->n_mixes++] = m;
    } break;
    default: break;
  }
  return rc;
}

static int is_descriptor(int type) {
  return type == IH_OBU_SEQUENCE_HEADER || type == IH_OBU_CODEC_CONFIG || type == IH_OBU_AUDIO_ELEMENT ||
         type == IH_OBU_MIX_PRESENTATION;
}

/* iamf_decoder_internal_read_descriptors_OBUs, IAMF_decoder.c:2784-2836 */
static uint32_t read_descriptors(IAMF_DecoderHandle h, const uint8_t *data, uint32_t size) {
  uint32_t pos = 0;
  ih_obu o;
  while (pos < size) {
    uint32_t used = ih_obu_split(data + pos, size - pos, &o);
    if (!used) break;
    if (o.redundant && !(~h->flags & IH_FLAG_DESCRIPTORS)) { pos += used; continue; }
    if (is_descriptor(o.type)) {
      if (db_add_descriptor(h, &o) == IAMF_OK) {
        switch (o.type) {
          case IH_OBU_SEQUENCE_HEADER: h->flags = IH_FLAG_MAGIC; break;
          case IH_OBU_CODEC_CONFIG: h->flags |= IH_FLAG_CODEC; break;
          case IH_OBU_AUDIO_ELEMENT: h->flags |= IH_FLAG_ELEMENT; break;
          case IH_OBU_MIX_PRESENTATION: h->flags |= IH_FLAG_MIX; break;
          default: break;
        }
      }
    } else {
      /* the first non-descriptor OBU after a complete descriptor set ends the configuration */
      if (!(~h->flags & IH_FLAG_DESCRIPTORS)) h->flags |= IH_FLAG_CONFIG;
      break;
    }
    pos += used;
  }
  return pos;
}

/* iamf_decoder_internal_init, IAMF_decoder.c:2752-2782 */
static int internal_init(IAMF_DecoderHandle h, const uint8_t *data, uint32_t size, uint32_t *rsize) {
  uint32_t pos = 0, used = 0;
  if (~h->flags & IH_FLAG_MAGIC) {
    ih_obu o;
    while (pos < size) {
      used = ih_obu_split(data, size, &o); /* (the reference re-splits the first OBU; so do we) */
      if (!used || o.type == IH_OBU_SEQUENCE_HEADER) break;
      pos += used;
    }
    if (pos > size) pos = size;   /* (a first OBU that is no sequence header is stepped over repeatedly: never past the buffer) */
  }
  if (used || (h->flags & IH_FLAG_MAGIC)) pos += read_descriptors(h, data + pos, size - pos);
  *rsize = pos;
  return (~h->flags & IH_FLAG_CONFIG) ? IAMF_ERR_BUFFER_TOO_SMALL : IAMF_OK;
}

/* ------------------------------------------------------------------ presentation ---- */
/* iamf_target_layout_matching_calculation, IAMF_decoder.c:2996-3028 */
static int layout_score(IAMF_DecoderHandle h, int type, int ss) {
  int s = 0;
  if (type == h->layout_type) {
    if (type == IAMF_LAYOUT_TYPE_BINAURAL) s = 100;
    else if (type == IAMF_LAYOUT_TYPE_LOUDSPEAKERS_SS_CONVENTION && ss == h->sound_system) s = 100;
  }
  if (!s) {
    int chs = 0;
    s = 50;
    if (type == IAMF_LAYOUT_TYPE_LOUDSPEAKERS_SS_CONVENTION) chs = IAMF_layout_sound_system_channels_count(ss);
    else if (type == IAMF_LAYOUT_TYPE_BINAURAL) chs = 2;
    if (h->out_channels < chs) s += chs - h->out_channels;
    else s -= h->out_channels - chs;
  }
  return s;
}

/* iamf_decoder_get_best_mix_presentation, IAMF_decoder.c:3083-3111 */
static const ih_mix *best_mix(IAMF_DecoderHandle h) {
  const ih_mix *mp = 0;
  if (h->n_mixes <= 0) return 0;
  if (h->n_mixes == 1) mp = &h->mixes[0];
  else if ((int64_t)h->mix_id >= 0) mp = db_mix(h, h->mix_id);
  if (!mp) {
    int best = 0;
    for (int i = 0; i < h->n_mixes; ++i) {
      int score = 0;
      for (int l = 0; l < h->mixes[i].n_layouts; ++l) {
        int s = layout_score(h, h->mixes[i].layouts[l].type, h->mixes[i].layouts[l].sound_system);
        if (s > score) score = s;
      }
      if (best < score) { best = score; mp = &h->mixes[i]; }
    }
  }
  return mp;
}

/* iamf_mix_presentation_get_best_loudness, IAMF_decoder.c:3030-3059 */
static float best_loudness(IAMF_DecoderHandle h, const ih_mix *m) {
  int score = 0, idx = -1;
  for (int i = 0; i < m->n_layouts; ++i) {
    int s = layout_score(h, m->layouts[i].type, m->layouts[i].sound_system);
    if (s > score) { score = s; idx = i; }
  }
  return idx >= 0 ? ih_q_to_float(m->layouts[idx].loud.integrated_loudness, 8) : 0.f;
}

/* iamf_stream_new + iamf_stream_set_output_layout, IAMF_decoder.c:1617-1825 */
static int stream_setup(IAMF_DecoderHandle h, ih_stream *st, const ih_element *el, const ih_codec *cc) {
  ih_codec_close(st);
  for (int k = 0; k < IH_MAX_SUBSTREAMS; ++k) pkt_drop(st, k);
  memset(st, 0, sizeof(*st));
  st->el = el;
  st->cc = cc;
  st->dmx_mode = -1;
  if (el->type == AUDIO_ELEMENT_CHANNEL_BASED) {
    int chs = 0, last = -1;
    if (el->n_layers <= 0) return IAMF_ERR_INVALID_PACKET;
    for (int i = 0; i < el->n_layers; ++i) {
      const ih_layer *l = &el->layers[i];
      if (l->layout < 0 || l->layout > IA_CHANNEL_LAYOUT_BINAURAL) return IAMF_ERR_UNIMPLEMENTED;
      int n = layer_new_channels(last, l->layout, &st->chs_order[chs], IAMFB_MAX_LAYOUT_CH - chs);
      if (n < 0) return IAMF_ERR_BUFFER_TOO_SMALL;
      chs += n;
      st->n_coupled += l->n_coupled;
      last = l->layout;
    }
    st->n_channels = el->n_sub + st->n_coupled;
    st->layer = el->n_layers - 1;
    /* the generic (non-TV) build picks the layer matching the playback layout, else the next larger one */
    if (el->n_layers > 1 && h->layout_type != IAMF_LAYOUT_TYPE_BINAURAL) {
      int found = 0;
      for (int i = 0; i < el->n_layers && !found; ++i)
        if (el->layers[i].layout <= 8 && k_layout_ss[el->layers[i].layout] == h->sound_system) { st->layer = i; found = 1; }
      const int playback = IAMF_layout_sound_system_channels_count((IAMF_SoundSystem)h->sound_system);
      for (int i = 0; i < el->n_layers && !found; ++i)
        if (k_layout_channels[el->layers[i].layout] > playback) { st->layer = i; found = 1; }
    }
    st->layout = el->layers[st->layer].layout;
    st->n_layout_ch = k_layout_channels[st->layout];
    st->n_decoded = 0;
    st->n_sub_used = 0;
    for (int i = 0; i <= st->layer; ++i) {
      st->n_decoded += el->layers[i].n_sub + el->layers[i].n_coupled;
      st->n_sub_used += el->layers[i].n_sub;
    }
  } else {
    st->n_channels = el->ambi_channels;
    st->n_coupled = el->ambi_coupled;
    st->n_decoded = el->ambi_sub + el->ambi_coupled;
    st->n_sub_used = el->ambi_sub;
    st->layer = 0;
  }
  return IAMF_OK;
}

static void fill_element_desc(IAMF_DecoderHandle h, const ih_stream *st, iamfb_element_desc *d) {
  const ih_element *el = st->el;
  memset(d, 0, sizeof(*d));
  if (el->type == AUDIO_ELEMENT_CHANNEL_BASED) {
    d->kind = IAMFB_EL_CHANNEL;
    d->n_in = st->n_decoded;
    /* iamf_stream_render: a one-channel element is rendered as mono whatever its layer says (IAMF_decoder.c:2593) */
    d->layout = st->layout;
    for (int i = 0; i < st->n_decoded && i < IAMFB_MAX_LAYOUT_CH; ++i) d->chs_in[i] = st->chs_order[i];
    /* iamf_stream_scale_demixer_configure, IAMF_decoder.c:2351-2401 */
    int n = 0;
    for (int l = 0; l <= st->layer; ++l) {
      const ih_layer *ly = &el->layers[l];
      if (!ly->out_gain_present) continue;
      const float g = ih_db2lin(ih_q_to_float(ly->out_gain_q, 8));
      for (int c = 0; c < 6; ++c) {
        if (!(ly->out_gain_flags & (0x20 >> c))) continue;   /* RSHIFT(c): bit 5-c */
        int ch = output_gain_channel(ly->layout, c);
        if (ch && n < IAMFB_MAX_LAYOUT_CH) { d->out_gain_ch[n] = ch; d->out_gain[n] = g; ++n; }
      }
    }
    d->n_out_gain = n;
    d->has_demix_info = el->has_demix;
    d->default_mode = el->has_demix ? el->dmx_mode : -1;
    d->default_w_idx = el->has_demix ? el->dmx_w : -1;
    d->first_layer_layout = el->layers[0].layout;
    d->selected_layer = st->layer;
    d->recon_present = el->layers[st->layer].recon_present;
    /* iamf_stream_renderer_enable_downmix, IAMF_decoder.c:2448-2478 */
    d->use_dmr = 0;
    if (h->layout_type == IAMF_LAYOUT_TYPE_LOUDSPEAKERS_SS_CONVENTION && el->has_demix) {
      const int out = k_ss_layout[h->sound_system];
      if (out >= 0 && dmr_pair_valid(st->layout, out)) { d->use_dmr = 1; d->dmr_out_layout = out; }
    }
  } else {
    d->kind = IAMFB_EL_SCENE;
    d->n_in = st->n_decoded;
    d->ambi_channels = el->ambi_channels;
    d->ambi_mode = el->ambi_mode;
    if (el->ambi_mode == 0) {
      for (int i = 0; i < el->ambi_channels && i < IAMFB_MAX_SCENE_CH; ++i) d->ambi_map[i] = el->ambi_map[i];
    } else {
      d->ambi_cols = el->ambi_sub + el->ambi_coupled;
      const int count = el->ambi_channels * d->ambi_cols;
      for (int i = 0; i < count && i < IAMFB_MAX_SCENE_CH * IAMFB_MAX_SCENE_CH; ++i)
        d->ambi_matrix[i] = ih_q_to_float((int16_t)((el->ambi_map[2 * i] << 8) | el->ambi_map[2 * i + 1]), 15);
    }
  }
}

static int engine_build(IAMF_DecoderHandle h) {
  iamfb_plan_desc *d = &h->desc;
  memset(d, 0, sizeof(*d));
  const ih_stream *s0 = &h->streams[0];
  d->frame_size = (int)s0->cc->frame_size;
  d->in_rate = s0->cc->rate;
  d->out_rate = (int)h->sampling_rate;
  d->n_elements = h->n_streams;
  for (int i = 0; i < h->n_streams; ++i) fill_element_desc(h, &h->streams[i], &d->el[i]);
  /* The binauraliser.  The reference compiles it out by default (DISABLE_BINAURALIZER 1, ae_rdr.h:67-69): binaural output
   * is then the stereo rows of the matrix tables, and that is the default here too.  IAMF_B200_BINAURALIZER=1 is the
   * run-time counterpart of building the reference with DISABLE_BINAURALIZER 0: scene-based elements always take the HRTF
   * renderer (IAMF_decoder.c:2606-2612), channel-based ones - scalable layers included: the engine de-mixes them in front
   * of the renderer - when the mix presentation says headphones_rendering_mode 1 (:2565-2573). */
  {
    const char *benv = getenv("IAMF_B200_BINAURALIZER");
    if (benv && atoi(benv) && h->layout_type == IAMF_LAYOUT_TYPE_BINAURAL && h->mix) {
      for (int i = 0; i < h->n_streams; ++i) {
        const ih_element *el = h->streams[i].el;
        int mode = 0;
        for (int k = 0; k < h->mix->n_elements; ++k)
          if (h->mix->el[k].element_id == el->id) mode = h->mix->el[k].headphones_mode;
        if (el->type != AUDIO_ELEMENT_CHANNEL_BASED || mode == 1) d->el[i].binaural_hrtf = 1;
      }
    }
  }
  d->target = h->layout_type == IAMF_LAYOUT_TYPE_BINAURAL ? IAMFB_TARGET_BINAURAL : h->sound_system;
  d->loudness_gain = h->norm_loudness ? ih_db2lin(h->norm_loudness - h->loudness) : 0.f;
  d->limiter = h->limiter_on;
  d->limiter_threshold_db = h->threshold_db;
  d->bit_depth = h->bit_depth ? (int)h->bit_depth : 16; /* bit_depth 0: rendered but never copied out (:121-167) */
  {
    /* default: bit-identical to the reference.  IAMF_B200_ARITH=fma lets the HOA matrix and the resampler FIR fuse multiply
     * and add (half the FP32 work, PCM within +-1 LSB of the reference's; include/iamf_b200.h, IAMFB_ARITH_FMA). */
    const char *aenv = getenv("IAMF_B200_ARITH");
    d->arithmetic = (aenv && (!strcmp(aenv, "fma") || !strcmp(aenv, "1"))) ? IAMFB_ARITH_FMA : IAMFB_ARITH_EXACT;
  }
  h->frame_size = d->frame_size;

  /* handles are dealt round the process's devices in the order they are configured */
  engine_release(h);
  devices_init();
  pthread_mutex_lock(&g_shared_mu);
  const int dev = g_devs[g_next_dev++ % g_ndev];
  pthread_mutex_unlock(&g_shared_mu);
  return engine_bind(h, dev);
}

/* the handle's own single-stream batch and pinned frame buffers: created by its first IAMF_decoder_decode (handles that
 * step through IAMF_decoder_decode_batch never need them) */
static int engine_private(IAMF_DecoderHandle h) {
  if (h->batch) return IAMF_OK;
  if (!h->plan) return IAMF_ERR_INTERNAL;
  const iamfb_plan_desc *d = &h->desc;
  ih_shared *sh = (ih_shared *)h->shared;
  pthread_mutex_lock(&sh->mu);
  const int rc = iamfb_batch_create(h->plan, 1, 1, &h->batch);
  pthread_mutex_unlock(&sh->mu);
  if (rc != IAMFB_OK) return IAMF_ERR_INTERNAL;
  const size_t N = (size_t)d->frame_size;
  for (int e = 0; e < h->n_streams; ++e) {
    h->in[e] = (float *)iamfb_host_alloc(sizeof(float) * N * (size_t)d->el[e].n_in);
    h->ramp[e] = (iamfb_gain_ramp *)iamfb_host_alloc(sizeof(iamfb_gain_ramp));
    if (!h->in[e] || !h->ramp[e]) return IAMF_ERR_ALLOC_FAIL;
  }
  h->out_ramp = (iamfb_gain_ramp *)iamfb_host_alloc(sizeof(iamfb_gain_ramp));
  h->pcm_stage = (uint8_t *)iamfb_host_alloc(h->pcm_stage_size);
  h->fp_stage = (iamfb_frame_params *)iamfb_host_alloc(sizeof(iamfb_frame_params));
  h->counts_stage = (int32_t *)iamfb_host_alloc(sizeof(int32_t));
  if (!h->out_ramp || !h->pcm_stage || !h->fp_stage || !h->counts_stage) return IAMF_ERR_ALLOC_FAIL;
  return IAMF_OK;
}

/* iamf_extra_data_init, IAMF_decoder.c:3620-3665 */
static void metadata_init(IAMF_DecoderHandle h) {
  IAMF_extradata *m = &h->metadata;
  const ih_mix *mix = h->mix;
  free(m->loudness_layout); free(m->loudness); free(m->param);
  memset(m, 0, sizeof(*m));
  m->output_sound_system = h->layout_type == IAMF_LAYOUT_TYPE_LOUDSPEAKERS_SS_CONVENTION ? (IAMF_SoundSystem)h->sound_system
                                                                                         : SOUND_SYSTEM_INVALID;
  m->bitdepth = h->bit_depth;
  m->sampling_rate = IH_OUTPUT_RATE;
  /* iamf_presentation_get_output_sound_mode (IAMF_decoder.c:1555-1578): every element contributes a mode - a scene-based one
   * the mode of the OUTPUT layout (:358-370), a channel-based one the mode of its own selected layer layout - and the
   * modes combine (:241-254: equal -> that mode, binaural with anything else -> n/a, otherwise multichannel) */
  {
    int mode = IAMF_SOUND_MODE_NONE;
    for (int i = 0; i < h->n_streams; ++i) {
      const ih_stream *st = &h->streams[i];
      int sm;
      if (st->el->type != AUDIO_ELEMENT_CHANNEL_BASED)
        sm = h->layout_type == IAMF_LAYOUT_TYPE_BINAURAL ? IAMF_SOUND_MODE_BINAURAL
             : h->sound_system == SOUND_SYSTEM_A         ? IAMF_SOUND_MODE_STEREO
                                                         : IAMF_SOUND_MODE_MULTICHANNEL;
      else
        sm = (st->layout == IA_CHANNEL_LAYOUT_MONO || st->layout == IA_CHANNEL_LAYOUT_STEREO) ? IAMF_SOUND_MODE_STEREO
             : st->layout == IA_CHANNEL_LAYOUT_BINAURAL                                          ? IAMF_SOUND_MODE_BINAURAL
                                                                                                 : IAMF_SOUND_MODE_MULTICHANNEL;
      if (mode == IAMF_SOUND_MODE_NONE) mode = sm;
      else if (sm == IAMF_SOUND_MODE_NONE || sm == mode) { /* unchanged */ }
      else if (mode == IAMF_SOUND_MODE_BINAURAL || sm == IAMF_SOUND_MODE_BINAURAL) mode = IAMF_SOUND_MODE_NA;
      else mode = IAMF_SOUND_MODE_MULTICHANNEL;
    }
    m->output_sound_mode = (IAMF_SoundMode)mode;
  }
  m->num_loudness_layouts = mix->n_layouts;
  if (mix->n_layouts > 0) {
    m->loudness_layout = (IAMF_Layout *)calloc((size_t)mix->n_layouts, sizeof(IAMF_Layout));
    m->loudness = (IAMF_LoudnessInfo *)calloc((size_t)mix->n_layouts, sizeof(IAMF_LoudnessInfo));
    for (int i = 0; m->loudness_layout && m->loudness && i < mix->n_layouts; ++i) {
      m->loudness_layout[i].type = (uint8_t)mix->layouts[i].type;
      if (mix->layouts[i].type == IAMF_LAYOUT_TYPE_LOUDSPEAKERS_SS_CONVENTION)
        m->loudness_layout[i].sound_system.sound_system = (uint8_t)mix->layouts[i].sound_system;
      m->loudness[i] = mix->layouts[i].loud;
      m->loudness[i].anchor_loudness = 0;
    }
  }
  for (int i = 0; i < h->n_streams; ++i)
    if (h->streams[i].demix) {
      m->num_parameters = 1;
      m->param = (IAMF_Param *)calloc(1, sizeof(IAMF_Param));
      if (m->param) { m->param->parameter_length = 8; m->param->parameter_definition_type = IAMF_PARAMETER_TYPE_DEMIXING; }
      break;
    }
}

/* iamf_decoder_enable_mix_presentation, IAMF_decoder.c:3113-3203 */
static int enable_mix(IAMF_DecoderHandle h, const ih_mix *mix) {
  h->mix = mix;
  h->n_streams = 0;
  h->out_gain_item = 0;
  h->info.max_frame_size = 0;
  for (int i = 0; i < mix->n_elements; ++i) {
    ih_element *el = db_element(h, mix->el[i].element_id);
    ih_codec *cc = el ? db_codec(h, el->codec_id) : 0;
    if (!el || !cc) continue;
    /* the mix-gain parameter may be shared by several elements */
    ih_param_item *pi = ih_param_find(h, mix->el[i].gain_def.id);
    if (!pi) {
      pi = ih_param_add(h, &mix->el[i].gain_def, IH_INVALID_ID, cc->rate);
      if (pi) pi->default_gain = ih_db2lin(ih_q_to_float(mix->el[i].gain_q, 8));
    }
    if (h->n_streams >= IAMFB_MAX_ELEMENTS) return IAMF_ERR_UNIMPLEMENTED;
    ih_stream *st = &h->streams[h->n_streams];
    int rc = stream_setup(h, st, el, cc);
    if (rc != IAMF_OK) return rc;
    st->mix_gain = pi;
    for (int k = 0; k < el->n_params; ++k) {
      ih_param_item *p = ih_param_find(h, el->params[k].id);
      if (p && p->type == IAMF_PARAMETER_TYPE_DEMIXING && !st->demix) st->demix = p;
      if (p && p->type == IAMF_PARAMETER_TYPE_RECON_GAIN && !st->recon) st->recon = p;
    }
    /* iamf_stream_new: max_frame_size (IAMF_decoder.c:1628-1630) */
    uint32_t mfs = 1024 < cc->frame_size ? (uint32_t)cc->frame_size * 6 : 6144;
    /* (never less than what one call can really write: a frame resampled to the output rate, or the flush of the
     * limiter delay + resampler tail - more than 6 frames' worth only for ratios the reference's own buffers do not hold) */
    if (cc->rate > 0 && h->sampling_rate && (int)h->sampling_rate != cc->rate) {
      const uint64_t rs = ((uint64_t)cc->frame_size * h->sampling_rate + cc->rate - 1) / (uint64_t)cc->rate + 2;
      const uint64_t fl = 240 + 64 + 256ull * h->sampling_rate / (uint64_t)cc->rate;
      if (rs > mfs) mfs = (uint32_t)rs;
      if (fl > mfs) mfs = (uint32_t)fl;
    }
    if (mfs > h->info.max_frame_size) h->info.max_frame_size = mfs;
    ++h->n_streams;
  }
  if (h->n_streams <= 0) return IAMF_ERR_INTERNAL;
  for (int i = 1; i < h->n_streams; ++i)
    if (h->streams[i].cc->frame_size != h->streams[0].cc->frame_size || h->streams[i].cc->rate != h->streams[0].cc->rate)
      return IAMF_ERR_UNIMPLEMENTED; /* the mixer needs equal frame sizes (IAMF_decoder.c:2702-2717) */
  ih_param_item *po = ih_param_find(h, mix->out_def.id);
  if (!po) po = ih_param_add(h, &mix->out_def, IH_INVALID_ID, (int)h->sampling_rate);
  if (po) {
    po->default_gain = ih_db2lin(ih_q_to_float(mix->out_q, 8));
    h->out_gain_item = po;
  }
  return IAMF_OK;
}

/* ------------------------------------------------------------------ open / close / setters ---- */
IAMF_DecoderHandle IAMF_decoder_open(void) {
  IAMF_DecoderHandle h = (IAMF_DecoderHandle)calloc(1, sizeof(struct IAMF_Decoder));
  if (!h) return 0;
  h->threshold_db = IH_LIMITER_DEFAULT_DB;
  h->loudness = 1.0f;
  h->sampling_rate = IH_OUTPUT_RATE;
  h->status = IH_STATUS_INIT;
  h->mix_id = IH_INVALID_ID;
  h->limiter_on = 1;
  h->group_owner = 1;
  h->group_size = 1;
  h->metadata.output_sound_mode = IAMF_SOUND_MODE_NONE;
  return h;
}

int IAMF_decoder_close(IAMF_DecoderHandle h) {
  if (h) {
    db_reset(h);
    engine_release(h);
    free(h);
  }
  return 0;
}

int IAMF_decoder_output_layout_set_sound_system(IAMF_DecoderHandle h, IAMF_SoundSystem ss) {
  if (!h || !ss_valid(ss)) return IAMF_ERR_BAD_ARG;
  if (h->layout_type == IAMF_LAYOUT_TYPE_LOUDSPEAKERS_SS_CONVENTION && h->sound_system == (int)ss) return IAMF_OK;
  h->layout_type = IAMF_LAYOUT_TYPE_LOUDSPEAKERS_SS_CONVENTION;
  h->sound_system = ss;
  h->need_configure |= IH_NEED_LAYOUT;
  return IAMF_OK;
}

int IAMF_decoder_output_layout_set_binaural(IAMF_DecoderHandle h) {
  if (!h) return IAMF_ERR_BAD_ARG;
  if (h->layout_type == IAMF_LAYOUT_TYPE_BINAURAL) return IAMF_OK;
  h->layout_type = IAMF_LAYOUT_TYPE_BINAURAL;
  h->need_configure |= IH_NEED_LAYOUT;
  return IAMF_OK;
}

int IAMF_decoder_set_mix_presentation_id(IAMF_DecoderHandle h, uint64_t id) {
  if (!h) return IAMF_ERR_BAD_ARG;
  if (h->mix_id == id) return IAMF_OK;
  h->mix_id = id;
  h->need_configure |= IH_NEED_MIX;
  return IAMF_OK;
}

char *IAMF_decoder_get_codec_capability(void) {
  /* "iamf.<primary>.<additional>.<codec>" list; what this build can core-decode (IAMF_decoder.c:4010-4071) */
  char *list = (char *)calloc(1024, 1);
  if (!list) return 0;
  if (ih_codec_supported(IAMF_CODEC_OPUS)) strcat(list, "iamf.001.001.Opus;");
  strcat(list, "iamf.001.001.ipcm");
  if (ih_codec_supported(IAMF_CODEC_FLAC)) strcat(list, ";iamf.001.001.fLaC");
  return list;
}

int IAMF_decoder_set_normalization_loudness(IAMF_DecoderHandle h, float loudness) {
  if (!h) return IAMF_ERR_BAD_ARG;
  h->norm_loudness = loudness;
  return IAMF_OK;
}
int IAMF_decoder_set_bit_depth(IAMF_DecoderHandle h, uint32_t bit_depth) {
  if (!h) return IAMF_ERR_BAD_ARG;
  h->bit_depth = bit_depth;
  return IAMF_OK;
}
int IAMF_decoder_peak_limiter_enable(IAMF_DecoderHandle h, uint32_t enable) {
  if (!h) return IAMF_ERR_BAD_ARG;
  h->limiter_on = enable ? 1 : 0;
  return IAMF_OK;
}
int IAMF_decoder_peak_limiter_set_threshold(IAMF_DecoderHandle h, float db) {
  if (!h) return IAMF_ERR_BAD_ARG;
  h->threshold_db = db;
  return IAMF_OK;
}
float IAMF_decoder_peak_limiter_get_threshold(IAMF_DecoderHandle h) { return h ? h->threshold_db : IH_LIMITER_DEFAULT_DB; }

int IAMF_decoder_set_sampling_rate(IAMF_DecoderHandle h, uint32_t rate) {
  static const uint32_t rates[] = {8000, 12000, 16000, 24000, 32000, 44100, 48000};
  if (!h) return IAMF_ERR_BAD_ARG;
  if (h->status != IH_STATUS_INIT) return IAMF_ERR_INVALID_STATE;
  for (unsigned i = 0; i < sizeof(rates) / sizeof(rates[0]); ++i)
    if (rates[i] == rate) { h->sampling_rate = rate; return IAMF_OK; }
  return IAMF_ERR_BAD_ARG;
}

IAMF_StreamInfo *IAMF_decoder_get_stream_info(IAMF_DecoderHandle h) { return &h->info; }

int IAMF_decoder_set_pts(IAMF_DecoderHandle h, int64_t pts, uint32_t time_base) {
  if (!h) return IAMF_ERR_BAD_ARG;
  h->pts = pts;
  h->pts_time_base = time_base;
  h->duration = 0;
  return IAMF_OK;
}

int IAMF_decoder_get_last_metadata(IAMF_DecoderHandle h, int64_t *pts, IAMF_extradata *md) {
  if (!h || !pts || !md) return IAMF_ERR_BAD_ARG;
  *pts = h->pts + ih_time_transform((int64_t)h->duration - h->last_frame_size, (int)h->sampling_rate, (int)h->pts_time_base);
  const IAMF_extradata *src = &h->metadata;
  *md = *src;
  md->loudness_layout = 0; md->loudness = 0; md->param = 0;
  if (src->num_loudness_layouts > 0) {
    md->loudness_layout = (IAMF_Layout *)calloc((size_t)src->num_loudness_layouts, sizeof(IAMF_Layout));
    md->loudness = (IAMF_LoudnessInfo *)calloc((size_t)src->num_loudness_layouts, sizeof(IAMF_LoudnessInfo));
    if (!md->loudness_layout || !md->loudness) return IAMF_ERR_ALLOC_FAIL;
    memcpy(md->loudness_layout, src->loudness_layout, sizeof(IAMF_Layout) * (size_t)src->num_loudness_layouts);
    memcpy(md->loudness, src->loudness, sizeof(IAMF_LoudnessInfo) * (size_t)src->num_loudness_layouts);
  }
  if (src->num_parameters) {
    md->param = (IAMF_Param *)calloc(src->num_parameters, sizeof(IAMF_Param));
    if (!md->param) return IAMF_ERR_ALLOC_FAIL;
    memcpy(md->param, src->param, sizeof(IAMF_Param) * src->num_parameters);
  }
  md->number_of_samples = (uint32_t)h->last_frame_size;
  return IAMF_OK;
}

/* ------------------------------------------------------------------ configure ---- */
/* iamf_decoder_internal_configure, IAMF_decoder.c:3759-3911 */
static int internal_configure(IAMF_DecoderHandle h, const uint8_t *data, uint32_t size, uint32_t *rsize) {
  int ret = IAMF_OK;
  if (!h) return IAMF_ERR_BAD_ARG;
  if (h->need_configure & IH_NEED_LAYOUT) {
    if (h->layout_type == IAMF_LAYOUT_TYPE_LOUDSPEAKERS_SS_CONVENTION)
      h->out_channels = IAMF_layout_sound_system_channels_count((IAMF_SoundSystem)h->sound_system);
    else if (h->layout_type == IAMF_LAYOUT_TYPE_BINAURAL) h->out_channels = 2;
    h->have_layout = 1;
  }
  if (data && size > 0) {
    if (h->status == IH_STATUS_INIT) h->status = IH_STATUS_CONFIGURE;
    else if (h->status == IH_STATUS_RECEIVE) h->status = IH_STATUS_RECONFIGURE;
    if (h->status == IH_STATUS_RECONFIGURE) {
      db_reset(h);
      h->status = IH_STATUS_CONFIGURE;
    }
    if (!h->have_layout) return IAMF_ERR_INTERNAL; /* the reference dereferences a null output layout here */
    ret = internal_init(h, data, size, rsize);
    if (ret == IAMF_OK) h->need_configure = 0;
  } else if (h->need_configure) {
    if (h->status < IH_STATUS_RECEIVE) return IAMF_ERR_BAD_ARG;
    if ((h->need_configure & IH_NEED_MIX) && (!h->mix || (h->mix_id != h->mix->id && !db_mix(h, h->mix_id)))) ret = IAMF_ERR_INTERNAL;
    h->need_configure = 0;
  } else {
    return IAMF_ERR_BAD_ARG;
  }
  if (ret == IAMF_OK) {
    const ih_mix *mix = best_mix(h);
    if (!mix) return IAMF_ERR_INVALID_PACKET;
    ret = enable_mix(h, mix);
    if (ret == IAMF_OK) {
      h->loudness = best_loudness(h, mix);
      ret = engine_build(h);
    }
    if (ret == IAMF_OK) {
      metadata_init(h);
      h->status = IH_STATUS_RECEIVE;
    }
    for (int i = 0; i < h->n_params; ++i) { /* iamf_database_parameters_clear_segments */
      ih_param_clear(&h->params[i]);
      h->params[i].duration = 0;
    }
  }
  return ret;
}

int IAMF_decoder_configure(IAMF_DecoderHandle h, const uint8_t *data, uint32_t size, uint32_t *rsize) {
  uint32_t rs = 0;
  int ret = internal_configure(h, data, size, &rs);
  if (rsize) { *rsize = rs; return ret; }
  if (ret == IAMF_ERR_BUFFER_TOO_SMALL && !(~h->flags & IH_FLAG_DESCRIPTORS)) {
    /* rsize == NULL: the buffer is the complete descriptor set (IAMF_decoder.c:3913-3933) */
    h->flags |= IH_FLAG_CONFIG;
    h->need_configure = IH_NEED_PRESENTATION;
    h->status = IH_STATUS_RECEIVE;
    ret = internal_configure(h, 0, 0, 0);
  }
  return ret;
}

/* ------------------------------------------------------------------ decode ---- */
static ih_stream *stream_of_element(IAMF_DecoderHandle h, uint64_t eid) {
  for (int i = 0; i < h->n_streams; ++i)
    if (h->streams[i].el->id == eid) return &h->streams[i];
  return 0;
}

/* iamf_stream_decoder_update_parameter, IAMF_decoder.c:2131-2151 */
static void stream_update_parameter(ih_stream *st, const ih_param_item *pi) {
  const uint64_t pts = st->timestamp + st->cc->frame_size / 2;
  const ih_segment *seg = ih_param_segment_at(pi, pts);
  if (pi->type == IAMF_PARAMETER_TYPE_DEMIXING) {
    st->dmx_mode = seg ? seg->dmx_mode : IAMF_ERR_INTERNAL;
  } else if (pi->type == IAMF_PARAMETER_TYPE_RECON_GAIN && seg) {
    /* only the list of the selected layer reaches the de-mixer (IAMF_decoder.c:2324-2343) */
    if (st->el->layers[st->layer].recon_present) {
      st->has_recon = 1;
      st->recon_flags = seg->rg[st->layer].flags;
      memcpy(st->recon_q, seg->rg[st->layer].q, sizeof(st->recon_q));
    }
  }
}

/* iamf_decoder_internal_parse_OBUs, IAMF_decoder.c:2871-2944.  returns bytes consumed; *run = every stream has a
 * packet for each of its sub-streams */
/* a sub-stream packet is either owned (malloc'ed copy: the caller's buffer need not outlive the call, IAMF_decoder.c:2113-2120)
 * or - inside a batch step, until the step ends - borrowed from the caller's buffer (bit k of pkt_borrowed) */
static void pkt_drop(ih_stream *st, int k) {
  if (!((st->pkt_borrowed >> k) & 1u)) free(st->pkt[k]);
  st->pkt_borrowed &= ~(1u << k);
  st->pkt[k] = 0;
  st->pkt_size[k] = 0;
}
/* packets still borrowed when a batch step ends (its buffer stopped inside a temporal unit) become owned copies */
static void pkt_own_all(ih_stream *st) {
  for (int k = 0; k < IH_MAX_SUBSTREAMS; ++k)
    if ((st->pkt_borrowed >> k) & 1u) {
      uint8_t *c = (uint8_t *)malloc(st->pkt_size[k] ? st->pkt_size[k] : 1);
      if (c) memcpy(c, st->pkt[k], st->pkt_size[k]);
      st->pkt[k] = c;
      st->pkt_borrowed &= ~(1u << k);
    }
}

static uint32_t parse_obus(IAMF_DecoderHandle h, const uint8_t *data, uint32_t size, int *run, int borrow) {
  uint32_t pos = 0;
  ih_obu o;
  *run = 0;
  while (pos < size) {
    uint32_t used = ih_obu_split(data + pos, size - pos, &o);
    if (!used) break;
    if (o.type == IH_OBU_PARAMETER_BLOCK) {
      const uint64_t pid = ih_obu_parameter_id(&o);
      ih_param_item *pi = ih_param_find(h, pid);
      if (pi) {
        int n_layers = 0, nseg = 0;
        unsigned recon_flags = 0;
        const ih_element *e = pi->parent != IH_INVALID_ID ? db_element(h, pi->parent) : 0;
        if (e && e->type == AUDIO_ELEMENT_CHANNEL_BASED) {
          n_layers = e->n_layers;
          for (int i = 0; i < e->n_layers; ++i)
            if (e->layers[i].recon_present) recon_flags |= 1u << i;
        }
        uint64_t id2;
        ih_segment *segs = ih_parse_parameter_block(&o, &id2, pi->def, n_layers, recon_flags, &nseg);
        if (o.redundant && pi->duration > 0) { /* iamf_database_parameter_add, IAMF_decoder.c:1045-1075 */
          while (segs) { ih_segment *n = segs->next; free(segs); segs = n; }
        } else {
          ih_param_push(pi, segs);
        }
        ih_stream *st = e ? stream_of_element(h, e->id) : 0;
        if (st) stream_update_parameter(st, pi);
      }
    } else if (o.type >= IH_OBU_AUDIO_FRAME && o.type <= IH_OBU_AUDIO_FRAME_ID17) {
      ih_reader r;
      uint64_t sid;
      ih_rd_init(&r, o.payload, o.payload_size);
      if (o.type == IH_OBU_AUDIO_FRAME) sid = ih_rd_leb128(&r);
      else sid = (uint64_t)(o.type - IH_OBU_AUDIO_FRAME_ID0);
      const uint8_t *payload = o.payload + ih_rd_tell(&r);
      const uint32_t psize = o.payload_size - ih_rd_tell(&r);
      for (int i = 0; i < h->n_streams; ++i) {   /* iamf_decoder_internal_deliver, IAMF_decoder.c:2946-2995 */
        ih_stream *st = &h->streams[i];
        int idx = -1;
        for (int k = 0; k < st->el->n_sub; ++k)
          if (st->el->sub_ids[k] == sid) { idx = k; break; }
        if (idx < 0) continue;
        if (idx == 0) {
          st->trimming_start = o.trim_start;
          st->trimming_end = o.trim_end;
          st->strim = o.trim_start;
          st->etrim = o.trim_end;
        }
        if (!st->pkt[idx]) ++st->pkt_count;
        pkt_drop(st, idx);
        if (borrow) {
          st->pkt[idx] = (uint8_t *)(uintptr_t)payload;
          st->pkt_borrowed |= 1u << idx;
        } else {
          st->pkt[idx] = (uint8_t *)malloc(psize ? psize : 1);
          if (st->pkt[idx]) memcpy(st->pkt[idx], payload, psize);
        }
        st->pkt_size[idx] = psize;
        break;
      }
      int all = 1;
      for (int i = 0; i < h->n_streams; ++i)
        if (h->streams[i].pkt_count != h->streams[i].el->n_sub) all = 0;
      if (all) { h->status = IH_STATUS_RUN; *run = 1; }
    } else if (o.type == IH_OBU_SEQUENCE_HEADER && !o.redundant) {
      h->status = IH_STATUS_RECONFIGURE;
      break;
    }
    pos += used;
    if (h->status == IH_STATUS_RUN) break;
  }
  return pos;
}

/* Host part of one temporal unit: core decode of every element into `in[e]`, per-frame parameters into *fp.
 * returns >0 frame ready (samples entering the engine after trimming), 0 dropped / nothing, <0 error.
 * (the loop body of iamf_decoder_internal_decode, IAMF_decoder.c:3336-3457, up to where samples are touched) */
static int prepare_frame(IAMF_DecoderHandle h, void *const in[], int s16, iamfb_frame_params *fp, iamfb_gain_ramp *const ramp[],
                         iamfb_gain_ramp *out_ramp, int *use_ramp, int *use_out_ramp) {
  const int N = h->frame_size;
  int lret = 1, real = 0;
  uint64_t frame_pts = 0;
  memset(fp, 0, sizeof(*fp));
  fp->out_gain = 1.0f;
  for (int e = 0; e < IAMFB_MAX_ELEMENTS; ++e) {
    fp->el[e].dmx_mode = -1; fp->el[e].mix_gain = 1.0f; use_ramp[e] = 0;
    if (ramp[e]) ramp[e]->n_segs = 0;
  }
  *use_out_ramp = 0;
  out_ramp->n_segs = 0;
  for (int s = 0; s < h->n_streams; ++s) {
    ih_stream *st = &h->streams[s];
    const ih_element *el = st->el;
    const uint64_t pts = st->timestamp;
    int ret;
    if (s == 0) frame_pts = pts;
    /* core decode: layers 0..layer of a channel-based element one after the other, each with its own coupled count
     * (iamf_stream_scale_decoder_decode :2276-2322); all sub-streams of a scene-based one (:2415-2446) */
    if (el->type == AUDIO_ELEMENT_CHANNEL_BASED) {
      int sub = 0, ch = 0;
      ret = 0;
      for (int l = 0; l <= st->layer; ++l) {
        const ih_layer *ly = &el->layers[l];
        ret = s16 ? ih_codec_decode_s16(st, sub, st->cc, &st->pkt[sub], &st->pkt_size[sub], ly->n_sub, ly->n_coupled, (int16_t *)in[s] + (size_t)ch * N, N)
                  : ih_codec_decode(st, sub, st->cc, &st->pkt[sub], &st->pkt_size[sub], ly->n_sub, ly->n_coupled, (float *)in[s] + (size_t)ch * N, N);
        if (ret < 0) break;
        sub += ly->n_sub;
        ch += ly->n_sub + ly->n_coupled;
      }
    } else {
      ret = s16 ? ih_codec_decode_s16(st, 0, st->cc, st->pkt, st->pkt_size, el->ambi_sub, el->ambi_coupled, (int16_t *)in[s], N)
                : ih_codec_decode(st, 0, st->cc, st->pkt, st->pkt_size, el->ambi_sub, el->ambi_coupled, (float *)in[s], N);
    }
    for (int k = 0; k < IH_MAX_SUBSTREAMS; ++k) pkt_drop(st, k);
    st->pkt_count = 0;
    if (ret > 0 && ret != N) ret = IAMF_ERR_INTERNAL; /* short frames (frame_padding) need a codec with delay: not supported */

    if (s < IAMFB_MAX_ELEMENTS) {
      fp->el[s].dmx_mode = (int8_t)(st->dmx_mode > -1 ? st->dmx_mode : -1);
      fp->el[s].has_recon = (uint8_t)st->has_recon;
      fp->el[s].recon_flags = (uint16_t)st->recon_flags;
      memcpy(fp->el[s].recon_gain, st->recon_q, 12);
      st->has_recon = 0;
    }
    if (s == 0) {
      fp->trim_start = (uint16_t)st->strim;
      fp->trim_end = (uint16_t)st->etrim;
    }
    int samples = ret;
    if (ret > 0) {
      if ((int)st->strim == N || (int)st->etrim == N) samples = 0; /* whole frame cut (:3354-3358) */
      else samples = N - (int)st->strim - (int)st->etrim;
      if (samples < 0) samples = 0;
    }
    if (!s && st->strim > 0) h->pts += ih_time_transform((int64_t)st->strim, st->cc->rate, (int)h->pts_time_base);
    if (samples <= 0) {
      st->timestamp += (uint64_t)N;
      lret = ret < 0 ? ret : 0;
      continue;
    }
    real = samples;
    if (st->mix_gain) { /* :3425-3433 */
      float g = 1.f;
      /* the gain time line is read at the frame's pts AFTER trimming: iamf_frame_trim does f->pts += strim (:1379)
       * before get_mix_gain_unit(f->pts, f->samples) (:3425-3427) */
      int kind = ih_mix_gain_unit(st->mix_gain, pts + (uint64_t)st->strim, samples, st->cc->rate, &g, ramp[s]);
      if (kind == 1) fp->el[s].mix_gain = g;
      else if (kind == 2) use_ramp[s] = 1;
      else if (kind < 0) lret = IAMF_ERR_UNIMPLEMENTED;   /* a frame spanning more than 8 gain segments */
    }
    if (el->type == AUDIO_ELEMENT_CHANNEL_BASED && h->metadata.param && st->dmx_mode >= 0)
      h->metadata.param->dmixp_mode = (uint32_t)st->dmx_mode;
    st->timestamp += (uint64_t)N;
  }
  if (lret <= 0) {
    /* the frame is decoded (de-mixer state advances on the device) but nothing is mixed or returned */
    return lret < 0 ? lret : 0;
  }
  if (h->out_gain_item) { /* :3463-3469 */
    float g = 1.f;
    /* the mixed frame carries the (trimmed) pts of the first element's frame (iamf_mixer_mix :2702-2733) */
    int kind = ih_mix_gain_unit(h->out_gain_item, frame_pts + (uint64_t)h->streams[0].strim, real, h->streams[0].cc->rate, &g, out_ramp);
    if (kind == 1) fp->out_gain = g;
    else if (kind == 2) *use_out_ramp = 1;
    else if (kind < 0) return IAMF_ERR_UNIMPLEMENTED;
  }
  ih_params_elapse(h, (uint64_t)real, (uint32_t)h->streams[0].cc->rate);
  return real;
}

static size_t pcm_bytes(IAMF_DecoderHandle h, int samples) { return (size_t)samples * (size_t)h->out_channels * (h->bit_depth / 8); }

/* iamf_decoder_internal_decode, IAMF_decoder.c:3303-3525 */
int IAMF_decoder_decode(IAMF_DecoderHandle h, const uint8_t *data, int32_t size, uint32_t *rsize, void *pcm) {
  if (!h) return IAMF_ERR_BAD_ARG;
  if (h->status != IH_STATUS_RECEIVE) return IAMF_ERR_INVALID_STATE;
  if (h->leader || h->group_size > 1) return IAMF_ERR_INVALID_STATE; /* grouped handles step through decode_batch */
  if (rsize) *rsize = 0;
  if (h->n_streams <= 0 || !h->plan) return IAMF_ERR_INTERNAL;
  {
    const int rc = engine_private(h);
    if (rc != IAMF_OK) return rc;
  }
  ih_shared *sh = (ih_shared *)h->shared;
  int real = 0;
  if (data && size > 0) {
    int run = 0;
    uint32_t used = parse_obus(h, data, (uint32_t)size, &run, 0);
    if (rsize) *rsize = used;
    if (h->status == IH_STATUS_RECONFIGURE) return IAMF_ERR_INVALID_STATE;
    if (h->status != IH_STATUS_RUN) return 0;
    int use_ramp[IAMFB_MAX_ELEMENTS], use_out_ramp = 0;
    void *inp[IAMFB_MAX_ELEMENTS] = {h->in[0], h->in[1]};
    int ready = prepare_frame(h, inp, 0, h->fp_stage, h->ramp, h->out_ramp, use_ramp, &use_out_ramp);
    iamfb_io io;
    memset(&io, 0, sizeof(io));
    for (int e = 0; e < h->n_streams; ++e) {
      io.in[e] = h->in[e];
      if (use_ramp[e]) io.gain_segs[e] = h->ramp[e];
    }
    if (use_out_ramp) io.out_gain_segs = h->out_ramp;
    if (ready <= 0) {
      /* dropped frame: keep the device-side parameter state in step (trim == frame size), return what decode did */
      if (ready == 0) {
        h->fp_stage->trim_start = (uint16_t)h->frame_size;
        h->fp_stage->trim_end = 0;
      }
    }
    if (ready >= 0) {
      io.params = h->fp_stage;
      io.pcm = h->pcm_stage;
      io.out_counts = h->counts_stage;
      pthread_mutex_lock(&sh->mu);
      const int rc = iamfb_batch_submit_host(h->batch, &io, 1);
      pthread_mutex_unlock(&sh->mu);
      if (rc != IAMFB_OK) { h->status = IH_STATUS_RECEIVE; return IAMF_ERR_INTERNAL; }
      real = h->counts_stage[0];
    }
    if (ready <= 0) {
      h->status = IH_STATUS_RECEIVE;
      return ready;
    }
    if (real > 0 && pcm && h->bit_depth) memcpy(pcm, h->pcm_stage, pcm_bytes(h, real));
  }
  if (!data) { /* iamf_delay_buffer_handle, :3250-3301 */
    pthread_mutex_lock(&sh->mu);
    const int rc = iamfb_batch_flush_host(h->batch, h->pcm_stage, h->counts_stage);
    pthread_mutex_unlock(&sh->mu);
    if (rc != IAMFB_OK) return IAMF_ERR_INTERNAL;
    real = h->counts_stage[0];
    if (real > 0 && pcm && h->bit_depth) memcpy(pcm, h->pcm_stage, pcm_bytes(h, real));
  }
  h->duration += (uint64_t)real;
  h->last_frame_size = real;
  h->status = IH_STATUS_RECEIVE;
  return real;
}

/* ------------------------------------------------------------------ additive batch extension ---- */
static int same_signature(const IAMF_DecoderHandle a, const IAMF_DecoderHandle b) {
  return memcmp(&a->desc, &b->desc, sizeof(a->desc)) == 0;
}

/* ---- a small pool of host threads for the per-handle part of a batch step (bitstream parsing + core decode of the
 * handles are independent of each other; they write disjoint slots of the group's pinned buffers) */
#include <pthread.h>
#include <unistd.h>
typedef struct {
  pthread_mutex_t mu;
  pthread_cond_t cv_work, cv_done;
  pthread_t th[64];
  int n_threads, started;
  unsigned long gen;
  void (*fn)(void *, int);
  void *arg;
  int next, total, pending;
} ih_pool;
static ih_pool g_pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER};
static pthread_mutex_t g_pool_call = PTHREAD_MUTEX_INITIALIZER;   /* one parallel-for at a time */

static void pool_drain(ih_pool *p) {   /* called with p->mu held */
  while (p->next < p->total) {
    const int lo = p->next;
    int hi = lo + 4;                   /* a few handles per grab */
    if (hi > p->total) hi = p->total;
    p->next = hi;
    pthread_mutex_unlock(&p->mu);
    for (int i = lo; i < hi; ++i) p->fn(p->arg, i);
    pthread_mutex_lock(&p->mu);
    p->pending -= hi - lo;
  }
}
static void *pool_worker(void *v) {
  ih_pool *p = (ih_pool *)v;
  unsigned long seen = 0;
  pthread_mutex_lock(&p->mu);
  for (;;) {
    while (p->gen == seen) pthread_cond_wait(&p->cv_work, &p->mu);
    seen = p->gen;
    pool_drain(p);
    if (p->pending == 0) pthread_cond_signal(&p->cv_done);
  }
  return 0;
}
static void pool_for(int total, void (*fn)(void *, int), void *arg) {
  ih_pool *p = &g_pool;
  int want = 1;
  const char *env = getenv("IAMF_B200_HOST_THREADS");
  if (env) want = atoi(env);
  else {
    long nc = sysconf(_SC_NPROCESSORS_ONLN);
    want = nc > 1 ? (int)nc : 1;
  }
  if (want > 64) want = 64;
  if (want <= 1 || total < 8) {
    for (int i = 0; i < total; ++i) fn(arg, i);
    return;
  }
  pthread_mutex_lock(&g_pool_call);
  pthread_mutex_lock(&p->mu);
  while (p->started < want - 1) {      /* the calling thread works too */
    if (pthread_create(&p->th[p->started], 0, pool_worker, p) != 0) break;
    pthread_detach(p->th[p->started]);
    ++p->started;
  }
  p->fn = fn; p->arg = arg; p->next = 0; p->total = total; p->pending = total;
  ++p->gen;
  pthread_cond_broadcast(&p->cv_work);
  pool_drain(p);
  while (p->pending > 0) pthread_cond_wait(&p->cv_done, &p->mu);
  pthread_mutex_unlock(&p->mu);
  pthread_mutex_unlock(&g_pool_call);
}

static int group_build(IAMF_DecoderHandle *hs, int n, int units) {
  IAMF_DecoderHandle L = hs[0];
  int s16 = 1;
  for (int i = 0; i < n; ++i) {
    if (!hs[i] || hs[i]->status != IH_STATUS_RECEIVE || !hs[i]->plan) return IAMF_ERR_INVALID_STATE;
    if (!same_signature(L, hs[i])) return IAMF_ERR_BAD_ARG;
    if (hs[i]->duration || hs[i]->leader) return IAMF_ERR_INVALID_STATE; /* group before the first decode call */
    for (int e = 0; e < hs[i]->n_streams; ++e)
      if (!ih_codec_is_s16(hs[i]->streams[e].cc)) s16 = 0;
  }
  /* the leader's context / plan serve the whole group; every member's private single-stream engine is released */
  const size_t N = (size_t)L->frame_size, F = (size_t)units;
  const size_t esz = s16 ? sizeof(int16_t) : sizeof(float);
  iamfb_batch *gb = 0;
  ih_shared *sh = (ih_shared *)L->shared;
  pthread_mutex_lock(&sh->mu);
  const int brc = iamfb_batch_create(L->plan, n, units, &gb);
  if (brc == IAMFB_OK && L->batch) iamfb_batch_destroy(L->batch);
  pthread_mutex_unlock(&sh->mu);
  if (brc != IAMFB_OK) return IAMF_ERR_INTERNAL;
  L->batch = gb;
  for (int e = 0; e < L->n_streams; ++e) {
    iamfb_host_free(L->in[e]); iamfb_host_free(L->ramp[e]);
    L->in[e] = (float *)iamfb_host_alloc(esz * N * (size_t)L->desc.el[e].n_in * (size_t)n * F);
    L->ramp[e] = (iamfb_gain_ramp *)iamfb_host_alloc(sizeof(iamfb_gain_ramp) * (size_t)n * F);
    if (!L->in[e] || !L->ramp[e]) return IAMF_ERR_ALLOC_FAIL;
  }
  iamfb_host_free(L->out_ramp); iamfb_host_free(L->pcm_stage); iamfb_host_free(L->fp_stage); iamfb_host_free(L->counts_stage);
  L->pcm_stage_size = iamfb_plan_out_stride_bytes(L->plan, units);
  L->out_ramp = (iamfb_gain_ramp *)iamfb_host_alloc(sizeof(iamfb_gain_ramp) * (size_t)n * F);
  L->pcm_stage = (uint8_t *)iamfb_host_alloc(L->pcm_stage_size * (size_t)n);
  L->fp_stage = (iamfb_frame_params *)iamfb_host_alloc(sizeof(iamfb_frame_params) * (size_t)n * F);
  L->counts_stage = (int32_t *)iamfb_host_alloc(sizeof(int32_t) * (size_t)n * F);
  if (!L->out_ramp || !L->pcm_stage || !L->fp_stage || !L->counts_stage) return IAMF_ERR_ALLOC_FAIL;
  L->group_size = n;
  L->group_index = 0;
  L->group_units = units;
  L->group_s16 = s16;
  for (int i = 1; i < n; ++i) {
    engine_release(hs[i]);
    hs[i]->group_owner = 0;
    hs[i]->leader = L;
    hs[i]->group_size = n;
    hs[i]->group_index = i;
    hs[i]->frame_size = L->frame_size;
  }
  return IAMF_OK;
}

typedef struct {
  IAMF_DecoderHandle *hs;
  const uint8_t *const *data;
  const int32_t *size;
  int units;
  int base;                      /* pool index 0 = handle `base` */
} ih_step_job;

/* phase 1 of a batch step for handle i: its temporal units of this call, parsed and core-decoded into the handle's slots
 * [i][f] of the group's pinned buffers */
static void step_handle(void *v, int idx) {
  const ih_step_job *job = (const ih_step_job *)v;
  const int i = job->base + idx;
  IAMF_DecoderHandle h = job->hs[i], L = job->hs[0];
  const size_t N = (size_t)L->frame_size;
  const int F = job->units, s16 = L->group_s16;
  const size_t esz = s16 ? sizeof(int16_t) : sizeof(float);
  uint32_t pos = 0;
  h->unit_used = 0;
  h->units_done = 0;
  for (int f = 0; f < F; ++f) {
    iamfb_frame_params *fp = &L->fp_stage[(size_t)i * F + f];
    memset(fp, 0, sizeof(*fp));
    fp->trim_start = 0xFFFF; /* "no frame for this stream in this step": the engine leaves its state untouched */
    h->unit_ret[f] = 0;
    h->unit_flags[f] = 0;
  }
  if (h->status != IH_STATUS_RECEIVE) { h->unit_ret[0] = IAMF_ERR_INVALID_STATE; return; }
  for (int f = 0; f < F && pos < (uint32_t)job->size[i]; ++f) {
    iamfb_frame_params *fp = &L->fp_stage[(size_t)i * F + f];
    void *in[IAMFB_MAX_ELEMENTS] = {0, 0};
    iamfb_gain_ramp *ramp[IAMFB_MAX_ELEMENTS] = {0, 0};
    for (int e = 0; e < L->n_streams; ++e) {
      in[e] = (uint8_t *)L->in[e] + ((size_t)i * F + f) * N * (size_t)L->desc.el[e].n_in * esz;
      ramp[e] = L->ramp[e] + ((size_t)i * F + f);
    }
    int run = 0, use_ramp[IAMFB_MAX_ELEMENTS] = {0, 0}, use_out = 0;
    uint32_t used = parse_obus(h, job->data[i] + pos, (uint32_t)job->size[i] - pos, &run, 1);
    pos += used;
    h->unit_used = pos;
    if (h->status == IH_STATUS_RECONFIGURE) { h->unit_ret[f] = IAMF_ERR_INVALID_STATE; break; }
    if (h->status != IH_STATUS_RUN) break;      /* the buffer ended inside a temporal unit: the rest comes with the next call */
    int ready = prepare_frame(h, in, s16, fp, ramp, L->out_ramp + ((size_t)i * F + f), use_ramp, &use_out);
    h->status = IH_STATUS_RECEIVE;
    ++h->units_done;
    if (ready < 0) { h->unit_ret[f] = ready; fp->trim_start = 0xFFFF; continue; }
    if (ready == 0) { fp->trim_start = (uint16_t)N; fp->trim_end = 0; }
    h->unit_ret[f] = ready > 0 ? 1 : 0;
    h->unit_flags[f] = (uint8_t)((use_ramp[0] ? 1 : 0) | (use_ramp[1] ? 2 : 0) | (use_out ? 4 : 0));
    if (!used) break;
  }
  for (int s = 0; s < h->n_streams; ++s) pkt_own_all(&h->streams[s]);
}

/* phase 3 of a batch step for handle i: its samples into the caller's buffer */
typedef struct {
  IAMF_DecoderHandle *hs;
  void *const *pcm;
  int *ret;
  uint32_t *rsize;
  int *units_done;
  int units, flush;
  int base;
} ih_out_job;
static size_t pcm_bytes(IAMF_DecoderHandle h, int samples);
static void out_handle(void *v, int idx) {
  const ih_out_job *job = (const ih_out_job *)v;
  const int i = job->base + idx;
  IAMF_DecoderHandle h = job->hs[i], L = job->hs[0];
  const int F = job->units;
  if (job->rsize) job->rsize[i] = job->flush ? 0 : h->unit_used;
  if (job->units_done) job->units_done[i] = job->flush ? 0 : h->units_done;
  int real = 0, err = 0;
  if (job->flush) real = L->counts_stage[i];
  else
    for (int f = 0; f < F; ++f) {
      if (h->unit_ret[f] < 0 && !err) err = h->unit_ret[f];
      if (h->unit_ret[f] > 0) {
        const int c = L->counts_stage[(size_t)i * F + f];
        real += c;
        h->last_frame_size = c;
      }
    }
  /* (a flush returns its rows at the pitch of a one-frame submit) */
  const size_t pitch = job->flush ? iamfb_plan_out_stride_bytes(L->plan, 1) : L->pcm_stage_size;
  if (real > 0 && job->pcm && job->pcm[i] && h->bit_depth) memcpy(job->pcm[i], L->pcm_stage + (size_t)i * pitch, pcm_bytes(h, real));
  h->duration += (uint64_t)real;
  if (job->flush) h->last_frame_size = real;
  job->ret[i] = (real == 0 && err) ? err : real;
}

typedef struct {
  ih_step_job step;
  ih_out_job out;
} ih_group_step;

static void group_fill(void *user, int s_lo, int s_cnt, iamfb_io *io) {
  ih_group_step *gs = (ih_group_step *)user;
  IAMF_DecoderHandle *hs = gs->step.hs, L = hs[0];
  const int F = gs->step.units;
  ih_step_job job = gs->step;
  job.base = s_lo;
  pool_for(s_cnt, step_handle, &job);
  int any_ramp[IAMFB_MAX_ELEMENTS] = {0, 0}, any_out_ramp = 0;
  for (int i = s_lo; i < s_lo + s_cnt; ++i)
    for (int f = 0; f < F; ++f) {
      const int fl = hs[i]->unit_flags[f];
      any_ramp[0] |= fl & 1; any_ramp[1] |= (fl >> 1) & 1; any_out_ramp |= (fl >> 2) & 1;
    }
  /* a segment array applies to the whole group of streams: frames without segments (n_segs 0; every slot without a
   * frame too) take the constant of their iamfb_frame_params on the device */
  for (int i = s_lo; i < s_lo + s_cnt; ++i)
    for (int f = 0; f < F; ++f) {
      const int fl = hs[i]->unit_flags[f];
      for (int e = 0; e < L->n_streams; ++e)
        if (!(fl & (1 << e))) L->ramp[e][(size_t)i * F + f].n_segs = 0;
      if (!(fl & 4)) L->out_ramp[(size_t)i * F + f].n_segs = 0;
    }
  for (int e = 0; e < L->n_streams; ++e) io->gain_segs[e] = any_ramp[e] ? L->ramp[e] : 0;
  io->out_gain_segs = any_out_ramp ? L->out_ramp : 0;
}

static void group_drain(void *user, int s_lo, int s_cnt) {
  ih_group_step *gs = (ih_group_step *)user;
  ih_out_job job = gs->out;
  job.base = s_lo;
  job.flush = 0;
  pool_for(s_cnt, out_handle, &job);
}

/* one step of the handles that share ONE device (device < 0: wherever the first handle was configured) */
static int batch_units(IAMF_DecoderHandle *hs, int n, const uint8_t *const *data, const int32_t *size, uint32_t *rsize,
                       void *const *pcm, int *ret, int max_units, int *units_done, int device) {
  if (!hs || n <= 0 || !data || !size || !ret || !hs[0] || max_units < 1 || max_units > 64) return IAMF_ERR_BAD_ARG;
  IAMF_DecoderHandle L = hs[0];
  if (L->group_size != n || L->leader || L->group_units != max_units) {
    if (L->group_size > 1 || L->leader) return IAMF_ERR_INVALID_STATE; /* group membership and step size are fixed */
    if (device >= 0 && L->status == IH_STATUS_RECEIVE && L->plan && !L->duration) {
      int brc = engine_bind(L, device);   /* the group lives where its leader's plan lives */
      if (brc != IAMF_OK) return brc;
    }
    int rc = group_build(hs, n, max_units);
    if (rc != IAMF_OK) return rc;
  }
  for (int i = 1; i < n; ++i)
    if (!hs[i] || hs[i]->leader != L || hs[i]->group_index != i) return IAMF_ERR_BAD_ARG;
  const int F = max_units;
  int n_flush = 0;
  for (int i = 0; i < n; ++i) n_flush += data[i] ? 0 : 1;
  if (n_flush && n_flush != n) return IAMF_ERR_UNIMPLEMENTED; /* a group flushes together */
  ih_shared *sh = (ih_shared *)L->shared;
  if (n_flush) {
    pthread_mutex_lock(&sh->mu);
    const int frc = iamfb_batch_flush_host(L->batch, L->pcm_stage, L->counts_stage);
    pthread_mutex_unlock(&sh->mu);
    if (frc != IAMFB_OK) return IAMF_ERR_INTERNAL;
  } else {
    /* The step runs group by group of handles (iamfb_batch_submit_host_hooks): phase 1 - parse + core decode of a group
     * into its slots of the pinned buffers, on the pool - right before the group's upload; phase 3 - its samples back into
     * the callers' buffers - once its download is done; the device works on the neighbouring groups meanwhile */
    ih_group_step gs;
    memset(&gs, 0, sizeof(gs));
    gs.step.hs = hs; gs.step.data = data; gs.step.size = size; gs.step.units = F;
    gs.out.hs = hs; gs.out.pcm = pcm; gs.out.ret = ret; gs.out.rsize = rsize; gs.out.units_done = units_done; gs.out.units = F;
    iamfb_io io;
    memset(&io, 0, sizeof(io));
    for (int e = 0; e < L->n_streams; ++e) io.in[e] = L->in[e];
    io.in_format = L->group_s16 ? IAMFB_IN_S16 : IAMFB_IN_F32;
    io.params = L->fp_stage;
    io.pcm = L->pcm_stage;
    io.out_counts = L->counts_stage;
    iamfb_chunk_hooks hk = {group_fill, group_drain, &gs};
    pthread_mutex_lock(&sh->mu);
    const int src = iamfb_batch_submit_host_hooks(L->batch, &io, F, &hk);
    pthread_mutex_unlock(&sh->mu);
    if (src != IAMFB_OK) return IAMF_ERR_INTERNAL;
    return IAMF_OK;
  }
  /* phase 3 (host, per handle, on the pool): hand every stream's samples back (the frames of a stream lie back to back in
   * its row of the PCM buffer) */
  {
    ih_out_job oj = {hs, pcm, ret, rsize, units_done, F, 1, 0};
    pool_for(n, out_handle, &oj);
  }
  return IAMF_OK;
}

/* ---- the public call: the handles split over the process's devices (handle i -> device i mod D), one host thread per
 * device stepping its share (its own context, group batch and pinned buffers); no exchange between the devices ---- */
typedef struct {
  IAMF_DecoderHandle *hs;
  const uint8_t **data;
  int32_t *size;
  uint32_t *rsize;
  void **pcm;
  int *ret, *units_done;
  int n, max_units, device, rc;
} ih_dev_job;
static void *dev_thread(void *v) {
  ih_dev_job *j = (ih_dev_job *)v;
  j->rc = batch_units(j->hs, j->n, (const uint8_t *const *)j->data, j->size, j->rsize, (void *const *)j->pcm, j->ret, j->max_units,
                      j->units_done, j->device);
  return 0;
}

int IAMF_decoder_decode_batch_units(IAMF_DecoderHandle *hs, int n, const uint8_t *const *data, const int32_t *size, uint32_t *rsize,
                                    void *const *pcm, int *ret, int max_units, int *units_done) {
  if (!hs || n <= 0 || !data || !size || !ret || !hs[0] || max_units < 1 || max_units > 64) return IAMF_ERR_BAD_ARG;
  devices_init();
  int D = g_ndev < n ? g_ndev : n;
  if (D <= 1) return batch_units(hs, n, data, size, rsize, pcm, ret, max_units, units_done, g_ndev > 1 || getenv("IAMF_B200_DEVICES") ? g_devs[0] : -1);
  /* device-major copies of the per-handle arrays (the handles' own buffers are not copied) */
  const size_t per = sizeof(IAMF_DecoderHandle) + sizeof(uint8_t *) + sizeof(int32_t) + sizeof(uint32_t) + sizeof(void *) + 2 * sizeof(int);
  uint8_t *mem = (uint8_t *)calloc((size_t)n, per);
  if (!mem) return IAMF_ERR_ALLOC_FAIL;
  IAMF_DecoderHandle *hs2 = (IAMF_DecoderHandle *)mem;
  const uint8_t **data2 = (const uint8_t **)(hs2 + n);
  void **pcm2 = (void **)(data2 + n);
  int32_t *size2 = (int32_t *)(pcm2 + n);
  uint32_t *rsize2 = (uint32_t *)(size2 + n);
  int *ret2 = (int *)(rsize2 + n), *units2 = ret2 + n;
  ih_dev_job job[IH_MAX_DEVICES];
  pthread_t th[IH_MAX_DEVICES];
  int at = 0;
  for (int d = 0; d < D; ++d) {
    ih_dev_job *j = &job[d];
    j->hs = hs2 + at; j->data = data2 + at; j->size = size2 + at; j->rsize = rsize2 + at; j->pcm = pcm2 + at; j->ret = ret2 + at;
    j->units_done = units2 + at; j->max_units = max_units; j->device = g_devs[d]; j->rc = IAMF_OK; j->n = 0;
    for (int i = d; i < n; i += D, ++at, ++j->n) {
      hs2[at] = hs[i]; data2[at] = data[i]; size2[at] = size[i]; pcm2[at] = pcm ? pcm[i] : 0;
    }
    if (!pcm) j->pcm = 0;
  }
  int started[IH_MAX_DEVICES] = {0};
  for (int d = 1; d < D; ++d) started[d] = pthread_create(&th[d], 0, dev_thread, &job[d]) == 0;
  dev_thread(&job[0]);
  int rc = job[0].rc;
  for (int d = 1; d < D; ++d) {
    if (started[d]) pthread_join(th[d], 0);
    else dev_thread(&job[d]);
    if (rc == IAMF_OK) rc = job[d].rc;
  }
  at = 0;
  for (int d = 0; d < D; ++d)
    for (int i = d; i < n; i += D, ++at) {
      ret[i] = ret2[at];
      if (rsize) rsize[i] = rsize2[at];
      if (units_done) units_done[i] = units2[at];
    }
  free(mem);
  return rc;
}

int IAMF_decoder_decode_batch(IAMF_DecoderHandle *hs, int n, const uint8_t *const *data, const int32_t *size, uint32_t *rsize,
                              void *const *pcm, int *ret) {
  return IAMF_decoder_decode_batch_units(hs, n, data, size, rsize, pcm, ret, 1, 0);
}
