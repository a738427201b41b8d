/*
 * iamf_obu_parse.c - OBU splitting and descriptor / parameter-block parsing for the drop-in host layer.
 * Wire syntax = the early IAMF draft the reference decoder reads (IAMF_OBU.c:79-138, 260-356, 391-607, 641-932,
 * 990-1254); written from that syntax description, not from the reference sources.
 */
#include <stdlib.h>
#include <string.h>

#include "iamf_host.h"

/* ------------------------------------------------------------------ reader ---- */
void ih_rd_init(ih_reader *r, const uint8_t *p, uint32_t size) {
  r->p = p;
  r->size = size;
  r->pos = 0;
  r->bit = 0;
}

uint32_t ih_rd_bits(ih_reader *r, int n) {
  uint32_t v = 0;
  while (n-- > 0) {
    uint32_t b = r->pos < r->size ? (r->p[r->pos] >> (7 - r->bit)) & 1u : 0u;
    v = (v << 1) | b;
    if (++r->bit == 8) {
      r->bit = 0;
      ++r->pos;
    }
  }
  return v;
}

void ih_rd_skip_bits(ih_reader *r, int n) {
  int b = r->bit + n;
  r->pos += b / 8;
  r->bit = b % 8;
}

void ih_rd_align(ih_reader *r) {
  if (r->bit) {
    r->bit = 0;
    ++r->pos;
  }
}

uint32_t ih_rd_u8(ih_reader *r) {
  ih_rd_align(r);
  uint32_t v = r->pos < r->size ? r->p[r->pos] : 0;
  ++r->pos;
  return v;
}

uint32_t ih_rd_u16(ih_reader *r) {
  uint32_t hi = ih_rd_u8(r);
  return (hi << 8) | ih_rd_u8(r);
}

uint64_t ih_rd_leb128(ih_reader *r) {
  /* at most 8 bytes are consumed (bitstream.c:133-152) */
  uint64_t v = 0;
  uint32_t i;
  ih_rd_align(r);
  if (r->pos >= r->size) return 0;
  for (i = 0; i < 8; ++i) {
    if (r->pos + i >= r->size) break;
    uint8_t byte = r->p[r->pos + i];
    v |= ((uint64_t)(byte & 0x7f)) << (7 * i);
    if (!(byte & 0x80)) break;
  }
  r->pos += i + 1;
  return v;
}

void ih_rd_bytes(ih_reader *r, uint8_t *dst, uint32_t n) {
  ih_rd_align(r);
  if (dst) {
    for (uint32_t i = 0; i < n; ++i) dst[i] = r->pos + i < r->size ? r->p[r->pos + i] : 0;
  }
  r->pos += n;
}

void ih_rd_cstring(ih_reader *r) {
  ih_rd_align(r);
  while (r->pos < r->size && r->p[r->pos]) ++r->pos;
  ++r->pos;
}

uint32_t ih_rd_tell(const ih_reader *r) { return r->bit ? r->pos + 1 : r->pos; }

/* ------------------------------------------------------------------ OBU header ---- */
uint32_t ih_obu_split(const uint8_t *data, uint32_t size, ih_obu *o) {
  ih_reader r;
  if (size < 2) return 0;
  ih_rd_init(&r, data, size);
  memset(o, 0, sizeof(*o));
  o->type = (int)ih_rd_bits(&r, 5);
  o->redundant = (int)ih_rd_bits(&r, 1);
  int trimming = (int)ih_rd_bits(&r, 1);
  int extension = (int)ih_rd_bits(&r, 1);
  uint64_t body = ih_rd_leb128(&r);
  if (body == UINT64_MAX || body + ih_rd_tell(&r) > size) return 0;
  o->total_size = ih_rd_tell(&r) + (uint32_t)body;
  if (trimming) {
    o->trim_end = ih_rd_leb128(&r);
    o->trim_start = ih_rd_leb128(&r);
  }
  if (extension) {
    uint64_t ext = ih_rd_leb128(&r);
    /* a malformed / truncated OBU: the optional header fields must end inside the OBU (IAMF_OBU.c:79-138 reads them
     * from the same bounded bit stream) */
    if (ext == UINT64_MAX || ih_rd_tell(&r) > o->total_size || ext > (uint64_t)(o->total_size - ih_rd_tell(&r))) return 0;
    ih_rd_bytes(&r, 0, (uint32_t)ext);
  }
  if (ih_rd_tell(&r) > o->total_size) return 0;
  o->payload = data + ih_rd_tell(&r);
  o->payload_size = o->total_size - ih_rd_tell(&r);
  return o->total_size;
}

/* ------------------------------------------------------------------ descriptors ---- */
static uint32_t fourcc(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

static int rd_be32(const uint8_t *p) { return (int)((uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]); }

int ih_parse_codec(const ih_obu *o, ih_codec *c) {
  ih_reader r;
  uint8_t cc4[4];
  ih_rd_init(&r, o->payload, o->payload_size);
  memset(c, 0, sizeof(*c));
  c->id = ih_rd_leb128(&r);
  ih_rd_bytes(&r, cc4, 4);
  c->frame_size = ih_rd_leb128(&r);
  (void)ih_rd_u16(&r); /* roll distance */
  c->conf_size = (int)o->payload_size - (int)ih_rd_tell(&r);
  if (c->conf_size < 0) return IAMF_ERR_INVALID_PACKET;
  const int full_conf = c->conf_size;
  if (c->conf_size > (int)sizeof(c->conf)) c->conf_size = sizeof(c->conf);
  ih_rd_bytes(&r, c->conf, (uint32_t)c->conf_size);
  switch (fourcc(cc4)) {
    case 0x6970636d: c->codec = IAMF_CODEC_PCM; break;  /* 'ipcm' */
    case 0x4f707573: c->codec = IAMF_CODEC_OPUS; break; /* 'Opus' */
    case 0x6d703461: c->codec = IAMF_CODEC_AAC; break;  /* 'mp4a' */
    case 0x664c6143: c->codec = IAMF_CODEC_FLAC; break; /* 'fLaC' */
    default: return IAMF_ERR_INVALID_PACKET;
  }
  if (!ih_codec_supported(c->codec)) return IAMF_ERR_UNIMPLEMENTED;
  /* sampling rate of the stream, iamf_codec_conf_get_sampling_rate IAMF_decoder.c:708-755 */
  c->rate = -1;
  if (c->codec == IAMF_CODEC_PCM && c->conf_size >= 6) c->rate = rd_be32(c->conf + 2);
  else if (c->codec == IAMF_CODEC_OPUS && c->conf_size >= 8) c->rate = rd_be32(c->conf + 4);
  else if (c->codec == IAMF_CODEC_FLAC) {
    /* (a decoder config with more metadata than the 64 bytes kept here is not supported) */
    if (full_conf > (int)sizeof(c->conf)) return IAMF_ERR_UNIMPLEMENTED;
    c->rate = ih_flac_rate(c->conf, c->conf_size);
  }
  /* a damaged config must not reach the buffer arithmetic: a rate of a few hertz against the 48 kHz output would make one
   * frame hundreds of thousands of samples long (the reference sizes its own buffers by max_frame_size alone) */
  if (c->frame_size < 1 || c->frame_size > 65536) return IAMF_ERR_INVALID_PACKET;
  if (c->rate != -1 && (c->rate < 8000 || c->rate > 192000)) return IAMF_ERR_UNIMPLEMENTED;
  return IAMF_OK;
}

static int parse_param_def(ih_reader *r, int type, ih_param_def *d) {
  memset(d, 0, sizeof(*d));
  d->type = type;
  d->id = ih_rd_leb128(r);
  d->rate = ih_rd_leb128(r);
  d->mode = (int)ih_rd_bits(r, 1);
  if (!d->mode) {
    d->duration = ih_rd_leb128(r);
    d->const_interval = ih_rd_leb128(r);
    if (!d->const_interval) {
      d->n_segments = (int)ih_rd_leb128(r);
      if (d->n_segments > 16) return IAMF_ERR_UNIMPLEMENTED;
      for (int i = 0; i < d->n_segments; ++i) d->seg_interval[i] = ih_rd_leb128(r);
    } else {
      d->n_segments = (int)((d->duration + d->const_interval - 1) / d->const_interval);
    }
  }
  return IAMF_OK;
}

int ih_parse_element(const ih_obu *o, ih_element *e) {
  ih_reader r;
  ih_rd_init(&r, o->payload, o->payload_size);
  memset(e, 0, sizeof(*e));
  e->id = ih_rd_leb128(&r);
  e->type = (int)ih_rd_bits(&r, 3);
  ih_rd_bits(&r, 5);
  e->codec_id = ih_rd_leb128(&r);
  e->n_sub = (int)ih_rd_leb128(&r);
  if (e->n_sub > IH_MAX_SUBSTREAMS) return IAMF_ERR_UNIMPLEMENTED;
  for (int i = 0; i < e->n_sub; ++i) e->sub_ids[i] = ih_rd_leb128(&r);
  int np = (int)ih_rd_leb128(&r);
  e->dmx_mode = e->dmx_w = -1;
  for (int i = 0; i < np; ++i) {
    uint64_t type = ih_rd_leb128(&r);
    if (type == IAMF_PARAMETER_TYPE_DEMIXING || type == IAMF_PARAMETER_TYPE_RECON_GAIN) {
      if (e->n_params >= 4) return IAMF_ERR_UNIMPLEMENTED;
      ih_param_def *d = &e->params[e->n_params++];
      int rc = parse_param_def(&r, (int)type, d);
      if (rc) return rc;
      if (type == IAMF_PARAMETER_TYPE_DEMIXING) {
        int mode = (int)ih_rd_bits(&r, 3);
        ih_rd_skip_bits(&r, 5);
        int w = (int)ih_rd_bits(&r, 4);
        ih_rd_skip_bits(&r, 4);
        if (!e->has_demix) { /* the first demixing definition provides the defaults (IAMF_decoder.c:1733-1741) */
          e->has_demix = 1;
          e->dmx_mode = mode;
          e->dmx_w = w;
        }
      }
    } else {
      uint64_t sz = ih_rd_leb128(&r);
      ih_rd_bytes(&r, 0, (uint32_t)sz);
    }
  }
  if (e->type == 0) {
    e->n_layers = (int)ih_rd_bits(&r, 3);
    ih_rd_skip_bits(&r, 5);
    if (e->n_layers > IH_MAX_LAYERS) return IAMF_ERR_INVALID_PACKET;
    for (int i = 0; i < e->n_layers; ++i) {
      ih_layer *l = &e->layers[i];
      l->layout = (int)ih_rd_bits(&r, 4);
      l->out_gain_present = (int)ih_rd_bits(&r, 1);
      l->recon_present = (int)ih_rd_bits(&r, 1);
      l->n_sub = (int)ih_rd_u8(&r);
      l->n_coupled = (int)ih_rd_u8(&r);
      if (l->out_gain_present) {
        l->out_gain_flags = (int)ih_rd_bits(&r, 6);
        l->out_gain_q = (int16_t)ih_rd_u16(&r);
      }
    }
    {   /* the layers' sub-streams are the element's sub-streams (they index st->pkt, IH_MAX_SUBSTREAMS entries) */
      int sum = 0;
      for (int i = 0; i < e->n_layers; ++i) {
        if (e->layers[i].n_coupled > e->layers[i].n_sub) return IAMF_ERR_INVALID_PACKET;
        sum += e->layers[i].n_sub;
      }
      if (sum != e->n_sub) return IAMF_ERR_INVALID_PACKET;
    }
  } else if (e->type == 1) {
    e->ambi_mode = (int)ih_rd_leb128(&r);
    if (e->ambi_mode == 0) {
      e->ambi_channels = (int)ih_rd_u8(&r);
      e->ambi_sub = (int)ih_rd_u8(&r);
      e->ambi_map_size = e->ambi_channels;
    } else if (e->ambi_mode == 1) {
      e->ambi_channels = (int)ih_rd_u8(&r);
      e->ambi_sub = (int)ih_rd_u8(&r);
      e->ambi_coupled = (int)ih_rd_u8(&r);
      e->ambi_map_size = 2 * e->ambi_channels * (e->ambi_sub + e->ambi_coupled);
    } else {
      return IAMF_ERR_INVALID_PACKET;
    }
    /* sub-stream counts index st->pkt / pkt_size / codec_state (IH_MAX_SUBSTREAMS entries) and the engine's 16-row
     * scene input: they must agree with the element's own sub-stream list (IAMF_OBU.c:512-607 reads them from the same
     * element) */
    if (e->ambi_sub != e->n_sub || e->ambi_coupled > e->ambi_sub || e->ambi_sub + e->ambi_coupled > IAMFB_MAX_SCENE_CH ||
        e->ambi_channels < 1 || e->ambi_channels > IAMFB_MAX_SCENE_CH)
      return IAMF_ERR_INVALID_PACKET;
    if (e->ambi_map_size > (int)sizeof(e->ambi_map)) return IAMF_ERR_UNIMPLEMENTED;
    ih_rd_bytes(&r, e->ambi_map, (uint32_t)e->ambi_map_size);
  } else {
    return IAMF_ERR_UNIMPLEMENTED;
  }
  return IAMF_OK;
}

int ih_parse_mix(const ih_obu *o, ih_mix *m) {
  ih_reader r;
  ih_rd_init(&r, o->payload, o->payload_size);
  memset(m, 0, sizeof(*m));
  m->id = ih_rd_leb128(&r);
  int labels = (int)ih_rd_leb128(&r);
  for (int i = 0; i < 2 * labels; ++i) ih_rd_cstring(&r); /* languages, then presentation labels */
  uint64_t sub = ih_rd_leb128(&r);
  if (sub != 1) return IAMF_ERR_INVALID_PACKET; /* one sub-mix only (IAMF_OBU.c:700-707) */
  m->n_elements = (int)ih_rd_leb128(&r);
  if (m->n_elements < 1 || m->n_elements > 2) return IAMF_ERR_INVALID_PACKET; /* :742-752 */
  for (int i = 0; i < m->n_elements; ++i) {
    m->el[i].element_id = ih_rd_leb128(&r);
    for (int k = 0; k < labels; ++k) ih_rd_cstring(&r);
    m->el[i].headphones_mode = (int)ih_rd_bits(&r, 2);
    uint64_t ext = ih_rd_leb128(&r);
    ih_rd_bytes(&r, 0, (uint32_t)ext);
    int rc = parse_param_def(&r, IAMF_PARAMETER_TYPE_MIX_GAIN, &m->el[i].gain_def);
    if (rc) return rc;
    m->el[i].gain_q = (int16_t)ih_rd_u16(&r);
  }
  int rc = parse_param_def(&r, IAMF_PARAMETER_TYPE_MIX_GAIN, &m->out_def);
  if (rc) return rc;
  m->out_q = (int16_t)ih_rd_u16(&r);
  m->n_layouts = (int)ih_rd_leb128(&r);
  if (m->n_layouts > 16) return IAMF_ERR_UNIMPLEMENTED;
  for (int i = 0; i < m->n_layouts; ++i) {
    m->layouts[i].type = (int)ih_rd_bits(&r, 2);
    if (m->layouts[i].type == IAMF_LAYOUT_TYPE_LOUDSPEAKERS_SS_CONVENTION) m->layouts[i].sound_system = (int)ih_rd_bits(&r, 4);
    ih_rd_align(&r);
    IAMF_LoudnessInfo *li = &m->layouts[i].loud;
    li->info_type = (uint8_t)ih_rd_u8(&r);
    li->integrated_loudness = (int16_t)ih_rd_u16(&r);
    li->digital_peak = (int16_t)ih_rd_u16(&r);
    if (li->info_type & 1) li->true_peak = (int16_t)ih_rd_u16(&r);
    if (li->info_type & 2) {
      int n = (int)ih_rd_u8(&r);
      for (int k = 0; k < n; ++k) { ih_rd_u8(&r); ih_rd_u16(&r); } /* anchored loudness: not needed for rendering */
    }
    if (li->info_type & ~3) {
      uint64_t sz = ih_rd_leb128(&r);
      ih_rd_bytes(&r, 0, (uint32_t)sz);
    }
  }
  return IAMF_OK;
}

uint64_t ih_obu_parameter_id(const ih_obu *o) {
  ih_reader r;
  ih_rd_init(&r, o->payload, o->payload_size);
  return ih_rd_leb128(&r);
}

/* parameter block -> linked list of segments (IAMF_OBU.c:990-1215) */
ih_segment *ih_parse_parameter_block(const ih_obu *o, uint64_t *pid_out, const ih_param_def *def, int n_layers,
                                     unsigned recon_present_flags, int *n_segments) {
  ih_reader r;
  ih_rd_init(&r, o->payload, o->payload_size);
  uint64_t pid = ih_rd_leb128(&r);
  if (pid_out) *pid_out = pid;
  *n_segments = 0;
  if (!def) return 0;
  uint64_t duration, const_iv;
  int nseg;
  if (!def->mode) {
    duration = def->duration;
    const_iv = def->const_interval;
    nseg = def->n_segments;
  } else {
    duration = ih_rd_leb128(&r);
    const_iv = ih_rd_leb128(&r);
    uint64_t ns = const_iv ? (duration + const_iv - 1) / const_iv : ih_rd_leb128(&r);
    if (duration == UINT64_MAX || const_iv == UINT64_MAX || ns > IH_MAX_SEGMENTS) return 0;
    nseg = (int)ns;
  }
  /* the segment count comes from the stream: every segment takes at least one payload byte (a demixing mode, a
   * leb128 animation type, a recon flag word), so a count beyond the bytes left - or beyond a sane limit - is a
   * malformed block, not a reason to allocate */
  if (nseg < 0 || nseg > IH_MAX_SEGMENTS || (uint32_t)nseg > o->payload_size) return 0;
  ih_segment *head = 0, *tail = 0;
  uint64_t left = duration, iv = 0;
  for (int i = 0; i < nseg; ++i) {
    if (ih_rd_tell(&r) > o->payload_size) break;   /* ran past the payload: stop (reads past the end return 0) */
    if (!const_iv) iv = def->mode ? ih_rd_leb128(&r) : (i < 16 ? def->seg_interval[i] : 0);
    uint64_t seg_iv = iv ? iv : (const_iv < left ? const_iv : left);
    left -= seg_iv;
    ih_segment *s = (ih_segment *)calloc(1, sizeof(*s));
    if (!s) break;
    s->interval = seg_iv;
    if (def->type == IAMF_PARAMETER_TYPE_MIX_GAIN) {
      s->anim = (int)ih_rd_leb128(&r);
      s->g_start = ih_db2lin(ih_q_to_float((int16_t)ih_rd_u16(&r), 8));
      if (s->anim != ANIMATION_TYPE_STEP) {
        s->g_end = ih_db2lin(ih_q_to_float((int16_t)ih_rd_u16(&r), 8));
        if (s->anim == ANIMATION_TYPE_BEZIER) {
          s->g_control = ih_db2lin(ih_q_to_float((int16_t)ih_rd_u16(&r), 8));
          s->g_ctime = ih_qf_to_float((uint8_t)ih_rd_u8(&r));
        }
      }
    } else if (def->type == IAMF_PARAMETER_TYPE_DEMIXING) {
      s->dmx_mode = (int)ih_rd_bits(&r, 3);
    } else if (def->type == IAMF_PARAMETER_TYPE_RECON_GAIN) {
      s->n_layers = n_layers;
      for (int k = 0; k < n_layers && k < IH_MAX_LAYERS; ++k) {
        if (!(recon_present_flags & (1u << k))) continue;
        s->rg[k].flags = (uint32_t)ih_rd_leb128(&r);
        int n = 0;
        for (uint32_t f = s->rg[k].flags; f; f &= f - 1) ++n;
        s->rg[k].n = n;
        for (int c = 0; c < n && c < 12; ++c) s->rg[k].q[c] = (uint8_t)ih_rd_u8(&r);
      }
    }
    if (tail) tail->next = s; else head = s;
    tail = s;
    ++*n_segments;
  }
  return head;
}
