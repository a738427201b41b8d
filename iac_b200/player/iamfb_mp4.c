/* iamfb_mp4.c - see iamfb_mp4.h */
#include "iamfb_mp4.h"

#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#define FOURCC(a, b, c, d) (((uint32_t)(a) << 24) | ((uint32_t)(b) << 16) | ((uint32_t)(c) << 8) | (uint32_t)(d))

static uint32_t be16(const uint8_t *p) { return ((uint32_t)p[0] << 8) | p[1]; }
static uint32_t be32(const uint8_t *p) { return (be16(p) << 16) | be16(p + 2); }
static uint64_t be64(const uint8_t *p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }

typedef struct {
  uint32_t type;
  const uint8_t *start, *body, *end;   /* header start, payload start, one past the box */
} box_t;

/* the box at cur (inside [cur, lim)); 0 when there is none or it does not fit */
static int box_at(const uint8_t *cur, const uint8_t *lim, box_t *b) {
  if (lim - cur < 8) return 0;
  uint64_t size = be32(cur);
  size_t hdr = 8;
  b->type = be32(cur + 4);
  if (size == 1) {
    if (lim - cur < 16) return 0;
    size = be64(cur + 8);
    hdr = 16;
  } else if (size == 0) {
    size = (uint64_t)(lim - cur);
  }
  if (size < hdr || size > (uint64_t)(lim - cur)) return 0;
  b->start = cur; b->body = cur + hdr; b->end = cur + size;
  return 1;
}

/* first child of `type` inside [p, lim) */
static int find_box(const uint8_t *p, const uint8_t *lim, uint32_t type, box_t *out) {
  box_t b;
  while (box_at(p, lim, &b)) {
    if (b.type == type) { *out = b; return 1; }
    p = b.end;
  }
  return 0;
}

typedef struct {
  uint32_t track_id, timescale;
  int64_t skip;
  int n_desc;
  struct { const uint8_t *obus; uint32_t size; } desc[8];
  box_t stts, stsc, stsz, stco, co64;
  int has_stts, has_stsc, has_stsz, has_stco, has_co64;
} trak_t;

static void read_stsd(const box_t *stsd, trak_t *t) {
  if (stsd->end - stsd->body < 8) return;
  uint32_t n = be32(stsd->body + 4);
  const uint8_t *p = stsd->body + 8;
  box_t e;
  for (uint32_t i = 0; i < n && box_at(p, stsd->end, &e); ++i, p = e.end) {
    /* AudioSampleEntry: reserved(6) data_reference_index(2) reserved(8) channelcount(2) samplesize(2) pre_defined(2)
     * reserved(2) samplerate(4) = 28 bytes, the descriptor OBUs follow */
    if (e.type != FOURCC('i', 'a', 'm', 'f') || e.end - e.body < 28 || t->n_desc >= 8) continue;
    t->desc[t->n_desc].obus = e.body + 28;
    t->desc[t->n_desc].size = (uint32_t)(e.end - e.body - 28);
    ++t->n_desc;
  }
}

static void read_trak(const box_t *trak, trak_t *t) {
  box_t b, c, d;
  memset(t, 0, sizeof(*t));
  if (find_box(trak->body, trak->end, FOURCC('t', 'k', 'h', 'd'), &b) && b.end - b.body >= 16) {
    const int v1 = b.body[0] == 1;
    if (b.end - b.body >= (v1 ? 24 : 16)) t->track_id = be32(b.body + (v1 ? 20 : 12));
  }
  if (find_box(trak->body, trak->end, FOURCC('e', 'd', 't', 's'), &b) && find_box(b.body, b.end, FOURCC('e', 'l', 's', 't'), &c) &&
      c.end - c.body >= 8) {
    const int v1 = c.body[0] == 1;
    const uint32_t n = be32(c.body + 4);
    const uint8_t *p = c.body + 8;
    int64_t start = 0;
    for (uint32_t i = 0; i < n && c.end - p >= (v1 ? 20 : 12); ++i, p += v1 ? 20 : 12)
      start = v1 ? (int64_t)be64(p + 8) : (int64_t)(int32_t)be32(p + 4);   /* media_time of the LAST entry (mp4demux.c:474-487) */
    if (start > 0) t->skip = start;
  }
  if (!find_box(trak->body, trak->end, FOURCC('m', 'd', 'i', 'a'), &b)) return;
  if (find_box(b.body, b.end, FOURCC('m', 'd', 'h', 'd'), &c) && c.end - c.body >= 20) {
    const int v1 = c.body[0] == 1;
    if (c.end - c.body >= (v1 ? 32 : 20)) t->timescale = be32(c.body + (v1 ? 20 : 12));
  }
  if (!find_box(b.body, b.end, FOURCC('m', 'i', 'n', 'f'), &c) || !find_box(c.body, c.end, FOURCC('s', 't', 'b', 'l'), &d)) return;
  if (find_box(d.body, d.end, FOURCC('s', 't', 's', 'd'), &b)) read_stsd(&b, t);
  t->has_stts = find_box(d.body, d.end, FOURCC('s', 't', 't', 's'), &t->stts);
  t->has_stsc = find_box(d.body, d.end, FOURCC('s', 't', 's', 'c'), &t->stsc);
  t->has_stsz = find_box(d.body, d.end, FOURCC('s', 't', 's', 'z'), &t->stsz);
  t->has_stco = find_box(d.body, d.end, FOURCC('s', 't', 'c', 'o'), &t->stco);
  t->has_co64 = find_box(d.body, d.end, FOURCC('c', 'o', '6', '4'), &t->co64);
}

static int push_sample(iamfb_mp4 *m, size_t *cap, uint64_t off, uint32_t size, uint32_t delta, uint32_t di) {
  if (off > m->size || size > m->size - off) return -3;
  if (m->n_samples == *cap) {
    size_t nc = *cap ? *cap * 2 : 1024;
    iamfb_mp4_sample *ns = (iamfb_mp4_sample *)realloc(m->samples, nc * sizeof(*ns));
    if (!ns) return -3;
    m->samples = ns; *cap = nc;
  }
  iamfb_mp4_sample *s = &m->samples[m->n_samples++];
  s->offset = off; s->size = size; s->delta = delta; s->desc_index = di;
  return 0;
}

/* stsc x stsz x stco -> (offset, size, description) of every sample; stts -> durations */
static int table_samples(iamfb_mp4 *m, const trak_t *t, size_t *cap) {
  if (!t->has_stsz || !t->has_stsc || !(t->has_stco || t->has_co64)) return 0;   /* no samples in moov: fragments may follow */
  const box_t *sz = &t->stsz, *sc = &t->stsc, *co = t->has_stco ? &t->stco : &t->co64;
  const int wide = !t->has_stco;
  if (sz->end - sz->body < 12 || sc->end - sc->body < 8 || co->end - co->body < 8) return -3;
  const uint32_t fixed = be32(sz->body + 4), n = be32(sz->body + 8);
  if (!fixed && (uint64_t)(sz->end - sz->body - 12) / 4 < n) return -3;
  const uint32_t n_sc = be32(sc->body + 4), n_co = be32(co->body + 4);
  if ((uint64_t)(sc->end - sc->body - 8) / 12 < n_sc || (uint64_t)(co->end - co->body - 8) / (wide ? 8 : 4) < n_co) return -3;
  /* durations */
  uint32_t stts_n = 0, stts_i = 0, stts_left = 0, stts_delta = 0;
  if (t->has_stts && t->stts.end - t->stts.body >= 8) {
    stts_n = be32(t->stts.body + 4);
    if ((uint64_t)(t->stts.end - t->stts.body - 8) / 8 < stts_n) return -3;
  }
  uint32_t s = 0, sc_i = 0;
  for (uint32_t chunk = 1; chunk <= n_co && s < n; ++chunk) {
    while (sc_i + 1 < n_sc && be32(sc->body + 8 + 12 * (sc_i + 1)) <= chunk) ++sc_i;
    if (!n_sc) return -3;
    const uint32_t per = be32(sc->body + 8 + 12 * sc_i + 4), di = be32(sc->body + 8 + 12 * sc_i + 8);
    uint64_t off = wide ? be64(co->body + 8 + 8 * (size_t)(chunk - 1)) : be32(co->body + 8 + 4 * (size_t)(chunk - 1));
    for (uint32_t k = 0; k < per && s < n; ++k, ++s) {
      const uint32_t size = fixed ? fixed : be32(sz->body + 12 + 4 * (size_t)s);
      while (!stts_left && stts_i < stts_n) {
        stts_left = be32(t->stts.body + 8 + 8 * (size_t)stts_i);
        stts_delta = be32(t->stts.body + 8 + 8 * (size_t)stts_i + 4);
        ++stts_i;
      }
      if (stts_left) --stts_left;
      int r = push_sample(m, cap, off, size, stts_delta, di);
      if (r) return r;
      off += size;
    }
  }
  return 0;
}

/* one movie fragment: the runs of the wanted track */
static int fragment_samples(iamfb_mp4 *m, const box_t *moof, uint32_t track_id, size_t *cap) {
  box_t traf;
  const uint8_t *p = moof->body;
  uint64_t next_in_moof = 0;   /* where the next run's data starts when it has no data_offset */
  while (box_at(p, moof->end, &traf)) {
    p = traf.end;
    if (traf.type != FOURCC('t', 'r', 'a', 'f')) continue;
    box_t tfhd;
    if (!find_box(traf.body, traf.end, FOURCC('t', 'f', 'h', 'd'), &tfhd) || tfhd.end - tfhd.body < 8) continue;
    const uint32_t tf = be32(tfhd.body) & 0xFFFFFF;
    if (be32(tfhd.body + 4) != track_id) continue;
    const uint8_t *q = tfhd.body + 8;
    uint64_t base = (uint64_t)(moof->start - m->data);
    uint32_t di = 1, def_dur = m->trex_duration, def_size = m->trex_size;
    if ((tf & 0x1) && tfhd.end - q >= 8) { base = be64(q); q += 8; }
    if ((tf & 0x2) && tfhd.end - q >= 4) { di = be32(q); q += 4; }
    if ((tf & 0x8) && tfhd.end - q >= 4) { def_dur = be32(q); q += 4; }
    if ((tf & 0x10) && tfhd.end - q >= 4) { def_size = be32(q); q += 4; }
    box_t trun;
    const uint8_t *r = traf.body;
    while (box_at(r, traf.end, &trun)) {
      r = trun.end;
      if (trun.type != FOURCC('t', 'r', 'u', 'n') || trun.end - trun.body < 8) continue;
      const uint32_t rf = be32(trun.body) & 0xFFFFFF, n = be32(trun.body + 4);
      const uint8_t *e = trun.body + 8;
      uint64_t off = next_in_moof ? next_in_moof : base;
      if ((rf & 0x1) && trun.end - e >= 4) { off = base + (int64_t)(int32_t)be32(e); e += 4; }
      if ((rf & 0x4) && trun.end - e >= 4) e += 4;
      const int per = ((rf & 0x100) ? 4 : 0) + ((rf & 0x200) ? 4 : 0) + ((rf & 0x400) ? 4 : 0) + ((rf & 0x800) ? 4 : 0);
      if (per && (uint64_t)(trun.end - e) / (uint64_t)per < n) return -3;
      for (uint32_t i = 0; i < n; ++i) {
        uint32_t dur = def_dur, size = def_size;
        if (rf & 0x100) { dur = be32(e); e += 4; }
        if (rf & 0x200) { size = be32(e); e += 4; }
        if (rf & 0x400) e += 4;
        if (rf & 0x800) e += 4;
        int rc = push_sample(m, cap, off, size, dur, di);
        if (rc) return rc;
        off += size;
      }
      next_in_moof = off;
    }
  }
  return 0;
}

int iamfb_mp4_open(iamfb_mp4 *m, const char *path) {
  memset(m, 0, sizeof(*m));
  int fd = open(path, O_RDONLY);
  if (fd < 0) return -1;
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size <= 0) { close(fd); return -1; }
  void *map = mmap(0, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) return -1;
  m->data = (const uint8_t *)map;
  m->size = (size_t)st.st_size;
  const uint8_t *lim = m->data + m->size;
  box_t moov, b;
  if (!find_box(m->data, lim, FOURCC('m', 'o', 'o', 'v'), &moov)) { iamfb_mp4_close(m); return -2; }
  if (find_box(moov.body, moov.end, FOURCC('m', 'v', 'h', 'd'), &b) && b.end - b.body >= 16) {
    const int v1 = b.body[0] == 1;
    if (b.end - b.body >= (v1 ? 28 : 16)) m->movie_timescale = be32(b.body + (v1 ? 20 : 12));
  }
  trak_t want;
  int found = 0;
  for (const uint8_t *p = moov.body; box_at(p, moov.end, &b); p = b.end) {
    if (b.type != FOURCC('t', 'r', 'a', 'k') || found) continue;
    trak_t t;
    read_trak(&b, &t);
    if (t.n_desc > 0) { want = t; found = 1; }
  }
  if (!found) { iamfb_mp4_close(m); return -2; }
  m->media_timescale = want.timescale;
  m->skip = want.skip;
  m->n_desc = want.n_desc;
  for (int i = 0; i < want.n_desc; ++i) { m->desc[i].obus = want.desc[i].obus; m->desc[i].size = want.desc[i].size; }
  /* defaults of the fragments (mvex / trex of the track) */
  box_t mvex;
  if (find_box(moov.body, moov.end, FOURCC('m', 'v', 'e', 'x'), &mvex))
    for (const uint8_t *p = mvex.body; box_at(p, mvex.end, &b); p = b.end)
      if (b.type == FOURCC('t', 'r', 'e', 'x') && b.end - b.body >= 24 && be32(b.body + 4) == want.track_id) {
        m->trex_duration = be32(b.body + 12);
        m->trex_size = be32(b.body + 16);
      }
  size_t cap = 0;
  int rc = table_samples(m, &want, &cap);
  for (const uint8_t *p = m->data; rc == 0 && box_at(p, lim, &b); p = b.end)
    if (b.type == FOURCC('m', 'o', 'o', 'f')) rc = fragment_samples(m, &b, want.track_id, &cap);
  if (rc) { iamfb_mp4_close(m); return rc; }
  return 0;
}

void iamfb_mp4_close(iamfb_mp4 *m) {
  if (m->data) munmap((void *)m->data, m->size);
  free(m->samples);
  memset(m, 0, sizeof(*m));
}
