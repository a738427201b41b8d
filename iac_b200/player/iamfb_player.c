/* iamfb_player.c - command-line player on top of the drop-in library (include/IAMF_decoder.h): IAMF bitstream or MP4 in,
 * WAV (+ optional .met extradata records) out.  Same options, output file names and output bytes as the reference's
 * test/tools/iamfplayer (player/iamfplayer.c: options :830-905, bitstream loop :529-660, MP4 loop :662-789, extradata
 * records :222-305), written from scratch around a memory-mapped input; the WAV writer and the MP4 reader are this
 * directory's own (iamfb_wav.c, iamfb_mp4.c).  One thing the reference's player cannot do: several input files at once -
 * their decoder handles then step together through IAMF_decoder_decode_batch_units, one device pass per step for all of
 * them (files of one pipeline signature; otherwise they are played one after the other). */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "IAMF_decoder.h"
#include "iamfb_mp4.h"
#include "iamfb_wav.h"

typedef struct {
  int input_mode, output_mode;      /* -i0 bitstream | -i1 mp4; -o2 wav */
  int layout_type;                  /* 2 sound system, 3 binaural, 0 unset */
  int sound_system;
  float peak, loudness;
  uint32_t bit_depth, rate, start_s;
  int metadata, no_limiter;
  uint64_t mix_id;
  int n_inputs;
  char **inputs;
} options;

static void usage(const char *me) {
  fprintf(stderr,
          "Usage:\n%s <options> <input file> [more input files]\noptions:\n"
          "-i[0-1]    0 : IAMF bitstream input.(default)\n           1 : mp4 input.\n"
          "-o2        2 : pcm output.\n"
          "-r [rate]    : audio signal sampling rate, 48000 is the default.\n"
          "-ts pos      : seek to a given position in seconds, which is valid when mp4 file is used as input.\n"
          "-s[0~12,b]   : output layout, the sound system A~J and extensions (0 A .. 9 J, 10 7.1.2, 11 3.1.2, 12 mono), b binaural.\n"
          "-p [dB]      : Peak threshold in dB.\n-l [LKFS]    : Normalization loudness in LKFS.\n-d           : Bit depth of pcm output.\n"
          "-mp [id]     : Set mix presentation id.\n-m           : Generate a metadata file with the suffix .met .\n"
          "-disable_limiter\n             : Disable peak limiter.\n"
          "several input files: decoded together, one GPU pass per step for all of them (IAMF_decoder_decode_batch_units)\n",
          me);
}

/* "ss<N>_<input name without directory and extension>" or "binaural_<...>" (iamfplayer.c:307-345) */
static int out_name(const options *o, const char *path, char *dst, size_t room, const char *ext) {
  const char *s = strrchr(path, '/');
  s = s ? s + 1 : path;
  const char *d = strrchr(path, '.');
  int n = o->layout_type == 2 ? snprintf(dst, room, "ss%d_", o->sound_system) : snprintf(dst, room, "binaural_");
  if (d && d > s) {
    size_t len = (size_t)(d - s);
    if (len > room - (size_t)n - 8) len = room - (size_t)n - 8;
    memcpy(dst + n, s, len);
    n += (int)len;
  }
  snprintf(dst + n, room - (size_t)n, "%s", ext);
  return n;
}

/* ---- .met: one record per decoded frame (layout documented at iamfplayer.c:222-262) ---- */
static size_t loudness_bytes(const IAMF_LoudnessInfo *l) {
  size_t n = 5;
  if (l->info_type & 1) n += 2;
  if (l->info_type & 2) n += 1 + (size_t)l->num_anchor_loudness * sizeof(anchor_loudness_t);   /* (reserved, not all written) */
  return n;
}
static int met_write(FILE *f, int64_t pts, const IAMF_extradata *md) {
  size_t data = 24;
  for (int i = 0; i < md->num_loudness_layouts; ++i) data += sizeof(IAMF_Layout) + loudness_bytes(&md->loudness[i]);
  data += 4 + sizeof(IAMF_Param) * md->num_parameters;
  const uint32_t hdr[5] = {(uint32_t)(20 + data), 1u, 0u, 0x7f000005u, (uint32_t)data};   /* nSize nVersion nPortIndex nType nDataSize */
  const size_t total = (hdr[0] + 8 + 3) & ~(size_t)3;
  uint8_t *buf = (uint8_t *)calloc(1, total + 16);
  if (!buf) return -1;
  size_t at = 0;
#define PUT(ptr, n) do { memcpy(buf + at, (ptr), (n)); at += (n); } while (0)
  PUT(&pts, 8);
  PUT(hdr, sizeof(hdr));
  PUT(&md->output_sound_system, 4); PUT(&md->number_of_samples, 4); PUT(&md->bitdepth, 4); PUT(&md->sampling_rate, 4);
  PUT(&md->output_sound_mode, 4); PUT(&md->num_loudness_layouts, 4);
  for (int i = 0; i < md->num_loudness_layouts; ++i) {
    const IAMF_LoudnessInfo *l = &md->loudness[i];
    PUT(&md->loudness_layout[i], sizeof(IAMF_Layout));
    PUT(&l->info_type, 1); PUT(&l->integrated_loudness, 2); PUT(&l->digital_peak, 2);
    if (l->info_type & 1) PUT(&l->true_peak, 2);
    if (l->info_type & 2) {
      PUT(&l->num_anchor_loudness, 1);
      for (int k = 0; k < l->num_anchor_loudness; ++k) { PUT(&l->anchor_loudness[k].anchor_element, 1); PUT(&l->anchor_loudness[k].anchored_loudness, 2); }
    }
  }
  PUT(&md->num_parameters, 4);
  for (uint32_t i = 0; i < md->num_parameters; ++i) PUT(&md->param[i], sizeof(IAMF_Param));
#undef PUT
  const int ok = fwrite(buf, 1, total, f) == total;
  free(buf);
  return ok ? 0 : -1;
}
static void met_release(IAMF_extradata *md) {
  free(md->loudness_layout);
  if (md->loudness) { free(md->loudness->anchor_loudness); free(md->loudness); }
  free(md->param);
  memset(md, 0, sizeof(*md));
}

/* ---- one input being played ---- */
typedef struct {
  const char *path;
  const uint8_t *data;      /* mapped bitstream (-i0) */
  size_t size, pos;
  iamfb_mp4 mp4;            /* -i1 */
  size_t next_sample;
  uint32_t last_desc;
  IAMF_DecoderHandle dec;
  iamfb_wav wav;
  FILE *met;
  int channels;
  void *pcm;
  size_t pcm_room;          /* bytes */
  long frames, samples;
  int configured, ended;
} input_t;

static IAMF_DecoderHandle open_decoder(const options *o, int *channels) {
  IAMF_DecoderHandle d = IAMF_decoder_open();
  if (!d) return 0;
  if (o->no_limiter) IAMF_decoder_peak_limiter_enable(d, 0);
  else IAMF_decoder_peak_limiter_set_threshold(d, o->peak);
  IAMF_decoder_set_normalization_loudness(d, o->loudness);
  IAMF_decoder_set_bit_depth(d, o->bit_depth);
  if (o->rate > 0 && IAMF_decoder_set_sampling_rate(d, o->rate) != IAMF_OK) {
    fprintf(stderr, "Invalid sampling rate %u\n", o->rate);
    IAMF_decoder_close(d);
    return 0;
  }
  if (o->layout_type == 2) {
    IAMF_decoder_output_layout_set_sound_system(d, (IAMF_SoundSystem)o->sound_system);
    *channels = IAMF_layout_sound_system_channels_count((IAMF_SoundSystem)o->sound_system);
  } else {
    IAMF_decoder_output_layout_set_binaural(d);
    *channels = IAMF_layout_binaural_channels_count();
  }
  if (o->mix_id != UINT64_MAX) IAMF_decoder_set_mix_presentation_id(d, o->mix_id);
  return d;
}

static int input_open(input_t *in, const options *o, const char *path) {
  memset(in, 0, sizeof(*in));
  in->path = path;
  in->dec = open_decoder(o, &in->channels);
  if (!in->dec) { fprintf(stderr, "IAMF decoder can't created.\n"); return -1; }
  if (o->input_mode == 1) {
    int rc = iamfb_mp4_open(&in->mp4, path);
    if (rc) { fprintf(stderr, "can not open mp4 file(%s): %s\n", path, rc == -1 ? strerror(errno) : rc == -2 ? "no IAMF audio track" : "malformed"); return -1; }
  } else {
    int fd = open(path, O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) { fprintf(stderr, "%s can't opened.\n", path); if (fd >= 0) close(fd); return -1; }
    in->size = (size_t)st.st_size;
    if (in->size) {
      void *p = mmap(0, in->size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (p == MAP_FAILED) { close(fd); fprintf(stderr, "%s can't mapped.\n", path); return -1; }
      in->data = (const uint8_t *)p;
    }
    close(fd);
  }
  char name[512];
  out_name(o, path, name, sizeof(name), ".wav");
  if (iamfb_wav_open(&in->wav, name, o->rate ? o->rate : 48000, o->bit_depth, (uint32_t)in->channels) != 0) {
    fprintf(stderr, "%s can't opened.\n", name);
    return -1;
  }
  if (o->metadata) {
    out_name(o, path, name, sizeof(name), ".met");
    in->met = fopen(name, "w+");
    if (!in->met) fprintf(stderr, "%s can't opened.\n", name);
  }
  return 0;
}

static void input_close(input_t *in) {
  iamfb_wav_close(&in->wav);
  if (in->met) fclose(in->met);
  if (in->dec) IAMF_decoder_close(in->dec);
  if (in->data) munmap((void *)in->data, in->size);
  iamfb_mp4_close(&in->mp4);
  free(in->pcm);
}

static int ensure_pcm(input_t *in, const options *o, int units) {
  const IAMF_StreamInfo *info = IAMF_decoder_get_stream_info(in->dec);
  const size_t need = (size_t)(o->bit_depth / 8) * info->max_frame_size * (size_t)in->channels * (size_t)units;
  if (need <= in->pcm_room) return 0;
  free(in->pcm);
  in->pcm = malloc(need ? need : 1);
  in->pcm_room = in->pcm ? need : 0;
  return in->pcm ? 0 : -1;
}

static void frame_out(input_t *in, const options *o, int samples, int frames) {
  in->frames += frames;
  in->samples += samples;
  iamfb_wav_write(&in->wav, in->pcm, (size_t)(o->bit_depth / 8) * (size_t)samples * (size_t)in->channels);
  if (o->metadata && in->met) {
    IAMF_extradata md;
    int64_t pts;
    if (IAMF_decoder_get_last_metadata(in->dec, &pts, &md) == IAMF_OK) met_write(in->met, pts, &md);
    met_release(&md);
  }
}

/* ---- configuration: descriptor OBUs at the head of a bitstream / in the sample entry of the MP4 track ---- */
static int bitstream_configure(input_t *in, const options *o) {
  uint32_t used = 0;
  const size_t left = in->size - in->pos;
  if (!in->configured) IAMF_decoder_set_pts(in->dec, 0, 90000);
  int rc = IAMF_decoder_configure(in->dec, in->data + in->pos, (uint32_t)(left > 0x40000000u ? 0x40000000u : left), &used);
  in->pos += used;
  if (rc != IAMF_OK) { fprintf(stderr, "errno: %d, fail to configure decoder.\n", rc); return rc; }
  in->configured = 1;
  return ensure_pcm(in, o, 1);
}

static int mp4_configure(input_t *in, const options *o) {
  iamfb_mp4 *m = &in->mp4;
  int64_t st;
  if (o->start_s) st = (int64_t)o->start_s * 90000;
  else {
    const double r = (double)m->skip * 90000;
    st = (int64_t)(r / (m->media_timescale ? m->media_timescale : 1) + 0.5f);
    printf("skip %d/%d pts is %" PRId64 "/90000\n", (int)m->skip, (int)m->media_timescale, st);
  }
  IAMF_decoder_set_pts(in->dec, -st, 90000);
  if (o->start_s) {
    /* walk whole samples until the start time is used up; the sample that crosses it is consumed too (mp4iamfpar.c:203-234) */
    int64_t left = (int64_t)o->start_s * m->movie_timescale + m->skip;
    while (left > 0) {
      if (in->next_sample >= m->n_samples) { fprintf(stderr, "invalid starting time for %s\n", in->path); return -1; }
      const int64_t d = m->samples[in->next_sample++].delta;
      if (left > d) left -= d;
      else break;
    }
  }
  uint32_t di = in->next_sample < m->n_samples ? m->samples[in->next_sample].desc_index : 1;
  if (di < 1 || (int)di > m->n_desc) di = 1;
  in->last_desc = di;
  int rc = IAMF_decoder_configure(in->dec, m->desc[di - 1].obus, m->desc[di - 1].size, 0);
  if (rc != IAMF_OK) { fprintf(stderr, "errno: %d, fail to configure decoder.\n", rc); return rc; }
  in->configured = 1;
  return ensure_pcm(in, o, 1);
}

/* next MP4 packet: the sample, with the descriptor OBUs of its sample entry in front when the entry changed */
static const uint8_t *mp4_packet(input_t *in, uint32_t *size, uint8_t **owned) {
  iamfb_mp4 *m = &in->mp4;
  *owned = 0;
  if (in->next_sample >= m->n_samples) return 0;
  const iamfb_mp4_sample *s = &m->samples[in->next_sample++];
  if (s->desc_index != in->last_desc && s->desc_index >= 1 && (int)s->desc_index <= m->n_desc) {
    const uint32_t dn = m->desc[s->desc_index - 1].size;
    uint8_t *buf = (uint8_t *)malloc((size_t)dn + s->size);
    if (!buf) return 0;
    memcpy(buf, m->desc[s->desc_index - 1].obus, dn);
    memcpy(buf + dn, m->data + s->offset, s->size);
    in->last_desc = s->desc_index;
    *owned = buf;
    *size = dn + s->size;
    return buf;
  }
  *size = s->size;
  return m->data + s->offset;
}

/* ---- one file, the way the reference's player steps it: one temporal unit per IAMF_decoder_decode ---- */
static int play_one(const options *o, const char *path) {
  input_t in;
  int ret = input_open(&in, o, path);
  if (ret == 0) ret = o->input_mode == 1 ? mp4_configure(&in, o) : bitstream_configure(&in, o);
  while (ret == 0 && !in.ended) {
    int n;
    if (o->input_mode == 1) {
      uint32_t size = 0;
      uint8_t *owned = 0;
      const uint8_t *pkt = mp4_packet(&in, &size, &owned);
      if (!pkt) in.ended = 1;
      n = pkt ? IAMF_decoder_decode(in.dec, pkt, (int32_t)size, 0, in.pcm) : IAMF_decoder_decode(in.dec, 0, 0, 0, in.pcm);
      free(owned);
    } else {
      uint32_t used = 0;
      const size_t left = in.size - in.pos;
      if (!left) in.ended = 1;
      n = left ? IAMF_decoder_decode(in.dec, in.data + in.pos, (int32_t)(left > 0x40000000u ? 0x40000000u : left), &used, in.pcm)
               : IAMF_decoder_decode(in.dec, 0, 0, &used, in.pcm);
      in.pos += used;
      if (n == IAMF_ERR_INVALID_STATE && left) {          /* new descriptors in the stream: configure again from here */
        printf("state change to invalid, need reconfigure.\n");
        if (bitstream_configure(&in, o) != 0) break;
        continue;
      }
      if (left && !used && n <= 0) in.ended = 1;          /* a truncated tail: nothing more can be decoded (flush follows) */
      if (in.ended && left) n = IAMF_decoder_decode(in.dec, 0, 0, &used, in.pcm);
    }
    if (n > 0) frame_out(&in, o, n, 1);
    else if (n < 0 && n != IAMF_ERR_INVALID_STATE && !in.ended) { ret = n; break; }
  }
  fprintf(stderr, "===================== Get %ld frames\n", in.frames);
  fprintf(stderr, "===================== Get %ld samples\n", in.samples);
  input_close(&in);
  return ret;
}

/* ---- several bitstream files together: K temporal units of every file per call, one device pass for all ---- */
#define BATCH_UNITS 8
static int play_batch(const options *o) {
  const int n = o->n_inputs;
  input_t *in = (input_t *)calloc((size_t)n, sizeof(*in));
  IAMF_DecoderHandle *hs = (IAMF_DecoderHandle *)calloc((size_t)n, sizeof(*hs));
  const uint8_t **data = (const uint8_t **)calloc((size_t)n, sizeof(*data));
  int32_t *size = (int32_t *)calloc((size_t)n, sizeof(*size));
  uint32_t *used = (uint32_t *)calloc((size_t)n, sizeof(*used));
  void **pcm = (void **)calloc((size_t)n, sizeof(*pcm));
  int *got = (int *)calloc((size_t)n, sizeof(*got)), *units = (int *)calloc((size_t)n, sizeof(*units));
  int ret = (in && hs && data && size && used && pcm && got && units) ? 0 : -1, opened = 0;
  for (int i = 0; ret == 0 && i < n; ++i, ++opened) {
    ret = input_open(&in[i], o, o->inputs[i]);
    if (ret == 0) ret = bitstream_configure(&in[i], o);
    if (ret == 0) ret = ensure_pcm(&in[i], o, BATCH_UNITS);
    hs[i] = in[i].dec;
    pcm[i] = in[i].pcm;
  }
  int flushing = 0;
  while (ret == 0) {
    int live = 0;
    for (int i = 0; i < n; ++i) {
      const size_t left = in[i].size - in[i].pos;
      data[i] = flushing ? 0 : in[i].data + in[i].pos;
      size[i] = flushing ? 0 : (int32_t)(left > 0x40000000u ? 0x40000000u : left);
      live += (!flushing && left && !in[i].ended) ? 1 : 0;
      if (in[i].ended) size[i] = 0;
    }
    if (!flushing && !live) { flushing = 1; continue; }
    int rc = IAMF_decoder_decode_batch_units(hs, n, data, size, used, pcm, got, BATCH_UNITS, units);
    if (rc != IAMF_OK) { ret = rc; break; }
    for (int i = 0; i < n; ++i) {
      if (!flushing) {
        in[i].pos += used[i];
        if (size[i] && !used[i] && got[i] <= 0) in[i].ended = 1;
      }
      if (got[i] > 0) frame_out(&in[i], o, got[i], flushing ? 1 : units[i]);
    }
    if (flushing) break;
  }
  for (int i = 0; i < opened; ++i) {
    fprintf(stderr, "%s: %ld frames, %ld samples\n", in[i].path, in[i].frames, in[i].samples);
    input_close(&in[i]);
  }
  free(in); free(hs); free(data); free(size); free(used); free(pcm); free(got); free(units);
  return ret;
}

int main(int argc, char **argv) {
  options o;
  memset(&o, 0, sizeof(o));
  o.peak = -1.f; o.bit_depth = 16; o.mix_id = UINT64_MAX; o.sound_system = -1;
  int probe = 0;
  o.inputs = (char **)calloc((size_t)argc, sizeof(char *));
  if (argc < 2 || !o.inputs) { usage(argv[0]); return -1; }
  for (int a = 1; a < argc; ++a) {
    const char *s = argv[a];
    if (s[0] != '-') { o.inputs[o.n_inputs++] = argv[a]; continue; }
    if (!strcmp(s, "-p") && a + 1 < argc) o.peak = strtof(argv[++a], 0);
    else if (!strcmp(s, "-l") && a + 1 < argc) o.loudness = strtof(argv[++a], 0);
    else if (!strcmp(s, "-d") && a + 1 < argc) o.bit_depth = (uint32_t)strtof(argv[++a], 0);
    else if (!strcmp(s, "-m")) o.metadata = 1;
    else if (!strcmp(s, "-ts") && a + 1 < argc) o.start_s = (uint32_t)strtoul(argv[++a], 0, 10);
    else if (!strcmp(s, "-r") && a + 1 < argc) o.rate = (uint32_t)strtoul(argv[++a], 0, 10);
    else if (!strcmp(s, "-mp") && a + 1 < argc) o.mix_id = strtoull(argv[++a], 0, 10);
    else if (!strcmp(s, "-disable_limiter")) o.no_limiter = 1;
    else if (!strcmp(s, "-probe")) probe = 1;
    else if (s[1] == 'h') { usage(argv[0]); return 0; }
    else if (s[1] == 'o') o.output_mode = atoi(s + 2);
    else if (s[1] == 'i') o.input_mode = atoi(s + 2);
    else if (s[1] == 's') {
      if (s[2] == 'b') o.layout_type = 3;
      else {
        o.sound_system = atoi(s + 2);
        if (o.sound_system >= 0 && o.sound_system < SOUND_SYSTEM_END && s[2]) o.layout_type = 2;
        else fprintf(stderr, "Invalid output layout of sound system %d\n", o.sound_system);
      }
    }
  }
  int rc = 0;
  if (probe) {   /* -probe: print the sample table the MP4 reader found (no decoder, no device needed) */
    for (int i = 0; i < o.n_inputs; ++i) {
      iamfb_mp4 m;
      const int r = iamfb_mp4_open(&m, o.inputs[i]);
      if (r) { printf("error %d\n", r); rc = -1; continue; }
      printf("samples %zu movie_timescale %u media_timescale %u skip %" PRId64 " descriptions %d", m.n_samples, m.movie_timescale, m.media_timescale, m.skip, m.n_desc);
      for (int k = 0; k < m.n_desc; ++k) printf(" %u", m.desc[k].size);
      printf("\n");
      for (size_t k = 0; k < m.n_samples; ++k)
        printf("%" PRIu64 " %u %u %u\n", m.samples[k].offset, m.samples[k].size, m.samples[k].delta, m.samples[k].desc_index);
      iamfb_mp4_close(&m);
    }
    free(o.inputs);
    return rc < 0 ? 1 : 0;
  }
  if (o.input_mode != 1 && o.start_s) { fprintf(stderr, "ERROR: -ts is valid when mp4 file is used as input.\n"); usage(argv[0]); rc = -1; }
  else if (!o.layout_type) { usage(argv[0]); fprintf(stderr, "invalid output sound system %d\n", o.sound_system); }
  else if (o.output_mode != 2 || o.input_mode < 0 || o.input_mode > 1) fprintf(stderr, "invalid output mode %d\n", o.output_mode);
  else if (!o.n_inputs) { usage(argv[0]); rc = -1; }
  else if (o.n_inputs > 1 && o.input_mode == 0 && !o.metadata) {
    rc = play_batch(&o);
    if (rc == IAMF_ERR_BAD_ARG) {   /* files of different pipeline signatures: one after the other */
      fprintf(stderr, "inputs differ in their rendering pipeline: played one by one\n");
      rc = 0;
      for (int i = 0; i < o.n_inputs; ++i) rc |= play_one(&o, o.inputs[i]);
    }
  } else {
    for (int i = 0; i < o.n_inputs; ++i) rc |= play_one(&o, o.inputs[i]);
  }
  free(o.inputs);
  return rc < 0 ? 1 : 0;
}
