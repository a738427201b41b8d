/* Mutation fuzzer for the host layer of the drop-in library (OBU parser, descriptor database, parameter blocks, codec glue)
 * and for the player's MP4 reader.  Test infrastructure: built with -fsanitize=address,undefined from the host sources
 * (tests/fuzz/run.sh) and run on the CPU box - without a GPU IAMF_decoder_configure fails after everything has been parsed
 * (no engine context), with tests/fuzz/engine_stub.c linked in place of libiamf_b200.so the decode path runs too.
 *
 *   fuzz_host <iterations> <seed> file...      (each file: descriptor OBUs followed by temporal units, or an .mp4)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "IAMF_decoder.h"
#include "iamfb_mp4.h"

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd(void) {
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (uint32_t)(rng_state >> 16);
}

static size_t mutate(const uint8_t *src, size_t n, uint8_t *dst, size_t cap) {
  memcpy(dst, src, n);
  size_t len = n;
  const int edits = 1 + (int)(rnd() % 4);
  for (int e = 0; e < edits && len > 0; ++e) {
    const uint32_t kind = rnd() % 11;
    const size_t at = (rnd() & 1) ? rnd() % len : rnd() % (len < 400 ? len : 400);    /* half of the edits in the descriptors */
    switch (kind) {
      case 0: dst[at] ^= (uint8_t)(1u << (rnd() % 8)); break;
      case 1: dst[at] = (uint8_t)rnd(); break;
      case 2: dst[at] = 0xff; break;
      case 3: dst[at] = 0x80; break;
      case 4: dst[at] = 0; break;
      case 5: len = at + 1; break;                                          /* truncate */
      case 6: {                                                             /* duplicate a span */
        size_t span = 1 + rnd() % 32;
        if (at + span > len) span = len - at;
        if (len + span <= cap) { memmove(dst + at + span, dst + at, len - at); len += span; }
        break;
      }
      case 8: {                                                             /* a size / count field blown up: 5-byte leb128 */
        static const uint8_t big[5] = {0xff, 0xff, 0xff, 0xff, 0x0f};
        size_t k = 1 + rnd() % 5;
        if (at + k > len) k = len - at;
        memcpy(dst + at, big + 5 - k, k);
        if (k > 1) dst[at] |= 0x80;
        break;
      }
      case 9: dst[at] = (uint8_t)(rnd() % 9); break;                       /* a small count / enum value */
      case 10: {                                                            /* a span copied over another place (OBUs re-ordered / repeated) */
        size_t span = 1 + rnd() % 64, to = rnd() % len;
        if (at + span > len) span = len - at;
        if (to + span > len) span = len - to;
        memmove(dst + to, dst + at, span);
        break;
      }
      default: {                                                            /* delete a span */
        size_t span = 1 + rnd() % 16;
        if (at + span > len) span = len - at;
        memmove(dst + at, dst + at + span, len - at - span);
        len -= span;
        break;
      }
    }
  }
  return len;
}

static long n_configured, n_units;
static void run_decoder(const uint8_t *data, size_t len) {
  IAMF_DecoderHandle h = IAMF_decoder_open();
  if (!h) return;
  IAMF_decoder_output_layout_set_sound_system(h, (IAMF_SoundSystem)(rnd() % 13));
  if (rnd() % 4 == 0) IAMF_decoder_output_layout_set_binaural(h);
  IAMF_decoder_set_bit_depth(h, 16);
  uint32_t used = 0;
  int rc = IAMF_decoder_configure(h, data, (uint32_t)len, &used);
  if (rc == IAMF_OK) {
    ++n_configured;
    /* the caller's buffer as the API asks for it: max_frame_size samples of the widest layout (24 channels), 32 bits */
    static uint8_t pcm[1 << 24];
    IAMF_StreamInfo *si = IAMF_decoder_get_stream_info(h);
    if (!si || (uint64_t)si->max_frame_size * 24 * 4 > sizeof(pcm)) { IAMF_decoder_close(h); return; }
    size_t off = used;
    for (int k = 0; k < 8 && off < len; ++k) {
      uint32_t rs = 0;
      int n = IAMF_decoder_decode(h, data + off, (int32_t)(len - off), &rs, pcm);
      if (n < 0 || rs == 0) break;
      if (n > 0) ++n_units;
      off += rs;
    }
    IAMF_decoder_decode(h, NULL, 0, NULL, pcm);
  }
  IAMF_decoder_close(h);
}

/* three handles configured from the pristine descriptors, stepped together through the batch entry on (differently)
 * damaged temporal units - ragged buffers, units that stop early, handles that run dry */
static void run_batch(const uint8_t *good, size_t good_len, const uint8_t *data, size_t len) {
  enum { NH = 3 };
  IAMF_DecoderHandle h[NH];
  uint32_t used = 0;
  int ok = 1;
  for (int i = 0; i < NH; ++i) {
    h[i] = IAMF_decoder_open();
    IAMF_decoder_output_layout_set_sound_system(h[i], SOUND_SYSTEM_A);
    IAMF_decoder_set_bit_depth(h[i], 16);
    if (IAMF_decoder_configure(h[i], good, (uint32_t)good_len, &used) != IAMF_OK) ok = 0;
  }
  if (ok && used < len) {
    static uint8_t pcm[NH][1 << 20];
    const uint8_t *d[NH];
    int32_t sz[NH];
    uint32_t rs[NH];
    void *out[NH];
    int ret[NH], done[NH];
    for (int i = 0; i < NH; ++i) {
      const size_t cut = i == 0 ? 0 : rnd() % (len - used);
      d[i] = data + used;
      sz[i] = (int32_t)(len - used - cut);
      out[i] = pcm[i];
    }
    for (int k = 0; k < 3; ++k) {
      if (IAMF_decoder_decode_batch_units(h, NH, d, sz, rs, out, ret, 2, done) != IAMF_OK) break;
      int any = 0;
      for (int i = 0; i < NH; ++i) {
        if (rs[i] > (uint32_t)sz[i]) abort();
        d[i] += rs[i]; sz[i] -= (int32_t)rs[i];
        any |= rs[i] != 0;
        if (ret[i] > 0) ++n_units;
      }
      if (!any) break;
    }
  }
  for (int i = 0; i < NH; ++i) IAMF_decoder_close(h[i]);
}

static void run_mp4(const char *path, const uint8_t *data, size_t len) {
  char tmp[256];
  snprintf(tmp, sizeof(tmp), "%s.fuzz.tmp", path);
  FILE *f = fopen(tmp, "wb");
  if (!f) return;
  fwrite(data, 1, len, f);
  fclose(f);
  iamfb_mp4 m;
  if (iamfb_mp4_open(&m, tmp) == 0) {
    volatile uint8_t sink = 0;
    for (int i = 0; i < m.n_desc; ++i)
      for (uint32_t k = 0; k < m.desc[i].size; ++k) sink ^= m.desc[i].obus[k];
    /* what the player does with the table: every sample it would hand to the decoder lies inside the mapped file */
    for (size_t i = 0; i < m.n_samples; ++i) {
      const iamfb_mp4_sample *sp = &m.samples[i];
      if (sp->offset > m.size || sp->size > m.size - sp->offset) abort();      /* the reader must have refused the file */
      if (sp->size) sink ^= m.data[sp->offset] ^ m.data[sp->offset + sp->size - 1];
    }
    (void)sink;
    iamfb_mp4_close(&m);
  }
  remove(tmp);
}

int main(int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s <iterations> <seed> file...\n", argv[0]); return 2; }
  const long iters = atol(argv[1]);
  rng_state ^= (uint64_t)atoll(argv[2]) * 0x9E3779B97F4A7C15ull;
  for (int fi = 3; fi < argc; ++fi) {
    FILE *f = fopen(argv[fi], "rb");
    if (!f) { perror(argv[fi]); return 2; }
    static uint8_t src[1 << 20], buf[(1 << 20) + 4096];
    const size_t n = fread(src, 1, sizeof(src), f);
    fclose(f);
    const size_t nl = strlen(argv[fi]);
    const int is_mp4 = nl > 4 && strcmp(argv[fi] + nl - 4, ".mp4") == 0;
    for (long it = 0; it < iters; ++it) {
      const size_t len = it == 0 ? (memcpy(buf, src, n), n) : mutate(src, n, buf, sizeof(buf));
      /* a heap copy of exactly len bytes: reads past the end are caught by the sanitizer */
      uint8_t *exact = (uint8_t *)malloc(len ? len : 1);
      memcpy(exact, buf, len);
      if (getenv("FUZZ_DUMP")) {      /* the input about to run, for reproducing a finding */
        FILE *df = fopen(getenv("FUZZ_DUMP"), "wb");
        if (df) { fwrite(exact, 1, len, df); fclose(df); }
      }
      if (is_mp4) run_mp4(argv[fi], exact, len);
      else if (it % 4 == 3 && len >= n) run_batch(src, n, exact, len);     /* (descriptors untouched only when nothing was cut in front) */
      else run_decoder(exact, len);
      free(exact);
    }
    fprintf(stderr, "%s: %ld inputs ok (%ld configured, %ld temporal units decoded so far)\n", argv[fi], iters, n_configured, n_units);
  }
  return 0;
}
