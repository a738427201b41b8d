/* iamfb_wav.c - see iamfb_wav.h */
#include "iamfb_wav.h"

#include <string.h>

static void le16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static void le32(uint8_t *p, uint32_t v) { le16(p, v); le16(p + 2, v >> 16); }

/* RIFF <len> WAVE | fmt  16 <PCM=1> <ch> <rate> <bytes/s> <bytes/frame> <bits> | data <len>   (dep_wavwriter.c:54-75) */
static int put_header(iamfb_wav *w) {
  uint8_t h[44];
  const uint32_t frame = w->bits / 8 * w->channels, len = (uint32_t)w->data_bytes;
  memcpy(h, "RIFF", 4); le32(h + 4, 4 + 8 + 16 + 8 + len); memcpy(h + 8, "WAVEfmt ", 8); le32(h + 16, 16);
  le16(h + 20, 1); le16(h + 22, w->channels); le32(h + 24, w->rate); le32(h + 28, frame * w->rate);
  le16(h + 32, frame); le16(h + 34, w->bits); memcpy(h + 36, "data", 4); le32(h + 40, len);
  return fwrite(h, 1, sizeof(h), w->f) == sizeof(h) ? 0 : -1;
}

int iamfb_wav_open(iamfb_wav *w, const char *path, uint32_t rate, uint32_t bits, uint32_t channels) {
  memset(w, 0, sizeof(*w));
  w->f = fopen(path, "wb");
  if (!w->f) return -1;
  w->rate = rate; w->bits = bits; w->channels = channels;
  return put_header(w);
}

int iamfb_wav_write(iamfb_wav *w, const void *pcm, size_t bytes) {
  if (!w->f) return -1;
  if (fwrite(pcm, 1, bytes, w->f) != bytes) return -1;
  w->data_bytes += bytes;
  return 0;
}

int iamfb_wav_close(iamfb_wav *w) {
  if (!w->f) return 0;
  int rc = fseek(w->f, 0, SEEK_SET) == 0 ? put_header(w) : -1;
  if (fclose(w->f) != 0) rc = -1;
  w->f = 0;
  return rc;
}
