/* iamfb_wav.h - canonical 44-byte RIFF/WAVE PCM writer of the player (the role dep_external/src/wav/dep_wavwriter.c plays
 * for the reference's iamfplayer: same header layout, so the files are byte-identical). */
#ifndef IAMFB_WAV_H_
#define IAMFB_WAV_H_
#include <stdint.h>
#include <stdio.h>

typedef struct iamfb_wav {
  FILE *f;
  uint32_t rate, bits, channels;
  uint64_t data_bytes;
} iamfb_wav;

int iamfb_wav_open(iamfb_wav *w, const char *path, uint32_t rate, uint32_t bits, uint32_t channels);
int iamfb_wav_write(iamfb_wav *w, const void *pcm, size_t bytes);
int iamfb_wav_close(iamfb_wav *w);   /* patches the two length fields */
#endif
