/* iamfb_mp4.h - the MP4 / fragmented-MP4 side of the player: finds the IAMF audio track and lists its samples (temporal
 * units).  Own design (the whole file is mapped and the box tree walked once into a flat sample table); the behaviour
 * follows what the reference's player does with such a file (test/tools/iamfplayer/src/mp4demux.c, mp4iamfpar.c):
 *   - the audio track is the first `trak` whose `stsd` holds an `iamf` sample entry (mp4demux.c:415-439);
 *   - the descriptor OBUs are everything behind the 28 bytes of AudioSampleEntry fields of that entry (:512-574);
 *   - `elst` media_time is the number of samples to skip at the start, in `mdhd` time-scale units (:454-494, :266-298);
 *   - samples come from stsc / stsz / stco (or co64), or from moof / traf / tfhd / trun fragments (:659-847, :907-1040);
 *   - when the sample-description index changes between chunks, the new entry's descriptor OBUs are handed to the decoder
 *     in front of the first sample that uses it (mp4iamfpar.c:138-168). */
#ifndef IAMFB_MP4_H_
#define IAMFB_MP4_H_
#include <stddef.h>
#include <stdint.h>

typedef struct iamfb_mp4_sample {
  uint64_t offset;        /* byte offset in the file */
  uint32_t size;
  uint32_t delta;         /* duration in media time-scale units (stts / trun) */
  uint32_t desc_index;    /* 1-based sample description index */
} iamfb_mp4_sample;

typedef struct iamfb_mp4 {
  const uint8_t *data;    /* the mapped file */
  size_t size;
  uint32_t movie_timescale, media_timescale;
  int64_t skip;           /* elst media_time of the track (0 when absent) */
  int n_desc;
  struct { const uint8_t *obus; uint32_t size; } desc[8];
  iamfb_mp4_sample *samples;
  size_t n_samples;
  uint32_t trex_size, trex_duration;   /* defaults of the fragments' samples (mvex / trex) */
} iamfb_mp4;

/* 0 on success; < 0: cannot open (-1), no IAMF track (-2), malformed (-3) */
int iamfb_mp4_open(iamfb_mp4 *m, const char *path);
void iamfb_mp4_close(iamfb_mp4 *m);
#endif
