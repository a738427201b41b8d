/*
 * ref_harness.c - TEST / BENCHMARK INFRASTRUCTURE ONLY.
 *
 * The CPU baseline of SURVEY 8(d): a C harness that drives a library exporting the reference's public API
 * (include/IAMF_decoder.h: the UNMODIFIED reference compiled into oracle/_ref/libiamf_ref.so) exactly like
 * test/tools/iamfplayer/player/iamfplayer.c:380-650 does - open, the player's setters, configure, decode per temporal
 * unit, flush, close - with one worker per host core, every worker rendering whole streams start to finish (streams
 * statically partitioned).  No Python, no copies of the PCM: what is timed is the reference.
 *
 *   int ref_harness_run(const ref_job *job, double *seconds, long long *samples)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "IAMF_decoder.h"

typedef struct ref_stream {
  const uint8_t *desc;      /* descriptor OBUs */
  int desc_len;
  int n_units;
  const uint8_t **units;    /* temporal units */
  const int *unit_len;
} ref_stream;

typedef struct ref_job {
  int n_streams;            /* distinct streams (cycled) */
  const ref_stream *streams;
  int renders;              /* stream renders in total, spread over the threads */
  int threads;
  int sound_system;         /* -1: binaural */
  int bit_depth, rate, limiter;
  float loudness, threshold_db;
  int out_channels;
} ref_job;

static int render_stream(const ref_job *job, const ref_stream *st, void *pcm, long long *samples) {
  IAMF_DecoderHandle h = IAMF_decoder_open();
  uint32_t rsize = 0;
  int ret;
  if (!h) return -1;
  IAMF_decoder_peak_limiter_set_threshold(h, job->threshold_db);
  IAMF_decoder_set_normalization_loudness(h, job->loudness);
  IAMF_decoder_set_bit_depth(h, (uint32_t)job->bit_depth);
  if (!job->limiter) IAMF_decoder_peak_limiter_enable(h, 0);
  if (job->rate) IAMF_decoder_set_sampling_rate(h, (uint32_t)job->rate);
  if (job->sound_system < 0) IAMF_decoder_output_layout_set_binaural(h);
  else IAMF_decoder_output_layout_set_sound_system(h, (IAMF_SoundSystem)job->sound_system);
  IAMF_decoder_set_pts(h, 0, 90000);
  {
    /* like the player, the block handed to configure runs past the descriptors (iamfplayer.c:574) */
    const int n = st->desc_len + (st->n_units ? st->unit_len[0] : 0);
    uint8_t *blob = (uint8_t *)malloc((size_t)n);
    memcpy(blob, st->desc, (size_t)st->desc_len);
    if (st->n_units) memcpy(blob + st->desc_len, st->units[0], (size_t)st->unit_len[0]);
    ret = IAMF_decoder_configure(h, blob, (uint32_t)n, &rsize);
    free(blob);
    if (ret != IAMF_OK) { IAMF_decoder_close(h); return -2; }
  }
  for (int u = 0; u < st->n_units; ++u) {
    ret = IAMF_decoder_decode(h, st->units[u], st->unit_len[u], &rsize, pcm);
    if (ret > 0) *samples += ret;
  }
  ret = IAMF_decoder_decode(h, 0, 0, &rsize, pcm);
  if (ret > 0) *samples += ret;
  IAMF_decoder_close(h);
  return 0;
}

/* one worker PROCESS per host core (fork): the reference allocates and frees several buffers per frame
 * (IAMF_decoder.c:2536-2651 and others), and worker threads of one process contend for the allocator - measured 15 %
 * slower than processes on 16 cores - so processes are what shows the reference at its best */
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

int ref_harness_run(const ref_job *job, double *seconds, long long *samples) {
  const int T = job->threads > 0 ? job->threads : 1;
  int go[2], res[2];
  struct timespec t0, t1;
  int err = 0;
  if (pipe(go) || pipe(res)) return -1;
  pid_t *pids = (pid_t *)calloc((size_t)T, sizeof(*pids));
  for (int t = 0; t < T; ++t) {
    const int first = (int)((long long)job->renders * t / T);
    const int count = (int)((long long)job->renders * (t + 1) / T) - first;
    pids[t] = fork();
    if (pids[t] == 0) {
      char c;
      long long n = 0, out[2];
      int bad = 0;
      void *pcm = malloc((size_t)4 * 6144 * 2 * 24);
      close(go[1]);
      close(res[0]);
      if (read(go[0], &c, 1) != 1) _exit(2);            /* start signal */
      for (int r = first; r < first + count; ++r)
        if (render_stream(job, &job->streams[r % job->n_streams], pcm, &n)) bad = 1;
      out[0] = n;
      out[1] = bad;
      if (write(res[1], out, sizeof(out)) != (ssize_t)sizeof(out)) _exit(3);
      _exit(0);
    }
    if (pids[t] < 0) err = 1;
  }
  close(go[0]);
  close(res[1]);
  usleep(20000);                                         /* let every worker reach its read */
  clock_gettime(CLOCK_MONOTONIC, &t0);
  {
    char c = 1;
    for (int t = 0; t < T; ++t)
      if (write(go[1], &c, 1) != 1) err = 1;
  }
  *samples = 0;
  for (int t = 0; t < T; ++t) {
    long long out[2] = {0, 1};
    if (read(res[0], out, sizeof(out)) != (ssize_t)sizeof(out)) err = 1;
    *samples += out[0];
    err |= (int)out[1];
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  for (int t = 0; t < T; ++t)
    if (pids[t] > 0) waitpid(pids[t], 0, 0);
  close(go[1]);
  close(res[0]);
  *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  free(pids);
  return err;
}
