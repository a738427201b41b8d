/*
 * oracle_pipeline.c - TEST INFRASTRUCTURE ONLY (see iamf_oracle.h).
 * Per-stream driver that chains the stage restatements in the order of the reference's
 * iamf_decoder_internal_decode (src/iamf_dec/IAMF_decoder.c:3303-3525) for everything after core decode:
 *   reconstruct (demix / ambisonics map) -> render -> trim -> element mix gain -> mix -> output mix gain
 *   -> resample -> loudness -> limiter -> quantise/interleave, plus the end-of-stream flush (:3250-3301).
 * Codec delay is 0 on this path (Opus/FLAC/ipcm), so demixer skip / renderer offset are 0.
 *
 * Known, documented divergence from the reference: output rows that its HOA renderer never writes (row 23 of sound
 * system H, h2m_rdr.c:1010,1114-1135) are zero here; in the reference they hold stale buffer contents, which are
 * zero for single-element presentations and garbage-dependent otherwise.
 */
#include <stdlib.h>
#include <string.h>

#include "iamf_oracle.h"

struct OrcStream {
  OrcStreamCfg cfg;
  OrcDemixer *dmx[2];
  OrcDownmixer *dmr[2];
  OrcHrtf *hrtf[2];
  OrcResampler *rs;
  OrcLimiter *lim;
  float *buf[2][2]; /* per element: reconstructed, rendered */
  float *mix, *tmp;
  int cap; /* floats per buffer */
};

OrcStream *orc_stream_open(const OrcStreamCfg *cfg) {
  OrcStream *s = (OrcStream *)calloc(1, sizeof(*s));
  s->cfg = *cfg;
  /* generous: resampled frame can be up to 2x+ the input frame (IAMF_decoder.c:3225-3226) */
  s->cap = ORC_MAX_OUT_CH * (cfg->frame_size * (cfg->out_rate / cfg->in_rate + 2) + ORC_LIM_MAX_DELAY);
  for (int e = 0; e < cfg->n_elements; ++e) {
    const OrcElementCfg *el = &cfg->el[e];
    s->buf[e][0] = (float *)calloc(s->cap, sizeof(float));
    s->buf[e][1] = (float *)calloc(s->cap, sizeof(float));
    if (el->type == ORC_EL_CHANNEL) {
      /* iamf_stream_scale_demixer_configure, IAMF_decoder.c:2351-2401 */
      OrcDemixer *d = orc_demixer_open(cfg->frame_size);
      orc_demixer_set_layout(d, el->layout);
      orc_demixer_set_channels_order(d, el->chs_in, el->n_in);
      orc_demixer_set_output_gain(d, el->out_gain_ch, el->out_gain, el->n_out_gain);
      if (el->has_demix_info) orc_demixer_set_demixing_info(d, el->default_mode, el->default_w_idx);
      { /* iamf_stream_scale_decoder_set_default_recon_gain, IAMF_decoder.c:2202-2236 */
        int chs[ORC_MAX_LAYOUT_CH], n = 0;
        float ones[ORC_MAX_LAYOUT_CH];
        uint32_t fl = 0;
        if (el->selected_layer > 0) {
          fl = orc_recon_flags(el->first_layer_layout, el->layout);
          n = orc_recon_order(el->layout, fl, chs);
          for (int i = 0; i < n; ++i) ones[i] = 1.f;
        }
        orc_demixer_set_recon_gain(d, n, chs, ones, fl);
      }
      /* first decode: iamf_stream_decoder_update_delay -> demixer_set_frame_offset(delay = 0), IAMF_decoder.c:2176-2185:
         this is what installs the Hann cross-fade windows */
      orc_demixer_set_frame_offset(d, 0);
      s->dmx[e] = d;
      if (el->use_dmr) { /* iamf_stream_renderer_enable_downmix, IAMF_decoder.c:2448-2478 */
        s->dmr[e] = orc_dmr_open(el->layout, el->dmr_out_layout);
        if (s->dmr[e]) orc_dmr_set_mode_weight(s->dmr[e], el->default_mode, el->default_w_idx);
      }
    }
  }
  for (int e = 0; e < cfg->n_elements; ++e)
    if (cfg->el[e].hrtf_taps) /* IAMF_element_renderer_init_M2B / _H2B, IAMF_decoder.c:2491-2508 */
      s->hrtf[e] = orc_hrtf_open(cfg->el[e].type == ORC_EL_CHANNEL ? orc_layout_channel_count(cfg->el[e].layout) : cfg->el[e].n_in,
                                 cfg->el[e].hrtf_taps);
  s->mix = (float *)calloc(s->cap, sizeof(float));
  s->tmp = (float *)calloc(s->cap, sizeof(float));
  if (cfg->in_rate != cfg->out_rate) s->rs = orc_resampler_open(cfg->out_channels, cfg->in_rate, cfg->out_rate, 4);
  if (cfg->limiter) {
    s->lim = (OrcLimiter *)malloc(sizeof(OrcLimiter));
    /* IAMF_decoder.c:3809-3815 with audio_defines.h:38-41; the rate is the requested OUTPUT rate */
    orc_limiter_init(s->lim, cfg->limiter_threshold_db, cfg->out_rate, cfg->out_channels, 0.001f, 0.200f, 240);
  }
  return s;
}

void orc_stream_close(OrcStream *s) {
  if (!s) return;
  for (int e = 0; e < 2; ++e) {
    orc_demixer_close(s->dmx[e]);
    orc_dmr_close(s->dmr[e]);
    orc_hrtf_close(s->hrtf[e]);
    free(s->buf[e][0]); free(s->buf[e][1]);
  }
  orc_resampler_close(s->rs);
  free(s->lim); free(s->mix); free(s->tmp); free(s);
}

int orc_stream_decode(OrcStream *s, float *const *in, const OrcFrameParams *fp, float out_gain_const,
                      const float *out_gain_ramp, int strim, int etrim, void *pcm) {
  const OrcStreamCfg *cfg = &s->cfg;
  int n = cfg->frame_size, co = cfg->out_channels;
  int nt = n; /* samples left after trimming */
  const float *rendered[2];
  /* a frame trimmed completely is decoded but dropped before rendering, IAMF_decoder.c:3354-3358 */
  int drop = (strim == n || etrim == n);

  for (int e = 0; e < cfg->n_elements; ++e) {
    const OrcElementCfg *el = &cfg->el[e];
    float *rec = s->buf[e][0], *ren = s->buf[e][1];
    if (el->type == ORC_EL_CHANNEL) {
      /* iamf_stream_scale_decoder_demix, IAMF_decoder.c:2324-2349 */
      if (fp[e].has_recon)
        orc_demixer_set_recon_gain(s->dmx[e], fp[e].n_recon, fp[e].recon_ch, fp[e].recon_gain, fp[e].recon_flags);
      if (fp[e].dmx_mode > -1) orc_demixer_set_demixing_info(s->dmx[e], fp[e].dmx_mode, -1);
      if (orc_demixer_demix(s->dmx[e], rec, in[e], n) < 0) return -1;
    } else if (el->ambi_mode == 2) {
      orc_ambisonics_projection(el->ambi_matrix, el->n_in /*rows = output channels*/, el->ambi_cols, in[e], rec, n);
    } else {
      orc_ambisonics_mono(el->ambi_map, el->n_in, in[e], rec, n);
    }
    if (drop) continue;
    /* iamf_stream_render, IAMF_decoder.c:2536-2651 */
    memset(ren, 0, sizeof(float) * co * n);
    if (s->hrtf[e]) { /* :2565-2573, :2606-2612: the binaural renderer takes the place of the matrix */
      orc_hrtf_render(s->hrtf[e], rec, ren, n);
    } else if (el->type == ORC_EL_CHANNEL) {
      if (s->dmr[e]) {
        if (fp[e].dmx_mode > -1) orc_dmr_set_mode_weight(s->dmr[e], fp[e].dmx_mode, -1);
        orc_dmr_downmix(s->dmr[e], rec, ren, 0, n, n);
      } else {
        orc_render_m2m(el->mat, el->mat_in, el->mat_out, rec, ren, n);
      }
    } else {
      orc_render_h2m(el->mat, el->mat_in, el->mat_out, el->lfe1, el->lfe2, rec, ren, n);
    }
    /* iamf_frame_trim on the rendered frame, IAMF_decoder.c:3387-3405 (codec delay 0) */
    if (strim || etrim) {
      nt = orc_frame_trim(ren, n, co, strim, etrim, 0);
      if (nt <= 0) return nt;
    }
    /* element mix gain over the remaining samples, IAMF_decoder.c:3425-3433 */
    if (fp[e].gain_ramp) orc_frame_gain_ramp(ren, nt, co, fp[e].gain_ramp);
    else orc_frame_gain_const(ren, nt, co, fp[e].gain_const);
    rendered[e] = ren;
  }
  if (drop) return 0;
  n = nt;

  orc_mix(s->mix, rendered, cfg->n_elements, n, co);                  /* :3459 */
  if (out_gain_ramp) orc_frame_gain_ramp(s->mix, n, co, out_gain_ramp); /* :3463-3469 */
  else orc_frame_gain_const(s->mix, n, co, out_gain_const);

  float *cur = s->mix, *other = s->tmp;
  int ns = n;
  if (s->rs) { /* :3474-3478 */
    ns = orc_resample(s->rs, cur, other, ns);
    float *t = cur; cur = other; other = t;
  }
  if (cfg->loudness_gain != 0.f) orc_loudness(cur, ns, co, cfg->loudness_gain); /* :3480-3484 */
  if (s->lim) { /* :3486-3490 */
    ns = orc_limiter_process(s->lim, cur, other, ns);
    float *t = cur; cur = other; other = t;
  }
  orc_plane2stride(pcm, cur, ns, co, cfg->bit_depth, co); /* :3497-3500 */
  return ns;
}

/* iamf_delay_buffer_handle, IAMF_decoder.c:3250-3301 */
int orc_stream_flush(OrcStream *s, void *pcm) {
  const OrcStreamCfg *cfg = &s->cfg;
  int co = cfg->out_channels;
  int fs = s->lim ? s->lim->delay_size : 0;
  if (!s->lim && !s->rs) return 0;
  float *in = (float *)calloc(s->cap, sizeof(float));
  float *out = (float *)calloc(s->cap, sizeof(float));
  if (s->rs) {
    int rsz = orc_resample(s->rs, 0, out, -1);
    fs += rsz;
    for (int c = 0; c < co; ++c) memcpy(in + c * fs, out + c * rsz, sizeof(float) * rsz);
    memset(out, 0, sizeof(float) * s->cap);
    if (!s->lim) memcpy(out, in, sizeof(float) * co * fs);
  }
  if (s->lim) fs = orc_limiter_process(s->lim, in, out, fs);
  orc_plane2stride(pcm, out, fs, co, cfg->bit_depth, co);
  free(in);
  free(out);
  return fs;
}
