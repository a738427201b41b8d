/*
 * iamf_oracle.h - CPU restatement of libiamf's post-decode rendering path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under iac_b200/ (the product) may include, link or dlopen anything in
 * oracle/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and
 * only as the checker.
 *
 * Every function is a plain-C (C99, scalar, strict IEEE: build with -ffp-contract=off, no -ffast-math) restatement of
 * the algorithm the reference implements; the reference location it follows is cited beside each declaration
 * (paths relative to the Samsung/iac tree).  Parity of this restatement against the *compiled, unmodified* reference
 * (oracle/_ref/libiamf_ref.so, built by oracle/Makefile) is pinned by tests/test_oracle_vs_ref.py and by the golden
 * fixtures under tests/golden/ that were produced by running that compiled reference (tools/make_golden.py).
 */
#ifndef IAMF_ORACLE_H_
#define IAMF_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* IAChannel ids, src/iamf_dec/IAMF_types.h:61-90 */
enum {
  ORC_CH_INVALID = 0, ORC_CH_L7 = 1, ORC_CH_R7 = 2, ORC_CH_C = 3, ORC_CH_LFE = 4, ORC_CH_SL7 = 5, ORC_CH_SR7 = 6,
  ORC_CH_BL7 = 7, ORC_CH_BR7 = 8, ORC_CH_HFL = 9, ORC_CH_HFR = 10, ORC_CH_HBL = 11, ORC_CH_HBR = 12,
  ORC_CH_MONO = 13, ORC_CH_L2 = 14, ORC_CH_R2 = 15, ORC_CH_TL = 16, ORC_CH_TR = 17, ORC_CH_L3 = 18, ORC_CH_R3 = 19,
  ORC_CH_SL5 = 20, ORC_CH_SR5 = 21, ORC_CH_HL = 22, ORC_CH_HR = 23, ORC_CH_COUNT = 24,
  ORC_CH_L5 = ORC_CH_L7, ORC_CH_R5 = ORC_CH_R7
};

/* IAChannelLayoutType, include/IAMF_defines.h:196-209 */
enum {
  ORC_LAYOUT_MONO = 0, ORC_LAYOUT_STEREO, ORC_LAYOUT_510, ORC_LAYOUT_512, ORC_LAYOUT_514, ORC_LAYOUT_710,
  ORC_LAYOUT_712, ORC_LAYOUT_714, ORC_LAYOUT_312, ORC_LAYOUT_BINAURAL, ORC_LAYOUT_COUNT
};

#define ORC_MAX_LAYOUT_CH 12
#define ORC_MAX_OUT_CH 24

/* ---- layout tables: src/iamf_dec/IAMF_utils.c:111-196 ---- */
int orc_layout_channel_count(int layout);
int orc_layout_channels(int layout, int *chs);             /* demixed (rendering) order   :117-133 */
int orc_layer_channels(int layout, int *chs);              /* first-layer transmission order :181-196 */
int orc_layout_surround(int layout);
int orc_layout_top(int layout);
/* src/iamf_dec/IAMF_decoder.c:450-531: channels a higher layer adds, in transmission order */
int orc_new_channels(int last_layout, int cur_layout, int *chs);
/* src/iamf_dec/IAMF_decoder.c:371-448 */
uint32_t orc_recon_flags(int l1, int l2);
int orc_recon_order(int layout, uint32_t flags, int *chs);
/* src/iamf_dec/IAMF_decoder.c:533-602 */
int orc_output_gain_channel(int layout, int gain_bit);

/* ---- scalars: src/common/fixedp11_5.c:45-99 ---- */
float orc_q_to_float(int16_t q, int frac);
float orc_qf_to_float(uint8_t q, int frac);
float orc_db2lin(float db);
float orc_get_w(int w_idx);
int orc_calc_w_idx(int offset, int prev);

/* ---- scalable-channel demixer: src/iamf_dec/demixer.c:74-664 ---- */
typedef struct OrcDemixer {
  int frame_size;
  int skip;
  float *hann, *start_win, *stop_win; /* :502-505, :537-563 */
  int layout;
  int chs_in[ORC_MAX_LAYOUT_CH], chs_out[ORC_MAX_LAYOUT_CH];
  int chs_count;
  int n_gain, gain_ch[ORC_MAX_LAYOUT_CH];
  float gain[ORC_MAX_LAYOUT_CH];
  int mode, last_mode, w_idx, last_w_idx;
  int n_recon, recon_ch[ORC_MAX_LAYOUT_CH];
  float recon_gain[ORC_MAX_LAYOUT_CH];
  uint32_t recon_flags;
  float last_sf[ORC_CH_COUNT], last_sfavg[ORC_CH_COUNT];
  float *scratch; /* 6 derived-channel slots like large_buffer :92 */
} OrcDemixer;

OrcDemixer *orc_demixer_open(int frame_size);                                    /* :477-525 */
void orc_demixer_close(OrcDemixer *d);
int orc_demixer_set_frame_offset(OrcDemixer *d, uint32_t offset);                /* :537-563 */
int orc_demixer_set_layout(OrcDemixer *d, int layout);                           /* :565-572 */
void orc_demixer_set_channels_order(OrcDemixer *d, const int *chs, int count);   /* :574-578 */
void orc_demixer_set_output_gain(OrcDemixer *d, const int *chs, const float *g, int count); /* :580-590 */
int orc_demixer_set_demixing_info(OrcDemixer *d, int mode, int w_idx);           /* :592-619 */
void orc_demixer_set_recon_gain(OrcDemixer *d, int count, const int *chs, const float *g, uint32_t flags); /* :621-634 */
int orc_demixer_demix(OrcDemixer *d, float *dst, float *src, uint32_t size);     /* :636-664 (src is modified, like the ref) */

/* ---- parametric down-mix renderer: src/iamf_dec/downmix_renderer.c:53-242 ---- */
typedef struct OrcDownmixer {
  int mode, w_idx;
  int chs_in[ORC_MAX_LAYOUT_CH], chs_out[ORC_MAX_LAYOUT_CH];
  int n_in, n_out;
  float alpha, beta, gamma, delta;
  int w_off;
  float tl_scale; /* deps[TL][1].s == gamma*w, :200-211 */
  int is_input[ORC_CH_COUNT];
} OrcDownmixer;
OrcDownmixer *orc_dmr_open(int in_layout, int out_layout);                        /* :131-176 (NULL if invalid pair) */
void orc_dmr_close(OrcDownmixer *d);
int orc_dmr_set_mode_weight(OrcDownmixer *d, int mode, int w_idx);               /* :180-216 */
int orc_dmr_downmix(OrcDownmixer *d, const float *in, float *out, uint32_t s, uint32_t duration, uint32_t size); /* :218-242 */

/* ---- matrix renderers ---- */
/* src/iamf_dec/m2m_rdr.c:1820-1840 : mat is [m_in][n_out] row-major */
void orc_render_m2m(const float *mat, int m_in, int n_out, const float *in, float *out, int nsamples);
/* src/iamf_dec/h2m_rdr.c:1088-1150 (DISABLE_LFE_HOA==1): mat is [n_out][m_in]; writes n_out(+lfe slots) planar rows
   of stride nsamples into out, which must hold out_channels rows (rows never written by the reference stay as they are) */
void orc_render_h2m(const float *mat, int m_in, int n_out, int lfe1, int lfe2, const float *in, float *out, int nsamples);
/* src/iamf_dec/IAMF_core_decoder.c:105-130 */
void orc_ambisonics_mono(const uint8_t *map, int channels, const float *in, float *out, int frame_size);
void orc_ambisonics_projection(const float *matrix, int rows, int cols, const float *in, float *out, int frame_size);

/* ---- gains / mixing / trimming / output: src/iamf_dec/IAMF_decoder.c ---- */
void orc_gain_linear(float s, float e, int d, int o, uint32_t l, float *g);                 /* :639-645 */
void orc_gain_bezier(float s, float e, int d, float c, int ct, int o, uint32_t l, float *g); /* :647-664 */
void orc_frame_gain_const(float *data, int samples, int channels, float gain);              /* :1392-1398 */
void orc_frame_gain_ramp(float *data, int samples, int channels, const float *gains);       /* :1401-1405 */
int orc_frame_trim(float *data, int samples, int channels, int start, int end, int start_ext); /* :1361-1381 */
void orc_mix(float *dst, const float *const *elems, int n_elems, int samples, int channels); /* :2702-2733 */
void orc_loudness(float *block, int frame_size, int channels, float gain);                  /* :3206-3221 */
void orc_plane2stride(void *dst, const float *src, int frame_size, int channels, uint32_t bit_depth, uint32_t stride); /* :100-167 */

/* ---- Speex resampler as configured by the reference (float build): src/iamf_dec/resample.c ---- */
typedef struct OrcResampler {
  uint32_t in_rate, out_rate, num_rate, den_rate;
  int quality;
  uint32_t nb_channels, filt_len, mem_alloc_size, buffer_size;
  int int_advance, frac_advance;
  float cutoff;
  uint32_t oversample;
  int use_direct;
  int32_t *last_sample;
  uint32_t *samp_frac_num;
  float *mem, *sinc_table;
  uint32_t sinc_table_length;
  int rest_flag; /* Samsung addition, speex_resampler.h */
} OrcResampler;
OrcResampler *orc_resampler_open(uint32_t channels, uint32_t in_rate, uint32_t out_rate, int quality); /* :703-775 + IAMF_decoder.c:1892-1909 (skip_zeros) */
void orc_resampler_close(OrcResampler *r);
/* planar in [ch][frame_size] -> planar out [ch][ret]; IAMF_decoder.c:3223-3248 + resample.c:917-998. frame_size<0 = flush (rest_flag 2) */
int orc_resample(OrcResampler *r, const float *in, float *out, int frame_size);
int orc_resampler_output_latency(const OrcResampler *r); /* :1102-1105 */

/* ---- peak limiter: src/iamf_dec/audio_effect_peak_limiter.c ---- */
#define ORC_LIM_MAX_DELAY 4096
typedef struct OrcLimiter {
  int init, padsize;
  float current_gain, target_start, target_end, attack_sec, release_sec, threshold, current_tc, inc_tc;
  int num_channels;
  float delay[ORC_MAX_OUT_CH][ORC_LIM_MAX_DELAY + 1];
  float peak[ORC_LIM_MAX_DELAY + 1];
  int entry, delay_size, peak_pos;
} OrcLimiter;
void orc_limiter_init(OrcLimiter *l, float threshold_db, int sample_rate, int channels, float atk, float rel, int delay); /* :73-92,211-235 */
int orc_limiter_process(OrcLimiter *l, const float *in, float *out, int frame_size); /* :94-204 */
OrcLimiter *orc_limiter_new(float threshold_db, int sample_rate, int channels, float atk, float rel, int delay);
void orc_limiter_free(OrcLimiter *l);

/* ---- whole path for one stream (driver restating IAMF_decoder.c:3303-3525 for the stages after core decode) ---- */
#define ORC_EL_CHANNEL 0
#define ORC_EL_SCENE 1
typedef struct OrcElementCfg {
  int type;                 /* ORC_EL_CHANNEL / ORC_EL_SCENE */
  int n_in;                 /* decoded (transmitted) channels */
  /* channel-based */
  int layout;               /* reconstructed layout */
  int chs_in[ORC_MAX_LAYOUT_CH];   /* transmission order */
  int n_out_gain; int out_gain_ch[ORC_MAX_LAYOUT_CH]; float out_gain[ORC_MAX_LAYOUT_CH];
  int has_demix_info; int default_mode; int default_w_idx;
  int first_layer_layout; int selected_layer;   /* default recon-gain list, IAMF_decoder.c:2202-2236 */
  int use_dmr; int dmr_out_layout;  /* DMRenderer instead of matrix, IAMF_decoder.c:2448-2478 */
  /* scene-based */
  int ambi_mode;            /* 1 mono map, 2 projection */
  uint8_t ambi_map[16];
  const float *ambi_matrix; /* column-major [cols][rows], q15->float */
  int ambi_cols;
  /* render matrix (M2M [n_in_layout][n_out] or H2M [n_out_mat][n_in]) */
  const float *mat; int mat_in, mat_out; int lfe1, lfe2;
  /* binaural HRTF rendering instead of the matrix (m2b_rdr.c / h2b_rdr.c; oracle_hrtf.c): one Q15 HRIR pair per
     renderer input channel, [mat_in][2][ORC_HRTF_TAPS]; NULL => off */
  const int16_t *hrtf_taps;
} OrcElementCfg;

#define ORC_HRTF_TAPS 256
typedef struct OrcHrtf OrcHrtf;
OrcHrtf *orc_hrtf_open(int n_ch, const int16_t *taps);
void orc_hrtf_close(OrcHrtf *h);
int32_t orc_hrtf_quantise(float x);
/* in [n_ch][n] planar -> out [2][n] planar; the filter state is carried from call to call */
void orc_hrtf_render(OrcHrtf *h, const float *in, float *out, int n);

typedef struct OrcStreamCfg {
  int frame_size, in_rate, out_rate;
  int n_elements; OrcElementCfg el[2];
  int out_channels;
  float loudness_gain;      /* 0 => stage skipped (normalization_loudness==0) */
  int limiter; float limiter_threshold_db;
  int bit_depth;            /* 16/24/32; 0 => float planar-interleaved debug output */
} OrcStreamCfg;

typedef struct OrcFrameParams {   /* per element, per frame */
  int dmx_mode;             /* -1: no demixing parameter this frame */
  int n_recon; int recon_ch[ORC_MAX_LAYOUT_CH]; float recon_gain[ORC_MAX_LAYOUT_CH]; uint32_t recon_flags; int has_recon;
  float gain_const; const float *gain_ramp;   /* element mix gain: ramp (frame_size floats) overrides const */
} OrcFrameParams;

typedef struct OrcStream OrcStream;
OrcStream *orc_stream_open(const OrcStreamCfg *cfg);
void orc_stream_close(OrcStream *s);
/* in[e] = planar decoded frame of element e ([n_in][frame_size], modified like the reference does); returns samples
   per channel written to pcm (interleaved, stride = out_channels), exactly like IAMF_decoder_decode. */
int orc_stream_decode(OrcStream *s, float *const *in, const OrcFrameParams *fp, float out_gain_const,
                      const float *out_gain_ramp, int strim, int etrim, void *pcm);
int orc_stream_flush(OrcStream *s, void *pcm); /* IAMF_decoder.c:3250-3301 */

#ifdef __cplusplus
}
#endif
#endif
