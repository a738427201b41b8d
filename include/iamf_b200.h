/*
 * iamf_b200.h - thin C ABI of the B200 (sm_100a) post-decode rendering engine.
 *
 * This is the seam the reference decoder's host code crosses right after core (codec) decode:
 *   src/iamf_dec/IAMF_core_decoder.c:269-294  iamf_core_decoder_decode  -> planar float frame per audio element
 * and everything the reference then does on the CPU up to the interleaved integer PCM it returns from
 *   src/iamf_dec/IAMF_decoder.c:3303-3525     iamf_decoder_internal_decode
 * runs behind these entry points as CUDA kernels, batched over many independent streams (decoder handles) and many
 * consecutive frames per stream.  Plain C types only; no torch / C++ types cross this boundary.
 *
 * Reference functions each entry point replaces (paths relative to the Samsung/iac tree):
 *   iamfb_plan_create      iamf_stream_scale_demixer_configure   IAMF_decoder.c:2351-2401
 *                          iamf_stream_renderer_open/_enable_downmix            :2448-2508
 *                          IAMF_element_renderer_get_M2M_matrix  m2m_rdr.c:1786-1804
 *                          IAMF_element_renderer_get_H2M_matrix  h2m_rdr.c:1070-1081
 *                          iamf_stream_resampler_open            IAMF_decoder.c:1892-1909 (+ resample.c:527-701,703-775)
 *                          audio_effect_peak_limiter_init        audio_effect_peak_limiter.c:73-92
 *   iamfb_batch_create     iamf_stream_decoder_open buffers      IAMF_decoder.c:2017-2103, demixer_open demixer.c:477-525
 *   iamfb_batch_submit     the per-frame body of iamf_decoder_internal_decode, IAMF_decoder.c:3336-3500:
 *                            demixer_set_recon_gain/_set_demixing_info/demixer_demixing   demixer.c:592-664
 *                            iamf_core_decoder_convert_mono/_projection                   IAMF_core_decoder.c:105-130
 *                            DMRenderer_set_mode_weight/_downmix                          downmix_renderer.c:180-242
 *                            IAMF_element_renderer_render_M2M / _render_H2M               m2m_rdr.c:1820, h2m_rdr.c:1088
 *                            iamf_frame_trim / iamf_frame_gain / iamf_mixer_mix           IAMF_decoder.c:1361-1408,2702-2733
 *                            iamf_resample (speex_resampler_process_interleaved_float)    IAMF_decoder.c:3223-3248
 *                            iamf_loudness_process                                        IAMF_decoder.c:3206-3221
 *                            audio_effect_peak_limiter_process_block                      audio_effect_peak_limiter.c:94-204
 *                            iamf_decoder_plane2stride_out                                IAMF_decoder.c:121-167
 *   iamfb_batch_flush      iamf_delay_buffer_handle              IAMF_decoder.c:3250-3301
 *
 * There is NO CPU fallback: every entry point that needs the device returns IAMFB_ERR_NO_DEVICE / IAMFB_ERR_CUDA
 * (and prints the CUDA error to stderr) when no sm_100a-capable GPU is usable.
 */
#ifndef IAMF_B200_H_
#define IAMF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IAMFB_MAX_ELEMENTS 2    /* IAMF_OBU.c:742-752: <= 2 audio elements per mix presentation */
#define IAMFB_MAX_LAYOUT_CH 12  /* IAMF_types.h:113 */
#define IAMFB_MAX_SCENE_CH 16   /* third-order ambisonics */
#define IAMFB_MAX_OUT_CH 24     /* audio_defines.h:58 */
#define IAMFB_LIMITER_DELAY 240 /* audio_defines.h:41 LIMITER_LookAhead */

enum {
  IAMFB_OK = 0,
  IAMFB_ERR_BAD_ARG = -1,      /* mirrors IAMF_ERR_BAD_ARG        IAMF_defines.h:181-190 */
  IAMFB_ERR_BUFFER_TOO_SMALL = -2,
  IAMFB_ERR_INTERNAL = -3,
  IAMFB_ERR_UNIMPLEMENTED = -6,
  IAMFB_ERR_ALLOC_FAIL = -7,
  IAMFB_ERR_NO_DEVICE = -100,  /* no usable CUDA device: the product has no CPU path */
  IAMFB_ERR_CUDA = -101
};

/* audio element kinds (IAMF_OBU.h AUDIO_ELEMENT_TYPE_*) */
enum { IAMFB_EL_CHANNEL = 0, IAMFB_EL_SCENE = 1 };

/* channel layouts = IAChannelLayoutType, IAMF_defines.h:196-209 */
enum {
  IAMFB_LAYOUT_MONO = 0, IAMFB_LAYOUT_STEREO, IAMFB_LAYOUT_510, IAMFB_LAYOUT_512, IAMFB_LAYOUT_514,
  IAMFB_LAYOUT_710, IAMFB_LAYOUT_712, IAMFB_LAYOUT_714, IAMFB_LAYOUT_312, IAMFB_LAYOUT_BINAURAL
};

/* playback targets = IAMF_SoundSystem (IAMF_defines.h:62-78) followed by binaural */
enum {
  IAMFB_TARGET_A = 0, IAMFB_TARGET_B, IAMFB_TARGET_C, IAMFB_TARGET_D, IAMFB_TARGET_E, IAMFB_TARGET_F,
  IAMFB_TARGET_G, IAMFB_TARGET_H, IAMFB_TARGET_I, IAMFB_TARGET_J, IAMFB_TARGET_712, IAMFB_TARGET_312,
  IAMFB_TARGET_MONO, IAMFB_TARGET_BINAURAL, IAMFB_TARGET_COUNT
};

/* IAChannel ids (IAMF_types.h:61-90) used in chs_in / gain / recon lists */
enum {
  IAMFB_CH_INVALID = 0, IAMFB_CH_L7 = 1, IAMFB_CH_R7, IAMFB_CH_C, IAMFB_CH_LFE, IAMFB_CH_SL7, IAMFB_CH_SR7,
  IAMFB_CH_BL7, IAMFB_CH_BR7, IAMFB_CH_HFL, IAMFB_CH_HFR, IAMFB_CH_HBL, IAMFB_CH_HBR, IAMFB_CH_MONO, IAMFB_CH_L2,
  IAMFB_CH_R2, IAMFB_CH_TL, IAMFB_CH_TR, IAMFB_CH_L3, IAMFB_CH_R3, IAMFB_CH_SL5, IAMFB_CH_SR5, IAMFB_CH_HL,
  IAMFB_CH_HR, IAMFB_CH_COUNT, IAMFB_CH_L5 = IAMFB_CH_L7, IAMFB_CH_R5 = IAMFB_CH_R7
};

/* One audio element of the enabled mix presentation, as the reference sets it up in iamf_stream_new /
 * iamf_stream_decoder_open / iamf_stream_renderer_open (IAMF_decoder.c:1617-1830, 2017-2103, 2480-2508). */
typedef struct iamfb_element_desc {
  int32_t kind;                              /* IAMFB_EL_CHANNEL | IAMFB_EL_SCENE */
  int32_t n_in;                              /* decoded channels handed over by core decode (transmission order) */
  /* -- channel based (scalable channel layout) -- */
  int32_t layout;                            /* reconstructed IAChannelLayoutType (ChannelLayerContext.layout) */
  int32_t chs_in[IAMFB_MAX_LAYOUT_CH];       /* IAChannel id of every decoded row (channels_order) */
  int32_t n_out_gain;                        /* output-gain list, demixer_set_output_gain demixer.c:580-590 */
  int32_t out_gain_ch[IAMFB_MAX_LAYOUT_CH];
  float out_gain[IAMFB_MAX_LAYOUT_CH];       /* linear (db2lin of the Q7.8 value) */
  int32_t has_demix_info;                    /* element carries a demixing parameter definition */
  int32_t default_mode, default_w_idx;       /* dmx_default_mode / dmx_default_w_idx */
  int32_t first_layer_layout;                /* conf_s[0].layout, used for the default recon-gain channel list
                                                (iamf_stream_scale_decoder_set_default_recon_gain :2202-2236) */
  int32_t selected_layer;                    /* ctx->layer (0 => no default recon list) */
  int32_t recon_present;                     /* the selected layer carries a recon-gain list (recon_gain_flag), i.e.
                                                demixer_set_recon_gain runs every frame (IAMF_decoder.c:2334-2343) */
  int32_t use_dmr;                           /* render with the parametric DMRenderer toward dmr_out_layout instead
                                                of the M2M matrix (iamf_stream_renderer_enable_downmix :2448-2478) */
  int32_t dmr_out_layout;
  /* -- scene based (ambisonics) -- */
  int32_t ambi_mode;                         /* 0 mono mapping, 1 projection (IAMF_OBU.h AMBISONICS_MODE_*) */
  int32_t ambi_channels;                     /* output_channel_count: 1, 4, 9 or 16 */
  uint8_t ambi_map[IAMFB_MAX_SCENE_CH];      /* mono: output channel i <- decoded row ambi_map[i] */
  int32_t ambi_cols;                         /* projection: substreams + coupled substreams */
  float ambi_matrix[IAMFB_MAX_SCENE_CH * IAMFB_MAX_SCENE_CH]; /* projection: [col][row], q_to_float(q15) */
  /* -- binaural target only -- */
  int32_t binaural_hrtf;                     /* 1: render this element with the HRTF renderer (256-tap HRIR pair per channel,
                                                iamfb_get_hrir) instead of the stereo rows of the matrix tables - what a
                                                reference built with DISABLE_BINAURALIZER == 0 does for channel-based
                                                elements with headphones_rendering_mode == 1 (IAMF_decoder.c:2565-2573,
                                                m2b_rdr.c:103-121) and for every scene-based element (:2606-2612,
                                                h2b_rdr.c:109-131).  0: the as-built behaviour (stereo matrices). */
} iamfb_element_desc;

typedef struct iamfb_plan_desc {
  int32_t frame_size;                        /* samples per frame of the codec config */
  int32_t in_rate, out_rate;                 /* stream rate, requested rate (resampler iff they differ) */
  int32_t n_elements;
  iamfb_element_desc el[IAMFB_MAX_ELEMENTS];
  int32_t target;                            /* IAMFB_TARGET_* */
  float loudness_gain;                       /* db2lin(normalization_loudness - ctx->loudness); 0 => stage off */
  int32_t limiter;                           /* IAMF_decoder_peak_limiter_enable */
  float limiter_threshold_db;                /* IAMF_decoder_peak_limiter_set_threshold (default -1 dBFS) */
  int32_t bit_depth;                         /* 16 | 24 | 32 ; 0 => float32 interleaved (test/debug output) */
  int32_t arithmetic;                        /* IAMFB_ARITH_EXACT (0, default): every expression in the reference's order and
                                                precision - the PCM is bit-identical to the reference decoder's.
                                                IAMFB_ARITH_FMA: the dense contraction of the path that is bound by the FP32
                                                pipe - the HOA-to-loudspeaker matrix (h2m_rdr.c:1088-1150) - fuses each
                                                multiply with its add (one rounding instead of two, same summation order):
                                                half the FP32 work; the PCM then agrees with the reference within +-1 LSB at
                                                16 bit, +-2 LSB at 24 bit (one 24-bit LSB is two float32 ulps near full
                                                scale) and 2e-7 of full scale as float (BASELINE's tolerance: 1e-5) instead
                                                of bit for bit.  Signatures without such a variant run the exact kernels
                                                (the resampler FIR was measured too: it is bound by shared-memory traffic,
                                                fusing gains nothing there). */
} iamfb_plan_desc;
enum { IAMFB_ARITH_EXACT = 0, IAMFB_ARITH_FMA = 1 };

/* Raw per-(stream,frame) parameters = what the parameter-block OBUs of one temporal unit resolve to before the
 * reference calls the stage functions (IAMF_decoder.c:2131-2151, 2324-2349, 3425-3469). */
typedef struct iamfb_frame_params {
  struct {
    int8_t dmx_mode;            /* demixing mode of this frame, -1 = none (ctx->dmx_mode <= INVALID_VALUE) */
    uint8_t has_recon;          /* a recon-gain list for the selected layer is present in this frame */
    uint16_t recon_flags;       /* recon_gain_flags of that layer */
    uint8_t recon_gain[12];     /* raw u8 gains, one per set flag bit in ascending bit order (gain = v/255) */
    float mix_gain;             /* element mix gain (linear) when no ramp array is supplied */
  } el[IAMFB_MAX_ELEMENTS];
  float out_gain;               /* output mix gain (linear) when no ramp array is supplied */
  uint16_t trim_start, trim_end;/* samples trimmed from this frame (OBU trimming, codec delay 0);
                                   trim_start == 0xFFFF: the stream has NO frame in this step (its state is untouched) */
} iamfb_frame_params;

/* Animated mix gain of one (stream, frame) as the parameter segments that cover the frame's samples after trimming
 * (iamf_database_parameter_get_mix_gain_unit, IAMF_decoder.c:857-982): the per-sample gains are evaluated ON THE DEVICE
 * with the reference's expressions (mix_gain_bezier_linear / _quad, :639-664) instead of travelling as N floats per frame.
 *   type 0 step:   g = start
 *   type 1 linear: g = start + (end - start) * i / interval                       (float)
 *   type 2 Bezier: alpha = interval - 2 ct;  a = alpha ? (sqrt(ct^2 + alpha i) - ct) / alpha : i / (2 ct)   (double / float)
 *                  g = (start + end - 2 control) a^2 + 2 a (control - start) + start                        (double)
 * for i = offset .. offset + count - 1 (position inside the segment).  n_segs == 0: the constant of iamfb_frame_params
 * applies to the frame. */
#define IAMFB_MAX_GAIN_SEGS 8
typedef struct iamfb_gain_seg {
  int32_t type, count, offset, interval, ct;
  float start, end, control;
} iamfb_gain_seg;
typedef struct iamfb_gain_ramp {
  int32_t n_segs, pad_[3];
  iamfb_gain_seg seg[IAMFB_MAX_GAIN_SEGS];
} iamfb_gain_ramp;

typedef struct iamfb_ctx iamfb_ctx;
typedef struct iamfb_plan iamfb_plan;
typedef struct iamfb_batch iamfb_batch;

/* Buffers of one submit.  All pointers are HOST pointers for iamfb_batch_submit_host and DEVICE pointers for
 * iamfb_batch_submit_device.  S = streams of the batch, F = n_frames of this call, N = frame_size.
 *   in[e]          float32 [S][F][n_in(e)][N]   decoded planar frames of element e (never modified); int16 of the same
 *                                               shape when in_format == IAMFB_IN_S16
 *   params         iamfb_frame_params [S][F]
 *   gain_ramp[e]   optional float32 [S][F][N]   per-sample element mix gains (animated mix gain), NULL = constants
 *   out_gain_ramp  optional float32 [S][F][N]
 *   gain_segs[e]   optional iamfb_gain_ramp [S][F]  the same as parameter segments, evaluated on the device (takes the
 *   out_gain_segs                                   place of the float arrays; k_gain_expand)
 *   pcm            bytes [S][out_stride_bytes]  interleaved PCM of each stream, out_counts tell how much is valid
 *   out_counts     int32 [S][F]                 samples per channel produced by each frame (what IAMF_decoder_decode
 *                                               returns for that temporal unit)
 */
enum { IAMFB_IN_F32 = 0, IAMFB_IN_S16 = 1 };

typedef struct iamfb_io {
  const float *in[IAMFB_MAX_ELEMENTS];
  const iamfb_frame_params *params;
  const float *gain_ramp[IAMFB_MAX_ELEMENTS];
  const float *out_gain_ramp;
  void *pcm;
  int32_t *out_counts;
  int32_t in_format;   /* IAMFB_IN_F32: `in` is float32 as above.  IAMFB_IN_S16: `in` points at int16 with the same
                          [S][F][n_in][N] shape - what Opus / AAC / 16-bit ipcm core decode produces BEFORE the codec glue
                          scales it by 1/32768 (opus/IAMF_opus_decoder.c:133-135); the scaling then happens on the
                          device (exact: a power of two), halving the host-to-device traffic */
  const iamfb_gain_ramp *gain_segs[IAMFB_MAX_ELEMENTS];
  const iamfb_gain_ramp *out_gain_segs;
} iamfb_io;

/* ---- context: one per GPU / host thread ---- */
int iamfb_device_count(void);   /* CUDA devices visible to the process (0 when there is none or no driver) */
int iamfb_ctx_create(int device, iamfb_ctx **ctx);
/* run on a caller-owned CUDA stream (cudaStream_t passed as void*), e.g. the framework's current stream */
int iamfb_ctx_set_stream(iamfb_ctx *ctx, void *cuda_stream);
int iamfb_ctx_synchronize(iamfb_ctx *ctx);
void iamfb_ctx_destroy(iamfb_ctx *ctx);
const char *iamfb_last_error(void);
const char *iamfb_version(void);

/* ---- plan: immutable pipeline signature + device-side constant tables ---- */
int iamfb_plan_create(iamfb_ctx *ctx, const iamfb_plan_desc *desc, iamfb_plan **plan);
void iamfb_plan_destroy(iamfb_plan *plan);
int iamfb_plan_out_channels(const iamfb_plan *plan);
/* diagnostic: which kernels serve this plan - IAMFB_PATH_MULTI (one kernel per stage), IAMFB_PATH_FUSED (k_fused: one
 * kernel per submit), IAMFB_PATH_STREAM (k_stream: pipelined per-stream kernel with one float32 stage, 16-bit channel-based
 * single-element signatures; trimmed / flushed streams of a submit still take k_fused), IAMFB_PATH_PIPE (k_pipe /
 * k_pipe_rs: the same pipeline for channel-based, scene-based, two-element and resampling signatures, any output depth,
 * two int16 stages or one float32 stage).  Where both exist float32 submits take k_stream and int16 submits k_pipe.
 * Results are bit-identical on every path. */
enum { IAMFB_PATH_MULTI = 0, IAMFB_PATH_FUSED = 1, IAMFB_PATH_STREAM = 2, IAMFB_PATH_PIPE = 3 };
int iamfb_plan_kernel_path(const iamfb_plan *plan);                      /* for float32 submits */
int iamfb_plan_kernel_path_fmt(const iamfb_plan *plan, int in_format);   /* IAMFB_IN_F32 | IAMFB_IN_S16 */
/* the arithmetic the plan's kernels really use: IAMFB_ARITH_FMA only when it was asked for AND the signature has such a
 * variant (third-order ambisonics -> sound system H); IAMFB_ARITH_EXACT otherwise */
int iamfb_plan_arithmetic(const iamfb_plan *plan);
/* upper bound of samples per channel one submit of n_frames can produce for one stream */
int iamfb_plan_max_out_samples(const iamfb_plan *plan, int n_frames);
/* bytes between consecutive streams in the pcm buffer for a submit of n_frames */
size_t iamfb_plan_out_stride_bytes(const iamfb_plan *plan, int n_frames);

/* ---- batch: n_streams independent streams sharing one plan; owns all per-stream device state ---- */
int iamfb_batch_create(iamfb_plan *plan, int n_streams, int max_frames_per_submit, iamfb_batch **batch);
void iamfb_batch_destroy(iamfb_batch *batch);
int iamfb_batch_reset(iamfb_batch *batch);  /* back to the state right after IAMF_decoder_configure */

/* device-resident: in / params / pcm / out_counts are device pointers; asynchronous on the context's stream */
int iamfb_batch_submit_device(iamfb_batch *batch, const iamfb_io *io, int n_frames);
/* host-resident: copies inputs H2D (pinned staging), runs, copies pcm + counts back, synchronises */
int iamfb_batch_submit_host(iamfb_batch *batch, const iamfb_io *io, int n_frames);
/* the same with the caller producing and consuming its host buffers group by group while the device works on the
 * neighbouring groups of streams: fill(user, s_lo, s_cnt, io) is called right before the inputs of the streams
 * [s_lo, s_lo + s_cnt) are uploaded - it writes their slices of in / params / ramps and may set or clear io->gain_ramp[] /
 * io->out_gain_ramp for this group -, drain(user, s_lo, s_cnt) once their pcm / out_counts slices are back in host memory.
 * (What the drop-in host layer uses to overlap bitstream parsing + core decode of one group of handles with the upload,
 * kernels and download of the others.) */
typedef struct iamfb_chunk_hooks {
  void (*fill)(void *user, int s_lo, int s_cnt, iamfb_io *io);
  void (*drain)(void *user, int s_lo, int s_cnt);
  void *user;
} iamfb_chunk_hooks;
int iamfb_batch_submit_host_hooks(iamfb_batch *batch, const iamfb_io *io, int n_frames, const iamfb_chunk_hooks *hooks);
/* end of stream (IAMF_decoder_decode(data == NULL)): resampler tail + limiter tail. pcm [S][stride for 1 frame],
 * out_counts [S]. */
int iamfb_batch_flush_device(iamfb_batch *batch, void *pcm, int32_t *out_counts);
int iamfb_batch_flush_host(iamfb_batch *batch, void *pcm, int32_t *out_counts);

/* pinned host memory for submit_host callers (plain malloc'ed memory works too, only slower) */
void *iamfb_host_alloc(size_t bytes);
void iamfb_host_free(void *p);

/* number of kernel launches issued by this library since the context was created (bench.py's gpu_launches) */
uint64_t iamfb_ctx_launch_count(const iamfb_ctx *ctx);

/* optional per-kernel timing with CUDA events recorded on the context's stream (used by bench.py for the roofline
 * figure; off by default).  iamfb_ctx_get_timing iterates index = 0,1,... until it returns non-zero. */
int iamfb_ctx_set_timing(iamfb_ctx *ctx, int enable);
int iamfb_ctx_get_timing(iamfb_ctx *ctx, int index, const char **name, double *total_ms, uint64_t *launches);
int iamfb_ctx_get_timing_median(iamfb_ctx *ctx, int index, double *median_ms);   /* median duration of one launch */

/* self-test (used by tests/): compares the branch-free division of k_stream's limiter scan with the IEEE division
 * `thr / w` for EVERY float w of the range the kernel uses it in (2^-60 .. 2^60); *mismatches receives the count of
 * differing bit patterns (0 expected).  thr must lie in 2^-20 .. 2^20 (outside it the kernel uses the plain division). */
int iamfb_selftest_quotient(iamfb_ctx *ctx, float thr, uint64_t *mismatches);

/* ---- table accessors (host side, no device needed) ---- */
int iamfb_target_channels(int target);                      /* IAMF_layout_sound_system_channels_count */
int iamfb_layout_channels(int layout, int32_t *chs);        /* rendering order, IAMF_utils.c:117-133 */
/* copies the [m][n] channel->channel matrix for (layout -> target); returns 0 or IAMFB_ERR_BAD_ARG */
int iamfb_get_m2m_matrix(int layout, int target, int32_t *m, int32_t *n, float *mat);
/* copies the [n][m] HOA->channel matrix for (order -> target) and its LFE slots */
int iamfb_get_h2m_matrix(int order, int target, int32_t *m, int32_t *n, int32_t *lfe1, int32_t *lfe2, float *mat);

/* the in-repo HRIR set of the binaural renderer (synthetic, tools/gen_hrir.py): copies the [2 ears][iamfb_hrir_taps()]
 * Q15 taps of IAChannel `index` (kind == IAMFB_EL_CHANNEL; the role BEAR's default.tf plays for m2b_rdr.c:56) or of
 * ambisonics channel `index` in ACN order (kind == IAMFB_EL_SCENE; h2b_rdr.c:60) */
int iamfb_get_hrir(int kind, int index, int16_t *taps);
int iamfb_hrir_taps(void);

#ifdef __cplusplus
}
#endif
#endif /* IAMF_B200_H_ */
