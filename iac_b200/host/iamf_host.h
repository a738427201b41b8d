/*
 * iamf_host.h - internal types of the drop-in libiamf.so host layer (plain C).
 *
 * The host layer does what stays on the CPU in the reference decoder: OBU parsing, the descriptor database, parameter
 * time lines, stream / layer selection and core (codec) decode.  Everything after core decode is handed to the CUDA
 * engine through include/iamf_b200.h.  Reference behaviour is cited as file:line of the Samsung/iac tree.
 */
#ifndef IAMF_HOST_H_
#define IAMF_HOST_H_

#include <stddef.h>
#include <stdint.h>

#include "IAMF_decoder.h"
#include "iamf_b200.h"

#define IH_MAX_CODECS 8
#define IH_MAX_ELEMENTS 8
#define IH_MAX_MIXES 8
#define IH_MAX_PARAMS 32
#define IH_MAX_LAYERS 6
#define IH_MAX_SUBSTREAMS 32
#define IH_MAX_SEGMENTS 4096 /* parameter-block segments accepted per block (a stream-supplied count) */
#define IH_INVALID_ID ((uint64_t)-1)

/* ---- byte / bit reader (MSB first; byte-aligned reads realign first, like bitstream.c:39-160) ---- */
typedef struct {
  const uint8_t *p;
  uint32_t size, pos;
  int bit; /* bits already consumed of p[pos] */
} ih_reader;

void ih_rd_init(ih_reader *r, const uint8_t *p, uint32_t size);
uint32_t ih_rd_bits(ih_reader *r, int n);
void ih_rd_skip_bits(ih_reader *r, int n);
void ih_rd_align(ih_reader *r);
uint32_t ih_rd_u8(ih_reader *r);
uint32_t ih_rd_u16(ih_reader *r);
uint64_t ih_rd_leb128(ih_reader *r);
void ih_rd_bytes(ih_reader *r, uint8_t *dst, uint32_t n);
void ih_rd_cstring(ih_reader *r);
uint32_t ih_rd_tell(const ih_reader *r);

/* ---- OBU ---- */
enum { IH_OBU_CODEC_CONFIG = 0, IH_OBU_AUDIO_ELEMENT, IH_OBU_MIX_PRESENTATION, IH_OBU_PARAMETER_BLOCK,
       IH_OBU_TEMPORAL_DELIMITER, IH_OBU_AUDIO_FRAME, IH_OBU_AUDIO_FRAME_ID0, IH_OBU_AUDIO_FRAME_ID17 = 23,
       IH_OBU_SEQUENCE_HEADER = 31 };

typedef struct {
  int type, redundant;
  uint64_t trim_start, trim_end;
  const uint8_t *payload;
  uint32_t payload_size;
  uint32_t total_size;
} ih_obu;

uint32_t ih_obu_split(const uint8_t *data, uint32_t size, ih_obu *o); /* IAMF_OBU.c:79-138; 0 = need more data */

/* ---- descriptors ---- */
typedef struct {
  int type; /* IAMF_ParameterType */
  uint64_t id, rate;
  int mode;
  uint64_t duration, const_interval;
  int n_segments;
  uint64_t seg_interval[16];
} ih_param_def;

typedef struct {
  uint64_t id;
  int codec; /* IAMF_CodecID */
  uint64_t frame_size;
  uint8_t conf[64];
  int conf_size;
  int rate;
} ih_codec;

typedef struct {
  int layout, out_gain_present, recon_present, n_sub, n_coupled;
  int out_gain_flags;
  int16_t out_gain_q;
} ih_layer;

typedef struct {
  uint64_t id;
  int type; /* 0 channel based, 1 scene based */
  uint64_t codec_id;
  int n_sub;
  uint64_t sub_ids[IH_MAX_SUBSTREAMS];
  int n_params;
  ih_param_def params[4];
  int has_demix, dmx_mode, dmx_w;
  int n_layers;
  ih_layer layers[IH_MAX_LAYERS];
  int ambi_mode, ambi_channels, ambi_sub, ambi_coupled;
  uint8_t ambi_map[512];
  int ambi_map_size;
} ih_element;

typedef struct {
  uint64_t id;
  int n_elements;
  struct {
    uint64_t element_id;
    int headphones_mode;
    ih_param_def gain_def;
    int16_t gain_q;
  } el[2];
  ih_param_def out_def;
  int16_t out_q;
  int n_layouts;
  struct {
    int type, sound_system;
    IAMF_LoudnessInfo loud;
  } layouts[16];
} ih_mix;

/* ---- parameter time lines (IAMF_decoder.c:760-1126) ---- */
typedef struct ih_segment {
  struct ih_segment *next;
  uint64_t interval;
  int anim;                      /* mix gain */
  float g_start, g_end, g_control, g_ctime;
  int dmx_mode;                  /* demixing */
  int n_layers;                  /* recon gain */
  struct { uint32_t flags; int n; uint8_t q[12]; } rg[IH_MAX_LAYERS];
} ih_segment;

typedef struct {
  uint64_t id, parent;
  int type, rate;
  const ih_param_def *def;
  uint64_t timestamp, duration, elapse;
  ih_segment *head, *tail;
  int use_default;
  float default_gain;
} ih_param_item;

/* ---- one enabled audio element (IAMF_Stream + IAMF_StreamDecoder) ---- */
typedef struct {
  const ih_element *el;
  const ih_codec *cc;
  int n_channels, n_coupled;
  int layer, layout, n_layout_ch;
  int chs_order[IAMFB_MAX_LAYOUT_CH];
  int n_decoded;                 /* channels decoded (layers 0..layer) */
  int n_sub_used;                /* substreams feeding them */
  uint64_t timestamp;
  uint64_t trimming_start, trimming_end;
  int dmx_mode;
  uint8_t *pkt[IH_MAX_SUBSTREAMS];
  uint32_t pkt_size[IH_MAX_SUBSTREAMS];
  int pkt_count;
  uint32_t pkt_borrowed;         /* bit k: pkt[k] points into the caller's buffer of the running batch step (not owned) */
  void *codec_state[IH_MAX_SUBSTREAMS]; /* per sub-stream core decoder (Opus), created on first use */
  uint64_t strim, etrim;
  ih_param_item *mix_gain, *demix, *recon;
  /* recon update delivered with the next frame */
  int has_recon;
  uint32_t recon_flags;
  uint8_t recon_q[12];
} ih_stream;

enum { IH_STATUS_INIT = 0, IH_STATUS_CONFIGURE, IH_STATUS_RECEIVE, IH_STATUS_RUN, IH_STATUS_RECONFIGURE };
enum { IH_FLAG_MAGIC = 1, IH_FLAG_CODEC = 2, IH_FLAG_ELEMENT = 4, IH_FLAG_MIX = 8, IH_FLAG_CONFIG = 16,
       IH_FLAG_DESCRIPTORS = 15 };
enum { IH_NEED_MIX = 1, IH_NEED_LAYOUT = 2, IH_NEED_PRESENTATION = 4 };

struct IAMF_Decoder {
  /* settings */
  float threshold_db, loudness, norm_loudness;
  uint32_t sampling_rate, bit_depth;
  int limiter_on;
  int layout_type, sound_system;      /* requested output */
  int have_layout;
  uint64_t mix_id;
  int need_configure, status;
  unsigned flags;
  int64_t pts;
  uint32_t pts_time_base;
  uint64_t duration;
  int last_frame_size;
  IAMF_StreamInfo info;
  /* database */
  int have_header;
  int n_codecs, n_elements, n_mixes, n_params;
  ih_codec codecs[IH_MAX_CODECS];
  ih_element elements[IH_MAX_ELEMENTS];
  ih_mix mixes[IH_MAX_MIXES];
  ih_param_item params[IH_MAX_PARAMS];
  /* presentation */
  const ih_mix *mix;
  int n_streams;
  ih_stream streams[IAMFB_MAX_ELEMENTS];
  ih_param_item *out_gain_item;
  int out_channels;
  IAMF_extradata metadata;
  /* engine */
  void *shared;                        /* the (context, plan) entry this handle borrows (iamf_decoder.c) */
  iamfb_ctx *ctx;
  iamfb_plan *plan;
  iamfb_batch *batch;
  iamfb_plan_desc desc;
  float *in[IAMFB_MAX_ELEMENTS];      /* pinned: decoded planar frame of each element */
  iamfb_gain_ramp *ramp[IAMFB_MAX_ELEMENTS], *out_ramp;   /* pinned: animated mix gains as parameter segments, per frame slot */
  uint8_t *pcm_stage;
  size_t pcm_stage_size;               /* bytes per stream */
  iamfb_frame_params *fp_stage;
  int32_t *counts_stage;
  int frame_size;
  /* batch extension: handles stepping together share the engine of the group leader */
  struct IAMF_Decoder *leader;
  int group_owner, group_size, group_index;
  int group_units;                     /* leader: temporal units per handle and call the group's buffers are sized for */
  int group_s16;                       /* leader: the group's decoded frames travel as int16 (IAMFB_IN_S16) */
  int unit_ret[64];                    /* decode_batch_units scratch: per unit of this handle, what prepare_frame returned */
  uint8_t unit_flags[64];
  uint32_t unit_used;                  /* bytes consumed by this call */
  int units_done;
};

/* iamf_obu_parse.c */
int ih_parse_codec(const ih_obu *o, ih_codec *c);
int ih_parse_element(const ih_obu *o, ih_element *e);
int ih_parse_mix(const ih_obu *o, ih_mix *m);
ih_segment *ih_parse_parameter_block(const ih_obu *o, uint64_t *pid_out, const ih_param_def *def, int n_layers,
                                     unsigned recon_present_flags, int *n_segments);
uint64_t ih_obu_parameter_id(const ih_obu *o);

/* iamf_timeline.c */
float ih_q_to_float(int16_t q, int frac);
float ih_qf_to_float(uint8_t q);
float ih_db2lin(float db);
int64_t ih_time_transform(int64_t t, int s1, int s2);
ih_param_item *ih_param_find(struct IAMF_Decoder *d, uint64_t pid);
ih_param_item *ih_param_add(struct IAMF_Decoder *d, const ih_param_def *def, uint64_t parent, int rate);
void ih_param_push(ih_param_item *pi, ih_segment *segs);
void ih_param_clear(ih_param_item *pi);
const ih_segment *ih_param_segment_at(const ih_param_item *pi, uint64_t pts);
void ih_params_elapse(struct IAMF_Decoder *d, uint64_t duration, uint32_t rate);
/* returns 0 = no unit, 1 = constant in *gain, 2 = animated: the covering parameter segments in *segs (evaluated on the
 * device), -1 = more segments than iamfb_gain_ramp holds */
int ih_mix_gain_unit(const ih_param_item *pi, uint64_t pt, int duration, int rate, float *gain, iamfb_gain_ramp *segs);

/* iamf_codec.c */
int ih_codec_supported(int codec);
/* decodes the packets of `n_sub` substreams (the first n_coupled are stereo) into planar float; returns samples */
int ih_codec_decode(ih_stream *st, int first_sub, const ih_codec *cc, uint8_t *const *pkt, const uint32_t *pkt_size,
                    int n_sub, int n_coupled, float *out, int frame_size);
int ih_flac_rate(const uint8_t *conf, int size);
int ih_flac_bits(const uint8_t *conf, int size);
int ih_codec_is_s16(const ih_codec *cc);
int ih_codec_decode_s16(ih_stream *st, int first_sub, const ih_codec *cc, uint8_t *const *pkt, const uint32_t *pkt_size,
                        int n_sub, int n_coupled, int16_t *out, int frame_size);
void ih_codec_close(ih_stream *st);

#endif
