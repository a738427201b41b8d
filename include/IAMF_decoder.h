/*
 * IAMF_decoder.h - public C API of the drop-in libiamf.so built by this repository.
 *
 * ABI-compatible with the reference decoder's public headers (Samsung/iac include/IAMF_decoder.h:60-239 and
 * include/IAMF_defines.h:62-209): same exported symbol names, argument types, enum values, struct layouts and return
 * conventions, so a program compiled against the reference headers (e.g. the stock test/tools/iamfplayer) runs
 * against this library unmodified.  Declarations are restated here from that contract (this file is not a copy of
 * the reference headers); one additive, batch-oriented entry point is declared at the end.
 *
 * Call protocol (IAMF_decoder.c:3726-3942 of the reference):
 *   open -> setters -> configure(descriptor OBUs [+ following data], &consumed) until IAMF_OK -> decode(...) per
 *   temporal unit (returns samples per channel written, 0 = need more data, <0 = error) -> decode(NULL) to flush the
 *   limiter / resampler delay -> close.
 */
#ifndef IAMF_DECODER_H
#define IAMF_DECODER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- IAMF_defines.h ------------------------------------------------------------------------------------------- */
typedef enum { AUDIO_ELEMENT_INVALID = -1, AUDIO_ELEMENT_CHANNEL_BASED, AUDIO_ELEMENT_SCENE_BASED, AUDIO_ELEMENT_COUNT } AudioElementType;
typedef enum AmbisonicsMode { AMBISONICS_MONO, AMBISONICS_PROJECTION } AmbisonicsMode;
typedef enum IAMF_LayoutType {
  IAMF_LAYOUT_TYPE_NOT_DEFINED = 0,
  IAMF_LAYOUT_TYPE_LOUDSPEAKERS_SS_CONVENTION = 2,
  IAMF_LAYOUT_TYPE_BINAURAL
} IAMF_LayoutType;

/* ITU-R BS.2051 sound systems A..J plus the three IAMF extensions */
typedef enum IAMF_SoundSystem {
  SOUND_SYSTEM_INVALID = -1,
  SOUND_SYSTEM_A,       /* 0+2+0 */
  SOUND_SYSTEM_B,       /* 0+5+0 */
  SOUND_SYSTEM_C,       /* 2+5+0 */
  SOUND_SYSTEM_D,       /* 4+5+0 */
  SOUND_SYSTEM_E,       /* 4+5+1 */
  SOUND_SYSTEM_F,       /* 3+7+0 */
  SOUND_SYSTEM_G,       /* 4+9+0 */
  SOUND_SYSTEM_H,       /* 9+10+3 */
  SOUND_SYSTEM_I,       /* 0+7+0 */
  SOUND_SYSTEM_J,       /* 4+7+0 */
  SOUND_SYSTEM_EXT_712, /* 2+7+0 */
  SOUND_SYSTEM_EXT_312, /* 2+3+0 */
  SOUND_SYSTEM_MONO,    /* 0+1+0 */
  SOUND_SYSTEM_END
} IAMF_SoundSystem;

typedef enum IAMF_ParameterType {
  IAMF_PARAMETER_TYPE_MIX_GAIN = 0,
  IAMF_PARAMETER_TYPE_DEMIXING,
  IAMF_PARAMETER_TYPE_RECON_GAIN
} IAMF_ParameterType;

typedef enum IAMF_AnimationType { ANIMATION_TYPE_INVALID = -1, ANIMATION_TYPE_STEP, ANIMATION_TYPE_LINEAR, ANIMATION_TYPE_BEZIER } IAMF_AnimationType;

/* one byte on the wire: layout_type(2) | sound_system(4) | reserved(2) */
typedef struct IAMF_Layout {
  union {
    struct { uint8_t reserved : 2; uint8_t sound_system : 4; uint8_t type : 2; } sound_system;
    struct { uint8_t reserved : 6; uint8_t type : 2; } binaural;
    struct { uint8_t reserved : 6; uint8_t type : 2; };
  };
} IAMF_Layout;

typedef struct _anchor_loudness_t { uint8_t anchor_element; int16_t anchored_loudness; } anchor_loudness_t;

typedef struct IAMF_LoudnessInfo {
  uint8_t info_type;
  int16_t integrated_loudness; /* Q7.8 LKFS */
  int16_t digital_peak;
  int16_t true_peak;
  uint8_t num_anchor_loudness;
  anchor_loudness_t *anchor_loudness;
} IAMF_LoudnessInfo;

typedef enum { IAMF_CODEC_UNKNOWN = 0, IAMF_CODEC_OPUS, IAMF_CODEC_AAC, IAMF_CODEC_FLAC, IAMF_CODEC_PCM, IAMF_CODEC_COUNT } IAMF_CodecID;

enum {
  IAMF_OK = 0,
  IAMF_ERR_BAD_ARG = -1,
  IAMF_ERR_BUFFER_TOO_SMALL = -2,
  IAMF_ERR_INTERNAL = -3,
  IAMF_ERR_INVALID_PACKET = -4,
  IAMF_ERR_INVALID_STATE = -5,
  IAMF_ERR_UNIMPLEMENTED = -6,
  IAMF_ERR_ALLOC_FAIL = -7
};

typedef enum {
  IA_CHANNEL_LAYOUT_INVALID = -1,
  IA_CHANNEL_LAYOUT_MONO = 0, IA_CHANNEL_LAYOUT_STEREO, IA_CHANNEL_LAYOUT_510, IA_CHANNEL_LAYOUT_512,
  IA_CHANNEL_LAYOUT_514, IA_CHANNEL_LAYOUT_710, IA_CHANNEL_LAYOUT_712, IA_CHANNEL_LAYOUT_714,
  IA_CHANNEL_LAYOUT_312, IA_CHANNEL_LAYOUT_BINAURAL, IA_CHANNEL_LAYOUT_COUNT
} IAChannelLayoutType;

/* ---- IAMF_decoder.h ------------------------------------------------------------------------------------------- */
typedef struct IAMF_StreamInfo { uint32_t max_frame_size; } IAMF_StreamInfo;
typedef struct IAMF_Decoder *IAMF_DecoderHandle;

IAMF_DecoderHandle IAMF_decoder_open(void);
int IAMF_decoder_close(IAMF_DecoderHandle handle);
int IAMF_decoder_configure(IAMF_DecoderHandle handle, const uint8_t *data, uint32_t size, uint32_t *rsize);
int IAMF_decoder_decode(IAMF_DecoderHandle handle, const uint8_t *data, int32_t size, uint32_t *rsize, void *pcm);
int IAMF_decoder_set_mix_presentation_id(IAMF_DecoderHandle handle, uint64_t id);
int IAMF_decoder_output_layout_set_sound_system(IAMF_DecoderHandle handle, IAMF_SoundSystem ss);
int IAMF_decoder_output_layout_set_binaural(IAMF_DecoderHandle handle);
int IAMF_layout_sound_system_channels_count(IAMF_SoundSystem ss);
int IAMF_layout_binaural_channels_count(void);
char *IAMF_decoder_get_codec_capability(void);                 /* malloc'ed, caller frees */
int IAMF_decoder_set_normalization_loudness(IAMF_DecoderHandle handle, float loudness);
int IAMF_decoder_set_bit_depth(IAMF_DecoderHandle handle, uint32_t bit_depth);
int IAMF_decoder_peak_limiter_enable(IAMF_DecoderHandle handle, uint32_t enable);
int IAMF_decoder_peak_limiter_set_threshold(IAMF_DecoderHandle handle, float db);
float IAMF_decoder_peak_limiter_get_threshold(IAMF_DecoderHandle handle);
int IAMF_decoder_set_sampling_rate(IAMF_DecoderHandle handle, uint32_t rate);
IAMF_StreamInfo *IAMF_decoder_get_stream_info(IAMF_DecoderHandle handle);

typedef struct IAMF_Param {
  int parameter_length;
  uint32_t parameter_definition_type;
  union { uint32_t dmixp_mode; };
} IAMF_Param;

typedef enum IAMF_SoundMode {
  IAMF_SOUND_MODE_NONE = -2, IAMF_SOUND_MODE_NA = -1, IAMF_SOUND_MODE_STEREO, IAMF_SOUND_MODE_MULTICHANNEL, IAMF_SOUND_MODE_BINAURAL
} IAMF_SoundMode;

typedef struct IAMF_extradata {
  IAMF_SoundSystem output_sound_system;
  uint32_t number_of_samples;
  uint32_t bitdepth;
  uint32_t sampling_rate;
  IAMF_SoundMode output_sound_mode;
  int num_loudness_layouts;
  IAMF_Layout *loudness_layout;     /* calloc'ed, caller frees */
  IAMF_LoudnessInfo *loudness;      /* calloc'ed, caller frees */
  uint32_t num_parameters;
  IAMF_Param *param;                /* calloc'ed, caller frees */
} IAMF_extradata;

int IAMF_decoder_set_pts(IAMF_DecoderHandle handle, int64_t pts, uint32_t time_base);
int IAMF_decoder_get_last_metadata(IAMF_DecoderHandle handle, int64_t *pts, IAMF_extradata *metadata);

/* ---- additive extension (not in the reference): step many handles in one device launch ------------------------- */
/* All handles must have been configured to the same pipeline signature (same descriptors and output settings).
 * data[i]/size[i] is the temporal unit of handle i (NULL = flush); pcm[i] receives its samples; ret[i] what the
 * per-handle call would have returned.  Returns IAMF_OK or the first hard error. */
int IAMF_decoder_decode_batch(IAMF_DecoderHandle *handles, int n, const uint8_t *const *data, const int32_t *size,
                              uint32_t *rsize, void *const *pcm, int *ret);

/* The same for up to max_units (1..64) temporal units per handle and call: data[i]/size[i] holds whole temporal units of
 * handle i back to back (what is left after max_units, or an incomplete last unit, is not consumed: rsize[i] tells);
 * pcm[i] receives the units' samples back to back; ret[i] = samples per channel written for handle i (or the error of
 * its first unit when nothing was written); units_done[i] (optional) = temporal units consumed.  The host part of every
 * handle (parsing, core decode) runs on a pool of host threads (IAMF_B200_HOST_THREADS, default: the online cores), the
 * device part once for the whole group.  The first call fixes the group and max_units. */
int IAMF_decoder_decode_batch_units(IAMF_DecoderHandle *handles, int n, const uint8_t *const *data, const int32_t *size,
                                    uint32_t *rsize, void *const *pcm, int *ret, int max_units, int *units_done);

#ifdef __cplusplus
}
#endif
#endif /* IAMF_DECODER_H */
