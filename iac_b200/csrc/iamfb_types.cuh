// iamfb_types.cuh - device-visible plan / state / per-frame records of the sm_100a rendering engine.
#pragma once
#include <cstdint>

#include "iamf_b200.h"

namespace iamfb {

constexpr int kMaxEl = IAMFB_MAX_ELEMENTS;
constexpr int kMaxOut = IAMFB_MAX_OUT_CH;
constexpr int kMaxRec = 16;            // reconstructed channels feeding a render matrix (12 layout / 16 HOA)
constexpr int kChCount = IAMFB_CH_COUNT;
constexpr int kLimDelay = IAMFB_LIMITER_DELAY;
constexpr int kMaxSub = 8;            // sub-chunks a submit is split into so that the limiter scan overlaps the rest
constexpr int kMaxRsHist = 256;       // upper bound of the resampler history (filt_len - 1 rounded up to 4) -> ratios down to 4:1

enum Renderer : int { kRdrM2M = 0, kRdrH2M = 1, kRdrDMR = 2 };

// Per-element constants.  Passed to kernels by value inside KernelPlan (constant bank).
struct ElPlan {
  int kind;                 // IAMFB_EL_*
  int n_in;                 // decoded rows
  int layout;               // channel based: reconstructed layout
  int n_rec;                // channels entering the renderer
  int renderer;             // Renderer
  int recon_present;        // demixer_set_recon_gain runs every frame
  signed char src_row[kChCount];   // IAChannel id -> decoded row or -1
  unsigned int gain_mask;          // IAChannel ids with an output gain
  float gain[kChCount];
  unsigned char need_s2, need_s3, need_s5, need_s7, need_h2, need_h4;   // derivation chain (demixer.c:127-378)
  unsigned char rec_ch[kMaxRec];   // layout order: slot m -> IAChannel id
  signed char slot_of[kChCount];   // IAChannel id -> layout slot or -1
  signed char recon_bit[kChCount]; // IAChannel id -> recon-gain flag bit of this layout or -1 (IAMF_decoder.c:409-448)
  // render matrix, output-major: out n = sum_m mat[n*n_rec + m] * x[m]
  int n_mat_out;
  signed char out_slot[kMaxOut];   // matrix output row -> output channel (H2M LFE slot shift, h2m_rdr.c:1114-1135)
  float mat[kMaxOut * kMaxRec];
  // fused kernel (resolved once the tile size is known): byte offset of every IAChannel's staged row inside the
  // block's input tile (absent channels point at an all-zero row), output gain per channel with 1.0 as the default
  // (multiplying by 1.0f is exact); the render matrix by OUTPUT CHANNEL (compressed rows, zeros dropped, inputs
  // ascending): entry q of row oc adds f_csr_val[q] * (staged/reconstructed row at byte offset f_csr_off[q])
  int f_src_off[kChCount];
  float f_gain[kChCount];
  int f_n_gain;                    // transmitted channels with an output gain: staged-row byte offset and gain
  int f_gain_off[IAMFB_MAX_LAYOUT_CH];
  float f_gain_val[IAMFB_MAX_LAYOUT_CH];
  int f_row_off;                   // byte offset of the element's first staged row
  // k_stream: byte offset of every IAChannel's row inside the staged tile ([n_in][240] floats), -1 when not transmitted
  int s_row_off[kChCount];
  unsigned short f_csr_ptr[kMaxOut + 1];
  int f_csr_off[kMaxOut * kMaxRec];
  float f_csr_val[kMaxOut * kMaxRec];
  // DMR (downmix_renderer.c)
  int dmr_n_out;
  unsigned char dmr_out_ch[IAMFB_MAX_LAYOUT_CH];
  unsigned int dmr_in_mask;        // IAChannel ids that are inputs of the down-mixer
  // scene
  int ambi_mode, ambi_cols;
  unsigned char ambi_map[IAMFB_MAX_SCENE_CH];
  float ambi_mat[IAMFB_MAX_SCENE_CH * IAMFB_MAX_SCENE_CH];   // [col][row]
};

struct KernelPlan {
  int frame_size, n_elements, out_channels;
  int overlap;              // recon-gain cross-fade length = frame_size/16 (demixer.c:539-540)
  int resample;             // 0/1
  int limiter;              // 0/1
  int hist;                 // history samples in front of the mixed time line (kLimDelay, or 0 without limiter)
  int bit_depth;
  float loud_gain;          // 0 => off, (1 => identity, skipped like IAMF_decoder.c:3211)
  float lim_thr;
  int lim_ja, lim_jr;       // first time index with T>=attack, T>=attack+release
  // resampler
  unsigned int rs_num, rs_den, rs_filt_len, rs_oversample;
  int rs_int_adv, rs_frac_adv, rs_direct;
  int rs_hist;              // input history kept in front of the pre-resample time line (>= filt_len - 1)
  ElPlan el[kMaxEl];
};

// Persistent per-stream state (device).  One struct per stream; small, read/written by the resolve + scan kernels.
struct ElState {
  int mode, w_idx;                 // demixer: demixing_mode, weight_state_idx
  int dmr_mode, dmr_w_idx;         // DMRenderer: mode, w_idx
  float dmr_tl;                    // DMRenderer deps[TL][1].s  (gamma * w)
  // recon gain, kept PER IAChannel (a list position of the reference maps to a channel through the flag bit order,
  // IAMF_decoder.c:409-448): the de-mixer's list (demixer.c chs_recon_gain_list) and the latest list received for the
  // selected layer (ChannelLayerContext conf_s[layer].recon_gain)
  unsigned int rflags;             // flags of the de-mixer's list
  unsigned int re_flags;           // flags of the latest received list
  float rgain[kChCount];           // de-mixer: recon gain of the channel
  float re_gain[kChCount];         // latest received gain of the channel
  float sfavg[kChCount];           // ch_last_sfavg
};

struct StreamState {
  ElState el[kMaxEl];
  // resampler (closed form)
  long long rs_in_total;           // input samples supplied so far
  long long rs_out_total;          // outputs emitted so far
  // limiter
  int lim_j;                       // integer time index since the last trigger, -1 = idle
  float lim_start, lim_end;
  int lim_pad;                     // priming samples still to drop (padsize)
  int lim_init;
};

// Resolved per-(stream, frame) record written by the resolve kernel.
struct ElFrame {
  int mode;
  float w;                         // get_w(weight_state_idx)
  float dmr_alpha, dmr_beta, dmr_gamma, dmr_delta, dmr_tl;
  unsigned int rmask;              // layout slots with a recon gain this frame
  float rlast[IAMFB_MAX_LAYOUT_CH], rcur[IAMFB_MAX_LAYOUT_CH];   // by layout slot; 1.0 in the slots outside rmask
  float gain;                      // element mix gain constant
  int pad_[3];                     // rlast / rcur start on 16-byte boundaries in every element of a FrameRec
};

struct __align__(16) FrameRec {
  ElFrame el[kMaxEl];
  float out_gain;
  int vstart, vlen;                // valid samples of this frame after trimming
  int t_off;                       // offset of the first valid sample on the stream's time line of this submit
};

// Per-stream per-submit scalars written by the resolve kernel.
struct SubmitRec {
  int in_len;                      // samples on the pre-resample / mixed time line this submit
  int lim_len;                     // samples entering the limiter stage this submit
  int out_len;                     // samples written to pcm
  int out_skip;                    // limiter priming samples dropped this submit
  int irregular;                   // some frame of this submit is trimmed / missing (or this is a flush): the stream
                                   // takes the sequential fused kernel instead of the pipelined one
  long long rs_out_first;          // absolute index of the first resampler output this submit
  int sub_off[kMaxSub + 1];        // limiter-stage sample offsets of the sub-chunk boundaries (sub_off[n_sub] == lim_len)
};

}  // namespace iamfb
