// iamfb_kernels.cuh - hand-written sm_100a kernels of the post-decode rendering path.
//
// Arithmetic contract: every expression below is evaluated in the SAME order and precision as the reference C code
// it replaces (strict IEEE-754 binary32, one double-precision expression in the S2->S3 de-mixer), so this TU is
// compiled with --fmad=false and uses true divisions.  Results are bit-identical to the reference, not merely
// within tolerance; the reference location is cited beside each block.
//
// Data layout in HBM (S streams, F frames per submit, N samples per frame):
//   in[e]      f32 [S][F][C_in][N]        decoded planar frames (as core decode hands them over)
//   tl_a       f32 [S][C_out][cap_a]      pre-resample time line (rs_hist history + F*N), only when resampling
//   tl_b       f32 [S][C_out][cap_b]      mixed time line entering the limiter: [hist | samples of this submit]
//   pk         f32 [S][cap_b]             per-instant cross-channel peak max_c |x|, same time axis as tl_b
//   gn         f32 [S][cap_b]             limiter gain per instant
//   pcm        int16/24/32 [S][stride]    interleaved output
#pragma once
#include <cuda_runtime.h>

#include "iamfb_types.cuh"

namespace iamfb {

// -------------------------------------------------------------------------------------------------------------------
// small device helpers
// -------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
  // streaming read: decoded PCM is touched exactly once -> do not allocate it in L1
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream1(const float *p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// demixing_type_mat (demixer.c:62-72): mode -> alpha, beta, gamma, delta, w step
static __constant__ float c_mix_alpha[8] = {1.0f, 0.707f, 1.0f, 0.f, 1.0f, 0.707f, 1.0f, 0.f};
static __constant__ float c_mix_beta[8] = {1.0f, 0.707f, 0.866f, 0.f, 1.0f, 0.707f, 0.866f, 0.f};
static __constant__ float c_mix_gamma[8] = {0.707f, 0.707f, 0.866f, 0.f, 0.707f, 0.707f, 0.866f, 0.f};
static __constant__ float c_mix_delta[8] = {0.707f, 0.707f, 0.866f, 0.f, 0.707f, 0.707f, 0.866f, 0.f};
static __constant__ int c_mix_woff[8] = {-1, -1, -1, 0, 1, 1, 1, 0};
// widx2w_table (fixedp11_5.c:81-82)
static __constant__ float c_w_table[11] = {0.0f, 0.0179f, 0.0391f, 0.0658f, 0.1038f, 0.25f, 0.3962f, 0.4342f, 0.4609f, 0.4821f, 0.5f};
// recon channel bit -> IAChannel per layout (IAMF_decoder.c:409-448)
static __constant__ unsigned char c_recon_map[9][12] = {
    {13, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},        {14, 0, 15, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {1, 3, 2, 20, 21, 0, 0, 0, 0, 0, 0, 4},       {1, 3, 2, 20, 21, 22, 23, 0, 0, 0, 0, 4},
    {1, 3, 2, 20, 21, 9, 10, 0, 0, 11, 12, 4},    {1, 3, 2, 5, 6, 0, 0, 7, 8, 0, 0, 4},
    {1, 3, 2, 5, 6, 22, 23, 7, 8, 0, 0, 4},       {1, 3, 2, 5, 6, 9, 10, 7, 8, 11, 12, 4},
    {18, 3, 19, 0, 0, 16, 17, 0, 0, 0, 0, 4}};

// -------------------------------------------------------------------------------------------------------------------
// K0: parameter resolve - one thread per stream walks the F frames of the submit and turns raw parameter-block
// values into the per-frame scalars the sample kernels need, carrying the metadata state machines of
//   demixer_set_demixing_info / calc_w            demixer.c:592-619, fixedp11_5.c:83-90
//   iamf_stream_scale_decoder_update_recon_gain   IAMF_decoder.c:2238-2274
//   demixer_set_recon_gain + dmx_rms smoothing    demixer.c:621-634, 443-475
//   DMRenderer_set_mode_weight                    downmix_renderer.c:180-216
//   frame trimming / time-line placement          IAMF_decoder.c:3354-3407
//   resampler output count (closed form)          resample.c:357-418, SURVEY 9.4-3
//   limiter priming (padsize)                     audio_effect_peak_limiter.c:185-201
// -------------------------------------------------------------------------------------------------------------------
struct ResolveArgs {
  const iamfb_frame_params *params;   // [S][F]
  StreamState *state;                 // [S]
  FrameRec *frames;                   // [S][F]
  SubmitRec *submit;                  // [S]
  int *gate, *gate_next;              // optional: *gate |= (some stream of this submit is irregular); *gate_next = 0 (the flag
                                      // of the submit after this one) - the multi-kernel launches behind k_pipe_rs return at once
                                      // when the flag is clear
  SubmitRec *submit_mk;               // optional [S]: the same records with the REGULAR streams' lengths zeroed - what the
                                      // multi-kernel path works from when k_pipe_rs renders the regular streams of the submit
  int32_t *out_counts;                // [S][F] (device copy of what decode() returns per frame)
  const float *qf_table;              // 256 entries: (float)(q / 255.0)  (fixedp11_5.c:53-55)
  int n_streams, n_frames;
  int flush;                          // end-of-stream pass: no frames, 240 limiter zeros (+ resampler tail)
  int n_sub;                          // sub-chunks of this submit
  int sub_frame[kMaxSub + 1];         // frame index where each sub-chunk starts (sub_frame[n_sub] == n_frames)
};

__device__ __forceinline__ long long rs_outputs_until(const KernelPlan &p, long long in_total) {
  // number of outputs n >= 0 with  filt_len/2 + int_adv*n + floor(frac_adv*n/den) < in_total
  long long lim = in_total - (long long)(p.rs_filt_len / 2);
  if (lim <= 0) return 0;
  // q(n) = (num*n) / den  (num = int_adv*den + frac_adv);   q(n) < lim  <=>  num*n < lim*den
  //   <=> n <= (lim*den - 1) / num
  long long num = (long long)p.rs_num, den = (long long)p.rs_den;
  return (lim * den - 1) / num + 1;
}

// One WARP per stream.  The raw parameter records of up to 32 frames are loaded at once, lane f holding the 12 words
// of frame f (one coalesced read per chunk instead of a round trip per frame), and handed to the whole warp frame by
// frame with shuffles.  Pass 1 walks the frames once per audio element: lane c owns IAChannel c - its recon gain as
// received, as used by the de-mixer, and its smoothed factor - while the scalar state machines (de-mixing mode, w
// index, down-mixer mode) run redundantly in every lane; each lane writes its own slot of the frame record.  Pass 2
// walks the frames for the stream-level bookkeeping (trimming, time-line placement, sample counts), every lane
// computing the same scalars and lane 0 storing them.
constexpr int kResolveThreads = 128;   // 4 streams per block
static __global__ void __launch_bounds__(kResolveThreads) k_resolve(const __grid_constant__ KernelPlan plan, ResolveArgs a) {
  // programmatic dependent launch: the kernel launched after this one (k_stream) may start its blocks right away - they
  // load their limiter history and curve, which this kernel does not touch, and wait (griddepcontrol.wait) before they
  // read anything written here
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ float s_qf[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_qf[i] = a.qf_table[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (kResolveThreads / 32) + (threadIdx.x >> 5);
  if (s >= a.n_streams) return;
  constexpr unsigned kFull = 0xffffffffu;
  StreamState &gst = a.state[s];
  const int N = plan.frame_size;
  const uint32_t *pwords = reinterpret_cast<const uint32_t *>(a.params + (size_t)s * a.n_frames);
  FrameRec *frames = a.frames + (size_t)s * a.n_frames;
  constexpr int kPW = (int)(sizeof(iamfb_frame_params) / 4);   // 12 words: el[0] 0-4, el[1] 5-9, out_gain 10, trims 11
  static_assert(kPW == 12 && kMaxEl == 2, "parameter record layout");
  // the chunk of frames [fb, fb + 32): lane l holds frame fb + l
  uint32_t pw[kPW];
  auto load_chunk = [&](int fb) {
    const int f = fb + lane;
#pragma unroll
    for (int i = 0; i < kPW; ++i) pw[i] = f < a.n_frames ? __ldg(pwords + (size_t)f * kPW + i) : 0u;
  };

  // ---------------------------------------------------------------- pass 1: per element
  const int c = lane;                       // this lane's IAChannel id (1 .. kChCount-1 are channels)
  const bool is_ch = c >= 1 && c < kChCount;
  for (int e = 0; e < plan.n_elements; ++e) {
    const ElPlan &ep = plan.el[e];
    if (ep.kind != IAMFB_EL_CHANNEL) {
      for (int fb = 0; fb < a.n_frames; fb += 32) {
        load_chunk(fb);
        const int f = fb + lane;
        if (f < a.n_frames && (pw[11] & 0xffffu) != 0xFFFFu) {
          ElFrame &ef = frames[f].el[e];
          ef.gain = __uint_as_float(pw[5 * e + 4]);
          ef.rmask = 0;
          ef.mode = 0;
          ef.w = 0.f;
        }
      }
      continue;
    }
    ElState &ges = gst.el[e];
    int mode = ges.mode, w_idx = ges.w_idx, dmr_mode = ges.dmr_mode, dmr_w_idx = ges.dmr_w_idx;
    float dmr_tl = ges.dmr_tl;
    unsigned int rflags = ges.rflags, re_flags = ges.re_flags;
    float rgain = 0.f, re_gain = 0.f, sfavg = 0.f;
    if (is_ch) { rgain = ges.rgain[c]; re_gain = ges.re_gain[c]; sfavg = ges.sfavg[c]; }
    const int rbit = is_ch ? ep.recon_bit[c] : -1;     // recon-gain flag bit of this lane's channel, or -1
    const int slot = is_ch ? ep.slot_of[c] : -1;       // its slot in the layout, or -1
    const bool tl_derived = !((ep.dmr_in_mask >> IAMFB_CH_TL) & 1u) && !((ep.dmr_in_mask >> IAMFB_CH_TR) & 1u);

    for (int fb = 0; fb < a.n_frames; fb += 32) {
      load_chunk(fb);
      const int nf = min(32, a.n_frames - fb);
      for (int fl = 0; fl < nf; ++fl) {
        const uint32_t w0 = __shfl_sync(kFull, pw[5 * e + 0], fl);
        const uint32_t g0 = __shfl_sync(kFull, pw[5 * e + 1], fl);
        const uint32_t g1 = __shfl_sync(kFull, pw[5 * e + 2], fl);
        const uint32_t g2 = __shfl_sync(kFull, pw[5 * e + 3], fl);
        const float mix_gain = __uint_as_float(__shfl_sync(kFull, pw[5 * e + 4], fl));
        const uint32_t trims = __shfl_sync(kFull, pw[11], fl);
        // "this stream has no frame in this step" (grouped handles stepping together, IAMF_decoder_decode_batch):
        // nothing is decoded, so no state machine advances
        if ((trims & 0xffffu) == 0xFFFFu) continue;
        ElFrame &ef = frames[fb + fl].el[e];
        // --- recon gain list of the selected layer (latest received wins), IAMF_decoder.c:2238-2274
        if ((w0 >> 8) & 0xffu) {
          const unsigned int fl_ = w0 >> 16;
          re_flags = fl_;
          if (rbit >= 0 && ((fl_ >> rbit) & 1u)) {
            const int pos = __popc(fl_ & ((1u << rbit) - 1u));          // list position = set bits below
            const unsigned int wsel = pos < 4 ? g0 : (pos < 8 ? g1 : g2);
            re_gain = s_qf[(wsel >> (8 * (pos & 3))) & 0xffu];
          }
        }
        // --- demixer_set_recon_gain, demixer.c:621-634 (called every frame when the layer has a recon list)
        if (ep.recon_present && re_flags) {
          rflags = re_flags;
          if (rbit >= 0 && ((re_flags >> rbit) & 1u)) rgain = re_gain;
        }
        // --- demixer_set_demixing_info(mode, -1), demixer.c:592-619
        const int m_in = (int)(signed char)(w0 & 0xffu);
        if (m_in >= 0 && m_in != 3 && m_in <= 6) {
          mode = m_in;
          w_idx = c_mix_woff[m_in] > 0 ? min(w_idx + 1, 10) : max(w_idx - 1, 0);
        }
        // --- dmx_rms factor update, demixer.c:443-475: sfavg = 0.25*sf + 0.75*last, for every channel of the list
        // (slots without a recon gain hold 1.0, which the branch-free multiply of k_stream relies on)
        const bool upd = rbit >= 0 && ((rflags >> rbit) & 1u);
        float last = sfavg, cur = sfavg;
        if (upd) {
          const float sf = rgain;
          const float Nf = 7.f;
          cur = (2 / (Nf + 1)) * sf + (1 - 2 / (Nf + 1)) * last;
          sfavg = cur;
        }
        const unsigned int rmask = __reduce_or_sync(kFull, (upd && slot >= 0) ? (1u << slot) : 0u);
        if (upd && slot >= 0) {
          ef.rlast[slot] = last;
          ef.rcur[slot] = cur;
        }
        if (lane < 2 * IAMFB_MAX_LAYOUT_CH && !((rmask >> (lane % IAMFB_MAX_LAYOUT_CH)) & 1u)) ef.rlast[lane] = 1.f;   // rlast | rcur are contiguous
        // --- DMRenderer_set_mode_weight(mode, -1), downmix_renderer.c:180-216
        if (ep.renderer == kRdrDMR) {
          if (m_in >= 0 && m_in != 3 && m_in < 7) {
            dmr_mode = m_in;
            dmr_w_idx = c_mix_woff[m_in] > 0 ? min(dmr_w_idx + 1, 10) : max(dmr_w_idx - 1, 0);
            if (tl_derived) dmr_tl = c_mix_gamma[m_in] * c_w_table[dmr_w_idx];
          }
        }
        if (lane == 0) {
          ef.mode = mode;
          ef.w = c_w_table[min(max(w_idx, 0), 10)];
          ef.gain = mix_gain;
          ef.rmask = rmask;
          if (ep.renderer == kRdrDMR) {
            const int dm = dmr_mode & 7;
            ef.dmr_alpha = c_mix_alpha[dm];
            ef.dmr_beta = c_mix_beta[dm];
            ef.dmr_gamma = c_mix_gamma[dm];
            ef.dmr_delta = c_mix_delta[dm];
            ef.dmr_tl = dmr_tl;
          }
        }
      }
    }
    if (lane == 0) {
      ges.mode = mode; ges.w_idx = w_idx; ges.dmr_mode = dmr_mode; ges.dmr_w_idx = dmr_w_idx; ges.dmr_tl = dmr_tl;
      ges.rflags = rflags; ges.re_flags = re_flags;
    }
    if (is_ch) { ges.rgain[c] = rgain; ges.re_gain[c] = re_gain; ges.sfavg[c] = sfavg; }
  }

  // ---------------------------------------------------------------- pass 2: stream-level bookkeeping
  long long rs_in_total = gst.rs_in_total;
  const long long rs_out0 = gst.rs_out_total;
  int lim_pad = gst.lim_pad, lim_init = gst.lim_init;
  const int pad_at_start = lim_pad;
  int t_off = 0, lim_in_total = 0, sub = 0, irregular = a.flush ? 1 : 0;
  SubmitRec sr;
  for (int fb = 0; fb < a.n_frames; fb += 32) {
    load_chunk(fb);
    const int nf = min(32, a.n_frames - fb);
    for (int fl = 0; fl < nf; ++fl) {
      const int f = fb + fl;
      const float out_gain = __uint_as_float(__shfl_sync(kFull, pw[10], fl));
      const uint32_t trims = __shfl_sync(kFull, pw[11], fl);
      while (sub <= a.n_sub && a.sub_frame[sub] == f) sr.sub_off[sub++] = lim_in_total;
      FrameRec &fr = frames[f];
      if ((trims & 0xffffu) == 0xFFFFu) {
        if (lane == 0) {
          fr.out_gain = 1.f;
          fr.vstart = 0;
          fr.vlen = 0;
          fr.t_off = t_off;
          if (a.out_counts) a.out_counts[(size_t)s * a.n_frames + f] = 0;
        }
        irregular = 1;
        continue;
      }
      // --- trimming: a fully trimmed frame is decoded (state above advances) but dropped, IAMF_decoder.c:3354-3358
      const int ts = (int)(trims & 0xffffu), te = (int)(trims >> 16);
      int vlen = N - ts - te;
      if (ts == N || te == N || vlen < 0) vlen = 0;
      if (lane == 0) {
        fr.out_gain = out_gain;
        fr.vstart = ts;
        fr.vlen = vlen;
        fr.t_off = t_off;
      }
      t_off += vlen;
      if (ts != 0 || vlen != N) irregular = 1;
      // --- per-frame sample count returned to the caller
      int cnt = vlen;
      if (vlen > 0 && plan.resample) {
        const long long before = rs_outputs_until(plan, rs_in_total);
        rs_in_total += vlen;
        const long long after = rs_outputs_until(plan, rs_in_total);
        cnt = (int)(after - before);
      }
      if (vlen > 0) {
        lim_in_total += cnt;
        if (plan.limiter && !lim_init) {
          if (lim_pad >= cnt) { lim_pad -= cnt; cnt = 0; }
          else { cnt -= lim_pad; lim_pad = 0; lim_init = 1; }
        }
      }
      if (lane == 0 && a.out_counts) a.out_counts[(size_t)s * a.n_frames + f] = cnt;
    }
  }

  sr.in_len = t_off;
  sr.rs_out_first = rs_out0;
  if (a.flush) {
    // iamf_delay_buffer_handle, IAMF_decoder.c:3250-3301: resampler fed filt_len/2 zeros (output capped at the
    // output latency), then the limiter is fed that tail followed by delaySize zeros.
    int tail = 0;
    if (plan.resample) {
      const long long before = rs_outputs_until(plan, rs_in_total);
      const long long after = rs_outputs_until(plan, rs_in_total + plan.rs_filt_len / 2);
      const long long lat = ((long long)(plan.rs_filt_len / 2) * plan.rs_den + (plan.rs_num >> 1)) / plan.rs_num;
      tail = (int)min(after - before, lat);
      rs_in_total += plan.rs_filt_len / 2;
    }
    int cnt = tail;
    if (plan.limiter) {
      cnt = tail + kLimDelay;
      lim_in_total = cnt;
      if (!lim_init) {
        if (lim_pad >= cnt) { lim_pad -= cnt; cnt = 0; }
        else { cnt -= lim_pad; lim_pad = 0; lim_init = 1; }
      }
    } else {
      lim_in_total = cnt;
    }
    sr.in_len = plan.resample ? (int)(plan.rs_filt_len / 2) : 0;
    if (lane == 0 && a.out_counts) a.out_counts[s] = cnt;
  }
  sr.lim_len = lim_in_total;
  sr.irregular = irregular;
  for (; sub <= kMaxSub; ++sub) sr.sub_off[sub] = lim_in_total;
  if (a.flush) sr.sub_off[0] = 0;
  sr.out_skip = pad_at_start - lim_pad;
  sr.out_len = lim_in_total - sr.out_skip;
  if (lane == 0) {
    gst.rs_in_total = rs_in_total;
    gst.rs_out_total = rs_out0 + (plan.resample ? (long long)lim_in_total - (a.flush && plan.limiter ? kLimDelay : 0) : 0);
    gst.lim_pad = lim_pad;
    gst.lim_init = lim_init;
    a.submit[s] = sr;
    if (a.gate) {
      if (irregular) atomicOr(a.gate, 1);
      if (s == 0) *a.gate_next = 0;
    }
    if (a.submit_mk) {
      if (!irregular) {
        sr.in_len = 0; sr.lim_len = 0; sr.out_len = 0; sr.out_skip = 0;
        for (int i = 0; i <= kMaxSub; ++i) sr.sub_off[i] = 0;
      }
      a.submit_mk[s] = sr;
    }
  }
}

// -------------------------------------------------------------------------------------------------------------------
// K1: reconstruct + render + gains (+ element sum) for one audio element.
//   thread = VEC consecutive samples of one (stream, frame); block = 128 threads = one 128*VEC-sample tile.
//   Everything between the decoded frame and the mixed time line stays in registers.
// -------------------------------------------------------------------------------------------------------------------
struct RenderArgs {
  const float *in;          // [S][F][n_in][N]
  const FrameRec *frames;   // [S][F]
  const float *gain_ramp;   // optional [S][F][N]
  const float *out_gain_ramp;
  const float *start_win;   // [overlap] hann[j]          (demixer.c:549-552)
  const float *stop_win;    // [overlap] hann[j+overlap]
  float *tl;                // destination time line [S][C_out][cap]
  float *pk;                // [S][cap] or nullptr
  int cap, hist;            // row stride and history offset of tl / pk
  int n_frames, e;          // frames per stream in this submit, element index
  int f_lo, nf;             // this launch covers frames [f_lo, f_lo + nf)
  int first, last;          // first element writes, later ones accumulate; the last applies output gain/loudness/peak
  int tiles_per_frame;
  const SubmitRec *only_irregular;   // non-null: render only the streams flagged irregular (the others are k_pipe_rs's)
  int n_blocks;                      // (stream, frame, tile) items of this launch
  const int *gate;                   // non-null: return at once when *gate == 0 (no irregular stream in this submit)
};

template <int VEC>
struct Vec {
  float v[VEC];
};

// SMEM: the source is a shared-memory tile, else global memory
template <int VEC, bool SMEM>
__device__ __forceinline__ Vec<VEC> load_row(const float *p, bool vec_ok, int valid) {
  Vec<VEC> r;
  if constexpr (VEC == 4) {
    if (vec_ok) {
      float4 t;
      if constexpr (SMEM) t = *reinterpret_cast<const float4 *>(p);
      else t = ldg_stream4(p);
      r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
      return r;
    }
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    if constexpr (SMEM) r.v[k] = k < valid ? p[k] : 0.f;
    else r.v[k] = k < valid ? ldg_stream1(p + k) : 0.f;
  }
  return r;
}

// Channel-based reconstruction of the VEC samples held by this thread: fills v[IAChannel][k].
// Follows dmx_gainup, dmx_s2..dmx_h4 and dmx_rms of demixer.c (skip == 0: codec delay is 0 on this path).
template <int VEC, bool SMEM>
__device__ __forceinline__ void reconstruct_channels(const ElPlan &ep, const ElFrame &ef, const float *src, int row_stride,
                                                     bool vec_ok, int valid, Vec<VEC> (&v)[kChCount]) {
  // transmitted channels -> IAChannel slots (static register indices, uniform predicates).  All loads are issued
  // before the first use so that a thread has its whole input (n_in x 16 B) in flight at once.
#pragma unroll
  for (int c = 1; c < kChCount; ++c) {
    int row = ep.src_row[c];
    if (row >= 0) {
      v[c] = load_row<VEC, SMEM>(src + (size_t)row * row_stride, vec_ok, valid);
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[c].v[k] = 0.f;
    }
  }
  if (ep.gain_mask) {   // dmx_gainup, demixer.c:421-430
#pragma unroll
    for (int c = 1; c < kChCount; ++c)
      if ((ep.gain_mask >> c) & 1u) {
        float g = ep.gain[c];
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[c].v[k] *= g;
      }
  }
  const int mode = ef.mode & 7;
  if (ep.need_s2) {   // R2 = 2*Mono - L2, demixer.c:136-138
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[IAMFB_CH_R2].v[k] = 2 * v[IAMFB_CH_MONO].v[k] - v[IAMFB_CH_L2].v[k];
  }
  if (ep.need_s3) {   // L3 = L2 - 0.707*C evaluated in double, demixer.c:165-168
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      double c = (double)v[IAMFB_CH_C].v[k];
      v[IAMFB_CH_L3].v[k] = (float)((double)v[IAMFB_CH_L2].v[k] - 0.707 * c);
      v[IAMFB_CH_R3].v[k] = (float)((double)v[IAMFB_CH_R2].v[k] - 0.707 * c);
    }
  }
  if (ep.need_s5) {   // Ls5 = (L3 - L5)/delta, demixer.c:213-218
    float d = c_mix_delta[mode];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      v[IAMFB_CH_SL5].v[k] = (v[IAMFB_CH_L3].v[k] - v[IAMFB_CH_L5].v[k]) / d;
      v[IAMFB_CH_SR5].v[k] = (v[IAMFB_CH_R3].v[k] - v[IAMFB_CH_R5].v[k]) / d;
    }
  }
  if (ep.need_s7) {   // Lb7 = (Ls5 - alpha*Lss7)/beta, demixer.c:262-269
    float al = c_mix_alpha[mode], be = c_mix_beta[mode];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      v[IAMFB_CH_BL7].v[k] = (v[IAMFB_CH_SL5].v[k] - v[IAMFB_CH_SL7].v[k] * al) / be;
      v[IAMFB_CH_BR7].v[k] = (v[IAMFB_CH_SR5].v[k] - v[IAMFB_CH_SR7].v[k] * al) / be;
    }
  }
  if (ep.need_h2) {   // Ltf2 = Ltf3 - delta*w*Ls5, demixer.c:318-323
    float dw = c_mix_delta[mode] * ef.w;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      v[IAMFB_CH_HL].v[k] = v[IAMFB_CH_TL].v[k] - dw * v[IAMFB_CH_SL5].v[k];
      v[IAMFB_CH_HR].v[k] = v[IAMFB_CH_TR].v[k] - dw * v[IAMFB_CH_SR5].v[k];
    }
  }
  if (ep.need_h4) {   // Ltb = (Ltf2 - Ltf4)/gamma, demixer.c:363-368
    float ga = c_mix_gamma[mode];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      v[IAMFB_CH_HBL].v[k] = (v[IAMFB_CH_HL].v[k] - v[IAMFB_CH_HFL].v[k]) / ga;
      v[IAMFB_CH_HBR].v[k] = (v[IAMFB_CH_HR].v[k] - v[IAMFB_CH_HFR].v[k]) / ga;
    }
  }
}

// value of IAChannel `ch` as the parametric down-mixer computes it (downmix_renderer.c:65-75,115-129):
// an input channel is passed through, everything else is the ordered two-term sum of its dependencies.
template <int VEC>
__device__ __forceinline__ void dmr_prepare(const ElPlan &ep, const ElFrame &ef, Vec<VEC> (&v)[kChCount]) {
  const unsigned int inm = ep.dmr_in_mask;
  auto is_in = [&](int c) { return (inm >> c) & 1u; };
  auto two = [&](int dst, int a, float sa, int b, float sb) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float sum = 0.f;
      sum += v[a].v[k] * sa;
      sum += v[b].v[k] * sb;
      v[dst].v[k] = sum;
    }
  };
  // dependency order: SL5 <- SL7,BL7 ; L3 <- L5,SL5 ; L2 <- L3,C ; HL <- HFL,HBL ; TL <- HL,SL5 ; Mono <- R2,L2
  if (!is_in(IAMFB_CH_SL5)) two(IAMFB_CH_SL5, IAMFB_CH_SL7, ef.dmr_alpha, IAMFB_CH_BL7, ef.dmr_beta);
  if (!is_in(IAMFB_CH_SR5)) two(IAMFB_CH_SR5, IAMFB_CH_SR7, ef.dmr_alpha, IAMFB_CH_BR7, ef.dmr_beta);
  if (!is_in(IAMFB_CH_L3)) two(IAMFB_CH_L3, IAMFB_CH_L5, 1.f, IAMFB_CH_SL5, ef.dmr_delta);
  if (!is_in(IAMFB_CH_R3)) two(IAMFB_CH_R3, IAMFB_CH_R5, 1.f, IAMFB_CH_SR5, ef.dmr_delta);
  if (!is_in(IAMFB_CH_L2)) two(IAMFB_CH_L2, IAMFB_CH_L3, 1.f, IAMFB_CH_C, 0.707f);
  if (!is_in(IAMFB_CH_R2)) two(IAMFB_CH_R2, IAMFB_CH_R3, 1.f, IAMFB_CH_C, 0.707f);
  if (!is_in(IAMFB_CH_HL)) two(IAMFB_CH_HL, IAMFB_CH_HFL, 1.f, IAMFB_CH_HBL, ef.dmr_gamma);
  if (!is_in(IAMFB_CH_HR)) two(IAMFB_CH_HR, IAMFB_CH_HFR, 1.f, IAMFB_CH_HBR, ef.dmr_gamma);
  if (!is_in(IAMFB_CH_TL)) two(IAMFB_CH_TL, IAMFB_CH_HL, 1.f, IAMFB_CH_SL5, ef.dmr_tl);
  if (!is_in(IAMFB_CH_TR)) two(IAMFB_CH_TR, IAMFB_CH_HR, 1.f, IAMFB_CH_SR5, ef.dmr_tl);
  if (!is_in(IAMFB_CH_MONO)) two(IAMFB_CH_MONO, IAMFB_CH_R2, 0.5f, IAMFB_CH_L2, 0.5f);
}

template <int VEC>
__device__ __forceinline__ Vec<VEC> pick_channel(const Vec<VEC> (&v)[kChCount], int ch) {
  Vec<VEC> r;
#pragma unroll
  for (int k = 0; k < VEC; ++k) r.v[k] = 0.f;
  // uniform switch with static register reads in every arm (avoids dynamically indexed registers)
#pragma unroll
  for (int c = 1; c < kChCount; ++c)
    if (ch == c) r = v[c];
  return r;
}

// LAYOUT >= 0: channel based element with that reconstructed layout (x[m] gathered with static indices);
// LAYOUT == -1: scene based element (NREC = ambisonics channel count).
// One thread's share of the work: VEC samples starting at frame offset i0 of (stream s, frame index sf).
// src points at the thread's first sample of decoded row 0, rows are row_stride floats apart.
template <int LAYOUT, int NREC, int VEC, bool SMEM>
__device__ __forceinline__ void render_thread(const KernelPlan &plan, const RenderArgs &a, int s, int sf, int i0,
                                              const float *src, int row_stride) {
  const int N = plan.frame_size;
  const FrameRec &fr = a.frames[sf];
  const int vstart = fr.vstart, vlen = fr.vlen;
  if (vlen <= 0) return;
  // samples of this thread that survive trimming
  const int lo = max(i0, vstart), hi = min(i0 + VEC, vstart + vlen);
  if (lo >= hi) return;
  const ElPlan &ep = plan.el[a.e];
  const ElFrame &ef = fr.el[a.e];
  const int valid = min(VEC, N - i0);
  const bool vec_ok = (VEC == 4) && ((N & 3) == 0) && valid == VEC;

  Vec<VEC> x[NREC];   // channels entering the renderer, in renderer order
  Vec<VEC> v[(LAYOUT >= 0) ? kChCount : 1];

  if constexpr (LAYOUT >= 0) {
    reconstruct_channels<VEC, SMEM>(ep, ef, src, row_stride, vec_ok, valid, v);
    // dmx_rms cross-fade, demixer.c:461-468: x *= last*stop[i] + cur*start[i]
    constexpr unsigned char kOrder[9][12] = {
        {13}, {14, 15}, {1, 2, 3, 4, 20, 21}, {1, 2, 3, 4, 20, 21, 22, 23}, {1, 2, 3, 4, 20, 21, 9, 10, 11, 12},
        {1, 2, 3, 4, 5, 6, 7, 8}, {1, 2, 3, 4, 5, 6, 7, 8, 22, 23}, {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12},
        {18, 19, 3, 4, 16, 17}};
#pragma unroll
    for (int m = 0; m < NREC; ++m) {
      x[m] = v[kOrder[LAYOUT][m]];
      if ((ef.rmask >> m) & 1u) {
        const float last = ef.rlast[m], cur = ef.rcur[m];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          int i = i0 + k;
          float st = 0.f, sw = 1.f;
          if (i < plan.overlap) { st = a.stop_win[i]; sw = a.start_win[i]; }
          float f = last * st + cur * sw;
          x[m].v[k] *= f;
        }
      }
    }
    if (ep.renderer == kRdrDMR) {
      // the down-mixer works on IAChannel slots: write the (recon-gained) layout channels back
#pragma unroll
      for (int m = 0; m < NREC; ++m) v[kOrder[LAYOUT][m]] = x[m];
      dmr_prepare<VEC>(ep, ef, v);
    }
  } else {
    // scene based: mono mapping is a row permutation, projection an ordered mat-vec (IAMF_core_decoder.c:105-130)
    if (ep.ambi_mode == 0) {
#pragma unroll
      for (int m = 0; m < NREC; ++m) x[m] = load_row<VEC, SMEM>(src + (size_t)ep.ambi_map[m] * row_stride, vec_ok, valid);
    } else {
#pragma unroll
      for (int m = 0; m < NREC; ++m)
#pragma unroll
        for (int k = 0; k < VEC; ++k) x[m].v[k] = .0f;
      for (int l = 0; l < ep.ambi_cols; ++l) {
        Vec<VEC> t = load_row<VEC, SMEM>(src + (size_t)l * row_stride, vec_ok, valid);
#pragma unroll
        for (int m = 0; m < NREC; ++m) {
          float c = ep.ambi_mat[l * NREC + m];
#pragma unroll
          for (int k = 0; k < VEC; ++k) x[m].v[k] += t.v[k] * c;
        }
      }
    }
  }

  // ---- render + gains, one output channel at a time so that only x[] stays live
  const int co = plan.out_channels;
  const size_t row0 = (size_t)s * co * a.cap;
  const int t0 = a.hist + fr.t_off + (i0 - vstart);        // time-line index of sample i0
  const bool st_vec = (VEC == 4) && lo == i0 && hi == i0 + VEC && ((t0 & 3) == 0) && ((a.cap & 3) == 0);
  Vec<VEC> eg, og;   // element / output gain per sample
  {
    const bool eg_on = a.gain_ramp || (ef.gain != 1.f && ef.gain > 0.f);   // iamf_frame_gain, IAMF_decoder.c:1392
    const bool og_on = a.out_gain_ramp || (fr.out_gain != 1.f && fr.out_gain > 0.f);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      int j = i0 + k - vstart;   // index inside the trimmed frame
      bool in_rng = (i0 + k) >= lo && (i0 + k) < hi;
      eg.v[k] = (a.gain_ramp && in_rng) ? a.gain_ramp[(size_t)sf * N + j] : (eg_on ? ef.gain : 1.f);
      og.v[k] = (a.out_gain_ramp && in_rng) ? a.out_gain_ramp[(size_t)sf * N + j] : (og_on ? fr.out_gain : 1.f);
    }
  }
  const bool loud_on = !plan.resample && plan.loud_gain != 0.f && plan.loud_gain != 1.0f;
  Vec<VEC> peak;
#pragma unroll
  for (int k = 0; k < VEC; ++k) peak.v[k] = 0.f;

  for (int oc = 0; oc < co; ++oc) {
    Vec<VEC> y;
#pragma unroll
    for (int k = 0; k < VEC; ++k) y.v[k] = 0.f;
    if (ep.renderer == kRdrDMR) {
      if constexpr (LAYOUT >= 0) {
        if (oc < ep.dmr_n_out) y = pick_channel<VEC>(v, ep.dmr_out_ch[oc]);
      }
    } else {
      // which matrix row lands on output channel oc (identity for M2M; LFE slots are zero for H2M)
      int n = ep.out_slot[oc];
      if (n >= 0) {
        // IAMF_element_renderer_render_M2M / _H2M: out = 0; out += mat*in over inputs ascending
        const float *mrow = ep.mat + n * NREC;
#pragma unroll
        for (int m = 0; m < NREC; ++m) {
          const float c = mrow[m];
          if (c != 0.f) {   // adding +-0 never changes the running sum (it starts at +0): exact skip
#pragma unroll
            for (int k = 0; k < VEC; ++k) y.v[k] += c * x[m].v[k];
          }
        }
      }
    }
    // element mix gain, IAMF_decoder.c:1392-1405
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      if (a.gain_ramp || eg.v[k] != 1.f) y.v[k] *= eg.v[k];

    float *dst = a.tl + row0 + (size_t)oc * a.cap + t0;
    // iamf_mixer_mix, IAMF_decoder.c:2719-2730: acc = 0; acc += e0; acc += e1
    if (a.first) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) y.v[k] = 0.f + y.v[k];
    } else {
      if (st_vec) {
        float4 p = *reinterpret_cast<const float4 *>(dst);
        y.v[0] = p.x + y.v[0]; y.v[1 % VEC] = p.y + y.v[1 % VEC]; y.v[2 % VEC] = p.z + y.v[2 % VEC]; y.v[3 % VEC] = p.w + y.v[3 % VEC];
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          if ((i0 + k) >= lo && (i0 + k) < hi) y.v[k] = dst[k] + y.v[k];
      }
    }
    if (a.last) {
      // output mix gain (IAMF_decoder.c:3463-3469) then loudness (IAMF_decoder.c:3480-3484) when not resampling
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        if (a.out_gain_ramp || og.v[k] != 1.f) y.v[k] *= og.v[k];
        if (loud_on) y.v[k] *= plan.loud_gain;
        peak.v[k] = fmaxf(peak.v[k], fabsf(y.v[k]));
      }
    }
    if (st_vec) {
      *reinterpret_cast<float4 *>(dst) = make_float4(y.v[0], y.v[1 % VEC], y.v[2 % VEC], y.v[3 % VEC]);
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        if ((i0 + k) >= lo && (i0 + k) < hi) dst[k] = y.v[k];
    }
  }
  if (a.last && a.pk) {
    float *pd = a.pk + (size_t)s * a.cap + t0;
    if (st_vec) {
      *reinterpret_cast<float4 *>(pd) = make_float4(peak.v[0], peak.v[1 % VEC], peak.v[2 % VEC], peak.v[3 % VEC]);
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        if ((i0 + k) >= lo && (i0 + k) < hi) pd[k] = peak.v[k];
    }
  }
}

// direct variant: every thread reads its samples straight from global memory (any frame size)
template <int LAYOUT, int NREC, int VEC>
static __global__ void __launch_bounds__(128, 4) k_render(const __grid_constant__ KernelPlan plan, RenderArgs a) {
  const int N = plan.frame_size;
  if (a.gate && *a.gate == 0) return;
  // (a launch for the irregular streams of a submit only uses a small grid and strides over the (stream, frame, tile) list)
  for (int bx = blockIdx.x; bx < a.n_blocks; bx += gridDim.x) {
    const int tile = bx % a.tiles_per_frame;
    const int sfl = bx / a.tiles_per_frame;
    const int s = sfl / a.nf;
    const int sf = s * a.n_frames + a.f_lo + sfl % a.nf;    // s * F + f
    const int i0 = (tile * 128 + threadIdx.x) * VEC;         // first sample of this thread inside the frame
    if (i0 >= N) continue;
    if (a.only_irregular && !a.only_irregular[s].irregular) continue;
    const float *src = a.in + (size_t)sf * plan.el[a.e].n_in * N + i0;
    render_thread<LAYOUT, NREC, VEC, false>(plan, a, s, sf, i0, src, N);
  }
}

// ---- bulk-copy (TMA) / mbarrier helpers used by the single-kernel paths
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// -------------------------------------------------------------------------------------------------------------------
// K2: Speex resampler, closed form.  Output n of a stream reads filt_len consecutive inputs ending at stream position
//   q(n) = filt_len/2 + int_adv*n + floor(frac_adv*n/den)  with phase  phi(n) = (frac_adv*n) mod den
// (resampler_basic_interpolate_single / _direct_single, resample.c:258-313,357-418; chunking of
// speex_resampler_process_float :917-972 does not change results).  Output is clamped to +-1 (:84,959), then the
// loudness gain is applied (IAMF_decoder.c:3480-3484) and the per-instant peak for the limiter is reduced.
// thread = one output instant of one stream, looping over channels (taps are reused across channels).
// -------------------------------------------------------------------------------------------------------------------
struct ResampleArgs {
  const float *src;        // tl_a [S][C][cap_a]: rs_hist history samples, then the inputs of this submit
  float *dst;              // tl_b [S][C][cap_b]
  float *pk;               // [S][cap_b]
  const SubmitRec *submit;
  const StreamState *state;   // state AFTER resolve (rs_in_total includes this submit)
  const float *sinc;       // table
  int cap_a, cap_b, hist_b;
  int max_out;             // grid covers this many outputs per stream
  int flush;
  int n_streams;
  const int *gate;         // non-null: return at once when *gate == 0
};

static __global__ void __launch_bounds__(128) k_resample(const __grid_constant__ KernelPlan plan, ResampleArgs a) {
  extern __shared__ float s_sinc[];
  const int use_direct = plan.rs_direct;
  const int tab_len = use_direct ? plan.rs_filt_len * plan.rs_den : plan.rs_filt_len * plan.rs_oversample + 8;
  const bool tab_smem = tab_len <= 12 * 1024;
  if (tab_smem)
    for (int i = threadIdx.x; i < tab_len; i += blockDim.x) s_sinc[i] = a.sinc[i];
  __syncthreads();
  const float *tab = tab_smem ? s_sinc : a.sinc;

  const int s = blockIdx.y;
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  const SubmitRec sr = a.submit[s];
  const int n_out = a.flush ? (sr.lim_len - (plan.limiter ? kLimDelay : 0)) : sr.lim_len;
  if (u >= n_out) return;
  const long long n = sr.rs_out_first + u;
  const int Nf = plan.rs_filt_len;
  // stream position of the last input sample this output needs, and where this submit's input starts
  const long long num = plan.rs_num, den = plan.rs_den;
  const long long q = (long long)(Nf / 2) + (n * num) / den;
  const unsigned int frac_num = (unsigned int)((n * (long long)plan.rs_frac_adv) % den);
  const long long in_start = a.state[s].rs_in_total - sr.in_len;     // stream position of tl_a[rs_hist]
  // first tap reads stream position q - (Nf-1)
  const long long p0 = q - (Nf - 1) - in_start + plan.rs_hist;             // index into tl_a row
  const int co = plan.out_channels;
  const bool loud_on = plan.loud_gain != 0.f && plan.loud_gain != 1.0f && !a.flush;
  float peak = 0.f;

  int offset = 0;
  float interp[4] = {0.f, 0.f, 0.f, 0.f};
  if (!use_direct) {
    offset = frac_num * plan.rs_oversample / plan.rs_den;
    const float frac = ((float)((frac_num * plan.rs_oversample) % plan.rs_den)) / plan.rs_den;
    // cubic_coef, resample.c:246-256 (interp[2] is a double expression rounded once)
    interp[0] = -0.16667f * frac + 0.16667f * frac * frac * frac;
    interp[1] = frac + 0.5f * frac * frac - 0.5f * frac * frac * frac;
    interp[3] = -0.33333f * frac + 0.5f * frac * frac - 0.16667f * frac * frac * frac;
    interp[2] = (float)(1. - (double)interp[0] - (double)interp[1] - (double)interp[3]);
  }

  for (int c = 0; c < co; ++c) {
    const float *row = a.src + ((size_t)s * co + c) * a.cap_a;
    float sum;
    if (use_direct) {
      const float *sinct = tab + (size_t)frac_num * Nf;
      sum = 0.f;
      for (int j = 0; j < Nf; ++j) {
        long long idx = p0 + j;
        float xin = idx >= 0 ? row[idx] : 0.f;
        sum += sinct[j] * xin;
      }
    } else {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const int os = plan.rs_oversample;
      for (int j = 0; j < Nf; ++j) {
        long long idx = p0 + j;
        float xin = idx >= 0 ? row[idx] : 0.f;
        const float *t = tab + 4 + (j + 1) * os - offset;
        a0 += xin * t[-2];
        a1 += xin * t[-1];
        a2 += xin * t[0];
        a3 += xin * t[1];
      }
      sum = interp[0] * a0 + interp[1] * a1 + interp[2] * a2 + interp[3] * a3;
    }
    // FLTADJUST, resample.c:84
    sum = (sum < -1.0f) ? -1.0f : ((sum > 1.0f) ? 1.0f : sum);
    if (loud_on) sum *= plan.loud_gain;
    a.dst[((size_t)s * co + c) * a.cap_b + a.hist_b + u] = sum;
    peak = fmaxf(peak, fabsf(sum));
  }
  if (a.pk) a.pk[(size_t)s * a.cap_b + a.hist_b + u] = peak;
}

// K2b: the interpolating resampler (the 44.1 -> 48 kHz case and every other non-integer-phase ratio) with everything
// on chip.  block = 128 consecutive outputs of one stream.  The four neighbouring taps an output needs per input
// sample are stored as ONE 16-byte item, tab4[offset][j] = sinc[4 + (j+1)*oversample - offset - {2,1,0,-1}] (rows padded
// to an odd number of items so that lanes with different phases hit different banks); the inputs the block's outputs
// span are staged per channel pair in shared memory, and the taps are reused for both channels.  The accumulation is
// the reference's: four accumulators per channel over j ascending, cubic blend, clamp (resample.c:357-418,84).
struct Resample2Args {
  ResampleArgs r;
  const float4 *tab4;      // [oversample][filt_len + 1]
  int span;                // staged inputs per channel: inputs spanned by 128 consecutive outputs + filt_len (multiple of 4)
};

static __global__ void __launch_bounds__(128) k_resample_interp(const __grid_constant__ KernelPlan plan, Resample2Args b) {
  extern __shared__ __align__(16) float rs_smem[];
  const ResampleArgs &a = b.r;
  if (a.gate && *a.gate == 0) return;
  const int Nf = plan.rs_filt_len, os = plan.rs_oversample;
  const int trow = Nf + 1;
  float4 *s_tab = reinterpret_cast<float4 *>(rs_smem);              // [os][Nf + 1]
  float *s_x = rs_smem + (size_t)4 * os * trow;                      // [2][span]
  for (int i = threadIdx.x; i < os * trow; i += blockDim.x) s_tab[i] = b.tab4[i];

  // (gridDim.y may be smaller than the stream count: the launch for the irregular streams of a submit strides over them)
  for (int s = blockIdx.y; s < a.n_streams; s += gridDim.y) {
  const int u0 = blockIdx.x * blockDim.x, u = u0 + threadIdx.x;
  const SubmitRec sr = a.submit[s];
  const int n_out = a.flush ? (sr.lim_len - (plan.limiter ? kLimDelay : 0)) : sr.lim_len;
  if (u0 >= n_out) continue;
  const long long num = plan.rs_num, den = plan.rs_den;
  const long long in_start = a.state[s].rs_in_total - sr.in_len;     // stream position of tl_a[rs_hist]
  // first input (index into the tl_a row) needed by the block's first output
  const long long nb = sr.rs_out_first + u0;
  const long long pb = (long long)(Nf / 2) + (nb * num) / den - (Nf - 1) - in_start + plan.rs_hist;
  const bool live = u < n_out;
  const long long n = sr.rs_out_first + (live ? u : n_out - 1);
  const long long q = (long long)(Nf / 2) + (n * num) / den;
  const unsigned int frac_num = (unsigned int)((n * (long long)plan.rs_frac_adv) % den);
  const int rel = (int)(q - (Nf - 1) - in_start + plan.rs_hist - pb);   // first tap inside the staged span
  const int offset = frac_num * plan.rs_oversample / plan.rs_den;
  const float frac = ((float)((frac_num * plan.rs_oversample) % plan.rs_den)) / plan.rs_den;
  // cubic_coef, resample.c:246-256 (interp[2] is a double expression rounded once)
  float interp[4];
  interp[0] = -0.16667f * frac + 0.16667f * frac * frac * frac;
  interp[1] = frac + 0.5f * frac * frac - 0.5f * frac * frac * frac;
  interp[3] = -0.33333f * frac + 0.5f * frac * frac - 0.16667f * frac * frac * frac;
  interp[2] = (float)(1. - (double)interp[0] - (double)interp[1] - (double)interp[3]);
  const float4 *trow_p = s_tab + (size_t)offset * trow;
  const int co = plan.out_channels;
  const bool loud_on = plan.loud_gain != 0.f && plan.loud_gain != 1.0f && !a.flush;
  float peak = 0.f;

  for (int c0 = 0; c0 < co; c0 += 2) {
    const int nc = min(2, co - c0);
    __syncthreads();                                                  // table staged / previous pair consumed
    for (int i = threadIdx.x; i < nc * b.span; i += blockDim.x) {
      const int cc = i / b.span, k = i - cc * b.span;
      const long long idx = pb + k;
      const float *row = a.src + ((size_t)s * co + c0 + cc) * a.cap_a;
      s_x[cc * b.span + k] = (idx >= 0 && idx < a.cap_a) ? row[idx] : 0.f;
    }
    __syncthreads();
    if (!live) continue;
    const float *x0 = s_x + rel, *x1 = s_x + (nc > 1 ? b.span : 0) + rel;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll 8
    for (int j = 0; j < Nf; ++j) {
      const float4 t = trow_p[j];
      const float xa = x0[j], xb = x1[j];
      a0 += xa * t.x; a1 += xa * t.y; a2 += xa * t.z; a3 += xa * t.w;
      b0 += xb * t.x; b1 += xb * t.y; b2 += xb * t.z; b3 += xb * t.w;
    }
    float sum0 = interp[0] * a0 + interp[1] * a1 + interp[2] * a2 + interp[3] * a3;
    float sum1 = interp[0] * b0 + interp[1] * b1 + interp[2] * b2 + interp[3] * b3;
    // FLTADJUST, resample.c:84
    sum0 = (sum0 < -1.0f) ? -1.0f : ((sum0 > 1.0f) ? 1.0f : sum0);
    sum1 = (sum1 < -1.0f) ? -1.0f : ((sum1 > 1.0f) ? 1.0f : sum1);
    if (loud_on) { sum0 *= plan.loud_gain; sum1 *= plan.loud_gain; }
    a.dst[((size_t)s * co + c0) * a.cap_b + a.hist_b + u] = sum0;
    peak = fmaxf(peak, fabsf(sum0));
    if (nc > 1) {
      a.dst[((size_t)s * co + c0 + 1) * a.cap_b + a.hist_b + u] = sum1;
      peak = fmaxf(peak, fabsf(sum1));
    }
  }
  if (live && a.pk) a.pk[(size_t)s * a.cap_b + a.hist_b + u] = peak;
  __syncthreads();                                                    // the staged inputs are consumed before the next stream's arrive
  }
}

// -------------------------------------------------------------------------------------------------------------------
// K3a: sliding maximum.  wm[k] = max(pk[k-240 .. k-1]) = the `peak` the reference limiter looks up for instant k
// (audio_effect_peak_limiter.c:109-133; the arg-max cache there is only an optimisation).
// Doubling in shared memory: windows of 1,2,4,...,128 then 240 = 128+64+32+16.
// block = (stream, 1024-instant tile); the time axis has kLimDelay history in front.
// -------------------------------------------------------------------------------------------------------------------
struct WmaxArgs {
  const float *pk;     // [S][cap]
  float *wm;           // [S][cap]   wm[hist + k] for k in [0, lim_len)
  const SubmitRec *submit;
  int cap, hist;
  int flush;
  int sub;             // sub-chunk processed by this launch
  int n_streams;
  const int *gate;     // non-null: return at once when *gate == 0
};

constexpr int kWmTile = 1024;

static __global__ void __launch_bounds__(256) k_window_max(const __grid_constant__ KernelPlan plan, WmaxArgs a) {
  __shared__ float sa[kWmTile + kLimDelay + 16];
  __shared__ float sb[kWmTile + kLimDelay + 16];
  if (a.gate && *a.gate == 0) return;
  for (int s = blockIdx.y; s < a.n_streams; s += gridDim.y) {
  __syncthreads();                                         // (the previous stream's scratch is consumed)
  const int len = a.submit[s].sub_off[a.sub + 1];          // instants [sub_off[sub], sub_off[sub+1]) of this submit
  const int k0 = a.submit[s].sub_off[a.sub] + blockIdx.x * kWmTile;
  if (k0 >= len) continue;
  const float *row = a.pk + (size_t)s * a.cap + a.hist;   // row[k] = pk of instant k of this submit
  const int span = kWmTile + kLimDelay;                    // instants k0-240 .. k0+1023
  for (int i = threadIdx.x; i < span + 16; i += blockDim.x) {
    int k = k0 - kLimDelay + i;
    sa[i] = (i < span && k < len) ? row[k] : 0.f;          // k >= -240 always inside the history
  }
  __syncthreads();
  // after the pass with step d: buf[i] = max(pk[i .. i+2d-1])
  float *src = sa, *dst = sb;
  // the 16/32/64 windows are snapshotted into registers as the passes go by: every thread owns output instants i = threadIdx.x + r*256 (r < 4) and needs
  //   W16[i+224], W32[i+192], W64[i+128], W128[i]   (window start index i corresponds to instant k0-240+i)
  float r16[4], r32[4], r64[4];
  for (int d = 1; d <= 64; d <<= 1) {
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
      int j = i + d;
      dst[i] = fmaxf(src[i], j < span ? src[j] : 0.f);
    }
    __syncthreads();
    float *t = src; src = dst; dst = t;
    // src now holds windows of 2d
    if (2 * d == 16) {
#pragma unroll
      for (int r = 0; r < 4; ++r) r16[r] = src[threadIdx.x + r * 256 + 224];
    } else if (2 * d == 32) {
#pragma unroll
      for (int r = 0; r < 4; ++r) r32[r] = src[threadIdx.x + r * 256 + 192];
    } else if (2 * d == 64) {
#pragma unroll
      for (int r = 0; r < 4; ++r) r64[r] = src[threadIdx.x + r * 256 + 128];
    }
  }
  // src holds windows of 128
  float *out = a.wm + (size_t)s * a.cap + a.hist;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int i = threadIdx.x + r * 256;
    int k = k0 + i;
    if (k < len) out[k] = fmaxf(fmaxf(src[i], r64[r]), fmaxf(r32[r], r16[r]));
  }
  }
}

// -------------------------------------------------------------------------------------------------------------------
// K3b: limiter gain recurrence (compute_target_gain, audio_effect_peak_limiter.c:237-271).
// The float time constant of the reference only ever takes the values T[j] = j-fold float accumulation of 1/fs from 0
// (it restarts at 0 on every trigger), so the state is the integer j; acc[j] = curve_accel(...) is precomputed on the
// host with the same libm.  One lane per stream; a warp stages 32x32 tiles through shared memory so that global
// accesses stay coalesced.  Tiles in which every lane is idle and below threshold are answered without the serial walk.
// -------------------------------------------------------------------------------------------------------------------
struct ScanArgs {
  const float *wm;      // [S][cap]
  float *gn;            // [S][cap]
  StreamState *state;
  const SubmitRec *submit;
  const float *acc;     // [jr + 1]
  int cap, hist, n_streams;
  int max_len;
  int sub;              // sub-chunk processed by this launch
  const int *gate;      // non-null: return at once when *gate == 0
};

// Warp-specialised block of 16 warps per 32 streams:
//   warp 0 ("scanner")  one lane per stream, runs ONLY the serial recurrence over 32x32 tiles held in shared memory;
//   warps 1-15 ("movers") stream the tiles: coalesced global loads of the sliding-max values one tile ahead (held in
//                        registers across the barrier), thr/peak (IEEE division, :259) and the per-stream tile maximum
//                        computed on the way into shared memory, and the coalesced write-back of the gains one tile
//                        behind.  The tiles are double buffered, one __syncthreads per tile.
// The serial loop is branch-free and keeps every memory access off the dependent chain:
//   * both the attack and the release candidate are formed every step and selected;
//   * the curve values for the next two steps (acc[j+1], acc[j+2]) live in registers and acc[j+3] is fetched two steps
//     ahead (after a trigger the indices restart at 1, 2 - constants);
//   * peak and thr/peak are fetched together with one 64-bit shared load.
// Tiles in which a lane is idle and below threshold cost it nothing; a tile where that holds for all 32 lanes is
// answered without the walk.
constexpr int kScanAccSmem = 50 * 1024;   // floats of the curve kept in shared memory (200 KB, opt-in dynamic smem)
constexpr int kScanThreads = 512;      // 1 scanner warp + 15 mover warps
constexpr int kScanMovers = kScanThreads / 32 - 1;

template <bool ACC_SMEM>
static __global__ void __launch_bounds__(kScanThreads) k_limiter_scan(const __grid_constant__ KernelPlan plan, ScanArgs a) {
  __shared__ float2 t_in[2][32][33];   // {peak, thr/peak}
  __shared__ float t_g[2][32][33];
  __shared__ float t_max[2][32];
  __shared__ int s_lo[32], s_len[32];
  extern __shared__ float s_acc[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s0 = blockIdx.x * 32;
  if (a.gate && *a.gate == 0) return;
  const int ja = plan.lim_ja, jr = plan.lim_jr;
  const float thr = plan.lim_thr;
  if (ACC_SMEM)
    for (int i = threadIdx.x; i < jr + 4; i += kScanThreads) s_acc[i] = a.acc[i];   // jr + 4 entries, zero padded
  if (warp == 0) {
    const int s = s0 + lane;
    const bool live = s < a.n_streams;
    const int lo = live ? a.submit[s].sub_off[a.sub] : 0;
    s_lo[lane] = lo;
    s_len[lane] = live ? a.submit[s].sub_off[a.sub + 1] - lo : 0;
  }
  __syncthreads();
  int max_len = s_len[lane];
#pragma unroll
  for (int o = 16; o; o >>= 1) max_len = max(max_len, __shfl_xor_sync(0xffffffffu, max_len, o));
  const int n_tiles = (max_len + 31) / 32;
  auto acc_at = [&](int i) -> float { return ACC_SMEM ? s_acc[i] : __ldg(a.acc + i); };

  // scanner state
  int j = -1, len = 0;
  float start = -1.f, end = -1.f, a1 = 0.f, a2 = 0.f;
  if (warp == 0) {
    const int s = s0 + lane;
    len = s_len[lane];
    if (s < a.n_streams) { j = a.state[s].lim_j; start = a.state[s].lim_start; end = a.state[s].lim_end; }
    if (j > jr) j = jr;
    a1 = acc_at(1);
    a2 = acc_at(2);
  }
  // mover state: rows r = warp-1, warp-1+15, ... (3 or 2 rows per mover warp), column = lane
  constexpr int kRowsPerMover = (32 + kScanMovers - 1) / kScanMovers;
  float pre[kRowsPerMover];
  auto mover_load = [&](int t) {
#pragma unroll
    for (int q = 0; q < kRowsPerMover; ++q) {
      const int r = (warp - 1) + kScanMovers * q;
      const int k = t * 32 + lane;
      pre[q] = (r < 32 && s0 + r < a.n_streams && k < s_len[r])
                   ? a.wm[(size_t)(s0 + r) * a.cap + a.hist + s_lo[r] + k] : 0.f;
    }
  };
  if (warp > 0 && n_tiles > 0) mover_load(0);

  for (int t = 0; t < n_tiles + 2; ++t) {
    if (warp > 0) {
      if (t < n_tiles) {
        // publish tile t (loaded last iteration), then start the loads of tile t+1
        const int b = t & 1;
#pragma unroll
        for (int q = 0; q < kRowsPerMover; ++q) {
          const int r = (warp - 1) + kScanMovers * q;
          const float v = pre[q];
          const float e = thr / v;
          // the scanner only needs to know whether ANY instant of the row's tile can cross the threshold at gain 1
          const bool hot = __any_sync(0xffffffffu, v * 1.0f > thr);
          if (r < 32) {
            t_in[b][r][lane] = make_float2(v, e);
            if (lane == 0) t_max[b][r] = hot ? 1.0f : 0.0f;
          }
        }
        if (t + 1 < n_tiles) mover_load(t + 1);
      }
      if (t >= 2) {
        // write back the gains of tile t-2
        const int b = t & 1;
#pragma unroll
        for (int q = 0; q < kRowsPerMover; ++q) {
          const int r = (warp - 1) + kScanMovers * q;
          const int k = (t - 2) * 32 + lane;
          if (r < 32 && s0 + r < a.n_streams && k < s_len[r])
            a.gn[(size_t)(s0 + r) * a.cap + a.hist + s_lo[r] + k] = t_g[b][r][lane];
        }
      }
    } else if (t >= 1 && t <= n_tiles) {
      const int b = (t - 1) & 1;
      const int k0 = (t - 1) * 32;
      const bool idle = (j < 0 || j >= jr);
      const bool quiet = idle && t_max[b][lane] == 0.0f;
      if (!__all_sync(0xffffffffu, quiet)) {
        const int nk = min(32, len - k0);
        // Speculative form of compute_target_gain (:237-265).  cont(m) = the gain m steps ahead if no trigger fires
        // until then; it only depends on the state at the last trigger (start S, end E, D = S-E, R = 1-E) and on the
        // time index, so it is formed ahead of time.  Each step computes the candidates for "a trigger fires now"
        // (index restarts: curve values a1..a3 are constants) next to them and selects - the dependent chain per
        // sample is g -> g-e -> a1*(g-e) -> g-(..) -> select.
        float S = start, E = end, D = start - end, R = 1.0f - end;
        auto cont = [&](int jpre, float p) -> float {   // gain of the step whose pre-increment index is jpre
          const float ga = S - p * D;
          const float gr = E + p * R;
          return (jpre >= 0 && jpre < jr) ? (jpre < ja ? ga : gr) : 1.0f;
        };
        const int jc = min(max(j, 0), jr);
        float g = cont(j, acc_at(min(jc + 1, jr + 3)));                              // gain of the coming step
        float n1 = cont(j < 0 ? j : j + 1, acc_at(min(jc + 2, jr + 3)));             // one step later, no trigger
        float n2 = cont(j < 0 ? j : j + 2, acc_at(min(jc + 3, jr + 3)));             // two steps later, no triggers
        float pq = acc_at(min(jc + 4, jr + 3));                                      // curve value for index j+4
        const float a3 = acc_at(3), a4 = acc_at(4 < jr + 3 ? 4 : jr + 3);
        // j is the pre-increment index of the coming step; saturate so that j+4 stays inside the padded table
#pragma unroll 8
        for (int i = 0; i < nk; ++i) {
          const float2 in = t_in[b][lane][i];
          const int jl = min(max(j, 0), jr);
          const float pl = acc_at(min(jl + 5, jr + 3));       // used next iteration (if no trigger fires now)
          const bool trig = in.x * g > thr;
          const float d = g - in.y;
          const float c1 = g - a1 * d, c2 = g - a2 * d, c3 = g - a3 * d;
          // no-trigger continuation three steps ahead (pre-increment index j+3), from the current trigger state
          const float n3 = cont(j < 0 ? j : j + 3, pq);
          t_g[b][lane][i] = g;
          // state after this step
          S = trig ? g : S;
          E = trig ? in.y : E;
          D = trig ? d : D;
          R = trig ? (1.0f - in.y) : R;
          const bool active = (j >= 0) && (j < jr);
          j = trig ? 0 : (active ? j + 1 : j);
          g = trig ? c1 : n1;
          n1 = trig ? c2 : n2;
          n2 = trig ? c3 : n3;
          pq = trig ? a4 : pl;
        }
        start = S;
        end = E;
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) t_g[b][lane][i] = 1.0f;
      }
    }
    __syncthreads();
  }
  if (warp == 0 && s0 + lane < a.n_streams) {
    const int s = s0 + lane;
    a.state[s].lim_j = j; a.state[s].lim_start = start; a.state[s].lim_end = end;
  }
}

// -------------------------------------------------------------------------------------------------------------------
// K3c: apply the gain to the 240-sample delayed signal, quantise and interleave
// (audio_effect_peak_limiter.c:139-147 + iamf_decoder_plane2stride_out / FLOAT2INT*, IAMF_decoder.c:100-167).
// thread = one output instant of one stream (all channels); the first kLimDelay priming outputs of a stream are
// dropped (padsize, :185-201).  Without a limiter the time line is quantised as is.
// -------------------------------------------------------------------------------------------------------------------
struct OutputArgs {
  const float *tl;     // [S][C][cap]; sample of instant k sits at hist + k, the delayed one at hist + k - delay
  const float *gn;     // [S][cap] or nullptr
  const SubmitRec *submit;
  void *pcm;
  size_t stride_bytes; // per stream
  int cap, hist;
  int sub;             // sub-chunk processed by this launch
  int n_streams;
  const int *gate;     // non-null: return at once when *gate == 0
};

__device__ __forceinline__ int quant16(float x) {
  x = x * 32768.f;
  x = x > -32768.f ? x : -32768.f;
  x = x < 32767.f ? x : 32767.f;
  return __float2int_rn(x);
}
__device__ __forceinline__ int quant24(float x) {
  x = x * 8388608.f;
  x = x > -8388608.f ? x : -8388608.f;
  x = x < 8388607.f ? x : 8388607.f;
  return __float2int_rn(x);
}
__device__ __forceinline__ int quant32(float x) {
  x = x * 2147483648.f;
  x = x > -2147483648.f ? x : -2147483648.f;
  x = x < 2147483647.f ? x : 2147483647.f;   // 2147483647.f == 2^31: positive full scale wraps like the reference
  return (int)__float2ll_rn(x);
}

// thread = 4 consecutive output samples of one stream (all channels): 16-byte loads along time per channel, and the
// 4 x C_out quantised values of a thread form one contiguous byte range of the interleaved output, written with the
// widest aligned stores available.
template <int BITS>
__device__ __forceinline__ void store_sample(char *out, size_t idx, float x) {
  if (BITS == 16) {
    ((int16_t *)out)[idx] = (int16_t)quant16(x);
  } else if (BITS == 24) {
    int v = quant24(x);
    unsigned char *p = (unsigned char *)out + idx * 3;
    p[0] = v & 0xff;
    p[1] = (v >> 8) & 0xff;
    p[2] = ((v >> 16) & 0x7f) | ((v >> 24) & 0x80);
  } else if (BITS == 32) {
    ((int32_t *)out)[idx] = quant32(x);
  } else {
    ((float *)out)[idx] = x;
  }
}

template <int BITS>
__device__ __forceinline__ void k_output_stream(const KernelPlan &plan, const OutputArgs &a, int s);
template <int BITS>
static __global__ void __launch_bounds__(256) k_output(const __grid_constant__ KernelPlan plan, OutputArgs a) {
  if (a.gate && *a.gate == 0) return;
  for (int s = blockIdx.y; s < a.n_streams; s += gridDim.y) k_output_stream<BITS>(plan, a, s);
}
template <int BITS>
__device__ __forceinline__ void k_output_stream(const KernelPlan &plan, const OutputArgs &a, int s) {
  const SubmitRec sr = a.submit[s];
  const int lo = sr.sub_off[a.sub], hi = sr.sub_off[a.sub + 1];
  // instants are grouped in fours aligned on the time line (so that the float4 loads are aligned)
  const int k4 = ((lo >> 2) + blockIdx.x * blockDim.x + threadIdx.x) << 2;
  if (k4 >= hi) return;
  const int co = plan.out_channels;
  const int delay = plan.limiter ? kLimDelay : 0;
  const int kbeg = max(max(k4, lo), sr.out_skip), kend = min(k4 + 4, hi);
  if (kbeg >= kend) return;
  char *out = (char *)a.pcm + (size_t)s * a.stride_bytes;
  const bool full = (kbeg == k4) && (kend == k4 + 4) && ((a.cap & 3) == 0) && (((a.hist - delay) & 3) == 0);
  float g[4] = {1.f, 1.f, 1.f, 1.f};
  if (a.gn) {
    const float *gp = a.gn + (size_t)s * a.cap + a.hist + k4;
    if (full && ((a.hist & 3) == 0)) {
      float4 t = *reinterpret_cast<const float4 *>(gp);
      g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) if (k4 + u >= kbeg && k4 + u < kend) g[u] = gp[u];
    }
  }
  const float *base = a.tl + (size_t)s * co * a.cap + a.hist + k4 - delay;
  const size_t o0 = (size_t)(k4 - sr.out_skip);   // output sample index of instant k4
  if (full && BITS == 16 && (co & 1) == 0 && (((o0 * co) & 1) == 0)) {
    // fast path: channel pairs packed into 32-bit words
    for (int c = 0; c < co; c += 2) {
      const float4 x0 = *reinterpret_cast<const float4 *>(base + (size_t)c * a.cap);
      const float4 x1 = *reinterpret_cast<const float4 *>(base + (size_t)(c + 1) * a.cap);
      const float v0[4] = {x0.x, x0.y, x0.z, x0.w}, v1[4] = {x1.x, x1.y, x1.z, x1.w};
      uint32_t *w = (uint32_t *)((int16_t *)out + o0 * co + c);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float y0 = v0[u], y1 = v1[u];
        if (a.gn) { y0 = y0 * g[u]; y1 = y1 * g[u]; }
        w[(size_t)u * (co >> 1)] = (uint32_t)(quant16(y0) & 0xffff) | ((uint32_t)quant16(y1) << 16);
      }
    }
    return;
  }
  for (int c = 0; c < co; ++c) {
    float v[4];
    if (full) {
      const float4 x = *reinterpret_cast<const float4 *>(base + (size_t)c * a.cap);
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (k4 + u >= kbeg && k4 + u < kend) ? base[(size_t)c * a.cap + u] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (k4 + u < kbeg || k4 + u >= kend) continue;
      float x = v[u];
      if (a.gn) x = x * g[u];
      store_sample<BITS>(out, (o0 + u) * co + c, x);
    }
  }
}

// -------------------------------------------------------------------------------------------------------------------
// K4: carry the history (limiter delay line / peak ring / resampler memory) to the front of the time line for the
// next submit.  Tiny: S x C x hist floats.
// -------------------------------------------------------------------------------------------------------------------
struct CarryArgs {
  float *tl;          // [S][rows][cap]
  const SubmitRec *submit;
  int rows, cap, hist;
  int use_in_len;     // 1: advance by in_len (pre-resample line), 0: by lim_len
  const int *gate;    // non-null: return at once when *gate == 0
};

static __global__ void __launch_bounds__(256) k_carry(CarryArgs a) {
  __shared__ float tmp[256];
  const int s = blockIdx.y, r = blockIdx.x;
  if (a.gate && *a.gate == 0) return;
  const int len = a.use_in_len ? a.submit[s].in_len : a.submit[s].lim_len;
  if (len == 0) return;
  float *row = a.tl + ((size_t)s * a.rows + r) * a.cap;
  // hist <= 256: stage through shared memory because source and destination overlap when len < hist
  int i = threadIdx.x;
  if (i < a.hist) tmp[i] = row[len + i];
  __syncthreads();
  if (i < a.hist) row[i] = tmp[i];
}

}  // namespace iamfb
