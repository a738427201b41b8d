// iamfb_pipe_rs.cuh - k_pipe_rs: the pipelined per-stream kernel of the RESAMPLING pipelines (configs[4] of BASELINE.json:
// stereo 44.1 -> 48 kHz, loudness, limiter, 16 bit).  k_pipe's skeleton (iamfb_pipe.cuh) with a resampling stage between the
// render and the limiter, everything between the decoded frames and the PCM on chip:
//
//     workers   out(t-1) | time-line store(t) | [input tiles -> render -> pre-resample ring] | FIR(t+1) | look-ahead max(t+1)
//     scanner   gain recurrence of tile t,  or - limiter idle - out(t-1) itself
//
// Two time axes.  INPUT tiles are 240 consecutive instants of the stream's decoded time line (they may span two frames:
// rows staged by bulk copies, one or two pieces per row); they are rendered exactly like k_pipe's tiles and land in a ring
// of pre-resample samples in shared memory (channel pairs interleaved, the first `mirror` entries repeated behind the end so
// that a filter window never wraps).  OUTPUT tiles are 240 consecutive resampler outputs = one limiter window; before the
// FIR of an output tile the workers render input tiles until the ring covers the last tap of its last output.
//
// The FIR is the reference's (resampler_basic_interpolate_single, resample.c:357-418, closed-form tap positions of SURVEY
// 9.4-3): four accumulators per channel over j ascending against the 4 neighbouring taps of the 8x oversampled sinc table,
// cubic blend, clamp to +-1 (:84,959), loudness (IAMF_decoder.c:3480-3484).  A thread owns 4 consecutive outputs; the two
// channels of a pair share every tap and go through the packed FP32 instructions (exact product + separate add, see
// pipe_mac2): 10 instructions per tap and output for 16 multiply-adds.
//
// Streams with trimmed / missing frames and flushes take the multi-kernel path (iamfb_kernels.cuh); both keep their
// histories in the same places (heads of the batch's time lines), so a stream may change path between submits.
#pragma once
#include "iamfb_pipe.cuh"

namespace iamfb {

struct PipeRsArgs {
  PipeArgs p;                   // (p.hist_y / p.hist_pk: heads of tl_b / pk, strides below)
  int hist_y_stride;            // floats between two channels' history rows (cap_b)
  int hist_pk_stride;           // floats between two streams' peak rows (cap_b)
  float *hist_rs;               // [S][co][hist_rs_stride]: the last rs_hist pre-resample samples of every channel
  int hist_rs_stride;           // (cap_a)
  const float4 *tab4;           // [oversample][tab_row]: the four neighbouring taps per (offset, input), tab_pad zero items in
                                // front of input 0 and behind input filt_len - 1
  int tab_row, tab_pad;         // items per row (odd: lanes of different phases hit different banks), zero items per side
  const float4 *interp4;        // [den]: the cubic interpolation weights of every phase (cubic_coef, resample.c:246-256)
  int ring, mirror;             // ring entries (multiple of 4), entries repeated behind the end (>= filt_len, multiple of 4)
  int off_pkr, off_ring, off_tab, off_stage;   // byte offsets of the shared-memory areas behind the time line
  int smem_bytes;
  // PRE variant (the resampler ran in front, k_resample_ls): the resampled, clamped, loudness-scaled time line
  const float *tl_pre;          // [S][co][tl_pre_stride], output u of this submit at tl_pre_off + u
  int tl_pre_stride, tl_pre_off;
};

// k_pipe_prerender: the render stage of k_pipe_rs on its own (split form) - the regular streams' decoded frames (float32 or
// int16, straight from memory) through reconstruction, the compile-time matrix and the element / output mix gains onto the
// pre-resample time line tl_a[S][co][cap] at hist + f * N, one thread per 4 instants.  Same expressions as render_in below.
struct PreRenderArgs {
  PipeArgs p;
  float *tl;
  int cap, hist;
};

template <class SIG>
__global__ void __launch_bounds__(128) k_pipe_prerender(const __grid_constant__ KernelPlan plan, PreRenderArgs b) {
  typedef typename SIG::E0 E0;
  constexpr int VEC = 4, CO = SIG::CO, NY = SIG::NY;
  typedef Vec<VEC> V;
  const PipeArgs &a = b.p;
  const int N = plan.frame_size;
  const int f = blockIdx.y, s = blockIdx.z;
  const int i0 = (blockIdx.x * 128 + threadIdx.x) * VEC;
  if (a.submit[s].irregular) return;          // rendered by the multi-kernel path (block-uniform)
  const bool live = i0 < N;
  const int i0r = live ? i0 : 0;
  const ElPlan &ep0 = plan.el[0];
  const int nin = ep0.n_in;
  const FrameRec &fr = a.frames[(size_t)s * a.n_frames + f];
  const char *rows = reinterpret_cast<const char *>(a.in[0]) + (((size_t)s * a.n_frames + f) * nin) * (size_t)N * SIG::kEsz + (size_t)i0r * SIG::kEsz;
  const bool fade_w = __any_sync(0xffffffffu, i0r < plan.overlap);
  V y[NY];
#pragma unroll
  for (int r = 0; r < NY; ++r) y[r] = vzero<VEC>();
  pipe_render_element<SIG, E0, VEC, NY>(plan, ep0, fr.el[0], rows, N * SIG::kEsz, i0r, fade_w, a.start_win, a.stop_win, y, a.neg_zero);
  const float eg = fr.el[0].gain, og = fr.out_gain;
  if (eg != 1.f && eg > 0.f) pipe_scale<SIG, E0, 0, VEC, NY>(y, eg);
  if (og != 1.f && og > 0.f) {
#pragma unroll
    for (int r = 0; r < NY; ++r)
#pragma unroll
      for (int q = 0; q < VEC; ++q) y[r].v[q] *= og;
  }
  if (!live) return;
#pragma unroll 1
  for (int c = 0; c < CO; ++c) {
    const int r = pipe_yrow_rt<SIG>(c);
    if (r < 0) continue;
    V v = y[0];
#pragma unroll
    for (int rr = 1; rr < NY; ++rr)
      if (rr == r) v = y[rr];
    float *dst = b.tl + ((size_t)s * CO + c) * b.cap + b.hist + (size_t)f * N + i0;
    __stcg(reinterpret_cast<float4 *>(dst), make_float4(v.v[0], v.v[1], v.v[2], v.v[3]));
  }
}

// PRE = true: the limiter half only - the resampler's outputs come from memory (tl_pre) instead of the ring + FIR
// output channel that owns row r of the time line (inverse of SIG::yrow)
template <class SIG>
__host__ __device__ constexpr int pipe_rs_row_channel(int r) {
  for (int c = 0; c < SIG::CO; ++c)
    if (SIG::yrow(c) == r) return c;
  return 0;
}

// f(row, channel) for every row of the time line, the channel a compile-time constant (a constexpr function called in a
// run-time expression would be evaluated on the device, where the tables it reads do not exist)
template <class SIG, int R, class F>
__device__ __forceinline__ void pipe_rs_rows(F &&f) {
  if constexpr (R < SIG::NY) {
    constexpr int c = pipe_rs_row_channel<SIG>(R);
    f(R, c);
    pipe_rs_rows<SIG, R + 1>(f);
  }
}

// resident blocks per SM the limiter half is compiled for: it holds no ring, table or stages, and with 14 blocks of 96
// threads a batch of 2048 streams is ONE wave (measured: 0.304 -> 0.265 ms per submit of configuration 5, spills included)
template <class SIG>
constexpr int pipe_rs_pre_minb() { return 2048 / SIG::kThreads < 14 ? 2048 / SIG::kThreads : 14; }
template <class SIG, bool PRE = false>
__global__ void __launch_bounds__(SIG::kThreads, PRE ? pipe_rs_pre_minb<SIG>() : SIG::kMinBlocks) k_pipe_rs(const __grid_constant__ KernelPlan plan, PipeRsArgs b) {
  typedef typename SIG::E0 E0;
  constexpr int VEC = SIG::VEC, NW = SIG::NW, CO = SIG::CO, NY = SIG::NY, WN = SIG::kWorkers, NS = SIG::kStages;
  constexpr int TL = kStreamTile, NP = (NY + 1) / 2;
  static_assert(!SIG::kTwo, "k_pipe_rs: one element");
  typedef Vec<VEC> V;
  const PipeArgs &a = b.p;
  extern __shared__ __align__(128) float fsm[];
  __shared__ __align__(8) uint64_t s_bar[NS];
  __shared__ __align__(8) uint64_t s_hbar;
  __shared__ __align__(16) FrameRec s_fr[2];
  __shared__ __align__(16) float s_es[2][32];
  __shared__ float s_acc[kStreamAccCache];
  __shared__ int s_hot[2][NW], s_apply[2];
  __shared__ float s_tot[NW];
  __shared__ int s_skip;
  const ElPlan &ep0 = plan.el[0];
  const int nin = ep0.n_in;
  float *Y = fsm;                    // [NY][2][TL]
  float *WM = Y + NY * 2 * TL;       // [2][TL]
  float *G = WM + 2 * TL;            // [2][TL]
  float *SA = G + 2 * TL;            // [2][TL]
  float *PKR = reinterpret_cast<float *>(reinterpret_cast<char *>(fsm) + b.off_pkr);     // [2][TL] peaks of the rendered output tiles
  float2 *RING = reinterpret_cast<float2 *>(reinterpret_cast<char *>(fsm) + b.off_ring); // [NP][ring + mirror] channel pairs
  const float4 *TAB = reinterpret_cast<const float4 *>(reinterpret_cast<char *>(fsm) + b.off_tab);   // [os][tab_row]
  char *ST = reinterpret_cast<char *>(fsm) + b.off_stage;
  const int s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const bool worker = tid < WN;
  const int N = plan.frame_size;
  const float thr = plan.lim_thr;
  const int bits = plan.bit_depth;
  const bool limiter = plan.limiter != 0;
  const int Nf = (int)plan.rs_filt_len, os = (int)plan.rs_oversample, den = (int)plan.rs_den;
  const int RH = plan.rs_hist;       // pre-resample history in front of this submit's first input instant
  const int RG = b.ring, RGM = b.ring + b.mirror;

  // ---- what the previous submit left: limiter delay line, peaks of its last window, resampler history; and the tap table
  if (tid == 0) {
    mbar_init(&s_hbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    uint32_t bytes = PRE ? 0u : (uint32_t)(os * b.tab_row * sizeof(float4));
    if (limiter) bytes += (uint32_t)((NY + 1) * kLimDelay * sizeof(float));
    mbar_expect_tx(&s_hbar, bytes);
    if (!PRE) bulk_g2s(const_cast<float4 *>(TAB), b.tab4, (uint32_t)(os * b.tab_row * sizeof(float4)), &s_hbar);
    if (limiter) {
#pragma unroll 1
      for (int c = 0; c < CO; ++c) {
        const int r = pipe_yrow_rt<SIG>(c);
        if (r >= 0) bulk_g2s(Y + (r * 2 + 1) * TL, a.hist_y + ((size_t)s * CO + c) * b.hist_y_stride, (uint32_t)(kLimDelay * sizeof(float)), &s_hbar);
      }
      bulk_g2s(PKR + TL, a.hist_pk + (size_t)s * b.hist_pk_stride, (uint32_t)(kLimDelay * sizeof(float)), &s_hbar);
    }
  }
  for (int i = tid; i < kStreamAccCache; i += SIG::kThreads) s_acc[i] = (limiter && i <= plan.lim_jr + 3) ? a.acc[i] : 0.f;
  // ring and stages start from zeros (whatever the filter's zero taps meet must be finite)
  if constexpr (!PRE) {
  for (int i = tid; i < NP * RGM; i += SIG::kThreads)
    if (i % RGM >= RH && !(i % RGM >= RG && i % RGM - RG < RH)) RING[i] = make_float2(0.f, 0.f);
  for (int i = tid; i < NS * a.stage_bytes / 4; i += SIG::kThreads) reinterpret_cast<float *>(ST)[i] = 0.f;
  // resampler history -> ring entries [0, RH) (pair-interleaved; mirrored below)
  for (int i = tid; i < NP * RH; i += SIG::kThreads) {
    const int p = i / RH, k = i - p * RH;
    float2 v = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int r = 2 * p + h;
      if (r >= NY) continue;
      int c = 0;
      for (int cc = 0; cc < CO; ++cc)
        if (pipe_yrow_rt<SIG>(cc) == r) c = cc;
      const float x = b.hist_rs[((size_t)s * CO + c) * b.hist_rs_stride + k];
      if (h == 0) v.x = x; else v.y = x;
    }
    RING[p * RGM + k] = v;
    if (k < b.mirror) RING[p * RGM + RG + k] = v;
  }
  }
  if constexpr (!PRE) asm volatile("griddepcontrol.wait;" ::: "memory");   // (PRE: launched behind finished kernels)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __syncthreads();
  mbar_wait(&s_hbar, 0u);
  if (limiter && tid < 32) {
    // suffix maxima of the peaks of the tile before this submit (tile -1, slot 1)
    const float *src = PKR + TL;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = lane < 30 ? src[8 * lane + i] : 0.f;
#pragma unroll
    for (int i = 6; i >= 0; --i) v[i] = fmaxf(v[i], v[i + 1]);
    float m = v[0];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float o = __shfl_down_sync(0xffffffffu, m, d);
      if (lane + d < 32) m = fmaxf(m, o);
    }
    float ex = __shfl_down_sync(0xffffffffu, m, 1);
    if (lane == 31) ex = 0.f;
    if (lane < 30) {
#pragma unroll
      for (int i = 0; i < 8; ++i) SA[TL + 8 * lane + i] = fmaxf(v[i], ex);
    }
  }
  const SubmitRec sr = a.submit[s];
  if (sr.irregular) return;            // rendered by the multi-kernel path right after (block-uniform)
  const int in_len = sr.in_len;        // input instants of this submit (= n_frames * N for a regular stream)
  const int L = sr.lim_len;            // resampler outputs of this submit = instants entering the limiter
  const int T = (L + TL - 1) / TL;     // output tiles
  const int K = (in_len + TL - 1) / TL;   // input tiles
  if (tid == 0) {
#pragma unroll
    for (int w = 0; w < NW; ++w) s_hot[0][w] = s_hot[1][w] = 0;
    s_apply[0] = s_apply[1] = 0;
    s_skip = sr.out_skip;
#pragma unroll
    for (int q = 0; q < NS; ++q) mbar_init(&s_bar[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // stream position of this submit's first input instant, first output: everything below is relative to them
  const long long in_start = a.state[s].rs_in_total - in_len;
  // first tap of output u (relative to this submit's first input instant) and its phase, for u = 0: carried tile by tile
  //   q(n) = Nf/2 + floor(n * num / den) is the LAST input of output n
  long long n0 = sr.rs_out_first;
  int base_rel = (int)((long long)(Nf / 2) + (n0 * (long long)plan.rs_num) / den - (Nf - 1) - in_start);
  int base_phi = (int)((n0 * (long long)plan.rs_frac_adv) % den);

  const int q0 = VEC * tid;
  const bool has = worker && q0 < TL;
  const int q0r = has ? q0 : TL - VEC;

  // ---- input side
  // tile k = input instants [240 k, 240 k + 240) of this submit: the rows of one frame, or the tail of one and the head
  // of the next; lane r of warp 0 issues row r's piece(s), thread 0 announces the bytes (and brings the records of the
  // frames that start inside the tile into the slot of their parity)
  auto issue_in = [&](int k) {
    if (tid < 32) {
      int s_it = s;
      asm volatile("" : "+r"(s_it));
      const int rel0 = k * TL;
      const int len = min(TL, in_len - rel0);
      const int f0 = rel0 / N, off0 = rel0 - f0 * N;
      const int lenA = min(len, N - off0), lenB = len - lenA;
      char *st = ST + (k % NS) * a.stage_bytes;
      uint64_t *bar = &s_bar[k % NS];
      const int esz = SIG::kEsz;
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        uint32_t bytes = (uint32_t)(len * esz * nin);
        if (off0 == 0) bytes += (uint32_t)sizeof(FrameRec);
        if (lenB > 0) bytes += (uint32_t)sizeof(FrameRec);
        mbar_expect_tx(bar, bytes);
        const FrameRec *fr = a.frames + (size_t)s_it * a.n_frames;
        if (off0 == 0) bulk_g2s(&s_fr[f0 & 1], fr + f0, (uint32_t)sizeof(FrameRec), bar);
        if (lenB > 0) bulk_g2s(&s_fr[(f0 + 1) & 1], fr + f0 + 1, (uint32_t)sizeof(FrameRec), bar);
      }
      __syncwarp();
      for (int r = lane; r < nin; r += 32) {
        const char *g = reinterpret_cast<const char *>(a.in[0]) + (((size_t)s_it * a.n_frames + f0) * nin + r) * (size_t)N * esz;
        bulk_g2s(st + r * a.row_bytes, g + (size_t)off0 * esz, (uint32_t)(lenA * esz), bar);
        if (lenB > 0) bulk_g2s(st + r * a.row_bytes + lenA * esz, g + (size_t)nin * N * esz, (uint32_t)(lenB * esz), bar);
      }
    }
  };
  // renders input tile k (k_pipe's render: reconstruction, matrix, element / output mix gains; NOT the loudness, which
  // follows the resampler) into the ring
  auto render_in = [&](int k) {
    const int rel = k * TL + q0r;
    const int f0 = (k * TL) / N;
    const int f = f0 + ((rel - f0 * N) >= N ? 1 : 0);
    const int i0 = rel - f * N;
    const char *st = ST + (k % NS) * a.stage_bytes;
    const FrameRec &fr = s_fr[f & 1];
    const char *rows = st + q0r * SIG::kEsz;
    const bool fade_w = __any_sync(0xffffffffu, i0 < plan.overlap);
    V y[NY];
#pragma unroll
    for (int r = 0; r < NY; ++r) y[r] = vzero<VEC>();
    pipe_render_element<SIG, E0, VEC, NY>(plan, ep0, fr.el[0], rows, a.row_bytes, i0, fade_w, a.start_win, a.stop_win, y, a.neg_zero);
    const float eg = fr.el[0].gain, og = fr.out_gain;
    if (eg != 1.f && eg > 0.f) pipe_scale<SIG, E0, 0, VEC, NY>(y, eg);
    if (og != 1.f && og > 0.f) {
#pragma unroll
      for (int r = 0; r < NY; ++r)
#pragma unroll
        for (int q = 0; q < VEC; ++q) y[r].v[q] *= og;
    }
    if (has) {
      int idx = (rel + RH) % RG;           // (a multiple of VEC: a thread's instants never wrap)
#pragma unroll
      for (int p = 0; p < NP; ++p) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          const float2 v = make_float2(y[2 * p].v[q], (2 * p + 1 < NY) ? y[(2 * p + 1 < NY) ? 2 * p + 1 : 0].v[q] : 0.f);
          RING[p * RGM + idx + q] = v;
          if (idx + q < b.mirror) RING[p * RGM + RG + idx + q] = v;
        }
      }
    }
  };

  V yh[NY];
  V pkh = vzero<VEC>();
#pragma unroll
  for (int r = 0; r < NY; ++r) yh[r] = vzero<VEC>();

  // ---- resampler: output tile tau, outputs u = 240 tau + q0 + k of this thread
  auto fir = [&](int tau, int t_rel, int t_phi) {
    // first tap / phase of the thread's first output from the tile's (t_rel, t_phi): phi advances by frac_adv per output
    const int fa = plan.rs_frac_adv, ia = plan.rs_int_adv;
    int phi[VEC], dk[VEC], st0;
    {
      const int adv = t_phi + q0r * fa;        // < den + 240 * den: fits
      int carry = adv / den;
      int ph = adv - carry * den;
      int pos = t_rel + q0r * ia + carry;
      st0 = (pos + RH) % RG;                   // ring entry of the first output's first tap (windows never wrap: mirror)
      const int pos0 = pos;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        phi[k] = ph;
        dk[k] = pos - pos0;                    // 0 .. (VEC - 1) * (int_adv + 1): how far output k's window starts behind output 0's
        ph += fa; pos += ia;
        if (ph >= den) { ph -= den; pos += 1; }
      }
    }
    const int u0 = tau * TL + q0r;
#pragma unroll
    for (int r = 0; r < NY; ++r) yh[r] = vzero<VEC>();
    const bool loud_on = plan.loud_gain != 0.f && plan.loud_gain != 1.0f;
    // One pass over the inputs the thread's outputs span: input i (from the first output's first tap) is loaded ONCE and
    // meets tap i - dk[k] of output k.  The tap rows carry `pad` zero items in front and behind, so every output takes part
    // in every step: x * 0 adds +-0 to an accumulator that started at +0 - the sums are the reference's sums over
    // j = 0 .. Nf-1 in its order, bit for bit (inputs are finite: stages and ring are cleared when the kernel starts).
    const int trow = b.tab_row, pad = b.tab_pad;
    const int steps = Nf + (VEC - 1) * (ia + 1);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      float acc[VEC][4][2];
#pragma unroll
      for (int k = 0; k < VEC; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[k][q][0] = acc[k][q][1] = 0.f;
      const float4 *tp[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) tp[k] = TAB + (phi[k] * os / den) * trow + pad - dk[k];
      const float2 *xp = RING + p * RGM + st0;
#pragma unroll 4
      for (int i = 0; i < steps; ++i) {
        const float2 x = xp[i];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float4 t = tp[k][i];
          pipe_mac2(acc[k][0][0], acc[k][0][1], x.x, x.y, t.x, a.neg_zero);
          pipe_mac2(acc[k][1][0], acc[k][1][1], x.x, x.y, t.y, a.neg_zero);
          pipe_mac2(acc[k][2][0], acc[k][2][1], x.x, x.y, t.z, a.neg_zero);
          pipe_mac2(acc[k][3][0], acc[k][3][1], x.x, x.y, t.w, a.neg_zero);
        }
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        // cubic_coef of the output's phase (resample.c:246-256), tabulated per phase at plan time with the same operations
        const float4 ci = __ldg(b.interp4 + phi[k]);
        const float interp[4] = {ci.x, ci.y, ci.z, ci.w};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (2 * p + h >= NY) continue;
          float sum = interp[0] * acc[k][0][h] + interp[1] * acc[k][1][h] + interp[2] * acc[k][2][h] + interp[3] * acc[k][3][h];
          sum = (sum < -1.0f) ? -1.0f : ((sum > 1.0f) ? 1.0f : sum);      // FLTADJUST, resample.c:84
          if (loud_on) sum *= plan.loud_gain;
          if (u0 + k >= L) sum = 0.f;                                      // beyond the submit's last output (ragged tile)
          yh[(2 * p + h < NY) ? 2 * p + h : 0].v[k] = sum;
        }
      }
    }
    V peak = vzero<VEC>();
#pragma unroll
    for (int r = 0; r < NY; ++r)
#pragma unroll
      for (int k = 0; k < VEC; ++k) peak.v[k] = fmaxf(peak.v[k], fabsf(yh[r].v[k]));
#pragma unroll
    for (int k = 0; k < VEC; ++k) pkh.v[k] = has ? peak.v[k] : 0.f;
    if (limiter && has) stsv<VEC>(PKR + (tau & 1) * TL + q0, peak);
  };

  // PRE: output tile tau of the resampled time line -> yh (loaded one tile ahead into yn: the load's latency hides behind
  // the previous tile's limiter work), peaks as above
  V yn[NY];
#pragma unroll
  for (int r = 0; r < NY; ++r) yn[r] = vzero<VEC>();
  auto pre_load = [&](int tau) {
    const int u0 = tau * TL + q0r;
    const bool vec = ((b.tl_pre_stride | b.tl_pre_off) & 3) == 0;
    // (row -> channel is a compile-time map: the loads land in yn[] without a dependent select, so that their latency hides
    // behind the limiter work of the tile in between)
    pipe_rs_rows<SIG, 0>([&](int r, int c) {
      const float *src = b.tl_pre + ((size_t)s * CO + c) * b.tl_pre_stride + b.tl_pre_off + u0;
      bool done = false;
      if (u0 >= L) {                 // (a ragged last tile: nothing is read behind the submit's last output)
        yn[r] = vzero<VEC>();
        done = true;
      }
      if constexpr (VEC == 4) {
        if (vec && !done) {
          const float4 f = __ldcg(reinterpret_cast<const float4 *>(src));
          yn[r].v[0] = f.x; yn[r].v[1] = f.y; yn[r].v[2] = f.z; yn[r].v[3] = f.w;
          done = true;
        }
      }
      if (!done) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) yn[r].v[k] = __ldcg(src + k);
      }
    });
  };
  auto pre_take = [&](int tau) {
    const int u0 = tau * TL + q0r;
    V peak = vzero<VEC>();
#pragma unroll
    for (int r = 0; r < NY; ++r)
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float v = (u0 + k < L) ? yn[r].v[k] : 0.f;        // beyond the submit's last output (ragged tile)
        yh[r].v[k] = v;
        peak.v[k] = fmaxf(peak.v[k], fabsf(v));
      }
#pragma unroll
    for (int k = 0; k < VEC; ++k) pkh.v[k] = has ? peak.v[k] : 0.f;
    if (limiter && has) stsv<VEC>(PKR + (tau & 1) * TL + q0, peak);
  };

  float pre[VEC], suf[VEC];
  auto wmax_scan = [&]() {
    float inc[VEC];
    inc[0] = pkh.v[0];
#pragma unroll
    for (int k = 1; k < VEC; ++k) inc[k] = fmaxf(inc[k - 1], pkh.v[k]);
    float sfx[VEC];
    sfx[VEC - 1] = pkh.v[VEC - 1];
#pragma unroll
    for (int k = VEC - 2; k >= 1; --k) sfx[k] = fmaxf(sfx[k + 1], pkh.v[k]);
    const float tot = inc[VEC - 1];
    sfx[0] = tot;
    float up = tot, dn = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float o = __shfl_up_sync(0xffffffffu, up, d);
      if (lane >= d) up = fmaxf(up, o);
      const float q = __shfl_down_sync(0xffffffffu, dn, d);
      if (lane + d < 32) dn = fmaxf(dn, q);
    }
    float ex = __shfl_up_sync(0xffffffffu, up, 1);
    if (lane == 0) ex = 0.f;
    float sx = __shfl_down_sync(0xffffffffu, dn, 1);
    if (lane == 31) sx = 0.f;
    pre[0] = ex;
#pragma unroll
    for (int k = 1; k < VEC; ++k) pre[k] = fmaxf(ex, inc[k - 1]);
#pragma unroll
    for (int k = 0; k < VEC; ++k) suf[k] = fmaxf(sfx[k], sx);
    if (lane == 31) s_tot[tid >> 5] = up;
  };
  auto wmax_combine = [&](int t) {
    const int bb = t & 1;
    const int wi = tid >> 5;
    float cp = 0.f, cs = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float o = s_tot[w];
      if (w < wi) cp = fmaxf(cp, o);
      if (w > wi) cs = fmaxf(cs, o);
    }
    int hot = 0;
    if (has) {
      const V A = ldsv<VEC>(SA + (bb ^ 1) * TL + q0);
      V W, Sx;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        W.v[k] = fmaxf(A.v[k], fmaxf(pre[k], cp));
        Sx.v[k] = fmaxf(suf[k], cs);
        hot |= (W.v[k] > thr) ? 1 : 0;
      }
      stsv<VEC>(WM + bb * TL + q0, W);
      stsv<VEC>(SA + bb * TL + q0, Sx);
    }
    const int any_hot = __any_sync(0xffffffffu, hot);
    if (lane == 0) s_hot[bb][wi] = any_hot;
  };
  // instants of output tile t that exist (the submit's last tile may be ragged)
  auto tile_len = [&](int t) { return min(TL, L - t * TL); };
  auto output_and_store = [&](int t, bool do_out, bool do_store) {
    const int bb = t & 1;
    float *yt = Y + (bb ^ 1) * TL + q0;
    if (limiter) {
      const int o0 = t * TL + q0 - *(volatile int *)&s_skip;
      const int nv = min(VEC, tile_len(t) - q0);
      if (do_out && has && o0 >= 0 && nv > 0 && s_apply[bb] != 0) {
        int s_it = s;
        asm volatile("" : "+r"(s_it));
        pipe_emit<SIG>(yt, G + bb * TL + q0, (char *)a.pcm + (size_t)s_it * a.stride_bytes, o0, nv, bits);
      }
    }
    if (do_store && has) {
#pragma unroll
      for (int r = 0; r < NY; ++r) stsv<VEC>(yt + r * 2 * TL, yh[r]);
    }
  };
  auto quiet_output = [&](int t) {
    const int bb = t & 1;
    const int skip = *(volatile int *)&s_skip;
    const int n = tile_len(t);
    int s_it = s;
    asm volatile("" : "+r"(s_it));
    char *out = (char *)a.pcm + (size_t)s_it * a.stride_bytes;
#pragma unroll 1
    for (int qd = lane; qd < TL / VEC; qd += 32) {
      const int o0 = t * TL + VEC * qd - skip;
      const int nv = min(VEC, n - VEC * qd);
      if (o0 < 0 || nv <= 0) continue;
      pipe_emit<SIG>(Y + (bb ^ 1) * TL + VEC * qd, nullptr, out, o0, nv, bits);
    }
  };

  __syncthreads();

  if (worker) {
    if constexpr (!PRE) {
#pragma unroll
      for (int q = 0; q < NS; ++q)
        if (q < K) issue_in(q);
    } else if (T > 0) {
      pre_load(0);
    }
    int filled = 0;                  // input tiles rendered into the ring so far
#pragma unroll 1
    for (int t = -1; t <= T; ++t) {
      if (t >= 0) output_and_store(t - 1, t >= 1, t < T);
      if constexpr (PRE) {
        if (t + 1 < T) {
          const int tau = t + 1;
          pre_take(tau);
          if (tau + 1 < T) pre_load(tau + 1);
          if (limiter) wmax_scan();
          asm volatile("bar.sync 1, %0;" ::"n"(WN) : "memory");
          if (limiter) wmax_combine(tau);
        }
      } else
      if (t + 1 < T) {
        const int tau = t + 1;
        // the tile's (first tap, phase) and the input its last output reaches
        const int ulast = min(TL, L - tau * TL) - 1;
        int need;
        {
          const long long adv = (long long)base_phi + (long long)ulast * plan.rs_frac_adv;
          need = base_rel + ulast * plan.rs_int_adv + (int)(adv / den) + Nf;      // input instants of this submit needed
        }
        const int kneed = min(K, (need + TL - 1) / TL);
        while (filled < kneed) {     // (block-uniform)
          mbar_wait(&s_bar[filled % NS], (uint32_t)(filled / NS) & 1u);
          render_in(filled);
          asm volatile("bar.sync 1, %0;" ::"n"(WN) : "memory");
          if (filled + NS < K) issue_in(filled + NS);
          ++filled;
        }
        fir(tau, base_rel, base_phi);
        {   // advance the tile base by 240 outputs
          const long long adv = (long long)base_phi + (long long)TL * plan.rs_frac_adv;
          const int carry = (int)(adv / den);
          base_phi = (int)(adv - (long long)carry * den);
          base_rel += TL * plan.rs_int_adv + carry;
        }
        if (limiter) wmax_scan();
        asm volatile("bar.sync 1, %0;" ::"n"(WN) : "memory");
        if (limiter) wmax_combine(tau);
      }
      if (!limiter && t >= 0 && t < T && has) {
        const int nv = min(VEC, tile_len(t) - q0);
        if (nv > 0) {
          int s_it = s;
          asm volatile("" : "+r"(s_it));
          pipe_emit<SIG>(Y + (t & 1) * TL + q0, nullptr, (char *)a.pcm + (size_t)s_it * a.stride_bytes, (long long)t * TL + q0, nv, bits);
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(SIG::kThreads) : "memory");
    }
    // input tiles no output of this submit reached yet (their samples are the next submit's history)
    while (!PRE && filled < K) {
      mbar_wait(&s_bar[filled % NS], (uint32_t)(filled / NS) & 1u);
      render_in(filled);
      asm volatile("bar.sync 1, %0;" ::"n"(WN) : "memory");
      if (filled + NS < K) issue_in(filled + NS);
      ++filled;
    }
  } else {
    int lj = -1, lS_i = 0, lE_i = 0;
    if (limiter) {
      const StreamState &st = a.state[s];
      lj = st.lim_j; lS_i = __float_as_int(st.lim_start); lE_i = __float_as_int(st.lim_end);
      if (lj > plan.lim_jr) lj = plan.lim_jr;
    }
    float lS = __int_as_float(lS_i), lE = __int_as_float(lE_i);
    bool in_run = false;
#pragma unroll 1
    for (int t = -1; t <= T; ++t) {
      if (limiter && t >= 0 && t < T) {
        const int bb = t & 1;
        const bool idle = lj < 0 || lj >= plan.lim_jr;
        int hot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) hot |= s_hot[bb][w];
        const bool run = hot != 0 || !idle;
        if (run) stream_scan(WM + bb * TL, G + bb * TL, &s_es[0][0], tile_len(t), lj, lS, lE, in_run, a.acc, s_acc, plan.lim_ja, plan.lim_jr, thr, lane);
        else {
          in_run = false;
          quiet_output(t);
        }
        if (lane == 0) s_apply[bb] = run ? 1 : 0;
      }
      asm volatile("bar.sync 2, %0;" ::"n"(SIG::kThreads) : "memory");
    }
    if (limiter && lane == 0) {
      StreamState &st = a.state[s];
      st.lim_j = lj; st.lim_start = lS; st.lim_end = lE;
    }
  }
  __syncthreads();
  // ---- what the next submit starts from
  // the last 240 limiter instants [L - 240, L): instant j sits in slot (j div 240) & 1 at offset j mod 240 (tile -1 = slot 1)
  if (limiter) {
    for (int i = tid; i < kLimDelay; i += SIG::kThreads) {
      const int j = L - kLimDelay + i;
      const int tile = j >= 0 ? j / TL : -1, off = j - tile * TL;
#pragma unroll 1
      for (int c = 0; c < CO; ++c) {
        const int r = pipe_yrow_rt<SIG>(c);
        if (r < 0) continue;
        a.hist_y[((size_t)s * CO + c) * b.hist_y_stride + i] = Y[(r * 2 + (tile & 1)) * TL + off];
      }
      a.hist_pk[(size_t)s * b.hist_pk_stride + i] = PKR[(tile & 1) * TL + off];
    }
  }
  // the last RH pre-resample samples: PRE - from the tail of this submit's inputs on the pre-resample time line to its head
  // (in_len >= one frame >= 240 > RH: source and destination do not overlap; the resampler in front has finished)
  if constexpr (PRE) {
    for (int i = tid; i < CO * RH; i += SIG::kThreads) {
      const int c = i / RH, k = i - c * RH;
      float *row = b.hist_rs + ((size_t)s * CO + c) * b.hist_rs_stride;
      row[k] = row[in_len + k];
    }
  }
  // otherwise the ring entries of the input instants [in_len - RH, in_len)
  if constexpr (!PRE)
  for (int i = tid; i < NP * RH; i += SIG::kThreads) {
    const int p = i / RH, k = i - p * RH;
    const float2 v = RING[p * RGM + (in_len + k) % RG];      // instant in_len - RH + k sits at entry (in_len - RH + k + RH) mod RG
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int r = 2 * p + h;
      if (r >= NY) continue;
      int c = 0;
      for (int cc = 0; cc < CO; ++cc)
        if (pipe_yrow_rt<SIG>(cc) == r) c = cc;
      b.hist_rs[((size_t)s * CO + c) * b.hist_rs_stride + k] = h == 0 ? v.x : v.y;
    }
  }
}

}  // namespace iamfb
