// iamfb_pipe.cuh - k_pipe: the software-pipelined per-stream kernel of every pipeline signature it is instantiated for
// (channel-based, scene-based and two-element mixes; int16 or float32 decoded input; any output depth).
//
// One thread block owns one stream for the whole submit: NW worker warps (VEC consecutive instants of every channel per
// thread) and one scanner warp.  A tile is exactly one limiter window (240 instants); per tile the two sides meet at one
// block-wide barrier:
//
//     workers   out(t-1) if limited | time-line store(t) | render(t+1) | look-ahead max(t+1)      copies of tiles t+2, t+3 in flight
//     scanner   gain recurrence of tile t,  or - limiter idle - out(t-1) itself
//     ------------------------------------ bar.sync ------------------------------------       once per 240 instants
//
// What is new against k_stream (iamfb_stream.cuh), whose stages and arithmetic it keeps expression by expression:
//   * the decoded rows are staged as the int16 the codec produced (IAMFB_IN_S16; x / 32768 is exact and happens on the way
//     into the registers), which halves the stage - so TWO stages fit where one float32 stage did and the copy of a tile
//     is in flight for a whole iteration instead of the ~1 us between "stage free" and "stage needed";
//   * every element kind: channel-based (de-mixing chain + compile-time channel->channel matrix), scene-based (compile-time
//     HOA->loudspeaker matrix incl. the LFE slot shift of h2m_rdr.c:1114-1150) and two elements summed (iamf_mixer_mix);
//   * output channels no matrix row ever writes (LFE slots of the HOA tables, row 23 of sound system H) take no space on
//     the time line;
//   * 16-, 24-, 32-bit and float output; worker warps / instants per thread chosen per signature (a 24-channel time line
//     leaves room for 3-4 streams per SM: more, lighter threads per stream keep the SM's issue slots busy).
// Streams with trimmed or missing frames, flushes and animated gains still take k_fused (they share the limiter history).
#pragma once
#include <cuda.h>

#include "iamfb_stream.cuh"

namespace iamfb {

constexpr int kPipeTargetCh[IAMFB_TARGET_COUNT] = {2, 6, 8, 10, 11, 12, 14, 24, 8, 12, 10, 6, 1, 2};

constexpr int h2m_find(int order, int out) {
  for (int i = 0; i < (int)(sizeof(k_h2m_index) / sizeof(k_h2m_index[0])); ++i)
    if (k_h2m_index[i].order == order && k_h2m_index[i].out == out) return i;
  return -1;
}
// matrix row that lands on output channel oc after the LFE slot shift (h2m_rdr.c:1114-1135), -1 for the slots the
// reference zeroes (:1137-1150) or never writes
constexpr int h2m_row_of_out(int idx, int oc) {
  const int n = k_h2m_index[idx].n, l1 = k_h2m_index[idx].lfe1, l2 = k_h2m_index[idx].lfe2;
  if (oc == l1 || oc == l2) return -1;
  int k = 0;
  for (int i = 0; i < n; ++i) {
    if (l1 >= 0 || l2 >= 0) {
      if (l1 == i) k++;
      if (l2 == i) k++;
    }
    if (k == oc) return i;
    k++;
  }
  return -1;
}

// compile-time view of one element's render matrix: coef(m, oc) = contribution of renderer input m to output channel oc
template <int L, int NREC, int TARGET>
struct PipeEl {
  static constexpr bool kScene = L < 0;
  static constexpr int kN = NREC;
  static constexpr int kLayout = L;
  static constexpr int kOrder = NREC == 1 ? 0 : (NREC == 4 ? 1 : (NREC == 9 ? 2 : 3));
  static constexpr int kIdx = kScene ? h2m_find(kOrder, TARGET) : m2m_find(L, TARGET);
  static constexpr int CO = kPipeTargetCh[TARGET];
  static constexpr uint32_t coef(int m, int oc) {
    if (kScene) {
      const int n = h2m_row_of_out(kIdx, oc);
      return n < 0 ? 0u : k_matrix_pool[k_h2m_index[kIdx].off + n * NREC + m];
    }
    return k_matrix_pool[k_m2m_index[kIdx].off + m * CO + oc];
  }
  static constexpr bool nz(int m, int oc) {
    const uint32_t b = coef(m, oc);
    return b != 0u && b != 0x80000000u;
  }
  static constexpr bool any(int oc) {
    for (int m = 0; m < NREC; ++m)
      if (nz(m, oc)) return true;
    return false;
  }
  static constexpr bool first_nz(int m, int oc) {
    for (int mm = 0; mm < m; ++mm)
      if (nz(mm, oc)) return false;
    return true;
  }
};
struct PipeNoEl {
  static constexpr bool kScene = false;
  static constexpr int kN = 0, kLayout = 0, CO = 0;
  static constexpr bool any(int) { return false; }
};

// One pipeline signature.  (L0, N0) / (L1, N1): layout (or -1 = scene-based) and renderer input count of the elements,
// N1 == 0 for a single element.  S16: the decoded rows are staged as int16.  NSTAGE input stages, NW worker warps of VEC
// instants per thread, MINB blocks per SM the register allocation aims at.
template <int L0, int N0, int L1, int N1, int TARGET, bool S16, int NSTAGE_, int NW_, int VEC_, int MINB_, bool FMA_ = false>
struct PipeSig {
  typedef PipeEl<L0, N0, TARGET> E0;
  typedef typename std::conditional<(N1 > 0), PipeEl<L1, (N1 > 0 ? N1 : 1), TARGET>, PipeNoEl>::type E1;
  static constexpr bool kTwo = N1 > 0;
  static constexpr bool kS16 = S16;
  static constexpr bool kFma = FMA_;   // IAMFB_ARITH_FMA: the dense contractions fuse multiply and add (tolerance mode)
  static constexpr int kStages = NSTAGE_, NW = NW_, VEC = VEC_, kMinBlocks = MINB_;
  static constexpr int kThreads = (NW_ + 1) * 32, kWorkers = NW_ * 32;
  static constexpr int CO = kPipeTargetCh[TARGET];
  static constexpr int kTarget = TARGET;
  static constexpr bool active(int oc) { return E0::any(oc) || E1::any(oc); }
  static constexpr int ny() {
    int n = 0;
    for (int oc = 0; oc < CO; ++oc) n += active(oc) ? 1 : 0;
    return n;
  }
  static constexpr int NY = ny();
  static constexpr uint32_t active_mask() {
    uint32_t m = 0;
    for (int oc = 0; oc < CO; ++oc) m |= active(oc) ? (1u << oc) : 0u;
    return m;
  }
  static constexpr uint32_t kActiveMask = active_mask();
  static constexpr int yrow(int oc) {   // row of output channel oc on the time line, -1 when nothing ever writes it
    if (!active(oc)) return -1;
    int n = 0;
    for (int c = 0; c < oc; ++c) n += active(c) ? 1 : 0;
    return n;
  }
  static constexpr int kEsz = S16 ? 2 : 4;
  // dense (HOA) matrices go through the packed FP32 instructions; the sparse channel matrices stay scalar
  static constexpr bool kPacked = (L0 < 0) || (N1 > 0 && L1 < 0);
  // 16-bit output of a thread's VEC instants x CO channels leaves as whole 16-byte pieces
  static constexpr bool kFast16 = (CO % 2 == 0) && ((VEC_ * CO) % 8 == 0);
  static_assert(NW_ * 32 * VEC_ >= kStreamTile, "workers must cover a tile");
  static_assert(E0::kIdx >= 0, "no such rendering matrix");
};

struct PipeArgs {
  const void *in[kMaxEl];       // [S][F][n_in][N] float32 or int16
  const FrameRec *frames;       // [S][F]
  const float *start_win, *stop_win;
  const SubmitRec *submit;      // [S]
  StreamState *state;           // [S]
  const float *acc;             // limiter curve by time index (jr + 4 entries)
  float *hist_y;                // [S][co][kLimDelay]  limiter delay line carried between submits
  float *hist_pk;               // [S][kLimDelay]      peak ring carried between submits
  void *pcm;
  size_t stride_bytes;
  int n_frames;
  int row_bytes;                // bytes between the staged rows of a tile (240 instants)
  int stage_bytes;              // bytes between two input stages (multiple of 128)
  unsigned tpf_magic;           // ceil(2^32 / tiles per frame)
  float neg_zero;               // -0.0f, opaque to the assembler (pipe_mul2)
  const float *gain_ramp[kMaxEl];   // optional [S][F][N]: animated element mix gains (per sample; k_gain_expand or the caller)
  const float *out_gain_ramp;       // optional [S][F][N]: animated output mix gain
};

template <int VEC>
__device__ __forceinline__ Vec<VEC> vzero() {
  Vec<VEC> r;
#pragma unroll
  for (int k = 0; k < VEC; ++k) r.v[k] = 0.f;
  return r;
}

// VEC consecutive instants of one staged row.  int16 -> float32: x / 32768 exactly (opus/IAMF_opus_decoder.c:133-135) - the
// biased sample goes into the mantissa of 2^23 (8388608 + u), and (8388608 + u) * 2^-15 - 257 = (u - 32768) / 32768 with
// every intermediate representable: one fused multiply-add per sample
__device__ __forceinline__ float pipe_s16_lo(uint32_t w) { return __fmaf_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7410)), 3.0517578125e-05f, -257.0f); }
__device__ __forceinline__ float pipe_s16_hi(uint32_t w) { return __fmaf_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7432)), 3.0517578125e-05f, -257.0f); }
template <int VEC, bool S16>
__device__ __forceinline__ Vec<VEC> pipe_ld(const char *p) {
  Vec<VEC> r;
  if constexpr (S16) {
    if constexpr (VEC == 4) {
      uint2 w = *reinterpret_cast<const uint2 *>(p);
      w.x ^= 0x80008000u; w.y ^= 0x80008000u;
      r.v[0] = pipe_s16_lo(w.x); r.v[1] = pipe_s16_hi(w.x); r.v[2] = pipe_s16_lo(w.y); r.v[3] = pipe_s16_hi(w.y);
    } else {
      static_assert(VEC == 2 || VEC == 4, "instants per thread");
      const uint32_t w = *reinterpret_cast<const uint32_t *>(p) ^ 0x80008000u;
      r.v[0] = pipe_s16_lo(w); r.v[1] = pipe_s16_hi(w);
    }
  } else {
    r = ldsv<VEC>(reinterpret_cast<const float *>(p));
  }
  return r;
}

template <int LAYOUT, int CH, int NREC, int VEC>
__device__ __forceinline__ void pipe_put(Vec<VEC> (&x)[NREC], const Vec<VEC> &v) {
  constexpr int slot = stream_slot_of(LAYOUT, CH);
  if constexpr (slot >= 0) x[slot] = v;
}

// x / d correctly rounded (stream_div of iamfb_stream.cuh for any VEC)
template <int VEC>
__device__ __forceinline__ Vec<VEC> pipe_div(const Vec<VEC> &x, float d, float r) {
  Vec<VEC> q;
  float amax = 0.f, amin = 3.0e38f;
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const float q0 = x.v[k] * r;
    const float rem = __fmaf_rn(-d, q0, x.v[k]);
    q.v[k] = __fmaf_rn(rem, r, q0);
    amax = fmaxf(amax, fabsf(x.v[k]));
    amin = fminf(amin, fabsf(x.v[k]));
  }
  // verified range of the three-operation form: 2^-100 <= |x| < 2^126 (one test for the thread's values)
  if (!(amin >= 7.888609052210118e-31f && amax < 8.507059173023462e37f) || d == 0.f) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) q.v[k] = stream_slow_div(x.v[k], d);
  }
  return q;
}

// adds input m of element E (value v) to the running sums of the output channels it feeds (compile-time coefficients;
// the first contribution to a channel initialises its sum)
// Two instants at a time on the packed FP32 pipe (FFMA2 / FADD2), bit for bit the scalar multiply and add:
//   p = fma(v, c, -0) is the correctly rounded product v * c (adding -0 changes neither the value nor the sign of a
//   zero), then y + p is rounded separately.  The -0 comes from a kernel parameter: with a literal the assembler folds the
//   fma back into a multiply and contracts multiply + add into one FFMA2 (it does so for f32x2 even with --fmad=false).
// Half the issue slots of FMUL + FADD per instant; the FP32 pipe itself is as fast either way (tools/experiments/f32x2_bench.cu).
__device__ __forceinline__ void pipe_mul2(float &p0, float &p1, float v0, float v1, float c, float nz) {
  asm("{.reg .b64 rv, rc, rz, rt; mov.b64 rv, {%2,%3}; mov.b64 rc, {%4,%4}; mov.b64 rz, {%5,%5}; fma.rn.f32x2 rt, rv, rc, rz; mov.b64 {%0,%1}, rt;}"
      : "=f"(p0), "=f"(p1) : "f"(v0), "f"(v1), "f"(c), "f"(nz));
}
__device__ __forceinline__ void pipe_mac2(float &y0, float &y1, float v0, float v1, float c, float nz) {
  asm("{.reg .b64 rv, rc, rz, rt, ry; mov.b64 rv, {%2,%3}; mov.b64 rc, {%4,%4}; mov.b64 rz, {%5,%5}; fma.rn.f32x2 rt, rv, rc, rz; "
      "mov.b64 ry, {%0,%1}; add.rn.f32x2 ry, ry, rt; mov.b64 {%0,%1}, ry;}"
      : "+f"(y0), "+f"(y1) : "f"(v0), "f"(v1), "f"(c), "f"(nz));
}
// IAMFB_ARITH_FMA: y += v * c with ONE rounding (not the reference's two)
__device__ __forceinline__ void pipe_fma2(float &y0, float &y1, float v0, float v1, float c) {
  asm("{.reg .b64 rv, rc, ry; mov.b64 rv, {%2,%3}; mov.b64 rc, {%4,%4}; mov.b64 ry, {%0,%1}; fma.rn.f32x2 ry, rv, rc, ry; mov.b64 {%0,%1}, ry;}"
      : "+f"(y0), "+f"(y1) : "f"(v0), "f"(v1), "f"(c));
}
template <class SIG, class E, int M, int OC, int VEC, int NYY>
__device__ __forceinline__ void pipe_mat_col(Vec<VEC> (&y)[NYY], const Vec<VEC> &v, float nz) {
  if constexpr (OC < SIG::CO) {
    if constexpr (E::nz(M, OC)) {
      constexpr int row = SIG::yrow(OC);
      const float c = __uint_as_float(E::coef(M, OC));
      if constexpr (SIG::kPacked && (VEC % 2 == 0)) {
#pragma unroll
        for (int k = 0; k < VEC; k += 2) {
          if constexpr (E::first_nz(M, OC)) pipe_mul2(y[row].v[k], y[row].v[k + 1], v.v[k], v.v[k + 1], c, nz);
          else if constexpr (SIG::kFma) pipe_fma2(y[row].v[k], y[row].v[k + 1], v.v[k], v.v[k + 1], c);
          else pipe_mac2(y[row].v[k], y[row].v[k + 1], v.v[k], v.v[k + 1], c, nz);
        }
      } else if constexpr (E::first_nz(M, OC)) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) y[row].v[k] = c * v.v[k];
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) y[row].v[k] += c * v.v[k];
      }
    }
    pipe_mat_col<SIG, E, M, OC + 1, VEC, NYY>(y, v, nz);
  }
}

// ---- one channel-based element: de-mixing chain + recon gain + render matrix (k_stream's render, any VEC / input type).
// rows = the element's staged rows at this thread's first instant; rowb = bytes between rows; y receives the element's
// rendered channels (time-line rows of SIG)
template <class SIG, class E, int VEC, int NYY>
__device__ __forceinline__ void pipe_render_channel(const KernelPlan &plan, const ElPlan &ep, const ElFrame &ef, const char *rows, int rowb,
                                                    int i0, bool fade_w, const float *start_win, const float *stop_win, Vec<VEC> (&y)[NYY], float nz) {
  constexpr int LAYOUT = E::kLayout, NREC = E::kN;
  constexpr bool S16 = SIG::kS16;
  typedef Vec<VEC> V;
  auto ld_ch = [&](int ch) -> V {   // a transmitted IAChannel (zeros when absent), with its output gain (dmx_gainup, demixer.c:421-430)
    const int off = ep.s_row_off[ch];   // byte offset of its staged row for this launch's input format, < 0 when not transmitted
    V r = vzero<VEC>();
    if (off >= 0) r = pipe_ld<VEC, S16>(rows + off);
    if ((ep.gain_mask >> ch) & 1u) {
      const float g = ep.gain[ch];
#pragma unroll
      for (int k = 0; k < VEC; ++k) r.v[k] *= g;
    }
    return r;
  };
  V xd[NREC];
  const bool s23 = (ep.need_s2 | ep.need_s3) != 0, s7h2 = (ep.need_s7 | ep.need_h2) != 0;
  if (s23 | s7h2 | ((ep.need_s5 | ep.need_h4) != 0)) {
    const int mode = ef.mode & 7;
    V pa = vzero<VEC>(), pb = vzero<VEC>();
    if (s23) {
      pa = ld_ch(IAMFB_CH_L2);
      pb = ld_ch(ep.need_s2 ? IAMFB_CH_MONO : IAMFB_CH_R2);
    } else if (ep.need_s5) {
      pa = ld_ch(IAMFB_CH_L3);
      pb = ld_ch(IAMFB_CH_R3);
    } else if (s7h2) {
      pa = ld_ch(IAMFB_CH_SL5);
      pb = ld_ch(IAMFB_CH_SR5);
    }
    if (ep.need_s2) {   // R2 = 2*Mono - L2, demixer.c:136-138
#pragma unroll
      for (int k = 0; k < VEC; ++k) pb.v[k] = 2 * pb.v[k] - pa.v[k];
      pipe_put<LAYOUT, IAMFB_CH_R2, NREC>(xd, pb);
    }
    if (ep.need_s3) {   // L3 = L2 - 0.707*C evaluated in double, demixer.c:165-168
      const V cc = ld_ch(IAMFB_CH_C);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const double c = (double)cc.v[k];
        pa.v[k] = (float)((double)pa.v[k] - 0.707 * c);
        pb.v[k] = (float)((double)pb.v[k] - 0.707 * c);
      }
      pipe_put<LAYOUT, IAMFB_CH_L3, NREC>(xd, pa);
      pipe_put<LAYOUT, IAMFB_CH_R3, NREC>(xd, pb);
    } else if (ep.need_s2 && ep.need_s5) {
      pa = ld_ch(IAMFB_CH_L3);
      pb = ld_ch(IAMFB_CH_R3);
    }
    if (ep.need_s5) {   // Ls5 = (L3 - L5)/delta, demixer.c:213-218
      const V l5 = ld_ch(IAMFB_CH_L5), r5 = ld_ch(IAMFB_CH_R5);
#pragma unroll
      for (int k = 0; k < VEC; ++k) { pa.v[k] = pa.v[k] - l5.v[k]; pb.v[k] = pb.v[k] - r5.v[k]; }
      pa = pipe_div<VEC>(pa, c_mix_delta[mode], c_mix_gd_r[mode]);
      pb = pipe_div<VEC>(pb, c_mix_delta[mode], c_mix_gd_r[mode]);
      pipe_put<LAYOUT, IAMFB_CH_SL5, NREC>(xd, pa);
      pipe_put<LAYOUT, IAMFB_CH_SR5, NREC>(xd, pb);
    } else if (s23 && s7h2) {
      pa = ld_ch(IAMFB_CH_SL5);
      pb = ld_ch(IAMFB_CH_SR5);
    }
    if (ep.need_h2 | ep.need_h4) {
      V ta = ld_ch(ep.need_h2 ? IAMFB_CH_TL : IAMFB_CH_HL);
      V tb = ld_ch(ep.need_h2 ? IAMFB_CH_TR : IAMFB_CH_HR);
      if (ep.need_h2) {   // Ltf2 = Ltf3 - delta*w*Ls5, demixer.c:318-323
        const float dw = c_mix_delta[mode] * ef.w;
#pragma unroll
        for (int k = 0; k < VEC; ++k) { ta.v[k] = ta.v[k] - dw * pa.v[k]; tb.v[k] = tb.v[k] - dw * pb.v[k]; }
        pipe_put<LAYOUT, IAMFB_CH_HL, NREC>(xd, ta);
        pipe_put<LAYOUT, IAMFB_CH_HR, NREC>(xd, tb);
      }
      if (ep.need_h4) {   // Ltb = (Ltf2 - Ltf4)/gamma, demixer.c:363-368
        const V hfl = ld_ch(IAMFB_CH_HFL), hfr = ld_ch(IAMFB_CH_HFR);
#pragma unroll
        for (int k = 0; k < VEC; ++k) { ta.v[k] = ta.v[k] - hfl.v[k]; tb.v[k] = tb.v[k] - hfr.v[k]; }
        pipe_put<LAYOUT, IAMFB_CH_HBL, NREC>(xd, pipe_div<VEC>(ta, c_mix_gamma[mode], c_mix_gd_r[mode]));
        pipe_put<LAYOUT, IAMFB_CH_HBR, NREC>(xd, pipe_div<VEC>(tb, c_mix_gamma[mode], c_mix_gd_r[mode]));
      }
    }
    if (ep.need_s7) {   // Lb7 = (Ls5 - alpha*Lss7)/beta, demixer.c:262-269
      const V sl7 = ld_ch(IAMFB_CH_SL7), sr7 = ld_ch(IAMFB_CH_SR7);
      const float al = c_mix_alpha[mode];
#pragma unroll
      for (int k = 0; k < VEC; ++k) { pa.v[k] = pa.v[k] - sl7.v[k] * al; pb.v[k] = pb.v[k] - sr7.v[k] * al; }
      pipe_put<LAYOUT, IAMFB_CH_BL7, NREC>(xd, pipe_div<VEC>(pa, c_mix_beta[mode], c_mix_beta_r[mode]));
      pipe_put<LAYOUT, IAMFB_CH_BR7, NREC>(xd, pipe_div<VEC>(pb, c_mix_beta[mode], c_mix_beta_r[mode]));
    }
  }
  // the layout's channels in layout order: derived value or staged row (plans with an output gain on a channel of the
  // layout itself take k_fused), recon gain (dmx_rms, demixer.c:461-468; 1.0 in the slots without one), matrix column
  auto column_value = [&](auto m_c) -> V {
    constexpr int m = decltype(m_c)::value;
    constexpr int ch = fused_order(LAYOUT, m);
    if constexpr (stream_derivable(ch)) {
      if (stream_derived(ep, ch)) return xd[m];
    }
    return pipe_ld<VEC, S16>(rows + ep.s_row_off[ch]);
  };
  if (fade_w) {
    const unsigned rmask = ef.rmask;
    V st = vzero<VEC>(), sw;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      sw.v[k] = 1.f;
      if (i0 + k < plan.overlap) { st.v[k] = stop_win[i0 + k]; sw.v[k] = start_win[i0 + k]; }
    }
    auto column = [&](auto m_c) {
      constexpr int m = decltype(m_c)::value;
      V v = column_value(m_c);
      if ((rmask >> m) & 1u) {
        const float lm = ef.rlast[m], cm = ef.rcur[m];
#pragma unroll
        for (int k = 0; k < VEC; ++k) v.v[k] *= lm * st.v[k] + cm * sw.v[k];
      }
      pipe_mat_col<SIG, E, m, 0, VEC, NYY>(y, v, nz);
    };
    stream_for<NREC>(column);
  } else {
    auto column = [&](auto m_c) {
      constexpr int m = decltype(m_c)::value;
      V v = column_value(m_c);
      const float cm = ef.rcur[m];
#pragma unroll
      for (int k = 0; k < VEC; ++k) v.v[k] *= cm;
      pipe_mat_col<SIG, E, m, 0, VEC, NYY>(y, v, nz);
    };
    stream_for<NREC>(column);
  }
}

// ---- one scene-based element with a mono channel mapping (IAMF_core_decoder.c:105-116): ambisonics channel m is decoded
// row ambi_map[m]; out = sum over m ascending (h2m_rdr.c:1103-1112), LFE slots shifted / zeroed at compile time
template <class SIG, class E, int VEC, int NYY>
__device__ __forceinline__ void pipe_render_scene(const ElPlan &ep, const char *rows, int rowb, Vec<VEC> (&y)[NYY], float nz) {
  auto column = [&](auto m_c) {
    constexpr int m = decltype(m_c)::value;
    const Vec<VEC> v = pipe_ld<VEC, SIG::kS16>(rows + ep.s_row_off[m]);   // (scene-based: indexed by ambisonics channel)
    pipe_mat_col<SIG, E, m, 0, VEC, NYY>(y, v, nz);
  };
  stream_for<E::kN>(column);
}

template <class SIG, class E, int VEC, int NYY>
__device__ __forceinline__ void pipe_render_element(const KernelPlan &plan, const ElPlan &ep, const ElFrame &ef, const char *rows, int rowb,
                                                    int i0, bool fade_w, const float *start_win, const float *stop_win, Vec<VEC> (&y)[NYY], float nz) {
  if constexpr (E::kScene) pipe_render_scene<SIG, E, VEC, NYY>(ep, rows, rowb, y, nz);
  else pipe_render_channel<SIG, E, VEC, NYY>(plan, ep, ef, rows, rowb, i0, fade_w, start_win, stop_win, y, nz);
}

// y[row of oc] += y1[row of oc] for the channels element 1 feeds (iamf_mixer_mix, IAMF_decoder.c:2719-2730: acc = 0;
// acc += e0; acc += e1 - a channel only one element feeds keeps that element's value: x + 0 == 0 + x == x up to the
// sign of a zero, settled where the float output needs it)
template <class SIG, int OC, int VEC, int NYY>
__device__ __forceinline__ void pipe_mix(Vec<VEC> (&y)[NYY], const Vec<VEC> (&y1)[NYY]) {
  if constexpr (OC < SIG::CO) {
    if constexpr (SIG::E1::any(OC)) {
      constexpr int row = SIG::yrow(OC);
      if constexpr (SIG::E0::any(OC)) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) y[row].v[k] = y[row].v[k] + y1[row].v[k];
      } else {
        y[row] = y1[row];
      }
    }
    pipe_mix<SIG, OC + 1, VEC, NYY>(y, y1);
  }
}
template <class SIG, class E, int OC, int VEC, int NYY>
__device__ __forceinline__ void pipe_scale_v(Vec<VEC> (&y)[NYY], const Vec<VEC> &g) {   // ... x its per-sample mix gains
  if constexpr (OC < SIG::CO) {
    if constexpr (E::any(OC)) {
      constexpr int row = SIG::yrow(OC);
#pragma unroll
      for (int k = 0; k < VEC; ++k) y[row].v[k] *= g.v[k];
    }
    pipe_scale_v<SIG, E, OC + 1, VEC, NYY>(y, g);
  }
}
template <int VEC>
__device__ __forceinline__ Vec<VEC> pipe_ldg(const float *p) {   // VEC consecutive floats from global memory (aligned)
  Vec<VEC> r;
  if constexpr (VEC == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
    r.v[0] = t.x; r.v[1] = t.y;
  }
  return r;
}
template <class SIG, class E, int OC, int VEC, int NYY>
__device__ __forceinline__ void pipe_scale(Vec<VEC> (&y)[NYY], float g) {   // the element's channels x its mix gain
  if constexpr (OC < SIG::CO) {
    if constexpr (E::any(OC)) {
      constexpr int row = SIG::yrow(OC);
#pragma unroll
      for (int k = 0; k < VEC; ++k) y[row].v[k] *= g;
    }
    pipe_scale<SIG, E, OC + 1, VEC, NYY>(y, g);
  }
}

// time-line row of output channel oc for a run-time (or unrolled) channel index
template <class SIG>
__device__ __forceinline__ int pipe_yrow_rt(int oc) {
  constexpr uint32_t m = SIG::kActiveMask;
  return ((m >> oc) & 1u) ? __popc(m & ((1u << oc) - 1u)) : -1;
}

// FLOAT2INT16/24/32 + interleave (IAMF_decoder.c:100-167) of n (<= VEC) instants starting at output index o0:
// ys = the thread's first instant in time-line slot 0 of row 0 (rows are 2 * TL floats apart), gp = gains or nullptr
template <class SIG>
__device__ __forceinline__ void pipe_emit(const float *ys, const float *gp, char *out, long long o0, int n, int bits) {
  constexpr int VEC = SIG::VEC, CO = SIG::CO, TL = kStreamTile;
  Vec<VEC> gg;
#pragma unroll
  for (int k = 0; k < VEC; ++k) gg.v[k] = 1.0f;
  if (gp) gg = ldsv<VEC>(gp);
  if constexpr (SIG::kFast16) {
    if (bits == 16 && n == VEC && ((((size_t)out) & 15) == 0)) {
      // (x * g) * 2^15 == x * (g * 2^15) bit for bit (a power-of-two scale commutes with the rounding of the product;
      // products small enough to be subnormal quantise to 0 either way)
#pragma unroll
      for (int k = 0; k < VEC; ++k) gg.v[k] *= 32768.f;
      uint32_t w[VEC * CO / 2];                      // [instant][channel pair]
#pragma unroll
      for (int c = 0; c < CO; c += 2) {
        Vec<VEC> v0 = vzero<VEC>(), v1 = vzero<VEC>();
        // (rows resolved at compile time through the unrolled channel index)
        const int r0 = pipe_yrow_rt<SIG>(c), r1 = pipe_yrow_rt<SIG>(c + 1);
        if (r0 >= 0) v0 = ldsv<VEC>(ys + r0 * 2 * TL);
        if (r1 >= 0) v1 = ldsv<VEC>(ys + r1 * 2 * TL);
#pragma unroll
        for (int k = 0; k < VEC; ++k) w[k * (CO / 2) + (c >> 1)] = stream_q16x2(v0.v[k] * gg.v[k], v1.v[k] * gg.v[k]);
      }
      uint32_t *dst = reinterpret_cast<uint32_t *>(reinterpret_cast<int16_t *>(out) + o0 * CO);
#pragma unroll
      for (int i = 0; i < VEC * CO / 2; i += 4) *reinterpret_cast<uint4 *>(dst + i) = make_uint4(w[i], w[i + 1], w[i + 2], w[i + 3]);
      return;
    }
  }
#pragma unroll 1
  for (int c = 0; c < CO; ++c) {
    const int r = pipe_yrow_rt<SIG>(c);
    Vec<VEC> v = vzero<VEC>();
    if (r >= 0) v = ldsv<VEC>(ys + r * 2 * TL);
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      if (k < n) store_any(out, (size_t)(o0 + k) * CO + c, gp ? v.v[k] * gg.v[k] : v.v[k], bits);
  }
}

// RAMPS: the variant launched for submits with animated mix gains (per-sample gain arrays).  A variant of its own because
// the mere presence of that code in the kernel costs the common case 1.5 - 3 % (measured on C3 / C4: code size, not the test)
template <class SIG, bool RAMPS = false>
__global__ void __launch_bounds__(SIG::kThreads, SIG::kMinBlocks)
k_pipe(const __grid_constant__ KernelPlan plan, PipeArgs a, const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1) {
  typedef typename SIG::E0 E0;
  typedef typename SIG::E1 E1;
  constexpr int VEC = SIG::VEC, NW = SIG::NW, CO = SIG::CO, NY = SIG::NY, WN = SIG::kWorkers, NS = SIG::kStages;
  constexpr int TL = kStreamTile;
  typedef Vec<VEC> V;
  extern __shared__ __align__(128) float fsm[];
  __shared__ __align__(8) uint64_t s_bar[NS];
  __shared__ __align__(8) uint64_t s_hbar;     // arrival of the previous submit's history
  __shared__ __align__(16) FrameRec s_fr[2];   // resolved parameters of the frames being rendered, by frame parity
  __shared__ __align__(16) float s_es[2][32];  // scanner: thr / peak of the steps of a burst
  __shared__ float s_acc[kStreamAccCache];
  __shared__ int s_hot[2][NW], s_apply[2];
  __shared__ float s_tot[NW];                  // maximum peak of each worker warp's part of the tile being rendered
  __shared__ int s_skip;
  const ElPlan &ep0 = plan.el[0];
  const int nin0 = ep0.n_in, nin1 = SIG::kTwo ? plan.el[1].n_in : 0;
  float *Y = fsm;                    // [NY][2][TL]  mixed time line, tile t in slot t & 1 (tile -1 = history in slot 1)
  float *WM = Y + NY * 2 * TL;       // [2][TL]      look-ahead maximum, tile t in slot t & 1
  float *G = WM + 2 * TL;            // [2][TL]      gains
  float *SA = G + 2 * TL;            // [2][TL]      suffix maxima of the peaks of tile t (instants r.. of the tile)
  // input stages (128-byte aligned): the decoded rows of a tile, nin0 + nin1 rows of TL instants
  char *ST = reinterpret_cast<char *>(fsm) + (((NY * 2 + 6) * TL * 4 + 127) & ~127);
  const int s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const bool worker = tid < WN;
  const int N = plan.frame_size;
  const int TPF = N / TL;
  const int T = a.n_frames * TPF;
  const unsigned tpf_magic = a.tpf_magic;        // ceil(2^32 / TPF): tau / TPF == umulhi(tau, magic) for tau < 2^16
  const float thr = plan.lim_thr;
  const int bits = plan.bit_depth;
  const bool limiter = plan.limiter != 0;

  // What the previous submit left - the limiter's delay line (tile -1: slot 1 of every time-line row) and the peaks of its
  // last window (into WM's slot 1 for now) - comes in by bulk copies on their own barrier, under everything up to the first
  // tile; under programmatic dependent launch all of this runs while k_resolve is still resolving this submit's frames
  if (limiter && tid == 0) {
    mbar_init(&s_hbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&s_hbar, (uint32_t)((NY + 1) * kLimDelay * sizeof(float)));
#pragma unroll 1
    for (int c = 0; c < CO; ++c) {
      const int r = pipe_yrow_rt<SIG>(c);
      if (r >= 0) bulk_g2s(Y + (r * 2 + 1) * TL, a.hist_y + ((size_t)s * CO + c) * kLimDelay, (uint32_t)(kLimDelay * sizeof(float)), &s_hbar);
    }
    bulk_g2s(WM + TL, a.hist_pk + (size_t)s * kLimDelay, (uint32_t)(kLimDelay * sizeof(float)), &s_hbar);
  }
  for (int i = tid; i < kStreamAccCache; i += SIG::kThreads) s_acc[i] = (limiter && i <= plan.lim_jr + 3) ? a.acc[i] : 0.f;
  // From here on k_resolve's results are needed; the launch after this one (k_fused for the irregular streams of the submit,
  // none of which are this kernel's) may start
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (limiter) {
    __syncthreads();                   // the barrier's initialisation is visible to every thread
    mbar_wait(&s_hbar, 0u);
    if (tid < 32) {
      // suffix maxima of the peaks of the tile before this submit (tile -1, slot 1): 8 instants per lane, 30 lanes
      const float *src = WM + TL;
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
  }
  if (a.submit[s].irregular) return;   // rendered by k_fused right after (block-uniform)
  if (tid == 0) {
#pragma unroll
    for (int w = 0; w < NW; ++w) s_hot[0][w] = s_hot[1][w] = 0;
    s_apply[0] = s_apply[1] = 0;
    s_skip = a.submit[s].out_skip;
#pragma unroll
    for (int b = 0; b < NS; ++b) mbar_init(&s_bar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const int q0 = VEC * tid;                       // this worker's first instant inside a tile
  const bool has = worker && q0 < TL;

  // thread 0: tile tau = the TL instants at offset t_off of frame f -> stage tau % NS: the frame's resolved parameters and
  // one tensor copy per element (a box of n_in rows x TL instants of the submit's input tensor map)
  auto issue = [&](int tau) {
    if (tid == 0) {
      int s_it = s;
      asm volatile("" : "+r"(s_it));   // (kept opaque: block-uniform pointers are rebuilt here, not carried in registers)
      const int f = TPF == 1 ? tau : (int)__umulhi((unsigned)tau, tpf_magic), t_off = (tau - f * TPF) * TL;
      char *st = ST + (tau % NS) * a.stage_bytes;
      uint64_t *bar = &s_bar[tau % NS];
      const int sf = s_it * a.n_frames + f;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      // the frame's record travels with the frame's first tile, into the slot of the frame's parity (the frame before last,
      // whose slot it takes, has been rendered completely by the time any tile of this frame is issued)
      mbar_expect_tx(bar, (uint32_t)(a.row_bytes * (nin0 + nin1) + (t_off == 0 ? sizeof(FrameRec) : 0)));
      if (t_off == 0) bulk_g2s(&s_fr[f & 1], a.frames + sf, (uint32_t)sizeof(FrameRec), bar);
      tensor_g2s_2d(st, &map0, t_off, sf * nin0, bar);
      if constexpr (SIG::kTwo) tensor_g2s_2d(st + nin0 * a.row_bytes, &map1, t_off, sf * nin1, bar);
    }
  };

  V yh[NY];
  V pkh = vzero<VEC>();                            // cross-channel peak of this thread's instants of the rendered tile
#pragma unroll
  for (int r = 0; r < NY; ++r) yh[r] = vzero<VEC>();

  // tile tau (frame f, offset t_off) is in its stage: leaves the mixed samples of this thread's instants in yh and their
  // cross-channel peak in pkh
  // (the lanes beyond the tile - 60..63 of the last warp when VEC = 4 - redo the tile's last piece instead of leaving: no
  // divergence in front of the warp-wide scans that follow, and the compiler keeps the shuffles free of re-convergence code)
  const int q0r = has ? q0 : TL - VEC;
  auto render = [&](int tau) {
    const int f = TPF == 1 ? tau : (int)__umulhi((unsigned)tau, tpf_magic), t_off = (tau - f * TPF) * TL;
    const char *st = ST + (tau % NS) * a.stage_bytes;
    const FrameRec &fr = s_fr[f & 1];
    const char *rows = st + q0r * SIG::kEsz;
    const int i0 = t_off + q0r;
    const bool fade_w = t_off + VEC * (tid & ~31) < plan.overlap;   // warp-uniform: some lane is inside the recon cross-fade
#pragma unroll
    for (int r = 0; r < NY; ++r) yh[r] = vzero<VEC>();
    pipe_render_element<SIG, E0, VEC, NY>(plan, ep0, fr.el[0], rows, a.row_bytes, i0, fade_w, a.start_win, a.stop_win, yh, a.neg_zero);
    // element mix gain (iamf_frame_gain IAMF_decoder.c:1392): a constant, skipped when it is 1 (or not positive) - or one
    // gain per sample (animated mix gain, :1395-1405), always applied
    size_t gidx = 0;
    bool ramp0 = false, ramp1 = false, og_ramp = false;
    if constexpr (RAMPS) {
      gidx = ((size_t)s * a.n_frames + f) * plan.frame_size + i0;
      ramp0 = a.gain_ramp[0] != nullptr; ramp1 = a.gain_ramp[kMaxEl - 1] != nullptr; og_ramp = a.out_gain_ramp != nullptr;   // (block-uniform)
      if (ramp0) pipe_scale_v<SIG, E0, 0, VEC, NY>(yh, pipe_ldg<VEC>(a.gain_ramp[0] + gidx));
    }
    if (!ramp0) {
      const float eg = fr.el[0].gain;
      if (eg != 1.f && eg > 0.f) pipe_scale<SIG, E0, 0, VEC, NY>(yh, eg);
    }
    if constexpr (SIG::kTwo) {
      V y1[NY];
#pragma unroll
      for (int r = 0; r < NY; ++r) y1[r] = vzero<VEC>();
      pipe_render_element<SIG, E1, VEC, NY>(plan, plan.el[1], fr.el[1], rows + nin0 * a.row_bytes, a.row_bytes, i0, fade_w, a.start_win,
                                            a.stop_win, y1, a.neg_zero);
      if constexpr (RAMPS) {
        if (ramp1) pipe_scale_v<SIG, E1, 0, VEC, NY>(y1, pipe_ldg<VEC>(a.gain_ramp[1] + gidx));
      }
      if (!ramp1) {
        const float eg = fr.el[1].gain;
        if (eg != 1.f && eg > 0.f) pipe_scale<SIG, E1, 0, VEC, NY>(y1, eg);
      }
      pipe_mix<SIG, 0, VEC, NY>(yh, y1);
    }
    // output mix gain (:3463-3469), loudness (:3480-3484, :3211) - each skipped when it is 1 - and the peak of every instant
    const float ogain = fr.out_gain;
    const bool og_on = ogain != 1.f && ogain > 0.f && !og_ramp;
    if constexpr (RAMPS) {
      if (og_ramp) {   // animated output mix gain (:3463-3469): one gain per sample, every channel
        const V og = pipe_ldg<VEC>(a.out_gain_ramp + gidx);
#pragma unroll
        for (int r = 0; r < NY; ++r)
#pragma unroll
          for (int k = 0; k < VEC; ++k) yh[r].v[k] *= og.v[k];
      }
    }
    const bool loud_on = plan.loud_gain != 0.f && plan.loud_gain != 1.0f;
    V peak = vzero<VEC>();
    if (og_on | loud_on | (bits == 0)) {          // (block-uniform; rarely taken)
#pragma unroll 1
      for (int pass = 0; pass < 3; ++pass) {
        // pass 0: output mix gain; 1: loudness; 2 (float output only): the sign of a zero is visible there - the reference's
        // sums start at +0 (m2m_rdr.c:1826, iamf_mixer_mix :2719), so a zero result is +0
        if (!(pass == 0 ? og_on : (pass == 1 ? loud_on : bits == 0))) continue;
        const float g = pass == 0 ? ogain : plan.loud_gain;
#pragma unroll
        for (int r = 0; r < NY; ++r)
#pragma unroll
          for (int k = 0; k < VEC; ++k) yh[r].v[k] = pass == 2 ? 0.f + yh[r].v[k] : yh[r].v[k] * g;
      }
    }
#pragma unroll
    for (int r = 0; r < NY; ++r)
#pragma unroll
      for (int k = 0; k < VEC; ++k) peak.v[k] = fmaxf(peak.v[k], fabsf(yh[r].v[k]));
#pragma unroll
    for (int k = 0; k < VEC; ++k) pkh.v[k] = has ? peak.v[k] : 0.f;
    // the peaks of the submit's last tile are the history the next submit starts from
    if (limiter && tau == T - 1 && has) stsv<VEC>(a.hist_pk + (size_t)s * kLimDelay + q0, peak);
  };

  // Look-ahead maximum of tile t: WM[r] = max(previous tile's instants r.., this tile's instants ..r-1) (van Herk /
  // Gil-Werman with blocks of one limiter window; peaks are >= 0, so 0 is the neutral element): both scans of this tile's
  // peaks run on the registers the render left them in - inside the thread, then across the warp by shuffles; the warps
  // exchange their totals through shared memory (s_tot) across the one worker barrier of the tile
  float pre[VEC], suf[VEC];
  auto wmax_scan = [&]() {
    float inc[VEC];                               // inclusive prefixes of the thread's instants
    inc[0] = pkh.v[0];
#pragma unroll
    for (int k = 1; k < VEC; ++k) inc[k] = fmaxf(inc[k - 1], pkh.v[k]);
    float sfx[VEC];                               // inclusive suffixes (sfx[0] is the total)
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
    float ex = __shfl_up_sync(0xffffffffu, up, 1);      // maximum of all earlier lanes
    if (lane == 0) ex = 0.f;
    float sx = __shfl_down_sync(0xffffffffu, dn, 1);    // maximum of all later lanes
    if (lane == 31) sx = 0.f;
    pre[0] = ex;
#pragma unroll
    for (int k = 1; k < VEC; ++k) pre[k] = fmaxf(ex, inc[k - 1]);
#pragma unroll
    for (int k = 0; k < VEC; ++k) suf[k] = fmaxf(sfx[k], sx);
    if (lane == 31) s_tot[tid >> 5] = up;                // the warp's total
  };
  auto wmax_combine = [&](int t) {
    const int b = t & 1;
    const int wi = tid >> 5;
    float cp = 0.f, cs = 0.f;                            // totals of the warps before / after this one
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float o = s_tot[w];
      if (w < wi) cp = fmaxf(cp, o);
      if (w > wi) cs = fmaxf(cs, o);
    }
    int hot = 0;
    if (has) {
      const V A = ldsv<VEC>(SA + (b ^ 1) * TL + q0);
      V W, Sx;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        W.v[k] = fmaxf(A.v[k], fmaxf(pre[k], cp));
        Sx.v[k] = fmaxf(suf[k], cs);
        hot |= (W.v[k] > thr) ? 1 : 0;
      }
      stsv<VEC>(WM + b * TL + q0, W);
      stsv<VEC>(SA + b * TL + q0, Sx);
    }
    const int any_hot = __any_sync(0xffffffffu, hot);
    if (lane == 0) s_hot[b][wi] = any_hot;
  };
  // Tile t leaves the limiter: instant k is (instant k of tile t-1) x gain[k] (delay line of 240,
  // audio_effect_peak_limiter.c:167-201), then quantise + interleave.  Tile t-1 sits in time-line slot (t-1) & 1 = the slot
  // tile t+1 (held in yh since it was rendered) goes to: every row is read, then overwritten.
  // Without a limiter there is no delay: tile t itself is written out (from yh's slot after the store).
  auto output_and_store = [&](int t, bool do_out, bool do_store) {
    const int b = t & 1;
    float *yt = Y + (b ^ 1) * TL + q0;
    if (limiter) {
      const int o0 = t * TL + q0 - *(volatile int *)&s_skip;
      // limiter priming: the first 240 instants are dropped (:180-189).  A tile the limiter left alone (gain 1 throughout)
      // has already been written out by the scanner warp (s_apply == 0)
      if (do_out && has && o0 >= 0 && s_apply[b] != 0) {
        int s_it = s;
        asm volatile("" : "+r"(s_it));
        pipe_emit<SIG>(yt, G + b * TL + q0, (char *)a.pcm + (size_t)s_it * a.stride_bytes, o0, VEC, bits);
      }
    }
    if (do_store && has) {
#pragma unroll
      for (int r = 0; r < NY; ++r) stsv<VEC>(yt + r * 2 * TL, yh[r]);
    }
  };
  // the same output stage on the scanner warp, for a tile the limiter leaves alone (gain 1.0 for every instant): lane l
  // takes the threads' pieces l, l + 32, ... one iteration EARLIER than the workers would, and the workers skip it
  auto quiet_output = [&](int t) {
    const int b = t & 1;
    const int skip = *(volatile int *)&s_skip;
    int s_it = s;
    asm volatile("" : "+r"(s_it));
    char *out = (char *)a.pcm + (size_t)s_it * a.stride_bytes;
#pragma unroll 1
    for (int qd = lane; qd < TL / VEC; qd += 32) {
      const int o0 = t * TL + VEC * qd - skip;
      if (o0 < 0) continue;
      pipe_emit<SIG>(Y + (b ^ 1) * TL + VEC * qd, nullptr, out, o0, VEC, bits);
    }
  };

  __syncthreads();                                 // history, curve cache, flags and the copy barriers are in place

  if (worker) {
    if (tid == 0) {
#pragma unroll
      for (int b = 0; b < NS; ++b)
        if (b < T) issue(b);
    }
#pragma unroll 1
    for (int t = -1; t <= T; ++t) {
      if (t >= 0) output_and_store(t - 1, t >= 1, t < T);
      if (t + 1 < T) {
        const int tau = t + 1;
        mbar_wait(&s_bar[tau % NS], (uint32_t)(tau / NS) & 1u);
        render(tau);
        if (limiter) wmax_scan();
        asm volatile("bar.sync 1, %0;" ::"n"(WN) : "memory");   // every worker is done with the stage, the warps' totals are posted
        if (tau + NS < T) issue(tau + NS);
        if (limiter) wmax_combine(tau);
      }
      if (!limiter && t >= 0 && t < T && has) {
        // no limiter: no delay line - tile t (stored above) is written out as it is
        int s_it = s;
        asm volatile("" : "+r"(s_it));
        pipe_emit<SIG>(Y + (t & 1) * TL + q0, nullptr, (char *)a.pcm + (size_t)s_it * a.stride_bytes, (long long)t * TL + q0, VEC, bits);
      }
      asm volatile("bar.sync 2, %0;" ::"n"(SIG::kThreads) : "memory");
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
        const int b = t & 1;
        const bool idle = lj < 0 || lj >= plan.lim_jr;
        int hot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) hot |= s_hot[b][w];
        const bool run = hot != 0 || !idle;
        if (run) stream_scan(WM + b * TL, G + b * TL, &s_es[0][0], TL, lj, lS, lE, in_run, a.acc, s_acc, plan.lim_ja, plan.lim_jr, thr, lane);
        else {
          in_run = false;
          quiet_output(t);
        }
        if (lane == 0) s_apply[b] = run ? 1 : 0;
      }
      asm volatile("bar.sync 2, %0;" ::"n"(SIG::kThreads) : "memory");
    }
    if (limiter && lane == 0) {
      StreamState &st = a.state[s];
      st.lim_j = lj; st.lim_start = lS; st.lim_end = lE;
    }
  }
  // the last 240 instants (= tile T-1) are the history of the next submit: one bulk copy per time-line row (rows no
  // matrix ever writes stay zero in the history, as batch_reset left them)
  if (limiter && tid == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll 1
    for (int c = 0; c < CO; ++c) {
      const int r = pipe_yrow_rt<SIG>(c);
      if (r < 0) continue;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.hist_y + ((size_t)s * CO + c) * kLimDelay),
                   "r"(smem_u32(Y + (r * 2 + ((T + 1) & 1)) * TL)), "r"((uint32_t)(kLimDelay * sizeof(float)))
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

}  // namespace iamfb
