// iamfb_fused.cuh - the fused per-stream kernel of the non-resampling pipelines.
//
// One thread block owns one stream for the whole submit and walks its frames tile by tile (a tile = up to `tile`
// consecutive samples of one frame).  Per tile, everything between the decoded planar input and the interleaved
// integer PCM stays on chip:
//
//   stage    the tile's decoded rows are brought into shared memory by bulk async copies (cp.async.bulk + mbarrier,
//            the TMA engine), issued one tile AHEAD: the copy of tile t+1 runs under the limiter / output phases of t
//   render   reconstruct + render + gains + element sum of ALL audio elements -> mixed samples in shared memory
//            (Y ring: 240 delayed samples + the tile) and the per-instant cross-channel peak (PK ring).  Channel
//            sources are read from the staged rows by (uniform) row index, so nothing is indexed dynamically in
//            registers: the reconstructed channels go back over the staged rows (each thread owns its four samples
//            of every row) and the render matrix is applied row by row from a compressed (zero-free) copy
//   wmax     240-sample sliding maximum of PK: windows of 8 in registers, then 16..128 by doubling with 16-byte
//            shared-memory accesses, 240 = two overlapping windows of 128                       (limiter look-ahead)
//   scan     the limiter's serial gain recurrence, run by warp 0 as a hybrid of
//              - a 32-wide parallel search for the next trigger while the gain follows its attack/release curve, and
//              - a speculative serial burst while the limiter re-triggers on every sample
//   output   delayed sample x gain -> quantise -> interleaved PCM, written straight from shared memory
//
// so HBM sees the algorithmic bytes only: the decoded input once and the PCM once (plus 240 samples of history per
// channel and stream carried between submits).  Arithmetic is the reference's, expression by expression (see the
// citations in iamfb_kernels.cuh); results are bit-identical to the multi-kernel path and to the oracle.
#pragma once
#include "iamfb_kernels.cuh"

namespace iamfb {

struct FusedArgs {
  const float *in[kMaxEl];      // [S][F][n_in][N]
  const FrameRec *frames;       // [S][F]
  const float *gain_ramp[kMaxEl];
  const float *out_gain_ramp;
  const float *start_win, *stop_win;
  const SubmitRec *submit;      // [S]
  StreamState *state;           // [S]
  const float *acc;             // limiter curve by time index (jr + 4 entries)
  float *hist_y;                // [S][co][kLimDelay]  limiter delay line carried between submits
  float *hist_pk;               // [S][kLimDelay]      peak ring carried between submits
  void *pcm;
  size_t stride_bytes;
  int n_frames, flush, tile;
  int only_irregular;           // 1 = only the streams k_stream / k_pipe left alone (trimmed / missing frames)
  int in_s16;                   // the decoded rows are int16 (IAMFB_IN_S16): widened in the stage right after the copy
};

constexpr int kWmPad = 320;     // scratch beyond the tile: 240 history + 64 (widest doubling step) + 16

// samples per thread in the render / output stages.  The per-tile latency of a block is one thread's serial
// instruction stream, so FEWER samples per thread (more threads per tile) shortens the critical path of a stream
// (template parameter VEC of everything below; the host picks it per plan together with the block size)

template <int VEC>
__device__ __forceinline__ Vec<VEC> ldsv(const float *p) {
  constexpr int kVec = VEC;
  Vec<VEC> r;
  if constexpr (kVec == 4) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    r.v[0] = t.x; r.v[1 % kVec] = t.y; r.v[2 % kVec] = t.z; r.v[3 % kVec] = t.w;
  } else if constexpr (kVec == 2) {
    const float2 t = *reinterpret_cast<const float2 *>(p);
    r.v[0] = t.x; r.v[1 % kVec] = t.y;
  } else {
    r.v[0] = *p;
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ void stsv(float *p, const Vec<VEC> &a) {
  constexpr int kVec = VEC;
  if constexpr (kVec == 4) *reinterpret_cast<float4 *>(p) = make_float4(a.v[0], a.v[1 % kVec], a.v[2 % kVec], a.v[3 % kVec]);
  else if constexpr (kVec == 2) *reinterpret_cast<float2 *>(p) = make_float2(a.v[0], a.v[1 % kVec]);
  else *p = a.v[0];
}
__device__ __forceinline__ const float *byte_off(const float *base, int off) {
  return reinterpret_cast<const float *>(reinterpret_cast<const char *>(base) + off);
}
__device__ __forceinline__ float *byte_off(float *base, int off) {
  return reinterpret_cast<float *>(reinterpret_cast<char *>(base) + off);
}

__device__ __forceinline__ constexpr int fused_order(int layout, int m) {   // IAMF_utils.c:117-133
  constexpr unsigned char kOrder[9][12] = {
      {13}, {14, 15}, {1, 2, 3, 4, 20, 21}, {1, 2, 3, 4, 20, 21, 22, 23}, {1, 2, 3, 4, 20, 21, 9, 10, 11, 12},
      {1, 2, 3, 4, 5, 6, 7, 8}, {1, 2, 3, 4, 5, 6, 7, 8, 22, 23}, {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12},
      {18, 19, 3, 4, 16, 17}};
  return kOrder[layout][m];
}

// x / d, correctly rounded (IEEE-754 binary32, round to nearest even), for the de-mixer's divisors.
// The reference divides by the float constants 1.0, 0.707 and 0.866 (demixer.c:62-72,213-218,262-269,363-368).
// With r = RN(1/d):  q0 = RN(x*r),  rem = x - d*q0 (exact, one FMA),  q = RN(q0 + rem*r)  equals RN(x/d) for EVERY
// finite x with 2^-100 <= |x| < 2^126 - verified exhaustively over all 2^32 inputs for each of the three divisors
// (tools/check_fast_div.c); outside that range (and for d == 0) the ordinary division is used, +-0 maps to q0 = +-0.
template <int VEC>
static __device__ __noinline__ void slow_div4(Vec<VEC> &q, const Vec<VEC> &x, float d) {
#pragma unroll 1
  for (int k = 0; k < VEC; ++k) q.v[k] = x.v[k] / d;
}
template <int VEC>
__device__ __forceinline__ Vec<VEC> exact_div4(const Vec<VEC> &x, float d, float r) {
  constexpr int kVec = VEC;
  typedef Vec<VEC> V4;
  V4 q;
  unsigned int bad = 0u;
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    const float q0 = x.v[k] * r;
    const float rem = __fmaf_rn(-d, q0, x.v[k]);
    q.v[k] = __fmaf_rn(rem, r, q0);
    const unsigned int ax = __float_as_uint(x.v[k]) & 0x7fffffffu;
    bad |= (ax - 0x0d800000u >= 0x7e800000u - 0x0d800000u) ? 1u : 0u;
  }
  if (bad | (d == 0.f ? 1u : 0u)) slow_div4<VEC>(q, x, d);
  return q;
}
static __constant__ float c_mix_beta_r[8] = {1.0f / 1.0f, 1.0f / 0.707f, 1.0f / 0.866f, 0.f, 1.0f / 1.0f, 1.0f / 0.707f, 1.0f / 0.866f, 0.f};
static __constant__ float c_mix_gd_r[8] = {1.0f / 0.707f, 1.0f / 0.707f, 1.0f / 0.866f, 0.f, 1.0f / 0.707f, 1.0f / 0.707f, 1.0f / 0.866f, 0.f};

// Channel-based reconstruction (demixer.c:127-378,421-475) from the staged rows.  in_q = the block's staged tile at
// this thread's four samples; every IAChannel has a byte offset to its row (an all-zero row when absent) and an
// output gain (1.0 by default).  Fills x[m] = layout channel m after recon gain.
template <int LAYOUT, int NREC, int VEC>
__device__ __forceinline__ void fused_reconstruct(const KernelPlan &plan, const FusedArgs &a, const ElPlan &ep,
                                                  const ElFrame &ef, const float *in_q, int i0, Vec<VEC> (&x)[NREC]) {
  constexpr int kVec = VEC;
  typedef Vec<VEC> V4;   // "the thread's samples" (historical name: four when VEC == 4)
  // dmx_gainup (demixer.c:421-430): the output gains scale the transmitted channels in place before anything reads them
  for (int i = 0; i < ep.f_n_gain; ++i) {
    float *p = byte_off(const_cast<float *>(in_q), ep.f_gain_off[i]);
    V4 r = ldsv<VEC>(p);
    const float g = ep.f_gain_val[i];
#pragma unroll
    for (int k = 0; k < kVec; ++k) r.v[k] *= g;
    stsv<VEC>(p, r);
  }
  // a transmitted channel (an all-zero row when absent)
  auto tx = [&](int ch) -> V4 {
    return ldsv<VEC>(byte_off(in_q, ep.f_src_off[ch]));
  };
  const int mode = ef.mode & 7;
  V4 dR2, dL3, dR3, dSL5, dSR5, dBL7, dBR7, dHL, dHR, dHBL, dHBR;
#pragma unroll
  for (int k = 0; k < kVec; ++k)
    dR2.v[k] = dL3.v[k] = dR3.v[k] = dSL5.v[k] = dSR5.v[k] = dBL7.v[k] = dBR7.v[k] = dHL.v[k] = dHR.v[k] = dHBL.v[k] = dHBR.v[k] = 0.f;
  if (ep.need_s2) {   // R2 = 2*Mono - L2, demixer.c:136-138
    const V4 mo = tx(IAMFB_CH_MONO), l2 = tx(IAMFB_CH_L2);
#pragma unroll
    for (int k = 0; k < kVec; ++k) dR2.v[k] = 2 * mo.v[k] - l2.v[k];
  }
  if (ep.need_s3) {   // L3 = L2 - 0.707*C evaluated in double, demixer.c:165-168
    const V4 l2 = tx(IAMFB_CH_L2), r2 = ep.need_s2 ? dR2 : tx(IAMFB_CH_R2), cc = tx(IAMFB_CH_C);
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      const double c = (double)cc.v[k];
      dL3.v[k] = (float)((double)l2.v[k] - 0.707 * c);
      dR3.v[k] = (float)((double)r2.v[k] - 0.707 * c);
    }
  }
  if (ep.need_s5) {   // Ls5 = (L3 - L5)/delta, demixer.c:213-218
    const V4 l3 = ep.need_s3 ? dL3 : tx(IAMFB_CH_L3), r3 = ep.need_s3 ? dR3 : tx(IAMFB_CH_R3);
    const V4 l5 = tx(IAMFB_CH_L5), r5 = tx(IAMFB_CH_R5);
    V4 nl, nr;
#pragma unroll
    for (int k = 0; k < kVec; ++k) { nl.v[k] = l3.v[k] - l5.v[k]; nr.v[k] = r3.v[k] - r5.v[k]; }
    dSL5 = exact_div4<VEC>(nl, c_mix_delta[mode], c_mix_gd_r[mode]);
    dSR5 = exact_div4<VEC>(nr, c_mix_delta[mode], c_mix_gd_r[mode]);
  }
  if (ep.need_s7) {   // Lb7 = (Ls5 - alpha*Lss7)/beta, demixer.c:262-269
    const V4 sl5 = ep.need_s5 ? dSL5 : tx(IAMFB_CH_SL5), sr5 = ep.need_s5 ? dSR5 : tx(IAMFB_CH_SR5);
    const V4 sl7 = tx(IAMFB_CH_SL7), sr7 = tx(IAMFB_CH_SR7);
    const float al = c_mix_alpha[mode];
    V4 nl, nr;
#pragma unroll
    for (int k = 0; k < kVec; ++k) { nl.v[k] = sl5.v[k] - sl7.v[k] * al; nr.v[k] = sr5.v[k] - sr7.v[k] * al; }
    dBL7 = exact_div4<VEC>(nl, c_mix_beta[mode], c_mix_beta_r[mode]);
    dBR7 = exact_div4<VEC>(nr, c_mix_beta[mode], c_mix_beta_r[mode]);
  }
  if (ep.need_h2) {   // Ltf2 = Ltf3 - delta*w*Ls5, demixer.c:318-323
    const V4 sl5 = ep.need_s5 ? dSL5 : tx(IAMFB_CH_SL5), sr5 = ep.need_s5 ? dSR5 : tx(IAMFB_CH_SR5);
    const V4 tl_ = tx(IAMFB_CH_TL), tr_ = tx(IAMFB_CH_TR);
    const float dw = c_mix_delta[mode] * ef.w;
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      dHL.v[k] = tl_.v[k] - dw * sl5.v[k];
      dHR.v[k] = tr_.v[k] - dw * sr5.v[k];
    }
  }
  if (ep.need_h4) {   // Ltb = (Ltf2 - Ltf4)/gamma, demixer.c:363-368
    const V4 hl = ep.need_h2 ? dHL : tx(IAMFB_CH_HL), hr = ep.need_h2 ? dHR : tx(IAMFB_CH_HR);
    const V4 hfl = tx(IAMFB_CH_HFL), hfr = tx(IAMFB_CH_HFR);
    V4 nl, nr;
#pragma unroll
    for (int k = 0; k < kVec; ++k) { nl.v[k] = hl.v[k] - hfl.v[k]; nr.v[k] = hr.v[k] - hfr.v[k]; }
    dHBL = exact_div4<VEC>(nl, c_mix_gamma[mode], c_mix_gd_r[mode]);
    dHBR = exact_div4<VEC>(nr, c_mix_gamma[mode], c_mix_gd_r[mode]);
  }
  // recon-gain cross-fade window of this thread's samples (dmx_rms, demixer.c:461-468): hann start / stop inside the
  // first frame_size/16 samples of the frame, 1 / 0 after
  V4 st, sw;
#pragma unroll
  for (int k = 0; k < kVec; ++k) { st.v[k] = 0.f; sw.v[k] = 1.f; }
  if (ef.rmask && i0 < plan.overlap) {
#pragma unroll
    for (int k = 0; k < kVec; ++k)
      if (i0 + k < plan.overlap) { st.v[k] = a.stop_win[i0 + k]; sw.v[k] = a.start_win[i0 + k]; }
  }
  // gather in layout order; a derived pair replaces the transmitted one exactly when its step ran
#pragma unroll
  for (int m = 0; m < NREC; ++m) {
    const int ch = fused_order(LAYOUT, m);
    bool der = false;
    V4 dv;
    switch (ch) {
      case IAMFB_CH_R2: der = ep.need_s2; dv = dR2; break;
      case IAMFB_CH_L3: der = ep.need_s3; dv = dL3; break;
      case IAMFB_CH_R3: der = ep.need_s3; dv = dR3; break;
      case IAMFB_CH_SL5: der = ep.need_s5; dv = dSL5; break;
      case IAMFB_CH_SR5: der = ep.need_s5; dv = dSR5; break;
      case IAMFB_CH_BL7: der = ep.need_s7; dv = dBL7; break;
      case IAMFB_CH_BR7: der = ep.need_s7; dv = dBR7; break;
      case IAMFB_CH_HL: der = ep.need_h2; dv = dHL; break;
      case IAMFB_CH_HR: der = ep.need_h2; dv = dHR; break;
      case IAMFB_CH_HBL: der = ep.need_h4; dv = dHBL; break;
      case IAMFB_CH_HBR: der = ep.need_h4; dv = dHBR; break;
      default: dv = dR2; break;
    }
    x[m] = der ? dv : tx(ch);
    if ((ef.rmask >> m) & 1u) {   // x *= last*stop[i] + cur*start[i]
      const float lastf = ef.rlast[m], cur = ef.rcur[m];
#pragma unroll
      for (int k = 0; k < kVec; ++k) {
        const float f = lastf * st.v[k] + cur * sw.v[k];
        x[m].v[k] *= f;
      }
    }
  }
}

// One element's contribution for the 4 samples [i0, i0+4) of this thread; the whole tile is rendered as if untrimmed
// (a trimmed tile is compacted afterwards), so every access is a whole aligned quad.
//   in_q  the block's staged tile at the thread's samples (rows tl floats apart)
//   yt    the thread's slots in the mixed time line (&Y[0][ring position], rows rs floats apart)
// DMR: the element is rendered by the parametric down-mixer (DMRenderer_downmix, downmix_renderer.c:218-242) instead of a
// matrix: the output channels are picked from the layout's channels and the ordered two-term sums of their dependencies
template <int LAYOUT, int NREC, int VEC, bool DMR = false>
__device__ __forceinline__ void fused_element(const KernelPlan &plan, const FusedArgs &a, int e, const FrameRec &fr,
                                              int sf, int i0, bool first, bool last, float *in_q, int tl, float *yt, int rs,
                                              float *pkt) {
  constexpr int kVec = VEC;
  typedef Vec<VEC> V4;
  const int N = plan.frame_size;
  const ElPlan &ep = plan.el[e];
  const ElFrame &ef = fr.el[e];
  const int co = plan.out_channels;
  float *ine = byte_off(in_q, ep.f_row_off);
  V4 dv[DMR ? kChCount : 1];   // DMR: every IAChannel's value (inputs, then the down-mixer's dependent channels)
  if constexpr (LAYOUT >= 0) {
    V4 x[NREC];
    fused_reconstruct<LAYOUT, NREC, VEC>(plan, a, ep, ef, in_q, i0, x);
    if constexpr (DMR) {
#pragma unroll
      for (int c = 0; c < kChCount; ++c)
#pragma unroll
        for (int k = 0; k < kVec; ++k) dv[c].v[k] = 0.f;
#pragma unroll
      for (int m = 0; m < NREC; ++m) dv[fused_order(LAYOUT, m)] = x[m];
      dmr_prepare<VEC>(ep, ef, dv);
    } else {
      // the reconstructed layout channels replace the staged rows (this thread's samples only; all its reads are done)
#pragma unroll
      for (int m = 0; m < NREC; ++m) stsv<VEC>(ine + (size_t)m * tl, x[m]);
    }
  } else {
    // scene based: a mono mapping is a row permutation (folded into the matrix offsets), a projection an ordered
    // mat-vec (IAMF_core_decoder.c:105-130) whose result replaces the staged rows
    if (ep.ambi_mode != 0) {
      V4 x[NREC];
#pragma unroll
      for (int m = 0; m < NREC; ++m)
#pragma unroll
        for (int k = 0; k < kVec; ++k) x[m].v[k] = .0f;
#pragma unroll 1
      for (int l = 0; l < ep.ambi_cols; ++l) {
        const V4 t = ldsv<VEC>(ine + (size_t)l * tl);
#pragma unroll
        for (int m = 0; m < NREC; ++m) {
          const float c = ep.ambi_mat[l * NREC + m];
#pragma unroll
          for (int k = 0; k < kVec; ++k) x[m].v[k] += t.v[k] * c;
        }
      }
#pragma unroll
      for (int m = 0; m < NREC; ++m) stsv<VEC>(ine + (size_t)m * tl, x[m]);
    }
  }

  // element mix gain / output mix gain: constants, or per-sample ramps (animated mix gain)
  const float *gr = a.gain_ramp[e], *ogr = last ? a.out_gain_ramp : nullptr;
  const bool eg_on = gr || (ef.gain != 1.f && ef.gain > 0.f);             // iamf_frame_gain, IAMF_decoder.c:1392
  const bool og_on = last && (ogr || (fr.out_gain != 1.f && fr.out_gain > 0.f));
  V4 eg, og;
#pragma unroll
  for (int k = 0; k < kVec; ++k) { eg.v[k] = ef.gain; og.v[k] = fr.out_gain; }
  if (gr || ogr) {
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      const int j = min(max(i0 + k - fr.vstart, 0), N - 1);   // samples outside the trimmed frame are discarded later
      if (gr) eg.v[k] = gr[(size_t)sf * N + j];
      if (ogr) og.v[k] = ogr[(size_t)sf * N + j];
    }
  }
  const bool loud_on = last && plan.loud_gain != 0.f && plan.loud_gain != 1.0f;
  V4 peak;
#pragma unroll
  for (int k = 0; k < kVec; ++k) peak.v[k] = 0.f;

  // render: out = 0; out += mat * in over inputs ascending (m2m_rdr.c:1820-1840, h2m_rdr.c:1103-1112)
#pragma unroll 1
  for (int oc = 0; oc < co; ++oc) {
    V4 y;
#pragma unroll
    for (int k = 0; k < kVec; ++k) y.v[k] = 0.f;
    const int q1 = ep.f_csr_ptr[oc + 1];
    if constexpr (DMR) {
      if (oc < ep.dmr_n_out) y = pick_channel<VEC>(dv, ep.dmr_out_ch[oc]);
    } else if constexpr (LAYOUT >= 0) {
      // channel matrices: a handful of entries per row
#pragma unroll 2
      for (int q = ep.f_csr_ptr[oc]; q < q1; ++q) {
        const float c = ep.f_csr_val[q];
        const V4 xm = ldsv<VEC>(byte_off(in_q, ep.f_csr_off[q]));
#pragma unroll
        for (int k = 0; k < kVec; ++k) y.v[k] += c * xm.v[k];
      }
    } else {
      // HOA matrices are dense (up to 16 entries per row): batch the loads of eight entries ahead of their use
      int q = ep.f_csr_ptr[oc];
      for (; q + 8 <= q1; q += 8) {
        float c[8];
        V4 xm[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i] = ep.f_csr_val[q + i]; xm[i] = ldsv<VEC>(byte_off(in_q, ep.f_csr_off[q + i])); }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int k = 0; k < kVec; ++k) y.v[k] += c[i] * xm[i].v[k];
      }
      for (; q < q1; ++q) {
        const float c = ep.f_csr_val[q];
        const V4 xm = ldsv<VEC>(byte_off(in_q, ep.f_csr_off[q]));
#pragma unroll
        for (int k = 0; k < kVec; ++k) y.v[k] += c * xm.v[k];
      }
    }
    // element mix gain, IAMF_decoder.c:1392-1405
    if (eg_on) {
#pragma unroll
      for (int k = 0; k < kVec; ++k) y.v[k] *= eg.v[k];
    }
    float *dst = yt + (size_t)oc * rs;
    // iamf_mixer_mix, IAMF_decoder.c:2719-2730: acc = 0; acc += e0; acc += e1.  (0 + e0 == e0 bit for bit unless e0 is
    // -0, and a sum that started at +0 is never -0 before it is scaled: only a caller-supplied ramp could make it so)
    if (first) {
      if (gr) {
#pragma unroll
        for (int k = 0; k < kVec; ++k) y.v[k] = 0.f + y.v[k];
      }
    } else {
      const V4 p = ldsv<VEC>(dst);
#pragma unroll
      for (int k = 0; k < kVec; ++k) y.v[k] = p.v[k] + y.v[k];
    }
    if (last) {
      // output mix gain (IAMF_decoder.c:3463-3469) then loudness (:3480-3484)
      if (og_on) {
#pragma unroll
        for (int k = 0; k < kVec; ++k) y.v[k] *= og.v[k];
      }
      if (loud_on) {
#pragma unroll
        for (int k = 0; k < kVec; ++k) y.v[k] *= plan.loud_gain;
      }
#pragma unroll
      for (int k = 0; k < kVec; ++k) peak.v[k] = fmaxf(peak.v[k], fabsf(y.v[k]));
    }
    stsv<VEC>(dst, y);
  }
  if (last && pkt) stsv<VEC>(pkt, peak);
}

static __device__ __noinline__ void store_any(char *out, size_t idx, float x, int bits) {
  if (bits == 16) store_sample<16>(out, idx, x);
  else if (bits == 24) store_sample<24>(out, idx, x);
  else if (bits == 32) store_sample<32>(out, idx, x);
  else store_sample<0>(out, idx, x);
}

// Limiter gain recurrence over the n instants of a tile, run by ONE warp with warp-uniform state (j, S, E).
//   j: number of time-constant increments since the last trigger (pre-increment index of the coming step),
//      j < 0 = never triggered, j >= jr = released (gain 1);  S, E = targetStartGain / targetEndGain.
//   wm[k] = look-ahead peak of instant k, ew[k] = thr / wm[k] (IEEE division, :259), g[k] receives the gain.
// compute_target_gain (audio_effect_peak_limiter.c:237-265): the gain of a step depends only on (S, E, j); a trigger
// (peak * gain > thr) restarts the curve from the current gain.  While no trigger fires the next 32 steps are
// evaluated in parallel, one per lane, and the first lane whose test fires is found with a ballot.  Right after a
// trigger the limiter fires again on every sample for as long as the peak stays in the look-ahead window (the attack
// curve never quite reaches thr/peak): that run is a strictly serial float recurrence  g' = g - acc[1]*(g - thr/peak)
// and is walked sixteen samples at a time speculatively, every lane holding the same values, with nothing but the
// three dependent float operations per sample on the critical path.

constexpr int kAccCache = 128;   // first entries of the limiter curve kept in shared memory (the ones right after a trigger)

static __device__ __noinline__ void fused_scan(const float *wm, const float *ew, float *g, int n, int &j, float &S, float &E,
                                           const float *__restrict__ acc, const float *acc_s, int ja, int jr, float thr, int lane) {
  const float a1 = acc_s[1];
  int pos = 0;
  while (pos < n) {
    // ---- parallel search for the next trigger
    const int k = pos + lane;
    const bool valid = k < n;
    const float p = valid ? wm[k] : 0.f;
    const int jj = j < 0 ? -1 : min(j + lane, jr);
    const bool active = jj >= 0 && jj < jr;
    const float ac = active ? (jj + 1 < kAccCache ? acc_s[jj + 1] : __ldg(acc + jj + 1)) : 0.f;
    const float ga = S - ac * (S - E);
    const float gr = E + ac * (1.0f - E);
    const float gk = active ? (jj < ja ? ga : gr) : 1.0f;
    const bool trig = valid && (p * gk > thr);
    const unsigned mask = __ballot_sync(0xffffffffu, trig);
    if (mask == 0u) {
      const int cnt = min(32, n - pos);
      if (valid) g[k] = gk;
      if (j >= 0) j = min(j + cnt, jr);
      pos += cnt;
      continue;
    }
    const int first = __ffs(mask) - 1;
    if (lane <= first) g[k] = gk;
    S = __shfl_sync(0xffffffffu, gk, first);
    E = __shfl_sync(0xffffffffu, valid ? ew[k] : 0.f, first);
    j = 0;
    pos += first + 1;
    // ---- serial run: the step after a trigger has gain S - acc[1]*(S - E); while it triggers again the state is
    // (S = that gain, E = thr/peak, j = 0) and the next step has the same form.  Up to 32 steps at a time are walked
    // speculatively - nothing but the three dependent float operations per step, thr/peak from 16-byte loads, gains
    // stored four at a time - and then checked in parallel, lane i testing step i; the first step that did not
    // trigger ends the run (its gain is still the right one: it only depends on the trigger before it).
    bool running = true;
    while (running && pos < n) {
      const int B = min(32, (n - pos) & ~3);
      if ((pos & 3) == 0 && B >= 4) {
        float sg = S, se = E;
#pragma unroll 1
        for (int i = 0; i < B; i += 4) {
          const float4 e4 = *reinterpret_cast<const float4 *>(ew + pos + i);
          const float g0 = sg - a1 * (sg - se);
          const float g1 = g0 - a1 * (g0 - e4.x);
          const float g2 = g1 - a1 * (g1 - e4.y);
          const float g3 = g2 - a1 * (g2 - e4.z);
          *reinterpret_cast<float4 *>(g + pos + i) = make_float4(g0, g1, g2, g3);
          sg = g3;
          se = e4.w;
        }
        __syncwarp();
        const bool mine = lane < B;
        const float gl = mine ? g[pos + lane] : 0.f;
        const float pl = mine ? wm[pos + lane] : 0.f;
        const unsigned ok = __ballot_sync(0xffffffffu, !mine || (pl * gl > thr));
        if (ok == 0xffffffffu) {
          S = sg;
          E = se;
          pos += B;
          continue;
        }
        const int f = __ffs(~ok) - 1;          // first step of the burst that did not trigger
        if (f > 0) { S = g[pos + f - 1]; E = ew[pos + f - 1]; }
        j = 1;                                 // the curve continues one increment after the last trigger
        pos += f + 1;
        running = false;
        continue;
      }
      // one step (unaligned position or fewer than four instants left)
      const float gi = S - a1 * (S - E);
      g[pos] = gi;
      if (wm[pos] * gi > thr) {
        S = gi;
        E = ew[pos];
      } else {
        j = 1;
        running = false;
      }
      pos += 1;
    }
  }
}

// VEC samples per thread in the render / output stages, THREADS per block.  The per-tile latency of a block is one
// thread's serial instruction stream: pipelines that fit many streams per SM (small tiles) run 4 samples per thread
// in 64-thread blocks (fewest instructions); pipelines whose rings are large (few streams per SM) spread a tile over
// more, lighter threads to keep the SM's schedulers fed.
template <int L0, int N0, int L1, int N1, int VEC, int THREADS, bool DMR0 = false>
static __global__ void __launch_bounds__(THREADS, (THREADS == 64 ? (DMR0 ? 5 : 7) : (THREADS == 128 ? 4 : 2))) k_fused(const __grid_constant__ KernelPlan plan, FusedArgs a) {
  constexpr int kVec = VEC;
  constexpr int kFusedThreads = THREADS;
  typedef Vec<VEC> V4;
  extern __shared__ __align__(128) float fsm[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ __align__(16) FrameRec s_fr;      // resolved parameters of the frame being rendered
  __shared__ int s_lim[4];
  __shared__ int s_next[2];                    // (frame, offset) of the tile after the current one
  __shared__ float s_acc[kAccCache];
  const int s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // co-resident blocks run their serial scans on different warp schedulers
  const int scan_warp = (blockIdx.x / 148) % (kFusedThreads / 32);
  const int N = plan.frame_size, co = plan.out_channels;
  const int H = plan.limiter ? kLimDelay : 0;
  const int TL = a.tile;
  const int C = H + TL;                         // capacity of the time-line rings (multiple of 4)
  const int nin0 = plan.el[0].n_in, nin1 = N1 > 0 ? plan.el[1].n_in : 0;
  float *IN = fsm;                              // [nin0 + nin1 + 1][TL] staged decoded rows + one all-zero row
  float *Y = IN + (size_t)(nin0 + nin1 + 1) * TL;   // [co][C] mixed time line, circular
  float *PK = Y + (size_t)co * C;               // [C] per-instant peak, circular
  float *WM = PK + C;                           // [TL] sliding maximum of the tile's instants
  float *SA = WM + TL;                          // [TL + kWmPad] scratch; later thr / WM
  float *SB = SA + TL + kWmPad;                 // [TL + kWmPad] scratch; later the gains
  float *EW = SA, *G = SB;

  const SubmitRec sr = a.submit[s];
  if (a.only_irregular && !sr.irregular) return;   // rendered by k_stream
  int lj = -1;
  float lS = -1.f, lE = -1.f;
  for (int i = tid; i < TL; i += kFusedThreads) IN[(size_t)(nin0 + nin1) * TL + i] = 0.f;
  if (plan.limiter) {
    for (int i = tid; i < kAccCache; i += kFusedThreads) s_acc[i] = i <= plan.lim_jr + 3 ? a.acc[i] : 0.f;
    const StreamState &st = a.state[s];
    lj = st.lim_j; lS = st.lim_start; lE = st.lim_end;
    if (lj > plan.lim_jr) lj = plan.lim_jr;
#pragma unroll 1
    for (int c = 0; c <= co; ++c) {
      const float *src = c < co ? a.hist_y + ((size_t)s * co + c) * kLimDelay : a.hist_pk + (size_t)s * kLimDelay;
      float *row = c < co ? Y + (size_t)c * C : PK;
      for (int i = tid; i < kLimDelay; i += kFusedThreads) row[i] = src[i];
    }
    for (int i = tid; i < TL; i += kFusedThreads) PK[H + i] = 0.f;
    for (int i = tid; i < TL + kWmPad; i += kFusedThreads) { SA[i] = 0.f; SB[i] = 0.f; }
  }
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- tile iterator over (frame, offset): only tiles with at least one sample left after trimming
  const int n_frames = a.flush ? 0 : a.n_frames;
  auto tile_range = [&](int f, int t_off, int &lo_t, int &hi_t) {
    const FrameRec &fr = a.frames[s * a.n_frames + f];
    const int vs = fr.vstart, vl = fr.vlen;
    lo_t = max(t_off, vs);
    hi_t = min(min(t_off + TL, N), vs + vl);
  };
  auto advance = [&](int &f, int &t_off) {      // next non-empty tile after (f, t_off); f == n_frames when done
    for (;;) {
      t_off += TL;
      if (t_off >= N) { t_off = 0; ++f; }
      if (f >= n_frames) return;
      int lo_t, hi_t;
      tile_range(f, t_off, lo_t, hi_t);
      if (hi_t > lo_t) return;
    }
  };
  auto issue = [&](int f, int t_off) {          // one thread: bulk copies of the tile's rows into IN
    const int len = min(TL, N - t_off);
    // int16 rows (IAMFB_IN_S16) land in the upper half of their float32 row and are widened in place after the copy
    const int esz = a.in_s16 ? 2 : 4;
    const uint32_t row_bytes = (uint32_t)(len * esz);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&s_bar, row_bytes * (uint32_t)(nin0 + nin1));
    const size_t sf = (size_t)s * a.n_frames + f;
    const char *g0 = reinterpret_cast<const char *>(a.in[0]) + (sf * nin0 * N + t_off) * esz;
    char *d0 = reinterpret_cast<char *>(IN) + (a.in_s16 ? TL * 2 : 0);
#pragma unroll 1
    for (int r = 0; r < nin0; ++r) bulk_g2s(d0 + (size_t)r * TL * 4, g0 + (size_t)r * N * esz, row_bytes, &s_bar);
    if constexpr (N1 > 0) {
      const char *g1 = reinterpret_cast<const char *>(a.in[1]) + (sf * nin1 * N + t_off) * esz;
#pragma unroll 1
      for (int r = 0; r < nin1; ++r) bulk_g2s(d0 + (size_t)(nin0 + r) * TL * 4, g1 + (size_t)r * N * esz, row_bytes, &s_bar);
    }
  };

  char *out = (char *)a.pcm + (size_t)s * a.stride_bytes;
  int lim_done = 0;                             // limiter-stage instants of this submit already processed
  uint32_t parity = 0;
  int w = H;                                    // ring position of the next tile's first instant (16-byte aligned)
  if (tid == 0) {
    int f0 = 0, t0 = -TL;
    if (n_frames > 0) {
      advance(f0, t0);
      if (f0 < n_frames) issue(f0, t0);
    }
    s_next[0] = f0; s_next[1] = t0;
  }
  __syncthreads();
  int f = s_next[0], t_off = s_next[1];
  int fr_loaded = -1;
  bool flush_pending = a.flush != 0;

  while (f < n_frames || flush_pending) {
    int n, shift = 0;
    if (flush_pending) {
      // end of stream: the limiter is fed 240 zeros (iamf_delay_buffer_handle, IAMF_decoder.c:3250-3301)
      n = min(kLimDelay - lim_done, TL);
      for (int i = tid; i < n; i += kFusedThreads) {
        int pos = w + i;
        if (pos >= C) pos -= C;
        for (int c = 0; c < co; ++c) Y[(size_t)c * C + pos] = 0.f;
        PK[pos] = 0.f;
      }
      if (lim_done + n >= kLimDelay) flush_pending = false;
    } else {
      const int sf = s * a.n_frames + f;
      if (fr_loaded != f) {
        // this frame's resolved parameters -> shared memory (one coalesced read instead of per-use global loads)
        const int *src = reinterpret_cast<const int *>(a.frames + sf);
        int *dst = reinterpret_cast<int *>(&s_fr);
        for (int i = tid; i < (int)(sizeof(FrameRec) / 4); i += kFusedThreads) dst[i] = src[i];
        fr_loaded = f;
        __syncthreads();
      }
      const FrameRec &fr = s_fr;
      const int lo_t = max(t_off, fr.vstart), hi_t = min(min(t_off + TL, N), fr.vstart + fr.vlen);
      n = hi_t - lo_t;
      shift = lo_t - t_off;
      // ---------------------------------------------------------------- render (the whole tile, as if untrimmed)
      mbar_wait(&s_bar, parity);
      parity ^= 1u;
      const int t_end = min(t_off + TL, N);
      if (a.in_s16) {
        // x / 32768 (opus/IAMF_opus_decoder.c:133-135; exact): every thread reads its samples of every row, then all write
        const int len8 = (t_end - t_off) >> 3;                       // (tiles of int16 plans are multiples of 8 samples)
        constexpr int kW = (1024 / 8 + kFusedThreads - 1) / kFusedThreads;
#pragma unroll 1
        for (int r = 0; r < nin0 + nin1; ++r) {
          float *row = IN + (size_t)r * TL;
          const int4 *src = reinterpret_cast<const int4 *>(reinterpret_cast<const char *>(row) + TL * 2);
          int4 v[kW];
#pragma unroll
          for (int u = 0; u < kW; ++u) {
            const int i = tid + u * kFusedThreads;
            v[u] = i < len8 ? src[i] : make_int4(0, 0, 0, 0);
          }
          __syncthreads();
#pragma unroll
          for (int u = 0; u < kW; ++u) {
            const int i = tid + u * kFusedThreads;
            if (i < len8) {
              const int w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
              float o[8];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                o[2 * k] = (float)(short)(w[k] & 0xffff) / 32768.f;
                o[2 * k + 1] = (float)(short)(w[k] >> 16) / 32768.f;
              }
              reinterpret_cast<float4 *>(row)[2 * i] = make_float4(o[0], o[1], o[2], o[3]);
              reinterpret_cast<float4 *>(row)[2 * i + 1] = make_float4(o[4], o[5], o[6], o[7]);
            }
          }
        }
        __syncthreads();
      }
      for (int i0 = t_off + tid * kVec; i0 < t_end; i0 += kFusedThreads * kVec) {
        const int q = i0 - t_off;                         // position inside the staged tile
        int pos = w + q;
        if (pos >= C) pos -= C;
        float *yt = Y + pos;
        float *pkt = plan.limiter ? PK + pos : nullptr;
        fused_element<L0, N0, VEC, DMR0>(plan, a, 0, fr, sf, i0, true, N1 == 0, IN + q, TL, yt, C, pkt);
        if constexpr (N1 > 0) fused_element<L1, N1, VEC>(plan, a, 1, fr, sf, i0, false, true, IN + q, TL, yt, C, pkt);
      }
      // a trimmed tile: move its surviving samples [lo_t, hi_t) to the front of the tile (iamf_frame_trim,
      // IAMF_decoder.c:1361-1381); rare (first / last frames), one row at a time through registers
      if (shift > 0) {
#pragma unroll 1
        for (int r = 0; r <= co; ++r) {
          if (r == co && !plan.limiter) break;
          float *row = r < co ? Y + (size_t)r * C : PK;
          __syncthreads();
          float t[1024 / kFusedThreads];
#pragma unroll
          for (int u = 0; u < 1024 / kFusedThreads; ++u) {
            const int i = tid + u * kFusedThreads;
            int pos = w + shift + i;
            if (pos >= C) pos -= C;
            t[u] = i < n ? row[pos] : 0.f;
          }
          __syncthreads();
#pragma unroll
          for (int u = 0; u < 1024 / kFusedThreads; ++u) {
            const int i = tid + u * kFusedThreads;
            int pos = w + i;
            if (pos >= C) pos -= C;
            if (i < n) row[pos] = t[u];
          }
        }
      }
    }
    __syncthreads();
    // the staged rows are consumed: one thread finds the next tile and starts its copy under the rest of this one
    if (!a.flush && tid == 0) {
      int f2 = f, t2 = t_off;
      advance(f2, t2);
      if (f2 < n_frames) issue(f2, t2);
      s_next[0] = f2; s_next[1] = t2;
    }
    const bool quads = (n & 3) == 0;           // every instant of the tile sits in a whole aligned quad
    bool apply_gain = false;
    if (plan.limiter) {
      // ---------------------------------------------------------------- sliding maximum over 240 instants
      // logical index i <-> ring position (w - 240 + i) mod C;  WM[k] = max(PK[k .. k+239]) for k < n
      bool hot = false;
      const float thr = plan.lim_thr;
      int base = w - kLimDelay;
      if (base < 0) base += C;
      if (n % kLimDelay == 0) {
        // van Herk / Gil-Werman with blocks of exactly one window: the logical array (240 of history, then the tile)
        // is cut into blocks B_j of 240; the window starting at 240j + r is the suffix of B_j from r plus the prefix
        // of B_j+1 up to r-1.  Every block's prefix / suffix maxima come from one warp-wide max-scan (8 instants per
        // lane, 30 lanes), so the whole stage needs one block barrier.  SA[240j + r] = max(B_j[r ..]),
        // SB[240j + r] = max(B_j+1[.. r-1]) (0 for r = 0: peaks are >= 0).
        const int nblk = n / kLimDelay;
        for (int job = warp; job < 2 * nblk; job += kFusedThreads / 32) {
          const bool suffix = job < nblk;
          const int j = suffix ? job : job - nblk + 1;       // block of the logical array
          float v[8];
          if (lane < 30) {
            int p0 = base + j * kLimDelay + 8 * lane;
            if (p0 >= C) p0 -= C;
            if (p0 >= C) p0 -= C;
            int p1 = p0 + 4;
            if (p1 >= C) p1 -= C;
            const float4 A = *reinterpret_cast<const float4 *>(PK + p0);
            const float4 B = *reinterpret_cast<const float4 *>(PK + p1);
            v[0] = A.x; v[1] = A.y; v[2] = A.z; v[3] = A.w; v[4] = B.x; v[5] = B.y; v[6] = B.z; v[7] = B.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
          }
          if (suffix) {
#pragma unroll
            for (int i = 6; i >= 0; --i) v[i] = fmaxf(v[i], v[i + 1]);      // v[i] = max of the lane's instants i..7
            float t = v[0];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const float o = __shfl_down_sync(0xffffffffu, t, d);
              if (lane + d < 32) t = fmaxf(t, o);
            }
            float ex = __shfl_down_sync(0xffffffffu, t, 1);                 // maximum of all later lanes
            if (lane == 31) ex = 0.f;
            if (lane < 30) {
              float *dst = SA + (j * kLimDelay + 8 * lane);
              *reinterpret_cast<float4 *>(dst) = make_float4(fmaxf(v[0], ex), fmaxf(v[1], ex), fmaxf(v[2], ex), fmaxf(v[3], ex));
              *reinterpret_cast<float4 *>(dst + 4) = make_float4(fmaxf(v[4], ex), fmaxf(v[5], ex), fmaxf(v[6], ex), fmaxf(v[7], ex));
            }
          } else {
#pragma unroll
            for (int i = 1; i < 8; ++i) v[i] = fmaxf(v[i], v[i - 1]);       // v[i] = max of the lane's instants 0..i
            float t = v[7];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const float o = __shfl_up_sync(0xffffffffu, t, d);
              if (lane >= d) t = fmaxf(t, o);
            }
            float ex = __shfl_up_sync(0xffffffffu, t, 1);                   // maximum of all earlier lanes
            if (lane == 0) ex = 0.f;
            if (lane < 30) {
              float *dst = SB + ((j - 1) * kLimDelay + 8 * lane);            // shifted by one: prefix up to r-1
              *reinterpret_cast<float4 *>(dst) = make_float4(ex, fmaxf(v[0], ex), fmaxf(v[1], ex), fmaxf(v[2], ex));
              *reinterpret_cast<float4 *>(dst + 4) = make_float4(fmaxf(v[3], ex), fmaxf(v[4], ex), fmaxf(v[5], ex), fmaxf(v[6], ex));
            }
          }
        }
        __syncthreads();
        for (int v4 = tid; 4 * v4 < n; v4 += kFusedThreads) {
          const float4 A = *reinterpret_cast<const float4 *>(SA + 4 * v4);
          const float4 B = *reinterpret_cast<const float4 *>(SB + 4 * v4);
          const float4 W = make_float4(fmaxf(A.x, B.x), fmaxf(A.y, B.y), fmaxf(A.z, B.z), fmaxf(A.w, B.w));
          *reinterpret_cast<float4 *>(WM + 4 * v4) = W;
          hot |= (W.x > thr) || (W.y > thr) || (W.z > thr) || (W.w > thr);
        }
      } else {
        // any other tile length (trimmed frames, frame sizes that are no multiple of 240): log-step doubling
        const int span4 = (n + kLimDelay + 3) >> 2;            // 16-byte items covering the window
        for (int v = tid; v < span4; v += kFusedThreads) {     // windows of 8, in registers
          int p0 = base + 4 * v;
          if (p0 >= C) p0 -= C;
          int p1 = p0 + 4;
          if (p1 >= C) p1 -= C;
          int p2 = p1 + 4;
          if (p2 >= C) p2 -= C;
          const float4 A = *reinterpret_cast<const float4 *>(PK + p0);
          const float4 B = *reinterpret_cast<const float4 *>(PK + p1);
          const float4 Cq = *reinterpret_cast<const float4 *>(PK + p2);
          const float m47 = fmaxf(fmaxf(B.x, B.y), fmaxf(B.z, B.w));
          const float s3 = A.w, s2 = fmaxf(A.z, s3), s1 = fmaxf(A.y, s2), s0 = fmaxf(A.x, s1);
          const float p9 = fmaxf(Cq.x, Cq.y), p10 = fmaxf(p9, Cq.z);
          *reinterpret_cast<float4 *>(SA + 4 * v) =
              make_float4(fmaxf(s0, m47), fmaxf(fmaxf(s1, m47), Cq.x), fmaxf(fmaxf(s2, m47), p9), fmaxf(fmaxf(s3, m47), p10));
        }
        __syncthreads();
        float *src = SA, *dst = SB;
#pragma unroll 1
        for (int d = 8; d <= 64; d <<= 1) {                    // windows of 16, 32, 64, 128
          for (int v = tid; v < span4; v += kFusedThreads) {
            const float4 A = *reinterpret_cast<const float4 *>(src + 4 * v);
            const float4 B = *reinterpret_cast<const float4 *>(src + 4 * v + d);
            *reinterpret_cast<float4 *>(dst + 4 * v) = make_float4(fmaxf(A.x, B.x), fmaxf(A.y, B.y), fmaxf(A.z, B.z), fmaxf(A.w, B.w));
          }
          __syncthreads();
          float *t = src; src = dst; dst = t;
        }
        // four passes: the windows of 128 are back in SA
        for (int v = tid; 4 * v < n; v += kFusedThreads) {     // 240 = two overlapping windows of 128
          const float4 A = *reinterpret_cast<const float4 *>(SA + 4 * v);
          const float4 B = *reinterpret_cast<const float4 *>(SA + 4 * v + 112);
          const float4 W = make_float4(fmaxf(A.x, B.x), fmaxf(A.y, B.y), fmaxf(A.z, B.z), fmaxf(A.w, B.w));
          *reinterpret_cast<float4 *>(WM + 4 * v) = W;
          const int left = n - 4 * v;
          hot |= (W.x > thr) || (left > 1 && W.y > thr) || (left > 2 && W.z > thr) || (left > 3 && W.w > thr);
        }
      }
      const int any_hot = __syncthreads_or(hot ? 1 : 0);
      // ---------------------------------------------------------------- gain recurrence
      const bool idle = lj < 0 || lj >= plan.lim_jr;         // block-uniform
      if (any_hot || !idle) {
        apply_gain = true;
        // thr / peak (targetEndGain of a trigger, :259): only tiles that can trigger need it
        for (int v = tid; 4 * v < n; v += kFusedThreads) {
          const float4 W = *reinterpret_cast<const float4 *>(WM + 4 * v);
          *reinterpret_cast<float4 *>(EW + 4 * v) = make_float4(thr / W.x, thr / W.y, thr / W.z, thr / W.w);
        }
        __syncthreads();
        if (warp == scan_warp) {
          fused_scan(WM, EW, G, n, lj, lS, lE, a.acc, s_acc, plan.lim_ja, plan.lim_jr, thr, lane);
          if (lane == 0) { s_lim[0] = lj; s_lim[1] = __float_as_int(lS); s_lim[2] = __float_as_int(lE); }
        }
        __syncthreads();
        lj = s_lim[0]; lS = __int_as_float(s_lim[1]); lE = __int_as_float(s_lim[2]);
      }
    }
    // ------------------------------------------------------------------ output: delayed sample x gain -> PCM
    {
      const long long base_o = (long long)lim_done - sr.out_skip;   // output index of instant 0 of this tile
      const int bits = plan.bit_depth;
      int rbase = w - H;                                            // ring position of the delayed instant 0
      if (rbase < 0) rbase += C;
      for (int k4 = tid * kVec; k4 < n; k4 += kFusedThreads * kVec) {
        const long long o0 = base_o + k4;
        const bool full = (k4 + kVec <= n) && o0 >= 0;
        int pos = rbase + k4;
        if (pos >= C) pos -= C;
        if (full && bits == 16 && (co & 1) == 0 && (((o0 * co) & 1) == 0)) {
          V4 gg;
#pragma unroll
          for (int u = 0; u < kVec; ++u) gg.v[u] = 1.f;
          if (apply_gain) gg = ldsv<VEC>(G + k4);
          uint32_t *wq = (uint32_t *)((int16_t *)out + o0 * co);
          const int half = co >> 1;
#pragma unroll 1
          for (int c = 0; c < co; c += 2) {
            const V4 v0 = ldsv<VEC>(Y + (size_t)c * C + pos), v1 = ldsv<VEC>(Y + (size_t)(c + 1) * C + pos);
#pragma unroll
            for (int u = 0; u < kVec; ++u) {
              float y0 = v0.v[u], y1 = v1.v[u];
              if (plan.limiter) { y0 = y0 * gg.v[u]; y1 = y1 * gg.v[u]; }
              wq[u * half + (c >> 1)] = (uint32_t)(quant16(y0) & 0xffff) | ((uint32_t)quant16(y1) << 16);
            }
          }
        } else {
#pragma unroll 1
          for (int c = 0; c < co; ++c) {
#pragma unroll 1
            for (int u = 0; u < kVec; ++u) {
              if (k4 + u >= n || o0 + u < 0) continue;
              int pu = pos + u;
              if (pu >= C) pu -= C;
              float x = Y[(size_t)c * C + pu];
              if (plan.limiter) x = x * (apply_gain ? G[k4 + u] : 1.0f);
              store_any(out, (size_t)(o0 + u) * co + c, x, bits);
            }
          }
        }
      }
    }
    lim_done += n;
    // ------------------------------------------------------------------ advance the ring
    if (quads && shift == 0) {
      w += n;
      if (w >= C) w -= C;
    } else if (plan.limiter) {
      // an irregular tile (trimmed, or not a multiple of four instants) would leave the ring misaligned: move the
      // last 240 instants to the front instead and restart at position 240
      __syncthreads();
      int from = w + n - kLimDelay;
      if (from < 0) from += C;
#pragma unroll 1
      for (int r = 0; r <= co; ++r) {
        float *row = (r < co) ? Y + (size_t)r * C : PK;
        float t[(kLimDelay + kFusedThreads - 1) / kFusedThreads];
#pragma unroll
        for (int u = 0; u < (kLimDelay + kFusedThreads - 1) / kFusedThreads; ++u) {
          const int i = tid + u * kFusedThreads;
          int pos = from + i;
          if (pos >= C) pos -= C;
          t[u] = i < kLimDelay ? row[pos] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < (kLimDelay + kFusedThreads - 1) / kFusedThreads; ++u) {
          const int i = tid + u * kFusedThreads;
          if (i < kLimDelay) row[i] = t[u];
        }
      }
      w = kLimDelay;
    }
    __syncthreads();
    if (!a.flush) { f = s_next[0]; t_off = s_next[1]; }
  }

  if (plan.limiter) {
    // the last 240 instants, in time order, are the history of the next submit
    int from = w - kLimDelay;
    if (from < 0) from += C;
#pragma unroll 1
    for (int c = 0; c <= co; ++c) {
      float *dst = c < co ? a.hist_y + ((size_t)s * co + c) * kLimDelay : a.hist_pk + (size_t)s * kLimDelay;
      const float *row = c < co ? Y + (size_t)c * C : PK;
      for (int i = tid; i < kLimDelay; i += kFusedThreads) {
        int pos = from + i;
        if (pos >= C) pos -= C;
        dst[i] = row[pos];
      }
    }
    if (tid == 0) {
      StreamState &st = a.state[s];
      st.lim_j = lj; st.lim_start = lS; st.lim_end = lE;
    }
  }
}

}  // namespace iamfb
