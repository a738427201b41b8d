// iamfb_resample_ls.cuh - k_resample_ls: the interpolating resampler (resampler_basic_interpolate_single, resample.c:357-418)
// with the LANES OF A WARP ON 32 DIFFERENT STREAMS.
//
// Why: in a block-per-stream FIR every lane owns outputs of its own phase, so every lane needs its own four-tap item per
// step - 16 bytes of shared-memory traffic per 8 multiply-adds, which is what bounds k_pipe_rs (shared-memory pipe 75 %
// busy, FP32 pipe 58 %).  Streams that run in step have the SAME phase at the same output index: with one
// stream per lane the tap items are warp-uniform (one broadcast read per warp instead of 32 distinct ones) and only the
// inputs are per lane - the loop is left with the FP32 pipe as its bound (an 8-byte input per lane and four broadcast 16-byte
// tap items per step: 10 shared-memory wavefronts against 64 cycles of FP32 work; measured: DESIGN.md 4.3).
//
// Work item = (group of 32 streams, kLsTeam chunks of `chunk` consecutive outputs, one per warp of a team).  The team stages, for each stream, the
// inputs the chunk spans (channel pairs interleaved, row stride odd: lanes hit different banks) from the pre-resample time
// line tl_a[S][co][cap_a] (k_pipe_prerender's output; the previous submit's last rs_hist inputs in front), then walks the chunk in
// groups of 4 consecutive outputs exactly like k_pipe_rs: one pass over the inputs of the four outputs, the rows of the
// tap table carrying zero items outside each output's window (adding +-0 to a sum that started at +0 is exact), four
// accumulators per channel over j ascending, packed exact multiply-add for the two channels of a pair, cubic blend with the
// per-phase weights tabulated at plan time, clamp to +-1 (resample.c:84), loudness.  Bit for bit the reference's sums.
//
// Lanes whose streams are at another phase (a stream that joined the batch later, lost a frame, ...) are served in further
// passes of the same warp (grouped by the phase of the chunk's first output); irregular streams of the submit (trims,
// missing frames) are skipped - the multi-kernel path renders them.  The kernel is persistent: one block per SM
// (the block shape and the shared staging areas: below) strides over the work items; k_pipe_rs<PRE> (the limiter half) follows.
// Measured with the limiter half BESIDE this kernel (chunks handed over through flags): no gain - the staging areas leave
// no shared memory for its blocks, and with fewer resampler warps the FP32 pipe idles more than the overlap saves.
#pragma once
#include "iamfb_pipe.cuh"

namespace iamfb {

struct ResampleLsArgs {
  const float *src;             // tl_a [S][co][cap_a]; this submit's first input instant at index rs_hist
  float *dst;                   // tl_b [S][co][cap_b]; output u of this submit at hist_b + u
  const SubmitRec *submit;      // [S] (irregular streams are skipped)
  const StreamState *state;     // [S]
  const float4 *tab4;           // [oversample][tab_row] tap items, tab_pad zero items on either side of a row
  const float4 *interp4;        // [den] cubic weights per phase
  int tab_row, tab_pad;
  int cap_a, cap_b, hist_b;
  int n_streams, co;
  int chunk;                    // outputs per work item (multiple of 4)
  int n_chunks;                 // chunks that cover the longest stream of the submit
  int span;                     // inputs staged per stream and chunk (row stride in float2, odd)
  float neg_zero;
};

// Block shape (compile-time; -D overrides exist for A/B builds, tools/gpu_variants.sh).  12 warps per SM in 3 TEAMS of 4: the
// warps of a team share ONE staging area that covers 4 x chunk consecutive outputs of 32 streams - neighbouring chunks'
// inputs overlap by the filter length, so one area for four chunks needs 45 % less shared memory per warp than an area
// each (which is what lets three warps per scheduler fit the SM), and a team's four warps sit on the four schedulers.
// Measured on C5 (ms per launch): an area per warp, 8 warps 1.36; teams of 2 (12 warps) 1.30; of 3 1.248; of 4 1.233;
// of 6 1.244; 16 warps in teams of 4 / 8: 1.28 / 1.29 (128 registers per thread and smaller chunks).  Also measured on the
// final shape: no start offset between the teams 1.240; the head and the tail of the tap walk peeled so that the zero pad
// items are skipped (4.5 % fewer multiply-adds, but two rolled loops around the unrolled one) 1.249.
#ifndef IAMFB_LS_TEAM
#define IAMFB_LS_TEAM 4
#endif
#ifndef IAMFB_LS_WARPS
#define IAMFB_LS_WARPS 12
#endif
constexpr int kLsTeam = IAMFB_LS_TEAM;      // warps that share one staging area
constexpr int kLsWarps = IAMFB_LS_WARPS;    // warps per block (one block per SM)
static_assert(kLsWarps % kLsTeam == 0 && kLsWarps / kLsTeam <= 15 && kLsWarps <= 32, "teams tile the block; one named barrier (1..15) per team");

__device__ __forceinline__ void ls_cp4(void *dst_smem, const void *src, bool valid) {   // 4-byte asynchronous copy, zeros when !valid
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1) k_resample_ls(const __grid_constant__ KernelPlan plan, ResampleLsArgs a) {
  extern __shared__ __align__(16) float ls_smem[];
  const int Nf = (int)plan.rs_filt_len, os = (int)plan.rs_oversample, den = (int)plan.rs_den;
  const int fa = plan.rs_frac_adv, ia = plan.rs_int_adv;
  const int trow = a.tab_row, pad = a.tab_pad;
  float4 *TAB = reinterpret_cast<float4 *>(ls_smem);                             // [os][trow]
  const int tab_items = os * trow;
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  float4 *CI = TAB + tab_items;                                                    // [den] cubic weights per phase
  const int pi = wi / kLsTeam, half = wi % kLsTeam;                                // team, this warp's part of the team's outputs
  float2 *X = reinterpret_cast<float2 *>(CI + den + (den + 3) / 4) + (size_t)pi * 32 * a.span;    // [32 streams][span] channel pairs
  for (int i = threadIdx.x; i < tab_items; i += NWARPS * 32) TAB[i] = a.tab4[i];
  for (int i = threadIdx.x; i < den; i += NWARPS * 32) CI[i] = a.interp4[i];
  int *OFFROW = reinterpret_cast<int *>(CI + den);                                 // [den] first item of the phase's tap row
  for (int i = threadIdx.x; i < den; i += NWARPS * 32) OFFROW[i] = (i * os / den) * trow + pad;
  __syncthreads();

  const int groups = (a.n_streams + 31) / 32;
  const int steps = Nf + 3 * (ia + 1);
  const bool loud_on = plan.loud_gain != 0.f && plan.loud_gain != 1.0f;
  const int co = a.co;
  const long long total = (long long)groups * a.n_chunks;       // (n_chunks counts the teams' items of kLsTeam x chunk outputs)
  const int teams = gridDim.x * (NWARPS / kLsTeam);
  // the teams take turns: every other team starts half an item late, so that one stages while the others keep the FP32
  // pipe busy - all items last the same, the offset persists
  if (pi & 1) {
    const long long t0 = clock64(), d = (long long)(a.chunk / 4) * steps * 64;
    while (clock64() - t0 < d) __nanosleep(200);
  }
  const int bar_id = 1 + pi;              // named barrier of the team (0 is the block's)

  for (long long item = (long long)blockIdx.x + (long long)pi * gridDim.x; item < total; item += teams) {
    const int c = (int)(item / groups), g = (int)(item - (long long)c * groups);     // chunks in order, groups fastest
    const int s = g * 32 + lane;
    const int us = c * kLsTeam * a.chunk; // first output of the team's item
    const int u0 = us + half * a.chunk;   // first output of this warp's part
    // this lane's stream: length; position of the item's first input (what is staged); position and phase of the part's
    // first output.  All warps of the team compute the same for the item.
    int L = 0, pos_s = 0, pos0 = 0, phi0 = -1;
    bool act_s = false;
    if (s < a.n_streams) {
      const SubmitRec sr = a.submit[s];
      if (!sr.irregular && us < sr.lim_len) {
        L = sr.lim_len;
        act_s = true;
        const long long in_start = a.state[s].rs_in_total - sr.in_len;
        // q(n) = Nf/2 + floor(n * num / den) is the LAST input of output n; index into the tl_a row
        const long long ns = sr.rs_out_first + us;
        pos_s = (int)((long long)(Nf / 2) + (ns * (long long)plan.rs_num) / den - (Nf - 1) - in_start + plan.rs_hist);
        if (u0 < L) {
          const long long n = sr.rs_out_first + u0;
          pos0 = (int)((long long)(Nf / 2) + (n * (long long)plan.rs_num) / den - (Nf - 1) - in_start + plan.rs_hist);
          phi0 = (int)((n * (long long)fa) % den);
        }
      }
    }
    if (__ballot_sync(0xffffffffu, act_s) == 0u) continue;       // (the same decision in every warp of the team)
    const unsigned todo = __ballot_sync(0xffffffffu, phi0 >= 0);
    for (int c0 = 0; c0 < co; c0 += 2) {
      const bool two = c0 + 1 < co;
      // ---- stage the inputs of the item: row st of X = stream 32 g + st, entries [pos_s(st), pos_s(st) + span); the warps
      // of the team bring the rows in turn
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kLsTeam * 32) : "memory");     // (the others have finished with the previous rows)
      // (asynchronous 4-byte copies, all of them in flight at once: the rows of 32 streams are 32 different places in memory)
      {
        // element index of the lane's own first staged input, < 0 for a lane without work
        const long long my_off = act_s ? ((long long)s * co + c0) * a.cap_a + pos_s : -1;
        const bool fast = __all_sync(0xffffffffu, !act_s || (pos_s >= 0 && pos_s + a.span <= a.cap_a));
        const uint32_t xb = (uint32_t)__cvta_generic_to_shared(X) + (uint32_t)lane * 8u;
        if (fast) {
#pragma unroll 4
          for (int st = half; st < 32; st += kLsTeam) {
            const long long o = __shfl_sync(0xffffffffu, my_off, st);
            const bool on = o >= 0;                                                   // (warp-uniform)
            const float *r0 = a.src + (on ? o : 0) + lane;
            const float *r1 = two ? r0 + a.cap_a : r0;
            uint32_t d = xb + (uint32_t)(st * a.span) * 8u;
            for (int k = lane; k < a.span; k += 32) {
              const int n0 = on ? 4 : 0, n1 = (on && two) ? 4 : 0;
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(r0), "r"(n0) : "memory");
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d + 4u), "l"(r1), "r"(n1) : "memory");
              r0 += 32; r1 += 32; d += 256u;
            }
          }
        } else {
#pragma unroll 1
          for (int st = half; st < 32; st += kLsTeam) {
            const int p0 = __shfl_sync(0xffffffffu, pos_s, st);
            const int on = __shfl_sync(0xffffffffu, act_s ? 1 : 0, st);
            if (!on) continue;                                                         // (warp-uniform)
            const float *r0 = a.src + ((size_t)(g * 32 + st) * co + c0) * a.cap_a;
            const float *r1 = two ? r0 + a.cap_a : r0;
            for (int k = lane; k < a.span; k += 32) {
              const int idx = p0 + k;
              const bool in = idx >= 0 && idx < a.cap_a;
              const int ic = in ? idx : 0;
              float2 *d = X + st * a.span + k;
              ls_cp4(&d->x, r0 + ic, in);
              ls_cp4(&d->y, r1 + ic, in && two);
            }
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kLsTeam * 32) : "memory");     // (every warp's rows are in)
      // ---- passes over the phases present in the warp (one when the streams run in step)
      unsigned left = todo;
      while (left) {
        const int leader = __ffs(left) - 1;
        const int phl = __shfl_sync(0xffffffffu, phi0, leader);
        const unsigned mine = __ballot_sync(0xffffffffu, phi0 == phl);
        left &= ~mine;
        if (phi0 != phl) continue;
        const float2 *xrow = X + lane * a.span;
        float *d0 = a.dst + ((size_t)s * co + c0) * a.cap_b + a.hist_b + u0;
        float *d1 = d0 + a.cap_b;
        const bool al4 = ((a.cap_b | a.hist_b) & 3) == 0;
        // walk the chunk in groups of four outputs; (rel, ph) = first tap (relative to pos0) and phase of the group's first output
        int rel = pos0 - pos_s, ph = phl;
        const int n_here = min(a.chunk, L - u0);
#pragma unroll 1
        for (int m = 0; m < n_here; m += 4) {
          int phi[4], dk[4];
          {
            int pp = ph, pos = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              phi[k] = pp;
              dk[k] = pos;
              pp += fa; pos += ia;
              if (pp >= den) { pp -= den; pos += 1; }
            }
            // (pp, pos) now describe the next group's first output
            const float4 *tp[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) tp[k] = TAB + OFFROW[phi[k]] - dk[k];
            float acc[4][4][2];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int q = 0; q < 4; ++q) acc[k][q][0] = acc[k][q][1] = 0.f;
            const float2 *xp = xrow + rel;
#pragma unroll 8
            for (int i = 0; i < steps; ++i) {
              const float2 x = xp[i];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float4 t = tp[k][i];
                pipe_mac2(acc[k][0][0], acc[k][0][1], x.x, x.y, t.x, a.neg_zero);
                pipe_mac2(acc[k][1][0], acc[k][1][1], x.x, x.y, t.y, a.neg_zero);
                pipe_mac2(acc[k][2][0], acc[k][2][1], x.x, x.y, t.z, a.neg_zero);
                pipe_mac2(acc[k][3][0], acc[k][3][1], x.x, x.y, t.w, a.neg_zero);
              }
            }
            float o[2][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // cubic_coef of the output's phase (resample.c:246-256), tabulated per phase at plan time
              const float4 ci = CI[phi[k]];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                float sum = ci.x * acc[k][0][h] + ci.y * acc[k][1][h] + ci.z * acc[k][2][h] + ci.w * acc[k][3][h];
                sum = (sum < -1.0f) ? -1.0f : ((sum > 1.0f) ? 1.0f : sum);      // FLTADJUST, resample.c:84
                if (loud_on) sum *= plan.loud_gain;
                if (m + k >= n_here) sum = 0.f;
                o[h][k] = sum;
              }
            }
            if (al4) {
              __stcg(reinterpret_cast<float4 *>(d0 + m), make_float4(o[0][0], o[0][1], o[0][2], o[0][3]));
              if (two) __stcg(reinterpret_cast<float4 *>(d1 + m), make_float4(o[1][0], o[1][1], o[1][2], o[1][3]));
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                __stcg(d0 + m + k, o[0][k]);
                if (two) __stcg(d1 + m + k, o[1][k]);
              }
            }
            rel += pos;
            ph = pp;
          }
        }
      }
    }
  }
}

}  // namespace iamfb
