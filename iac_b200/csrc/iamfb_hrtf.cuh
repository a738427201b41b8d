// iamfb_hrtf.cuh - binaural (HRTF) rendering of an audio element on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// What it replaces: IAMF_element_renderer_render_M2B (m2b_rdr.c:103-121: one HRIR pair per loudspeaker channel, BEAR) and
// IAMF_element_renderer_render_H2B (h2b_rdr.c:109-131: one HRIR pair per ambisonics channel, Resonance Audio) - both hand
// the planar frame [C][N] to a closed library and get [2][N] back, the filter state living inside the library.  Here:
//
//     out[ear][t] = sum_c sum_{k=0..255} h[c][ear][k] * x[c][t - k]                  (256-tap FIR per channel and ear)
//
// as EXACT integer arithmetic: the HRIR set is Q15 int16 (tools/gen_hrir.py), the decoded samples are fixed point too
// (int16 as the codecs produce it, or float32 rounded to Q20), both are split into 8-bit limbs, every limb pair is one
// int8 tensor-core product accumulated in int32 (TMEM), and the epilogue recombines the limb classes in int64 and rounds
// ONCE to float32.  The result does not depend on the order of accumulation, so the CPU oracle (oracle/oracle_hrtf.c) is
// matched bit for bit.
//
// The contraction.  Time is cut into blocks of 64 instants.  Output block b of a stream needs input blocks b-4 .. b:
//     out[64 b + i] = sum_{q=0..4} sum_{j=0..63} h[64 q + i - j] * x[64 (b - q) + j]            (h = 0 outside 0..255)
// i.e. D[(i, ear)][b] += A_q[(i, ear)][j] * X[j][b - q]: M = 128 rows (64 instants x 2 ears), N = the blocks of a tile
// (<= 128 consecutive blocks of one stream), K = 64 per (q, channel).  Neither operand is ever materialised:
//   * A_q is Toeplitz.  With the instants of a block stored in REVERSE order (j' = 63 - j) an 8 x 16 core matrix of the
//     K-major no-swizzle layout holds G_ear[base + (r >> 1) + u] (r = row: instant pair r >> 1, ear r & 1; u = byte) with
//     base = 64 q + 4 i4 + 16 kc: it depends on the SUM of the row group, the K core and q only.  One table of 96 core
//     matrices (12 KB) per (channel, limb) therefore serves all five q, all sixteen row groups and all four K cores: the
//     shared-memory descriptor just starts at core 16 q + 8 kk and strides 1 core per row group, 4 cores per K core
//     (the core matrices of an operand overlap - the descriptor is an address generator, nothing more).
//   * X for shift q is the same staged rows, started 4 - q rows further down: the blocks b0-4 .. b0+NB-1 of a channel are
//     staged once ([limb][kc][NB + 4 rows][16 bytes]) and serve all five q.
// Per channel a tile costs 20 * (limb pairs) MMAs of 128 x NB x 32 and moves 12 KB * 2 (tables, L2-resident) + the
// channel's samples once.
//
// Kernel shape (persistent, one CTA per SM, 576 threads): warp 0 = producer (bulk copies into a 4-stage ring, one stage
// per channel), warp 1 = MMA issuer (one elected lane), warps 2-17 = epilogue (TMEM -> registers -> int64 -> float32 ->
// the element's binaural frame [S][F][2][N]).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "iamf_b200.h"

namespace iamfb {

constexpr int kHrTaps = 256;
constexpr int kHrBlock = 64;                 // instants per time block (= K of one (q, channel) chunk)
constexpr int kHrShifts = kHrTaps / kHrBlock + 1;   // input blocks an output block depends on
constexpr int kHrHist = kHrTaps;             // history instants kept in front of a stream's plane (4 blocks)
constexpr int kHrCores = 96;                 // core matrices of one (channel, limb) Toeplitz table
constexpr int kHrTabBytes = kHrCores * 128;  // 12288
constexpr int kHrHLimbs = 2;                 // Q15 int16 taps: low limb unsigned, high limb signed
constexpr int kHrMaxXLimbs = 3;
constexpr int kHrStages = 4;
constexpr int kHrEpiGroups = 4;              // epilogue: column groups, each served by four warps (one per TMEM lane quarter)
constexpr int kHrThreads = 64 + kHrEpiGroups * 128;
constexpr int kHrMaxNB = 128;                // blocks per tile (N of the MMA)

// bytes of one pipeline stage for tiles of nb blocks and nl sample limbs
__host__ __device__ constexpr int hrtf_x_bytes(int nb, int nl) { return nl * 4 * (nb + 4) * 16; }
__host__ __device__ constexpr int hrtf_stage_bytes(int nb, int nl) { return kHrHLimbs * kHrTabBytes + ((hrtf_x_bytes(nb, nl) + 127) & ~127); }

struct HrtfGemmArgs {
  const uint8_t *tab;        // [C][2 limbs][96 cores][128]            Toeplitz core tables of the element's channels
  const uint8_t *planes;     // [S][C][NL][4 kc][NBP][16]              limb planes of the element (k_hrtf_prep)
  float *out;                // [S][F][2][N] float32                   the element's binaural frames
  const int *n_present;      // [S]   frames present in this submit
  const short *frame_of_slot;// [S][F] frame index of the k-th present frame
  int S, C, NL;              // streams (of this launch), channels, sample limbs
  int NB, NT, NBP;           // blocks per tile, tiles per stream, blocks per plane row (4 + NT * NB)
  int F, N;                  // frames per submit, frame size
  int x_shift;               // out = sum * 2^-(x_shift + 15)
  // RAW variant (16-bit PCM that reaches the renderer untouched): the limb rows are made inside the kernel from the decoded
  // frames themselves - no limb planes, no k_hrtf_prep
  const int16_t *raw_in;     // [S][F][n_in][N]
  int n_in;
  int row[IAMFB_MAX_SCENE_CH];   // decoded row of renderer input c
  const int *hist_in;        // [S][C][256] Q20: the previous submit's last instants
  int *hist_out;             // [S][C][256]
};
constexpr int kHrConvThreads = 128;          // RAW variant: converter warps (int16 rows -> limb rows in operand order)
__host__ __device__ constexpr int hrtf_raw_bytes(int nb) { return ((nb + 4) * kHrBlock * 2 + 127) & ~127; }

namespace hr {
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t *b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void bar_expect(uint64_t *b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void bar_wait(uint64_t *b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nHR_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra HR_DONE;\nbra HR_WAIT;\nHR_DONE:\n}\n" ::"r"(s32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes),
               "r"(s32(b))
               : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes); lbo = bytes between the two K cores of an
// MMA, sbo = bytes between 8-row groups (cute::UMMA::SmemDescriptor, version 1 = sm_100)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor for kind::i8: int32 accumulate, K-major A and B, M = 128
__device__ __forceinline__ uint32_t instr_desc_i8(bool a_signed, bool b_signed, int n) {
  return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
}  // namespace hr

// dynamic shared memory: [stages][table 2 x 12288 | X limbs]; static: barriers + the TMEM base address
template <int NL, bool RAW>
static __global__ void __launch_bounds__(kHrThreads + (RAW ? kHrConvThreads : 0), 1) k_hrtf_gemm(const HrtfGemmArgs a) {
  extern __shared__ __align__(128) uint8_t hr_smem[];
  __shared__ __align__(8) uint64_t s_full[kHrStages], s_empty[kHrStages], s_tfull, s_tempty;
  __shared__ __align__(8) uint64_t s_rfull[2], s_rempty[2];   // RAW: the two raw-row buffers behind the stages
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NB = a.NB, C = a.C;
  const int stage_bytes = hrtf_stage_bytes(NB, NL);
  const int xrow = (NB + 4) * 16;                 // bytes of one (limb, kc) row group of a stage
  const int n_tiles = a.S * a.NT;

  if (threadIdx.x == 0) {
    // (RAW: a stage is full when its tables have landed AND the converter warps have written its limb rows)
    for (int i = 0; i < kHrStages; ++i) { hr::bar_init(&s_full[i], RAW ? 2 : 1); hr::bar_init(&s_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { hr::bar_init(&s_rfull[i], 1); hr::bar_init(&s_rempty[i], 1); }
    hr::bar_init(&s_tfull, 1);
    hr::bar_init(&s_tempty, 4 * kHrEpiGroups);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hr::s32(&s_tmem)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  if (warp == 0) {
    // ===== producer: one stage per (tile, channel): the channel's two Toeplitz tables + its sample limbs for blocks b0-4 ..
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int s = tile / a.NT, b0 = (tile - s * a.NT) * NB;
        if (a.n_present[s] * a.N <= b0 * kHrBlock) continue;      // nothing of this stream reaches the tile
        for (int c = 0; c < C; ++c, ++it) {
          const int st = it % kHrStages;
          hr::bar_wait(&s_empty[st], ((it / kHrStages) & 1u) ^ 1u);
          uint8_t *dst = hr_smem + (size_t)st * stage_bytes;
          if constexpr (RAW) {
            hr::bar_expect(&s_full[st], (uint32_t)(kHrHLimbs * kHrTabBytes));
            hr::bulk_load(dst, a.tab + (size_t)c * kHrHLimbs * kHrTabBytes, kHrHLimbs * kHrTabBytes, &s_full[st]);
            // the channel's decoded int16 row for the instants [64 b0 - 256, 64 (b0 + NB)) of the stream's time line, frame
            // piece by frame piece (present frames only; the instants before the submit come from the history)
            const int rb = it & 1;
            hr::bar_wait(&s_rempty[rb], ((it >> 1) & 1u) ^ 1u);
            uint8_t *raw = hr_smem + (size_t)kHrStages * stage_bytes + (size_t)rb * hrtf_raw_bytes(NB);
            const int len = a.n_present[s] * a.N;
            const int t0 = b0 * kHrBlock - kHrHist, t1 = min(len, (b0 + NB) * kHrBlock);
            int t = max(t0, 0);
            hr::bar_expect(&s_rfull[rb], (uint32_t)(max(t1 - t, 0) * 2));
            while (t < t1) {
              const int slot = t / a.N, off = t - slot * a.N;
              const int cnt = min(a.N - off, t1 - t);
              const int f = a.frame_of_slot[(size_t)s * a.F + slot];
              hr::bulk_load(raw + (size_t)(t - t0) * 2, a.raw_in + (((size_t)s * a.F + f) * a.n_in + a.row[c]) * a.N + off, (uint32_t)(cnt * 2),
                            &s_rfull[rb]);
              t += cnt;
            }
          } else {
            hr::bar_expect(&s_full[st], (uint32_t)(kHrHLimbs * kHrTabBytes + NL * 4 * xrow));
            hr::bulk_load(dst, a.tab + (size_t)c * kHrHLimbs * kHrTabBytes, kHrHLimbs * kHrTabBytes, &s_full[st]);
            const uint8_t *src = a.planes + ((size_t)(s * C + c) * NL * 4) * a.NBP * 16 + (size_t)b0 * 16;
            uint8_t *xd = dst + kHrHLimbs * kHrTabBytes;
            for (int r = 0; r < NL * 4; ++r) hr::bulk_load(xd + r * xrow, src + (size_t)r * a.NBP * 16, (uint32_t)xrow, &s_full[st]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    if (lane == 0) {
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int s = tile / a.NT, b0 = (tile - s * a.NT) * NB;
        if (a.n_present[s] * a.N <= b0 * kHrBlock) continue;
        hr::bar_wait(&s_tempty, (tl & 1u) ^ 1u);                    // the epilogue has drained the accumulators
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t touched = 0;                                      // limb classes that hold a partial sum already
        for (int c = 0; c < C; ++c, ++it) {
          const int st = it % kHrStages;
          hr::bar_wait(&s_full[st], (it / kHrStages) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tab = hr::s32(hr_smem + (size_t)st * stage_bytes);
          const uint32_t xs = tab + kHrHLimbs * kHrTabBytes;
          for (int q = 0; q < kHrShifts; ++q)
            for (int kk = 0; kk < 2; ++kk)
              for (int xl = 0; xl < NL; ++xl) {
                const uint64_t bd = hr::smem_desc(xs + (uint32_t)((xl * 4 + 2 * kk) * xrow + (4 - q) * 16), (uint32_t)xrow, 128u);
                for (int hl = 0; hl < kHrHLimbs; ++hl) {
                  const uint64_t ad = hr::smem_desc(tab + (uint32_t)(hl * kHrTabBytes + (16 * q + 8 * kk) * 128), 512u, 128u);
                  const int cls = xl + hl;
                  hr::mma_i8(tmem + (uint32_t)(cls * NB), ad, bd, hr::instr_desc_i8(hl == kHrHLimbs - 1, xl == NL - 1, NB), (touched >> cls) & 1u);
                  touched |= 1u << cls;
                }
              }
          hr::mma_commit(&s_empty[st]);                            // the stage is free once these MMAs have read it
        }
        hr::mma_commit(&s_tfull);                                  // accumulators complete
        ++tl;
      }
    }
  } else if (RAW && warp >= kHrThreads / 32) {
    // ===== converters (RAW): the staged int16 row of a channel -> its two byte planes in operand order ([limb][kc][block][16],
    // instants of a block reversed: byte u of (kc, block) = instant 63 - 16 kc - u), 16 instants per thread and step; they also
    // carry the filter history: the submit's last 256 instants go out as Q20 for the next submit, the previous submit's come in
    if constexpr (RAW) {
      const int ct = threadIdx.x - kHrThreads;
      // streams without a frame in this submit keep their history
      for (int s = blockIdx.x; s < a.S; s += gridDim.x)
        if (a.n_present[s] == 0)
          for (int i = ct; i < C * kHrHist; i += kHrConvThreads) a.hist_out[(size_t)s * C * kHrHist + i] = a.hist_in[(size_t)s * C * kHrHist + i];
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int s = tile / a.NT, b0 = (tile - s * a.NT) * NB;
        const int len = a.n_present[s] * a.N;
        if (len <= b0 * kHrBlock) continue;
        const int t0 = b0 * kHrBlock - kHrHist;
        for (int c = 0; c < C; ++c, ++it) {
          const int st = it % kHrStages, rb = it & 1;
          hr::bar_wait(&s_rfull[rb], (it >> 1) & 1u);
          const uint8_t *raw = hr_smem + (size_t)kHrStages * stage_bytes + (size_t)rb * hrtf_raw_bytes(NB);
          uint8_t *xs = hr_smem + (size_t)st * stage_bytes + kHrHLimbs * kHrTabBytes;
          const size_t hbase = ((size_t)s * C + c) * kHrHist;
          for (int g = ct; g < (NB + 4) * 4; g += kHrConvThreads) {
            const int tau = t0 + 16 * g;
            uint32_t w[8];
            if (tau < 0) {
              const int4 *h = reinterpret_cast<const int4 *>(a.hist_in + hbase + kHrHist + tau);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int4 v = h[k];
                w[2 * k] = ((uint32_t)(v.x >> 5) & 0xffffu) | ((uint32_t)(v.y >> 5) << 16);
                w[2 * k + 1] = ((uint32_t)(v.z >> 5) & 0xffffu) | ((uint32_t)(v.w >> 5) << 16);
              }
            } else {
              const uint4 v0 = *reinterpret_cast<const uint4 *>(raw + 32 * g), v1 = *reinterpret_cast<const uint4 *>(raw + 32 * g + 16);
              w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
            }
            const int p0 = tau + kHrHist - len;            // index on the next history of this group's first instant
            if (p0 >= 0 && p0 < kHrHist && tau < len) {
              int4 *h = reinterpret_cast<int4 *>(a.hist_out + hbase + p0);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                h[k] = make_int4((int)(short)(w[2 * k] & 0xffffu) << 5, (int)(short)(w[2 * k] >> 16) << 5, (int)(short)(w[2 * k + 1] & 0xffffu) << 5,
                                 (int)(short)(w[2 * k + 1] >> 16) << 5);
            }
            const int bp = g >> 2, kc = 3 - (g & 3);
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              lo[q] = __byte_perm(w[7 - 2 * q], w[6 - 2 * q], 0x4602);
              hi[q] = __byte_perm(w[7 - 2 * q], w[6 - 2 * q], 0x5713);
            }
            *reinterpret_cast<uint4 *>(xs + (size_t)((0 * 4 + kc) * (NB + 4) + bp) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4 *>(xs + (size_t)((1 * 4 + kc) * (NB + 4) + bp) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the tensor core reads the rows through the async proxy
          asm volatile("bar.sync 3, %0;" ::"n"(kHrConvThreads) : "memory");
          if (ct == 0) {
            hr::bar_arrive(&s_full[st]);
            hr::bar_arrive(&s_rempty[rb]);
          }
        }
      }
    }
  } else {
    // ===== epilogue: 16 warps = 4 column groups x 4 lane quarters (a warp reads the TMEM lanes 32 (warp % 4) .. only):
    // row R = 2 i + ear of a quarter of the tile's columns (= blocks).  One warp per quarter took longer than the tile's
    // MMAs (every warp is latency-bound on its own dependent instructions); sixteen bring the drain under a fifth of them
    const int quarter = warp & 3, cg = (warp - 2) >> 2;
    const int R = quarter * 32 + lane, i = R >> 1, ear = R & 1;
    const float scale = __int_as_float((127 - (a.x_shift + 15)) << 23);   // 2^-(x_shift + 15)
    const int cols = NB / kHrEpiGroups;                                   // columns of this warp (NB is a multiple of 16)
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int s = tile / a.NT, b0 = (tile - s * a.NT) * NB;
      const int len = a.n_present[s] * a.N;                         // instants of this stream in the submit
      if (len <= b0 * kHrBlock) continue;
      hr::bar_wait(&s_tfull, tl & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const short *fos = a.frame_of_slot + (size_t)s * a.F;
      const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cg * cols);
      // d0 + d1 2^8 + d2 2^16 (+ d3 2^24) exactly, as hi 2^16 + lo with hi = d2 + (d1 >> 8) (+ d3 2^8), lo = d0 + ((d1 & 255) << 8);
      // int64 -> float32 rounds once
      auto value = [&](const int32_t (&d)[NL + 1][16], int k) -> float {
        const int lo = d[0][k] + ((d[1][k] & 255) << 8);
        long long hi = (long long)(d[2][k] + (d[1][k] >> 8));
        if constexpr (NL == 3) hi += (long long)d[3][k] << 8;
        return (float)(hi * 65536 + (long long)lo) * scale;
      };
      if ((a.N & (kHrBlock - 1)) == 0) {
        // frames of whole blocks: a column lies in one frame for every row - the frame walk is warp-uniform
        const int bpf = a.N / kHrBlock;                                   // blocks per frame
        int col = b0 + cg * cols;                                          // block index on the stream's time line
        int slot = col / bpf, brem = col - slot * bpf;
        const int nblk = len / kHrBlock;
        float *fbase = a.out + (((size_t)s * a.F + (col < nblk ? fos[slot] : 0)) * 2 + ear) * a.N + i;
        for (int n0 = 0; n0 < cols; n0 += 16) {
          int32_t d[NL + 1][16];
#pragma unroll
          for (int cls = 0; cls <= NL; ++cls) hr::tmem_ld16(lane_base + (uint32_t)(cls * NB + n0), d[cls]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            if (col < nblk) fbase[brem * kHrBlock] = value(d, k);
            ++col;
            if (++brem == bpf) {
              brem = 0;
              ++slot;
              fbase = a.out + (((size_t)s * a.F + (col < nblk ? fos[slot] : 0)) * 2 + ear) * a.N + i;
            }
          }
        }
      } else {
        // any frame size (a multiple of 16): every row walks the frames on its own
        int tau = (b0 + cg * cols) * kHrBlock + i;
        int slot = tau / a.N, rem = tau - slot * a.N;
        float *dst = a.out + (((size_t)s * a.F + (tau < len ? fos[slot] : 0)) * 2 + ear) * a.N + rem;
        for (int n0 = 0; n0 < cols; n0 += 16) {
          int32_t d[NL + 1][16];
#pragma unroll
          for (int cls = 0; cls <= NL; ++cls) hr::tmem_ld16(lane_base + (uint32_t)(cls * NB + n0), d[cls]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            if (tau < len) *dst = value(d, k);
            tau += kHrBlock;
            rem += kHrBlock;
            dst += kHrBlock;
            if (rem >= a.N) {             // next present frame
              rem -= a.N;
              ++slot;
              dst = a.out + (((size_t)s * a.F + (tau < len ? fos[slot] : 0)) * 2 + ear) * a.N + rem;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) hr::bar_arrive(&s_tempty);
      ++tl;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

// ---- k_hrtf_index: which frames of the submit feed the renderer.  A frame that is absent (trim_start == 0xFFFF) or trimmed
// completely (dropped before rendering, IAMF_decoder.c:3354-3358) never reaches the binaural renderer; the present frames of
// a stream are rendered as one continuous signal (slot k = the k-th present frame).
struct HrtfIndexArgs {
  const iamfb_frame_params *params;   // [S][F]
  int *n_present;                     // [S]
  short *frame_of_slot;               // [S][F]
  short *slot_of_frame;               // [S][F]  (-1: not rendered)
  int S, F, N;
};
static __global__ void __launch_bounds__(128) k_hrtf_index(const HrtfIndexArgs a) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.S) return;
  int n = 0;
  for (int f = 0; f < a.F; ++f) {
    const iamfb_frame_params &p = a.params[(size_t)s * a.F + f];
    const bool present = p.trim_start != 0xFFFF && p.trim_start != a.N && p.trim_end != a.N;
    a.slot_of_frame[(size_t)s * a.F + f] = present ? (short)n : (short)-1;
    if (present) a.frame_of_slot[(size_t)s * a.F + n++] = (short)f;
  }
  a.n_present[s] = n;
}

// ---- k_hrtf_prep: decoded rows -> the renderer's input channels (channel order / output gain of the de-mixer,
// demixer.c:421-430,636-664; ambisonics channel mapping or projection, IAMF_core_decoder.c:105-130) -> Q20 fixed point ->
// limb planes in the operand order of k_hrtf_gemm ([limb][kc][block][16 bytes], instants of a block reversed), with the
// last 256 instants of the previous submit in front.  One thread = 16 consecutive instants of one channel.
struct HrtfPrepArgs {
  const void *in;            // [S][F][n_in][N] float32 or int16
  uint8_t *planes;           // [S][C][NL][4][NBP][16]
  const int *hist_in;        // [S][C][256] Q20, the previous submit's last instants
  int *hist_out;             // [S][C][256]
  const int *n_present;      // [S]
  const short *frame_of_slot;// [S][F]
  int C, n_in, NL, NBP, F, N;
  int mode;                  // 0: row (x gain), 1: projection
  int row[IAMFB_MAX_SCENE_CH];
  float gain[IAMFB_MAX_SCENE_CH];       // 1.0 = none (the multiply is skipped, as dmx_gainup skips unflagged channels)
  int proj_cols;
  float proj[IAMFB_MAX_SCENE_CH * IAMFB_MAX_SCENE_CH];   // [col][C]
};
__device__ __forceinline__ int hrtf_quantise(float x) {     // oracle_hrtf.c: orc_hrtf_quantise
  float v = x * 1048576.0f;
  v = fminf(fmaxf(v, -8388607.0f), 8388607.0f);
  return __float2int_rn(v);
}
template <bool S16>
static __global__ void __launch_bounds__(256) k_hrtf_prep(const HrtfPrepArgs a) {
  const int sc = blockIdx.y, s = sc / a.C, c = sc - s * a.C;
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;        // 16-instant group of the plane (16 history groups first)
  const int len = a.n_present[s] * a.N;
  if (gi >= (kHrHist + len) / 16) return;
  int xq[16];
  if (gi < kHrHist / 16) {
    const int4 *h = reinterpret_cast<const int4 *>(a.hist_in + (size_t)sc * kHrHist + 16 * gi);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int4 v = h[k]; xq[4 * k] = v.x; xq[4 * k + 1] = v.y; xq[4 * k + 2] = v.z; xq[4 * k + 3] = v.w; }
  } else {
    const int tau = 16 * gi - kHrHist, slot = tau / a.N, i = tau - slot * a.N;
    const int f = a.frame_of_slot[(size_t)s * a.F + slot];
    const size_t frame = ((size_t)s * a.F + f) * a.n_in * a.N + i;
    auto load16 = [&](int row, float (&v)[16]) {
      if constexpr (S16) {
        const int4 *p = reinterpret_cast<const int4 *>(reinterpret_cast<const int16_t *>(a.in) + frame + (size_t)row * a.N);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int4 w = p[h];
          const int ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            v[8 * h + 2 * k] = (float)(short)(ww[k] & 0xffff) / 32768.f;
            v[8 * h + 2 * k + 1] = (float)(short)(ww[k] >> 16) / 32768.f;
          }
        }
      } else {
        const float4 *p = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(a.in) + frame + (size_t)row * a.N);
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float4 w = p[k]; v[4 * k] = w.x; v[4 * k + 1] = w.y; v[4 * k + 2] = w.z; v[4 * k + 3] = w.w; }
      }
    };
    float v[16];
    if (a.mode == 0) {
      load16(a.row[c], v);
      const float g = a.gain[c];
      if (g != 1.0f) {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] *= g;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = .0f;
      for (int l = 0; l < a.proj_cols; ++l) {
        float t[16];
        load16(l, t);
        const float m = a.proj[l * a.C + c];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] += t[k] * m;
      }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) xq[k] = hrtf_quantise(v[k]);
  }
  // the last 256 instants (old history included when the submit is shorter) are the next submit's history
  {
    const int p0 = 16 * gi - len;                 // index on the next history of this group's first instant
    if (p0 >= 0) {
      int4 *h = reinterpret_cast<int4 *>(a.hist_out + (size_t)sc * kHrHist + p0);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = make_int4(xq[4 * k], xq[4 * k + 1], xq[4 * k + 2], xq[4 * k + 3]);
    }
  }
  // limbs, instants reversed inside the block: byte u of (kc, block) = instant 63 - 16 kc - u
  const int bp = gi >> 2, kc = 3 - (gi & 3);
  const int shift = a.NL == 2 ? 5 : 0;            // 16-bit content travels as two limbs (Q15), anything else as three (Q20)
  for (int l = 0; l < a.NL; ++l) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t r = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) r |= (uint32_t)(((xq[15 - (4 * q + u)] >> shift) >> (8 * l)) & 255) << (8 * u);
      w[q] = r;
    }
    *reinterpret_cast<uint4 *>(a.planes + ((((size_t)sc * a.NL + l) * 4 + kc) * a.NBP + bp) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---- host: Toeplitz core tables of one channel.  taps = [2 ears][256] Q15; dst = [2 limbs][96 cores][8 rows][16 bytes]
inline void hrtf_build_table(const int16_t *taps, uint8_t *dst) {
  for (int hl = 0; hl < kHrHLimbs; ++hl)
    for (int cb = 0; cb < kHrCores; ++cb)
      for (int r = 0; r < 8; ++r)
        for (int u = 0; u < 16; ++u) {
          const int m = 4 * cb + (r >> 1) + u - (kHrBlock - 1);     // tap index
          const int ear = r & 1;
          const int v = (m >= 0 && m < kHrTaps) ? taps[ear * kHrTaps + m] : 0;
          dst[((hl * kHrCores + cb) * 8 + r) * 16 + u] = (uint8_t)(hl == 0 ? (v & 255) : ((v >> 8) & 255));
        }
}

}  // namespace iamfb
