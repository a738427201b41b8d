// iamfb_internal.h - host-side pieces shared by the translation units of libiamf_b200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

#include "iamf_b200.h"
#include "iamfb_types.cuh"

int iamfb_fail(int code, const char *fmt, ...);
#define fail iamfb_fail
#define CU(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return fail(IAMFB_ERR_CUDA, "%s -> %s", #call, cudaGetErrorString(e_));     \
  } while (0)

struct KernelTimer {
  const char *name;
  double total_ms;
  uint64_t launches;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
  std::vector<float> samples;   // per-launch durations (the median is robust against a launch that was pre-empted)
};

constexpr int kMaxChunks = 32;

struct iamfb_ctx {
  int device;
  int n_sm;                // multiprocessors of the device
  cudaStream_t stream;
  bool own_stream;
  uint64_t launches;
  bool timing;
  std::vector<KernelTimer> timers;
  std::vector<cudaEvent_t> event_pool;
  // second stream for the latency-bound limiter scan of the multi-kernel path, so that it overlaps the bandwidth-bound
  // kernels of the neighbouring sub-chunks; ordered against the main stream with events
  cudaStream_t aux;
  cudaEvent_t ev_w[iamfb::kMaxSub], ev_s[iamfb::kMaxSub];
  // host-resident submits: copy engines run on their own streams so that the upload of one group of streams, the
  // kernels of the previous group and the download of the one before overlap (PCIe is full duplex)
  cudaStream_t h2d, d2h;
  cudaEvent_t ev_up[kMaxChunks], ev_done[kMaxChunks], ev_free;
  cudaEvent_t ev_back[kMaxChunks];   // a group's PCM is back in host memory (iamfb_batch_submit_host_hooks)
};

// optional per-kernel CUDA-event timing (bench.py's roofline leg); events are recorded on the launching stream
struct ScopedKernelTimer {
  iamfb_ctx *ctx;
  KernelTimer *t = nullptr;
  cudaEvent_t a = nullptr, b = nullptr;
  static cudaEvent_t get(iamfb_ctx *c) {
    cudaEvent_t e;
    if (!c->event_pool.empty()) { e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
  }
  cudaStream_t st;
  ScopedKernelTimer(iamfb_ctx *c, const char *name, cudaStream_t stream = nullptr) : ctx(c), st(stream ? stream : c->stream) {
    if (!c->timing) return;
    for (auto &k : c->timers) if (!strcmp(k.name, name)) t = &k;
    if (!t) { c->timers.push_back(KernelTimer{name, 0.0, 0, {}, {}}); t = &c->timers.back(); }
    a = get(c); b = get(c);
    cudaEventRecord(a, st);
  }
  ~ScopedKernelTimer() {
    if (!t) return;
    cudaEventRecord(b, st);
    t->pending.emplace_back(a, b);
    ++t->launches;
  }
};

// ---- k_pipe (iamfb_pipe.cuh): the signatures it is instantiated for.
//   X(id, L0, N0, L1, N1, TARGET, NW, VEC, MINB)   (L = layout, -1 = scene-based; N = renderer inputs; N1 = 0: one element)
// Channel-based sources of 7.1.4 / 7.1 / 5.1.4 / 5.1 / stereo to stereo, 5.1 and the binaural target as the reference
// builds it; third-order ambisonics to sound system H; 7.1.4 + first-order ambisonics mixes.
#define IAMFB_PIPE_SIGS(X)                                                                                              \
  X(0, 7, 12, 0, 0, 1, 2, 4, 7)   X(1, 1, 2, 0, 0, 0, 2, 4, 7)   X(2, 7, 12, 0, 0, 0, 2, 4, 7)   X(3, 7, 12, 0, 0, 13, 2, 4, 7) \
  X(4, 5, 8, 0, 0, 1, 2, 4, 7)    X(5, 4, 10, 0, 0, 1, 2, 4, 7)  X(6, 2, 6, 0, 0, 1, 2, 4, 7)    X(7, 2, 6, 0, 0, 0, 2, 4, 7)   \
  X(8, 1, 2, 0, 0, 1, 2, 4, 7)    X(9, 1, 2, 0, 0, 13, 2, 4, 7)                                                                \
  X(10, -1, 16, 0, 0, 7, 4, 2, 3) X(11, 7, 12, -1, 4, 13, 2, 4, 7) X(12, -1, 4, 7, 12, 7, 4, 2, 3)                             \
  X(13, 7, 12, 0, 0, 1, 4, 2, 7) X(14, -1, 16, 0, 0, 7, 2, 4, 3) X(15, 1, 2, 1, 2, 13, 2, 4, 7)
#define IAMFB_PIPE_GROUP_OF(id) ((id) == 0 ? 0 : ((id) == 10 ? 1 : ((id) == 12 ? 2 : ((id) == 11 ? 3 : ((id) == 13 ? 4 : ((id) == 14 ? 5 : (4 + (id) % 3)))))))
constexpr int kPipeGroups = 7;
// IAMFB_ARITH_FMA variants (same columns): the dense contraction of the signature fuses multiply and add
#define IAMFB_PIPE_SIGS_FMA(X) X(48, -1, 16, 0, 0, 7, 4, 2, 3)   /* (the 2 x 4 thread shape measured 18 % slower) */
constexpr int kPipeFmaFirstId = 48;

// ---- k_pipe_rs (iamfb_pipe_rs.cuh): resampling pipelines.  X(id, L0, N0, TARGET, NW, VEC, MINB): one channel-based element
#define IAMFB_PIPE_RS_SIGS(X) X(32, 1, 2, 0, 2, 4, 4) X(33, 7, 12, 1, 2, 4, 4) X(34, 1, 2, 0, 4, 2, 7) X(35, 1, 2, 0, 4, 2, 5) X(36, 1, 2, 0, 2, 4, 6)

struct PipeSigInfo { int id, l0, n0, l1, n1, target, nw, vec, fma; };
// signature serving (element kinds / layouts, target), or nullptr; fma: prefer the IAMFB_ARITH_FMA variant when there is one
const PipeSigInfo *iamfb_pipe_find(int l0, int n0, int l1, int n1, int target, bool fma);

namespace iamfb { struct PipeArgs; }
// launches k_pipe<sig, s16> over S streams; returns IAMFB_OK or an error (the launch itself is checked by the caller)
int iamfb_pipe_launch(iamfb_ctx *ctx, int sig_id, bool s16, const iamfb::KernelPlan &kp, const iamfb::PipeArgs &pa, int S, size_t smem,
                      const CUtensorMap &m0, const CUtensorMap &m1);
namespace iamfb { struct PipeRsArgs; }
int iamfb_pipe_rs_launch(iamfb_ctx *ctx, int sig_id, bool s16, const iamfb::KernelPlan &kp, const iamfb::PipeRsArgs &pa, int S);
// split form of the resampling pipelines: k_pipe_prerender + k_resample_ls (one stream per lane) + k_pipe_rs<PRE> (the limiter half)
int iamfb_pipe_rs_lim_launch(iamfb_ctx *ctx, int sig_id, const iamfb::KernelPlan &kp, const iamfb::PipeRsArgs &pa, int S);
namespace iamfb { struct ResampleLsArgs; struct PreRenderArgs; }
int iamfb_pipe_prerender_launch(iamfb_ctx *ctx, int sig_id, bool s16, const iamfb::KernelPlan &kp, const iamfb::PreRenderArgs &pa, int S, int F);
int iamfb_resample_ls_blocks_resident(int smem_bytes);
int iamfb_resample_ls_launch(iamfb_ctx *ctx, const iamfb::KernelPlan &kp, const iamfb::ResampleLsArgs &a, int blocks, int smem_bytes);

// ---- binaural HRTF front end (iamfb_hrtf.cu)
struct iamfb_hrtf_front;
struct iamfb_hrtf_batch;
int iamfb_hrtf_front_create(const iamfb_plan_desc *d, iamfb_plan_desc *back, iamfb_hrtf_front **out);
void iamfb_hrtf_front_destroy(iamfb_hrtf_front *h);
int iamfb_hrtf_batch_create(const iamfb_hrtf_front *h, int S, int Fmax, iamfb_hrtf_batch **out);
int iamfb_hrtf_batch_reset(const iamfb_hrtf_front *h, iamfb_hrtf_batch *b, cudaStream_t st);
void iamfb_hrtf_batch_destroy(iamfb_hrtf_batch *b);
// demixed[e] != null: element e enters the renderer from that float32 [S][F][C][N] buffer (its layout channels, de-mixed)
int iamfb_hrtf_run(iamfb_ctx *ctx, const iamfb_hrtf_front *h, iamfb_hrtf_batch *b, const iamfb_io *io, int F, int s_lo, int s_cnt,
                   iamfb_io *out_io, const float *const *demixed);
bool iamfb_hrtf_needs_demix(const iamfb_hrtf_front *h, int e);
