// k_pipe_rs instantiations (resampling pipelines) and their launcher
#include <cuda.h>
#include <cuda_runtime.h>

#include <type_traits>

#include "iamf_b200.h"
#include "iamfb_internal.h"
#include "iamfb_kernels.cuh"
#include "iamfb_fused.cuh"
#include "iamfb_matrices.inc"
#include "iamfb_stream.cuh"
#include "iamfb_pipe.cuh"
#include "iamfb_pipe_rs.cuh"
#include "iamfb_resample_ls.cuh"

using namespace iamfb;

namespace {
template <class SIG, bool PRE = false>
int launch_rs(iamfb_ctx *ctx, const KernelPlan &kp, const PipeRsArgs &pa, int S, bool pdl = true) {
  CU(cudaFuncSetAttribute(k_pipe_rs<SIG, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pa.smem_bytes));
  ScopedKernelTimer tm_(ctx, PRE ? "k_pipe_rs_lim" : "k_pipe_rs");
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(S);
  lc.blockDim = dim3(SIG::kThreads);
  lc.dynamicSmemBytes = (size_t)pa.smem_bytes;
  lc.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;   // may start under k_resolve (griddepcontrol in the kernels)
  lc.attrs = at;
  lc.numAttrs = 1;
  CU(cudaLaunchKernelEx(&lc, k_pipe_rs<SIG, PRE>, kp, pa));
  return IAMFB_OK;
}
}  // namespace

// the limiter half of a resampling pipeline behind k_resample_ls
int iamfb_pipe_rs_lim_launch(iamfb_ctx *ctx, int sig_id, const KernelPlan &kp, const PipeRsArgs &pa, int S) {
#define X(id, L0, N0, T, NW, VEC, MINB)                                                                     \
  if (sig_id == id) return launch_rs<PipeSig<L0, N0, 0, 0, T, false, 2, NW, VEC, MINB>, true>(ctx, kp, pa, S, false);
  IAMFB_PIPE_RS_SIGS(X)
#undef X
  return fail(IAMFB_ERR_INTERNAL, "no k_pipe_rs signature %d", sig_id);
}

// the render stage on its own: regular streams -> pre-resample time line
int iamfb_pipe_prerender_launch(iamfb_ctx *ctx, int sig_id, bool s16, const KernelPlan &kp, const PreRenderArgs &pa, int S, int F) {
  const dim3 grid((unsigned)((kp.frame_size / 4 + 127) / 128), (unsigned)F, (unsigned)S);
  ScopedKernelTimer tm_(ctx, "k_pipe_prerender");
#define X(id, L0, N0, T, NW, VEC, MINB)                                                                          \
  if (sig_id == id) {                                                                                            \
    if (s16) k_pipe_prerender<PipeSig<L0, N0, 0, 0, T, true, 2, NW, VEC, MINB>><<<grid, 128, 0, ctx->stream>>>(kp, pa);  \
    else k_pipe_prerender<PipeSig<L0, N0, 0, 0, T, false, 2, NW, VEC, MINB>><<<grid, 128, 0, ctx->stream>>>(kp, pa);     \
    return IAMFB_OK;                                                                                             \
  }
  IAMFB_PIPE_RS_SIGS(X)
#undef X
  return fail(IAMFB_ERR_INTERNAL, "no k_pipe_rs signature %d", sig_id);
}

int iamfb_resample_ls_blocks_resident(int smem_bytes) {
  int n = 0;
  if (cudaFuncSetAttribute(k_resample_ls<kLsWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_resample_ls<kLsWarps>, kLsWarps * 32, (size_t)smem_bytes) != cudaSuccess) return 0;
  return n;
}

int iamfb_resample_ls_launch(iamfb_ctx *ctx, const KernelPlan &kp, const ResampleLsArgs &a, int blocks, int smem_bytes) {
  CU(cudaFuncSetAttribute(k_resample_ls<kLsWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  ScopedKernelTimer tm_(ctx, "k_resample_ls");
  k_resample_ls<kLsWarps><<<blocks, kLsWarps * 32, (size_t)smem_bytes, ctx->stream>>>(kp, a);
  return IAMFB_OK;
}

int iamfb_pipe_rs_launch(iamfb_ctx *ctx, int sig_id, bool s16, const KernelPlan &kp, const PipeRsArgs &pa, int S) {
#define X(id, L0, N0, T, NW, VEC, MINB)                                                                     \
  if (sig_id == id)                                                                                         \
    return s16 ? launch_rs<PipeSig<L0, N0, 0, 0, T, true, 2, NW, VEC, MINB>>(ctx, kp, pa, S)                \
               : launch_rs<PipeSig<L0, N0, 0, 0, T, false, 2, NW, VEC, MINB>>(ctx, kp, pa, S);
  IAMFB_PIPE_RS_SIGS(X)
#undef X
  return fail(IAMFB_ERR_INTERNAL, "no k_pipe_rs signature %d", sig_id);
}
