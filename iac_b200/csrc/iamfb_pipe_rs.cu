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

using namespace iamfb;

namespace {
template <class SIG>
int launch_rs(iamfb_ctx *ctx, const KernelPlan &kp, const PipeRsArgs &pa, int S) {
  CU(cudaFuncSetAttribute(k_pipe_rs<SIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, pa.smem_bytes));
  ScopedKernelTimer tm_(ctx, "k_pipe_rs");
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(S);
  lc.blockDim = dim3(SIG::kThreads);
  lc.dynamicSmemBytes = (size_t)pa.smem_bytes;
  lc.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;   // may start under k_resolve (griddepcontrol in the kernels)
  lc.attrs = at;
  lc.numAttrs = 1;
  CU(cudaLaunchKernelEx(&lc, k_pipe_rs<SIG>, kp, pa));
  return IAMFB_OK;
}
}  // namespace

int iamfb_pipe_rs_launch(iamfb_ctx *ctx, int sig_id, bool s16, const KernelPlan &kp, const PipeRsArgs &pa, int S) {
#define X(id, L0, N0, T, NW, VEC, MINB)                                                                     \
  if (sig_id == id)                                                                                         \
    return s16 ? launch_rs<PipeSig<L0, N0, 0, 0, T, true, 2, NW, VEC, MINB>>(ctx, kp, pa, S)                \
               : launch_rs<PipeSig<L0, N0, 0, 0, T, false, 2, NW, VEC, MINB>>(ctx, kp, pa, S);
  IAMFB_PIPE_RS_SIGS(X)
#undef X
  return fail(IAMFB_ERR_INTERNAL, "no k_pipe_rs signature %d", sig_id);
}
