// micro-benchmark: throughput of the bit-exact multiply-add pair as scalar FMUL+FADD vs packed FFMA2(v,c,-0)+FADD2
// build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -O3 -o f32x2_bench f32x2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define NACC 16
__device__ __forceinline__ void mac2(float &y0, float &y1, float v0, float v1, float c, float nz) {
  asm("{.reg .b64 rv, rc, rz, rt, ry; mov.b64 rv, {%2,%3}; mov.b64 rc, {%4,%4}; mov.b64 rz, {%5,%5}; fma.rn.f32x2 rt, rv, rc, rz; "
      "mov.b64 ry, {%0,%1}; add.rn.f32x2 ry, ry, rt; mov.b64 {%0,%1}, ry;}"
      : "+f"(y0), "+f"(y1) : "f"(v0), "f"(v1), "f"(c), "f"(nz));
}
template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, const float *in, int iters, float nz) {
  float y[NACC][2];
  float v0 = in[threadIdx.x], v1 = in[threadIdx.x + 128];
#pragma unroll
  for (int i = 0; i < NACC; ++i) y[i][0] = y[i][1] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      const float c = 0.0123f * (i + 1);
      if (MODE == 0) { y[i][0] += c * v0; y[i][1] += c * v1; }
      else mac2(y[i][0], y[i][1], v0, v1, c, nz);
    }
    v0 += 1e-7f; v1 -= 1e-7f;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += y[i][0] + y[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float *out, *in;
  cudaMalloc(&out, 148 * 16 * 128 * 4); cudaMalloc(&in, 256 * 4); cudaMemset(in, 0, 256 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 20000;
  for (int blocks_per_sm : {1, 2, 4, 8}) for (int mode = 0; mode < 2; ++mode) {
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(a);
      if (mode == 0) k<0><<<148 * blocks_per_sm, 128>>>(out, in, iters, -0.0f); else k<1><<<148 * blocks_per_sm, 128>>>(out, in, iters, -0.0f);
      cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    }
    double macs = (double)148 * blocks_per_sm * 128 * iters * NACC * 2;
    printf("warps/SM %2d mode %s: %.3f ms, %.1f GMAC/s, %.2f MAC/clk/SM (at 1.965 GHz)\n", blocks_per_sm * 4, mode ? "packed" : "scalar", ms, macs / ms / 1e6,
           macs / (ms * 1e-3) / 148 / 1.965e9);
  }
  float h[4]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); printf("%g\n", h[0]);
  return 0;
}
