// Standalone check of k_hrtf_gemm (iac_b200/csrc/iamfb_hrtf.cuh) against an int64 CPU convolution, and its timing.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I iac_b200/csrc -o tools/experiments/hrtf_gemm_test tools/experiments/hrtf_gemm_test.cu
//   ./hrtf_gemm_test [S] [C] [F] [NL]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "iamfb_hrtf.cuh"

using namespace iamfb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static uint32_t rng_state = 12345;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

int main(int argc, char **argv) {
  const int S = argc > 1 ? atoi(argv[1]) : 64, C = argc > 2 ? atoi(argv[2]) : 12, F = argc > 3 ? atoi(argv[3]) : 8, NL = argc > 4 ? atoi(argv[4]) : 2;
  const int N = 960, T = F * N;
  const int x_bits = NL == 2 ? 16 : 24, x_shift = NL == 2 ? 15 : 20;
  int NB = (T + 63) / 64;
  NB = (NB + 63) & ~63;
  if (NB > kHrMaxNB) NB = kHrMaxNB;
  const int NT = (T + NB * 64 - 1) / (NB * 64), NBP = 4 + NT * NB;
  printf("S %d C %d F %d NL %d  NB %d NT %d NBP %d\n", S, C, F, NL, NB, NT, NBP);
  // taps, samples (history + body), per stream a number of present frames
  std::vector<int16_t> taps((size_t)C * 2 * kHrTaps);
  for (auto &t : taps) t = (int16_t)((int)(rnd() % 65535) - 32767);
  std::vector<int> x((size_t)S * C * (kHrHist + T));
  const int lim = 1 << (x_bits - 1);
  for (auto &v : x) v = (int)(rnd() % (2 * lim)) - lim;
  std::vector<int> n_present(S);
  std::vector<short> fos((size_t)S * F);
  for (int s = 0; s < S; ++s) {
    n_present[s] = (s % 5 == 3) ? F / 2 : F;
    for (int k = 0; k < F; ++k) fos[(size_t)s * F + k] = (short)((s % 5 == 3) ? (2 * k < F ? 2 * k : 0) : k);   // every other frame present
  }
  std::vector<uint8_t> tab((size_t)C * kHrHLimbs * kHrTabBytes);
  for (int c = 0; c < C; ++c) hrtf_build_table(&taps[(size_t)c * 2 * kHrTaps], &tab[(size_t)c * kHrHLimbs * kHrTabBytes]);
  std::vector<uint8_t> planes((size_t)S * C * NL * 4 * NBP * 16, 0);
  for (int s = 0; s < S; ++s)
    for (int c = 0; c < C; ++c)
      for (int p = 0; p < kHrHist + T; ++p) {
        const int v = x[((size_t)s * C + c) * (kHrHist + T) + p];
        const int bp = p / 64, j = p % 64, jr = 63 - j, kc = jr / 16, u = jr % 16;
        for (int l = 0; l < NL; ++l)
          planes[((((size_t)(s * C + c) * NL + l) * 4 + kc) * NBP + bp) * 16 + u] = (uint8_t)((v >> (8 * l)) & 255);
      }
  uint8_t *d_tab, *d_planes;
  float *d_out;
  int *d_np;
  short *d_fos;
  CK(cudaMalloc(&d_tab, tab.size()));
  CK(cudaMalloc(&d_planes, planes.size()));
  CK(cudaMalloc(&d_out, sizeof(float) * S * F * 2 * N));
  CK(cudaMalloc(&d_np, sizeof(int) * S));
  CK(cudaMalloc(&d_fos, sizeof(short) * S * F));
  CK(cudaMemcpy(d_tab, tab.data(), tab.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_planes, planes.data(), planes.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_np, n_present.data(), sizeof(int) * S, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_fos, fos.data(), sizeof(short) * S * F, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0xff, sizeof(float) * S * F * 2 * N));
  HrtfGemmArgs a;
  a.tab = d_tab; a.planes = d_planes; a.out = d_out; a.n_present = d_np; a.frame_of_slot = d_fos;
  memset(&a, 0, sizeof(a));
  a.tab = d_tab; a.planes = d_planes; a.out = d_out; a.n_present = d_np; a.frame_of_slot = d_fos;
  a.S = S; a.C = C; a.NL = NL; a.NB = NB; a.NT = NT; a.NBP = NBP; a.F = F; a.N = N; a.x_shift = x_shift;
  const int smem = kHrStages * hrtf_stage_bytes(NB, NL);
  auto kern = NL == 2 ? k_hrtf_gemm<2, false> : k_hrtf_gemm<3, false>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = S * NT < sms ? S * NT : sms;
  printf("smem %d grid %d\n", smem, grid);
  kern<<<grid, kHrThreads, smem>>>(a);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> out((size_t)S * F * 2 * N);
  CK(cudaMemcpy(out.data(), d_out, sizeof(float) * out.size(), cudaMemcpyDeviceToHost));
  // reference on a few streams
  long long bad = 0, checked = 0;
  const float scale = 1.0f / (float)(1ll << (x_shift + 15));
  for (int s = 0; s < S; s += (S > 8 ? S / 8 : 1)) {
    const int len = n_present[s] * N;
    for (int ear = 0; ear < 2; ++ear)
      for (int tau = 0; tau < len; ++tau) {
        long long acc = 0;
        for (int c = 0; c < C; ++c) {
          const int *xs = &x[((size_t)s * C + c) * (kHrHist + T) + kHrHist + tau];
          const int16_t *h = &taps[((size_t)c * 2 + ear) * kHrTaps];
          for (int k = 0; k < kHrTaps; ++k) acc += (long long)h[k] * xs[-k];
        }
        const float ref = (float)acc * scale;
        const int slot = tau / N, f = fos[(size_t)s * F + slot];
        const float got = out[(((size_t)s * F + f) * 2 + ear) * N + tau % N];
        ++checked;
        if (memcmp(&ref, &got, 4) != 0) {
          if (bad < 10) printf("mismatch s %d ear %d tau %d: ref %.9g got %.9g\n", s, ear, tau, ref, got);
          ++bad;
        }
      }
  }
  printf("checked %lld, mismatches %lld\n", checked, bad);
  // timing
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) kern<<<grid, kHrThreads, smem>>>(a);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) kern<<<grid, kHrThreads, smem>>>(a);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  double audio = 0;
  for (int s = 0; s < S; ++s) audio += n_present[s] * N / 48000.0;
  const double macs = (double)S * NT * C * 20 * NL * 2 * 128.0 * NB * 32;
  printf("%.3f ms per launch, %.0f audio-s/s, %.1f int8 TOPS issued\n", ms, audio / (ms * 1e-3), 2 * macs / (ms * 1e-3) / 1e12);
  return bad ? 2 : 0;
}
