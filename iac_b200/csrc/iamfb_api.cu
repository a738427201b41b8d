// iamfb_api.cu - host side of the C ABI declared in include/iamf_b200.h: plans, batches, kernel sequencing.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

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

// ---------------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
int iamfb_fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  fprintf(stderr, "[iamf_b200] error %d: %s\n", code, g_err);
  return code;
}

extern "C" const char *iamfb_last_error(void) { return g_err; }
extern "C" const char *iamfb_version(void) { return "iamf_b200 0.1 (sm_100a)"; }

// ---------------------------------------------------------------------------------------------------------------------
// static tables (host)
// ---------------------------------------------------------------------------------------------------------------------
// IAMF_utils.c:111-133
static const int k_layout_count[10] = {1, 2, 6, 8, 10, 8, 10, 12, 6, 2};
static const unsigned char k_layout_order[10][12] = {
    {13}, {14, 15}, {1, 2, 3, 4, 20, 21}, {1, 2, 3, 4, 20, 21, 22, 23}, {1, 2, 3, 4, 20, 21, 9, 10, 11, 12},
    {1, 2, 3, 4, 5, 6, 7, 8}, {1, 2, 3, 4, 5, 6, 7, 8, 22, 23}, {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12},
    {18, 19, 3, 4, 16, 17}, {14, 15}};
static const int k_layout_surround[10] = {1, 2, 5, 5, 5, 7, 7, 7, 3, 2};
static const int k_layout_top[10] = {0, 0, 0, 2, 4, 0, 2, 4, 2, 0};
// IAMF_decoder.c:208-219 (+LFE) / 3998-4008
static const int k_target_channels[IAMFB_TARGET_COUNT] = {2, 6, 8, 10, 11, 12, 14, 24, 8, 12, 10, 6, 1, 2};

extern "C" int iamfb_target_channels(int target) {
  return (target >= 0 && target < IAMFB_TARGET_COUNT) ? k_target_channels[target] : 0;
}
extern "C" int iamfb_layout_channels(int layout, int32_t *chs) {
  if (layout < 0 || layout > 9) return 0;
  if (chs)
    for (int i = 0; i < k_layout_count[layout]; ++i) chs[i] = k_layout_order[layout][i];
  return k_layout_count[layout];
}

static inline float bits2f(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}

extern "C" int iamfb_get_m2m_matrix(int layout, int target, int32_t *m, int32_t *n, float *mat) {
  for (size_t i = 0; i < sizeof(k_m2m_index) / sizeof(k_m2m_index[0]); ++i)
    if (k_m2m_index[i].in == layout && k_m2m_index[i].out == target) {
      if (m) *m = k_m2m_index[i].m;
      if (n) *n = k_m2m_index[i].n;
      if (mat)
        for (int k = 0; k < k_m2m_index[i].m * k_m2m_index[i].n; ++k) mat[k] = bits2f(k_matrix_pool[k_m2m_index[i].off + k]);
      return IAMFB_OK;
    }
  return IAMFB_ERR_BAD_ARG;
}

extern "C" int iamfb_get_h2m_matrix(int order, int target, int32_t *m, int32_t *n, int32_t *lfe1, int32_t *lfe2, float *mat) {
  for (size_t i = 0; i < sizeof(k_h2m_index) / sizeof(k_h2m_index[0]); ++i)
    if (k_h2m_index[i].order == order && k_h2m_index[i].out == target) {
      if (m) *m = k_h2m_index[i].m;
      if (n) *n = k_h2m_index[i].n;
      if (lfe1) *lfe1 = k_h2m_index[i].lfe1;
      if (lfe2) *lfe2 = k_h2m_index[i].lfe2;
      if (mat)
        for (int k = 0; k < k_h2m_index[i].m * k_h2m_index[i].n; ++k) mat[k] = bits2f(k_matrix_pool[k_h2m_index[i].off + k]);
      return IAMFB_OK;
    }
  return IAMFB_ERR_BAD_ARG;
}

// ---------------------------------------------------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------------------------------------------------
struct iamfb_plan {
  iamfb_ctx *ctx;
  iamfb_plan_desc desc;
  KernelPlan kp;
  int tmpl[kMaxEl];       // render kernel variant per element
  // device constant arrays
  float *d_start_win, *d_stop_win;   // recon cross-fade windows [overlap]
  float *d_qf;                       // 256
  float *d_sinc;                     // resampler table
  int sinc_len;
  float4 *d_tab4;                    // interpolating resampler: the four neighbouring taps per (offset, input) as one item
  int rs_span;                       // inputs staged per block of 128 outputs
  float *d_acc;                      // limiter acceleration curve by time index
  // initial per-stream state (host copy)
  StreamState init_state;
  // fused single-kernel path (non-resampling pipelines whose element signature is instantiated)
  bool fused;
  int fused_variant;       // 0: 4 samples per thread, 64 threads; 1: 2 samples, 128 threads; 2: 1 sample, 256 threads
  int fused_tile;          // samples per tile
  size_t fused_smem;       // dynamic shared memory per block
  // register-resident pipelined kernel (k_stream) for the channel-based single-element signatures it is instantiated
  // for; streams with trims / flushes / animated gains still take k_fused
  bool stream;
  int stream_sig;          // layout * 16 + target
  size_t stream_smem;
  // k_pipe (iamfb_pipe.cuh): double-buffered int16 / float32 staging, channel-based, scene-based and two-element signatures
  bool pipe;
  int pipe_sig;
  bool s16_native;         // int16 submits (IAMFB_IN_S16) are staged as they are by k_pipe / k_fused
  // k_pipe_rs (iamfb_pipe_rs.cuh): the resampling pipelines' regular streams; irregular ones and flushes: multi-kernel path
  bool rs_pipe;
  int rs_pipe_sig, rs_ring, rs_mirror;
  bool fma;                // IAMFB_ARITH_FMA asked for (the signature's fused variant is used where there is one)
  float4 *d_interp4;       // k_pipe_rs: cubic interpolation weights per phase
  float4 *d_tab4p;         // k_pipe_rs: the tap items with rs_tab_pad zero items on either side of every row
  int rs_tab_row, rs_tab_pad;
  // split form of the resampling pipelines (k_pipe_prerender + k_resample_ls + k_pipe_rs<PRE>) instead of the one kernel k_pipe_rs
  bool rs_split;
  int rs_ls_chunk;         // outputs per work item of k_resample_ls (upper bound unless fixed)
  bool rs_ls_chunk_fixed;
  // binaural HRTF front end (iamfb_hrtf.cu): non-null when an element is rendered through it; kp / desc then describe the
  // pipeline BEHIND it (those elements as 2-channel pass-through elements fed with float32 binaural frames)
  iamfb_hrtf_front *hrtf;
  int in_rows[kMaxEl];     // rows per frame of the CALLER's input of every element
  // HRTF elements whose layout channels the de-mixer derives (scalable layers, recon gain): de-mixed in front of the renderer
  // from their OWN plan / per-stream state (kp and the batch's state describe the pipeline behind the front end)
  KernelPlan *kp_front;
  bool front_demix[kMaxEl];
  int front_tmpl[kMaxEl];
  StreamState init_state_front;
};

static int rs_ls_smem(const iamfb_plan *p, int chunk, int *span_out);
static const int kLsSmemMax = 226 * 1024;   // dynamic shared memory of a k_resample_ls block (one block per SM; the SM offers 227 KB)

struct iamfb_batch {
  iamfb_plan *plan;
  int S, Fmax;
  int cap_a, cap_b;
  size_t out_stride;       // bytes per stream for Fmax frames
  StreamState *d_state;
  FrameRec *d_frames;
  SubmitRec *d_submit;
  int *d_gate;             // [2] "some stream of the submit is irregular", by submit parity (see ResolveArgs::gate)
  unsigned submit_seq;
  SubmitRec *d_submit_mk;  // resampling plans served by k_pipe_rs: the records the multi-kernel path works from (regular streams zeroed)
  float *d_tl_a, *d_tl_b, *d_pk, *d_wm, *d_gn;
  float *d_hist_y, *d_hist_pk;   // fused path: limiter delay line / peak ring carried between submits
  // staging for the host-resident path
  float *d_in[kMaxEl];
  int16_t *d_in16[kMaxEl];   // int16 uploads (IAMFB_IN_S16)
  float *d_wide[kMaxEl];     // float32 copy of an int16 submit for the kernels that stage float32
  float *d_ramp[kMaxEl];
  float *d_oramp;
  iamfb_frame_params *d_params;
  char *d_pcm;
  int32_t *d_counts;
  size_t stage_frames;     // frames the staging buffers are sized for (0 = not allocated)
  iamfb_hrtf_batch *hrtf;  // per-stream buffers of the binaural HRTF front end
  StreamState *d_state_f;   // HRTF front de-mixing: its own state / resolved frames / submit records, the de-mixed channels
  FrameRec *d_frames_f;
  SubmitRec *d_submit_f;
  float *d_demixed[kMaxEl];
  float *d_seg_ramp[kMaxEl], *d_seg_oramp;   // per-sample gains expanded from gain segments (k_gain_expand)
  iamfb_gain_ramp *d_segs[kMaxEl + 1];       // host-resident submits: the uploaded segment records
};

// ---------------------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int iamfb_ctx_create(int device, iamfb_ctx **out) {
  if (!out) return fail(IAMFB_ERR_BAD_ARG, "ctx_create: null out");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(IAMFB_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU path",
                e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(IAMFB_ERR_BAD_ARG, "device %d out of range (%d devices)", device, n);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(IAMFB_ERR_NO_DEVICE, "device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
  iamfb_ctx *c = new iamfb_ctx();
  c->device = device;
  c->n_sm = prop.multiProcessorCount;
  c->launches = 0;
  c->timing = false;
  c->own_stream = true;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
  for (int i = 0; i < kMaxSub; ++i) {
    CU(cudaEventCreateWithFlags(&c->ev_w[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_s[i], cudaEventDisableTiming));
  }
  CU(cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
  for (int i = 0; i < kMaxChunks; ++i) {
    CU(cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
  }
  CU(cudaEventCreateWithFlags(&c->ev_free, cudaEventDisableTiming));
  for (int i = 0; i < kMaxChunks; ++i) CU(cudaEventCreateWithFlags(&c->ev_back[i], cudaEventDisableTiming | cudaEventBlockingSync));
  *out = c;
  return IAMFB_OK;
}

extern "C" int iamfb_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

extern "C" int iamfb_ctx_set_stream(iamfb_ctx *c, void *stream) {
  if (!c) return fail(IAMFB_ERR_BAD_ARG, "null ctx");
  if (c->own_stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)stream;
  c->own_stream = false;
  return IAMFB_OK;
}

extern "C" int iamfb_ctx_synchronize(iamfb_ctx *c) {
  if (!c) return fail(IAMFB_ERR_BAD_ARG, "null ctx");
  CU(cudaStreamSynchronize(c->stream));
  return IAMFB_OK;
}

extern "C" void iamfb_ctx_destroy(iamfb_ctx *c) {
  if (!c) return;
  if (c->own_stream) cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->aux);
  cudaStreamDestroy(c->h2d);
  cudaStreamDestroy(c->d2h);
  for (int i = 0; i < kMaxSub; ++i) { cudaEventDestroy(c->ev_w[i]); cudaEventDestroy(c->ev_s[i]); }
  for (int i = 0; i < kMaxChunks; ++i) { cudaEventDestroy(c->ev_up[i]); cudaEventDestroy(c->ev_done[i]); cudaEventDestroy(c->ev_back[i]); }
  cudaEventDestroy(c->ev_free);
  delete c;
}

extern "C" uint64_t iamfb_ctx_launch_count(const iamfb_ctx *c) { return c ? c->launches : 0; }

extern "C" int iamfb_ctx_set_timing(iamfb_ctx *c, int enable) {
  if (!c) return fail(IAMFB_ERR_BAD_ARG, "null ctx");
  c->timing = enable != 0;
  if (enable) c->timers.clear();
  return IAMFB_OK;
}

extern "C" int iamfb_ctx_get_timing(iamfb_ctx *c, int index, const char **name, double *total_ms, uint64_t *launches) {
  if (!c || index < 0 || index >= (int)c->timers.size()) return IAMFB_ERR_BAD_ARG;
  KernelTimer &k = c->timers[index];
  if (!k.pending.empty()) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->aux));
    for (auto &p : k.pending) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, p.first, p.second);
      k.total_ms += ms;
      k.samples.push_back(ms);
      c->event_pool.push_back(p.first);
      c->event_pool.push_back(p.second);
    }
    k.pending.clear();
  }
  if (name) *name = k.name;
  if (total_ms) *total_ms = k.total_ms;
  if (launches) *launches = k.launches;
  return IAMFB_OK;
}

extern "C" int iamfb_ctx_get_timing_median(iamfb_ctx *c, int index, double *median_ms) {
  double total = 0;
  uint64_t n = 0;
  const char *name = nullptr;
  int r = iamfb_ctx_get_timing(c, index, &name, &total, &n);   // (drains the pending events)
  if (r) return r;
  std::vector<float> v = c->timers[index].samples;
  if (v.empty()) { if (median_ms) *median_ms = 0; return IAMFB_OK; }
  std::nth_element(v.begin(), v.begin() + v.size() / 2, v.end());
  if (median_ms) *median_ms = v[v.size() / 2];
  return IAMFB_OK;
}

extern "C" void *iamfb_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    fail(IAMFB_ERR_ALLOC_FAIL, "cudaMallocHost(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
extern "C" void iamfb_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------------------------------------------------------------
// plan construction (host arithmetic uses the same libm calls as the reference so that the tables are bit-identical)
// ---------------------------------------------------------------------------------------------------------------------
namespace {

// Kaiser window tables and quality map of the Speex resampler (resample.c:105-207); only quality 4 is used by the
// decoder (IAMF_decoder.c:57 SPEEX_RESAMPLER_QUALITY) but Q0..Q8 share the code path.
const double kKaiser8[36] = {
    0.99635258, 1.00000000, 0.99635258, 0.98548012, 0.96759014, 0.94302200, 0.91223751, 0.87580811, 0.83439927,
    0.78875245, 0.73966538, 0.68797126, 0.63451750, 0.58014482, 0.52566725, 0.47185369, 0.41941150, 0.36897272,
    0.32108304, 0.27619388, 0.23465776, 0.19672670, 0.16255380, 0.13219758, 0.10562887, 0.08273982, 0.06335451,
    0.04724088, 0.03412321, 0.02369490, 0.01563093, 0.00959968, 0.00527363, 0.00233883, 0.00050000, 0.00000000};
constexpr int kQ4Len = 64, kQ4Over = 8, kWinOver = 32;
constexpr float kQ4DownBw = 0.921f, kQ4UpBw = 0.940f;

double kaiser_window(float x) {   // compute_func, resample.c:210-229
  float y = x * kWinOver;
  int ind = (int)floor(y);
  float frac = (y - ind);
  double i3 = -0.1666666667 * frac + 0.1666666667 * (frac * frac * frac);
  double i2 = frac + 0.5 * (frac * frac) - 0.5 * (frac * frac * frac);
  double i0 = -0.3333333333 * frac + 0.5 * (frac * frac) - 0.1666666667 * (frac * frac * frac);
  double i1 = 1.f - i3 - i2 - i0;
  return i0 * kKaiser8[ind] + i1 * kKaiser8[ind + 1] + i2 * kKaiser8[ind + 2] + i3 * kKaiser8[ind + 3];
}

float windowed_sinc(float cutoff, float x, int N) {   // sinc, resample.c:233-244
  float xx = x * cutoff;
  if (fabs(x) < 1e-6) return cutoff;
  if (fabs(x) > .5 * N) return 0;
  return cutoff * sin(M_PI * xx) / (M_PI * xx) * kaiser_window(fabs(2. * x / N));
}

uint32_t gcd_u32(uint32_t a, uint32_t b) {
  while (b) { uint32_t t = a % b; a = b; b = t; }
  return a;
}

// update_filter, resample.c:527-640 for a freshly created quality-4 resampler
void build_resampler(KernelPlan &kp, std::vector<float> &table, uint32_t in_rate, uint32_t out_rate) {
  uint32_t g = gcd_u32(in_rate, out_rate);
  kp.rs_num = in_rate / g;
  kp.rs_den = out_rate / g;
  kp.rs_int_adv = kp.rs_num / kp.rs_den;
  kp.rs_frac_adv = kp.rs_num % kp.rs_den;
  uint32_t oversample = kQ4Over, filt_len = kQ4Len;
  float cutoff;
  if (kp.rs_num > kp.rs_den) {
    cutoff = kQ4DownBw * kp.rs_den / kp.rs_num;
    uint32_t major = filt_len / kp.rs_den, remain = filt_len % kp.rs_den;
    filt_len = remain * kp.rs_num / kp.rs_den + major * kp.rs_num;
    filt_len = ((filt_len - 1) & (~0x7U)) + 8;
    if (2 * kp.rs_den < kp.rs_num) oversample >>= 1;
    if (4 * kp.rs_den < kp.rs_num) oversample >>= 1;
    if (8 * kp.rs_den < kp.rs_num) oversample >>= 1;
    if (16 * kp.rs_den < kp.rs_num) oversample >>= 1;
    if (oversample < 1) oversample = 1;
  } else {
    cutoff = kQ4UpBw;
  }
  kp.rs_filt_len = filt_len;
  kp.rs_oversample = oversample;
  kp.rs_direct = filt_len * kp.rs_den <= filt_len * oversample + 8;
  if (kp.rs_direct) {
    table.resize((size_t)filt_len * kp.rs_den);
    for (uint32_t i = 0; i < kp.rs_den; ++i)
      for (int32_t j = 0; j < (int32_t)filt_len; ++j)
        table[i * filt_len + j] = windowed_sinc(cutoff, ((j - (int32_t)filt_len / 2 + 1) - ((float)i) / kp.rs_den), filt_len);
  } else {
    table.resize((size_t)filt_len * oversample + 8);
    for (int32_t i = -4; i < (int32_t)(oversample * filt_len + 4); ++i)
      table[i + 4] = windowed_sinc(cutoff, (i / (float)oversample - filt_len / 2), filt_len);
  }
}

// curve_accel, audio_effect_peak_limiter.c:267-271
float accel_curve(float x) {
  if (1.0 < x) return 1.0f;
  if (x < 0) return 0.0f;
  return 1.0f - powf(x - 1, 2.0);
}

// The limiter's float time constant restarts at 0 on every trigger and then only accumulates incTC, so it walks a fixed
// sequence T[j]; tabulate curve_accel along it (compute_target_gain, audio_effect_peak_limiter.c:237-256).
void build_limiter(KernelPlan &kp, std::vector<float> &acc, float thr_db, int rate) {
  const float atk = 0.001f, rel = 0.200f;   // audio_defines.h:39-40
  const float inc = (float)1 / (float)rate;
  kp.lim_thr = pow(10, thr_db / 20);         // double pow rounded to float on store (:79)
  acc.clear();
  acc.push_back(0.f);
  float t = 0.0f;
  int j = 0;
  kp.lim_ja = -1;
  for (;;) {
    if (t < atk) {
      t += inc; ++j;
      acc.push_back(accel_curve(t / atk));
    } else if (t < rel + atk) {
      if (kp.lim_ja < 0) kp.lim_ja = j;
      t += inc; ++j;
      acc.push_back(accel_curve((t - atk) / rel));
    } else {
      break;
    }
    if (j > (1 << 22)) break;
  }
  if (kp.lim_ja < 0) kp.lim_ja = j;
  kp.lim_jr = j;
  for (int k = 0; k < 3; ++k) acc.push_back(0.f);   // padding: the scan kernel prefetches up to index j+3
}

template <typename T>
int upload(T **dst, const T *src, size_t n) {
  CU(cudaMalloc((void **)dst, n * sizeof(T) > 0 ? n * sizeof(T) : sizeof(T)));
  if (n) CU(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return IAMFB_OK;
}

// dependency resolution of dmx_channel (demixer.c:380-419) on the host: which derivations does the layout need?
struct Avail {
  bool have[kChCount];
  ElPlan *ep;
  bool s2() {
    if (!have[IAMFB_CH_L2]) return false;
    if (have[IAMFB_CH_R2]) return true;
    if (!have[IAMFB_CH_MONO]) return false;
    ep->need_s2 = 1; have[IAMFB_CH_R2] = true; return true;
  }
  bool s3() {
    if (have[IAMFB_CH_R3]) return true;
    if (!s2()) return false;
    if (!have[IAMFB_CH_C]) return false;
    ep->need_s3 = 1; have[IAMFB_CH_L3] = have[IAMFB_CH_R3] = true; return true;
  }
  bool s5() {
    if (have[IAMFB_CH_SR5]) return true;
    if (!s3()) return false;
    if (!have[IAMFB_CH_L5] || !have[IAMFB_CH_R5]) return false;
    ep->need_s5 = 1; have[IAMFB_CH_SL5] = have[IAMFB_CH_SR5] = true; return true;
  }
  bool s7() {
    if (have[IAMFB_CH_BR7]) return true;
    if (!s5()) return false;
    if (!have[IAMFB_CH_SL7] || !have[IAMFB_CH_SR7]) return false;
    ep->need_s7 = 1; have[IAMFB_CH_BL7] = have[IAMFB_CH_BR7] = true; return true;
  }
  bool h2() {
    if (have[IAMFB_CH_HR]) return true;
    if (!have[IAMFB_CH_TL] || !have[IAMFB_CH_TR]) return false;
    if (!s5()) return false;
    ep->need_h2 = 1; have[IAMFB_CH_HL] = have[IAMFB_CH_HR] = true; return true;
  }
  bool h4() {
    if (have[IAMFB_CH_HBR]) return true;
    if (!h2()) return false;
    if (!have[IAMFB_CH_HFR] || !have[IAMFB_CH_HFL]) return false;
    ep->need_h4 = 1; have[IAMFB_CH_HBL] = have[IAMFB_CH_HBR] = true; return true;
  }
  bool channel(int ch) {
    if (have[ch]) return true;
    switch (ch) {
      case IAMFB_CH_R2: return s2();
      case IAMFB_CH_L3: case IAMFB_CH_R3: return s3();
      case IAMFB_CH_SL5: case IAMFB_CH_SR5: return s5();
      case IAMFB_CH_BL7: case IAMFB_CH_BR7: return s7();
      case IAMFB_CH_HL: case IAMFB_CH_HR: return h2();
      case IAMFB_CH_HBL: case IAMFB_CH_HBR: return h4();
      default: return false;
    }
  }
};

int recon_flags_default(int l1, int l2) {   // iamf_recon_channels_get_flags, IAMF_decoder.c:371-407
  if (l1 == l2) return 0;
  int s1 = k_layout_surround[l1], s2 = k_layout_surround[l2], t1 = k_layout_top[l1], t2 = k_layout_top[l2];
  int f = 0;
  if (s1 != s2) {
    if (s2 <= 3) f |= (1 << 0) | (1 << 2);
    else if (s2 == 5) f |= (1 << 3) | (1 << 4);
    else if (s2 == 7) f |= (1 << 7) | (1 << 8);
  }
  if (t2 != t1 && t2 == 4) f |= (1 << 9) | (1 << 10);
  if (s2 == 5 && t1 && t2 == t1) f |= (1 << 5) | (1 << 6);
  return f;
}

const unsigned char k_recon_map_host[9][12] = {
    {13, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},        {14, 0, 15, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {1, 3, 2, 20, 21, 0, 0, 0, 0, 0, 0, 4},       {1, 3, 2, 20, 21, 22, 23, 0, 0, 0, 0, 4},
    {1, 3, 2, 20, 21, 9, 10, 0, 0, 11, 12, 4},    {1, 3, 2, 5, 6, 0, 0, 7, 8, 0, 0, 4},
    {1, 3, 2, 5, 6, 22, 23, 7, 8, 0, 0, 4},       {1, 3, 2, 5, 6, 9, 10, 7, 8, 11, 12, 4},
    {18, 3, 19, 0, 0, 16, 17, 0, 0, 0, 0, 4}};

const float k_mix_gamma_host[8] = {0.707f, 0.707f, 0.866f, 0.f, 0.707f, 0.707f, 0.866f, 0.f};
const float k_w_host[11] = {0.0f, 0.0179f, 0.0391f, 0.0658f, 0.1038f, 0.25f, 0.3962f, 0.4342f, 0.4609f, 0.4821f, 0.5f};

bool valid_dmr_pair(int in, int out) {   // DMRenderer_open validity, downmix_renderer.c:77-91,131-139
  if (in == out || in < 0 || in > 8 || out < 0 || out > 8) return false;
  int s1 = k_layout_surround[in], s2 = k_layout_surround[out], t1 = k_layout_top[in], t2 = k_layout_top[out];
  if (t1 && !t2) return false;
  return !(s1 < s2 || t1 < t2);
}

}  // namespace

static int build_element(const iamfb_plan_desc &d, int e, KernelPlan &kp, int &tmpl, StreamState &init) {
  const iamfb_element_desc &ed = d.el[e];
  ElPlan &ep = kp.el[e];
  ElState &es = init.el[e];
  memset(&ep, 0, sizeof(ep));
  ep.kind = ed.kind;
  ep.n_in = ed.n_in;
  for (int c = 0; c < kChCount; ++c) ep.src_row[c] = -1;
  for (int i = 0; i < kMaxOut; ++i) ep.out_slot[i] = -1;
  const int co = kp.out_channels;

  if (ed.kind == IAMFB_EL_CHANNEL) {
    if (ed.layout < 0 || ed.layout > 9) return fail(IAMFB_ERR_BAD_ARG, "element %d: bad layout %d", e, ed.layout);
    const int lay = ed.layout == IAMFB_LAYOUT_BINAURAL ? IAMFB_LAYOUT_STEREO : ed.layout;
    ep.layout = lay;
    ep.n_rec = k_layout_count[lay];
    if (ed.first_layer_layout < 0 || ed.first_layer_layout > 9) return fail(IAMFB_ERR_BAD_ARG, "element %d: bad first-layer layout %d", e, ed.first_layer_layout);
    if (ed.use_dmr && (ed.dmr_out_layout < 0 || ed.dmr_out_layout > 9)) return fail(IAMFB_ERR_BAD_ARG, "element %d: bad down-mix layout %d", e, ed.dmr_out_layout);
    if (ed.n_in != ep.n_rec)   // demixer_demixing: chs_count must equal the layout's channel count (demixer.c:640-641)
      return fail(IAMFB_ERR_BAD_ARG, "element %d: %d decoded channels but layout %d has %d", e, ed.n_in, lay, ep.n_rec);
    Avail av;
    memset(&av, 0, sizeof(av));
    av.ep = &ep;
    for (int i = 0; i < ed.n_in; ++i) {
      int ch = ed.chs_in[i];
      if (ch <= 0 || ch >= kChCount) return fail(IAMFB_ERR_BAD_ARG, "element %d: bad channel id %d", e, ch);
      ep.src_row[ch] = (signed char)i;
      av.have[ch] = true;
    }
    for (int i = 0; i < ed.n_out_gain; ++i) {
      int ch = ed.out_gain_ch[i];
      if (ch <= 0 || ch >= kChCount) continue;
      if (ep.src_row[ch] < 0) continue;           // dmx_gainup skips channels without data
      if ((ep.gain_mask >> ch) & 1u) return fail(IAMFB_ERR_UNIMPLEMENTED, "element %d: channel %d listed twice in output gain", e, ch);
      ep.gain_mask |= 1u << ch;
      ep.gain[ch] = ed.out_gain[i];
    }
    for (int c = 0; c < kChCount; ++c) ep.slot_of[c] = ep.recon_bit[c] = -1;
    for (int b = 0; b < 12; ++b)
      if (k_recon_map_host[lay][b]) ep.recon_bit[k_recon_map_host[lay][b]] = (signed char)b;
    for (int m = 0; m < ep.n_rec; ++m) {
      ep.rec_ch[m] = k_layout_order[lay][m];
      ep.slot_of[ep.rec_ch[m]] = (signed char)m;
      if (!av.channel(ep.rec_ch[m]))
        return fail(IAMFB_ERR_BAD_ARG, "element %d: layout channel %d cannot be reconstructed from the decoded channels", e, ep.rec_ch[m]);
    }
    ep.recon_present = ed.recon_present;
    tmpl = lay;
    // demixer initial state: iamf_stream_scale_demixer_configure, IAMF_decoder.c:2351-2401
    es.mode = 0;
    es.w_idx = 0;
    if (ed.has_demix_info) {
      int mode = ed.default_mode, w = ed.default_w_idx;
      if (!(mode < 0 || mode == 3 || mode > 6)) {
        if (w < 0 || w > 10) {
          es.mode = mode;
          es.w_idx = (mode >= 4) ? 1 : 0;
        } else {
          es.mode = mode;
          es.w_idx = w;
        }
      }
    }
    for (int c = 0; c < kChCount; ++c) es.sfavg[c] = 1.0f;
    if (ed.selected_layer > 0) {   // iamf_stream_scale_decoder_set_default_recon_gain, :2202-2236
      int fl = recon_flags_default(ed.first_layer_layout, lay);
      int n = 0;
      for (int b = 0; b < 12; ++b)
        if (fl & (1 << b)) { es.rgain[k_recon_map_host[lay][b]] = 1.f; ++n; }
      if (fl) es.rflags = fl;
    }
    es.re_flags = 0;

    if (ed.use_dmr) {
      if (!valid_dmr_pair(lay, ed.dmr_out_layout))
        return fail(IAMFB_ERR_BAD_ARG, "element %d: layouts %d -> %d are not a valid parametric down-mix", e, lay, ed.dmr_out_layout);
      if (k_layout_count[ed.dmr_out_layout] != co)
        return fail(IAMFB_ERR_BAD_ARG, "element %d: down-mix layout has %d channels, target %d", e, k_layout_count[ed.dmr_out_layout], co);
      ep.renderer = kRdrDMR;
      ep.dmr_n_out = k_layout_count[ed.dmr_out_layout];
      for (int i = 0; i < ep.dmr_n_out; ++i) ep.dmr_out_ch[i] = k_layout_order[ed.dmr_out_layout][i];
      for (int m = 0; m < ep.n_rec; ++m) ep.dmr_in_mask |= 1u << ep.rec_ch[m];
      // DMRenderer_open + DMRenderer_set_mode_weight(default_mode, default_w_idx)
      es.dmr_mode = -1;
      es.dmr_w_idx = -1;
      es.dmr_tl = 0.f;
      int mode = ed.default_mode;
      if (mode >= 0 && mode != 3 && mode < 7) {
        es.dmr_mode = mode;
        bool tl_derived = !((ep.dmr_in_mask >> IAMFB_CH_TL) & 1u) && !((ep.dmr_in_mask >> IAMFB_CH_TR) & 1u);
        int w = ed.default_w_idx;
        if (w < 0 || w > 10) {
          int nw = (mode >= 4) ? 0 : 0;   // calc_w from w_idx -1: min(-1+1,10)=0 or max(-2,0)=0
          es.dmr_w_idx = nw;
          if (tl_derived) es.dmr_tl = k_mix_gamma_host[mode] * k_w_host[nw];
        } else {
          es.dmr_w_idx = w;
          if (tl_derived) es.dmr_tl = k_mix_gamma_host[mode] * k_w_host[w];
        }
      }
    } else {
      ep.renderer = kRdrM2M;
      int32_t m = 0, n = 0;
      std::vector<float> mat(24 * 16);
      // iamf_stream_render: a one-channel element renders as IAMF_MONO, else by its layout (IAMF_decoder.c:2593-2598)
      if (iamfb_get_m2m_matrix(ed.layout, d.target, &m, &n, mat.data()) != IAMFB_OK)
        return fail(IAMFB_ERR_BAD_ARG, "element %d: no channel matrix for layout %d -> target %d", e, ed.layout, d.target);
      if (m != ep.n_rec || n != co) return fail(IAMFB_ERR_INTERNAL, "matrix shape %dx%d vs %dx%d", m, n, ep.n_rec, co);
      ep.n_mat_out = n;
      for (int o = 0; o < n; ++o) {
        ep.out_slot[o] = (signed char)o;
        for (int i = 0; i < m; ++i) ep.mat[o * ep.n_rec + i] = mat[i * n + o];   // [in][out] -> output-major
      }
    }
  } else if (ed.kind == IAMFB_EL_SCENE) {
    int nch = ed.ambi_channels;
    if (ed.n_in < 1 || ed.n_in > IAMFB_MAX_SCENE_CH) return fail(IAMFB_ERR_BAD_ARG, "element %d: %d decoded rows of a scene-based element", e, ed.n_in);
    int order = nch == 1 ? 0 : nch == 4 ? 1 : nch == 9 ? 2 : nch == 16 ? 3 : -1;   // iamf_stream_ambisionisc_order :2403-2413
    if (order < 0) return fail(IAMFB_ERR_BAD_ARG, "element %d: %d ambisonics channels", e, nch);
    ep.n_rec = nch;
    ep.renderer = kRdrH2M;
    ep.ambi_mode = ed.ambi_mode;
    if (ed.ambi_mode == 0) {
      for (int i = 0; i < nch; ++i) {
        if (ed.ambi_map[i] >= ed.n_in) return fail(IAMFB_ERR_BAD_ARG, "element %d: ambisonics map[%d]=%d >= %d rows", e, i, ed.ambi_map[i], ed.n_in);
        ep.ambi_map[i] = ed.ambi_map[i];
      }
    } else {
      if (ed.ambi_cols <= 0 || ed.ambi_cols > 16 || ed.ambi_cols > ed.n_in) return fail(IAMFB_ERR_BAD_ARG, "element %d: projection columns %d", e, ed.ambi_cols);
      ep.ambi_cols = ed.ambi_cols;
      for (int l = 0; l < ed.ambi_cols; ++l)
        for (int r = 0; r < nch; ++r) ep.ambi_mat[l * nch + r] = ed.ambi_matrix[l * nch + r];
    }
    int32_t m = 0, n = 0, l1 = -1, l2 = -1;
    std::vector<float> mat(24 * 16);
    if (iamfb_get_h2m_matrix(order, d.target, &m, &n, &l1, &l2, mat.data()) != IAMFB_OK)
      return fail(IAMFB_ERR_BAD_ARG, "element %d: no HOA matrix for order %d -> target %d", e, order, d.target);
    if (m != nch) return fail(IAMFB_ERR_INTERNAL, "HOA matrix inputs %d vs %d", m, nch);
    ep.n_mat_out = n;
    for (int i = 0; i < n * m; ++i) ep.mat[i] = mat[i];
    // LFE slot shift (h2m_rdr.c:1114-1135); LFE rows themselves are zero (DISABLE_LFE_HOA, :1137-1150); rows the
    // reference never writes (row 23 of sound system H) are zero here.
    int map[24];
    for (int i = 0; i < n; ++i) map[i] = i;
    if (l1 >= 0 || l2 >= 0) {
      int k = 0;
      for (int i = 0; i < n; ++i) {
        if (l1 == i) k++;
        if (l2 == i) k++;
        map[i] = k;
        k++;
      }
    }
    for (int i = 0; i < n; ++i)
      if (map[i] < co) ep.out_slot[map[i]] = (signed char)i;
    // a row later zeroed as an LFE slot: the reference moves first, then zeroes lfe1/lfe2 (:1137-1150)
    if (l1 >= 0 && l1 < co) ep.out_slot[l1] = -1;
    if (l2 >= 0 && l2 < co) ep.out_slot[l2] = -1;
    tmpl = 10 + order;
  } else {
    return fail(IAMFB_ERR_BAD_ARG, "element %d: unknown kind %d", e, ed.kind);
  }
  return IAMFB_OK;
}

// element-signature combinations the fused kernel is instantiated for: every single-element pipeline, and the
// two-element pairs of the named configurations (7.1.4 + first-order ambisonics in either order); anything else
// runs on the multi-kernel path
static int fused_variant(const int *tmpl, int n_elements, bool dmr0 = false) {
  if (n_elements == 1 && dmr0) return (tmpl[0] >= 2 && tmpl[0] <= 8) ? 200 + tmpl[0] : -1;   // parametric down-mixer (layouts with surrounds)
  if (n_elements == 1) return tmpl[0];
  if (n_elements == 2 && tmpl[0] == 7 && tmpl[1] == 11) return 100;
  if (n_elements == 2 && tmpl[0] == 11 && tmpl[1] == 7) return 101;
  if (n_elements == 2 && tmpl[0] == 1 && tmpl[1] == 1) return 102;    // two binaural frames behind the HRTF front end
  return -1;
}

static int launch_fused(iamfb_ctx *ctx, const iamfb_plan *p, const FusedArgs &fa, int S) {
  const KernelPlan &kp = p->kp;
  cudaStream_t st = ctx->stream;
  const size_t smem = p->fused_smem;
#define FLAUNCH(L0, N0, L1, N1, VEC, THREADS)                                                                         \
  {                                                                                                                   \
    CU(cudaFuncSetAttribute(k_fused<L0, N0, L1, N1, VEC, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    ScopedKernelTimer tm_(ctx, "k_fused");                                                                            \
    cudaLaunchConfig_t lc_ = {};                                                                                      \
    lc_.gridDim = dim3(S); lc_.blockDim = dim3(THREADS); lc_.dynamicSmemBytes = smem; lc_.stream = st;                \
    cudaLaunchAttribute at_[1];                                                                                       \
    at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                   \
    /* the launch for the streams k_stream left alone touches none of k_stream's streams: it may run beside it */    \
    at_[0].val.programmaticStreamSerializationAllowed = fa.only_irregular ? 1 : 0;                                    \
    lc_.attrs = at_; lc_.numAttrs = 1;                                                                                \
    CU(cudaLaunchKernelEx(&lc_, k_fused<L0, N0, L1, N1, VEC, THREADS>, kp, fa));                                      \
  }
#define FCASE(ID, L0, N0, L1, N1) \
  case ID: FLAUNCH(L0, N0, L1, N1, 4, 64) break;
#define FLAUNCH_DMR(L0, N0)                                                                                           \
  {                                                                                                                   \
    CU(cudaFuncSetAttribute(k_fused<L0, N0, 0, 0, 4, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    ScopedKernelTimer tm_(ctx, "k_fused");                                                                            \
    k_fused<L0, N0, 0, 0, 4, 64, true><<<S, 64, smem, st>>>(kp, fa);                                                  \
  }
#define FCASE_DMR(ID, L0, N0) \
  case ID: FLAUNCH_DMR(L0, N0) break;
  // scene-based pipelines with many output channels fit few streams per SM: lighter threads, more of them
#define FCASE3(ID, L0, N0, L1, N1)                                    \
  case ID:                                                            \
    if (p->fused_variant == 1) FLAUNCH(L0, N0, L1, N1, 2, 128)        \
    else FLAUNCH(L0, N0, L1, N1, 4, 64)                               \
    break;
  switch (fused_variant(p->tmpl, kp.n_elements, kp.el[0].renderer == kRdrDMR)) {
    FCASE_DMR(202, 2, 6) FCASE_DMR(203, 3, 8) FCASE_DMR(204, 4, 10) FCASE_DMR(205, 5, 8) FCASE_DMR(206, 6, 10) FCASE_DMR(207, 7, 12) FCASE_DMR(208, 8, 6)
    FCASE(0, 0, 1, 0, 0) FCASE(1, 1, 2, 0, 0) FCASE(2, 2, 6, 0, 0) FCASE(3, 3, 8, 0, 0) FCASE(4, 4, 10, 0, 0)
    FCASE(5, 5, 8, 0, 0) FCASE(6, 6, 10, 0, 0) FCASE(7, 7, 12, 0, 0) FCASE(8, 8, 6, 0, 0)
    FCASE3(10, -1, 1, 0, 0) FCASE3(11, -1, 4, 0, 0) FCASE3(12, -1, 9, 0, 0) FCASE3(13, -1, 16, 0, 0)
    FCASE(100, 7, 12, -1, 4) FCASE(101, -1, 4, 7, 12) FCASE(102, 1, 2, 1, 2)
    default: return fail(IAMFB_ERR_INTERNAL, "no fused kernel variant");
  }
#undef FCASE3
#undef FLAUNCH
#undef FCASE
  cudaError_t e_ = cudaGetLastError();
  if (e_ != cudaSuccess) return fail(IAMFB_ERR_CUDA, "launch of k_fused failed: %s", cudaGetErrorString(e_));
  ++ctx->launches;
  return IAMFB_OK;
}

// staged-row byte offsets and the row-compressed render matrix of the fused kernels, for a tile of tl samples
static void fill_fused_offsets(KernelPlan &kp, int tl, int nin) {
  const int co = kp.out_channels;
  int row_base = 0;
  for (int e = 0; e < kp.n_elements; ++e) {
    ElPlan &ep = kp.el[e];
    ep.f_row_off = row_base * tl * 4;
    for (int c = 0; c < kChCount; ++c) {
      ep.f_src_off[c] = (ep.kind == IAMFB_EL_CHANNEL && ep.src_row[c] >= 0) ? (row_base + ep.src_row[c]) * tl * 4 : nin * tl * 4;
      ep.f_gain[c] = ((ep.gain_mask >> c) & 1u) ? ep.gain[c] : 1.0f;
    }
    ep.f_n_gain = 0;
    if (ep.kind == IAMFB_EL_CHANNEL)
      for (int c = 1; c < kChCount; ++c)
        if (((ep.gain_mask >> c) & 1u) && ep.src_row[c] >= 0 && ep.f_n_gain < IAMFB_MAX_LAYOUT_CH) {
          ep.f_gain_off[ep.f_n_gain] = ep.f_src_off[c];
          ep.f_gain_val[ep.f_n_gain] = ep.gain[c];
          ++ep.f_n_gain;
        }
    int q = 0;
    for (int oc = 0; oc < co; ++oc) {
      ep.f_csr_ptr[oc] = (unsigned short)q;
      const int n = ep.out_slot[oc];
      if (n < 0) continue;
      for (int m = 0; m < ep.n_rec; ++m) {
        const float c = ep.mat[n * ep.n_rec + m];
        if (c == 0.f) continue;       // adding +-0 never changes the running sum (it starts at +0): exact skip
        // reconstructed channel m is written back over staged row m of the element; a mono-mapped ambisonics
        // channel is read straight from the decoded row it maps to
        const int xrow = (ep.kind == IAMFB_EL_SCENE && ep.ambi_mode == 0) ? ep.ambi_map[m] : m;
        ep.f_csr_off[q] = (row_base + xrow) * tl * 4;
        ep.f_csr_val[q] = c;
        ++q;
      }
    }
    for (int oc = co; oc <= kMaxOut; ++oc) ep.f_csr_ptr[oc] = (unsigned short)q;
    row_base += ep.n_in;
  }
}

// (layout, target) pairs k_stream is instantiated for
// (7.1.4 / 7.1 / 5.1.4 / 5.1 / stereo sources to stereo, 5.1 and the binaural target as the reference builds it)
#define IAMFB_STREAM_SIGS(X) X(7, 1) X(1, 0) X(7, 0) X(7, 13) X(5, 1) X(4, 1) X(2, 1) X(2, 0) X(1, 1) X(1, 13)
static bool stream_sig_exists(int layout, int target) {
#define X(L, T) if (layout == L && target == T) return true;
  IAMFB_STREAM_SIGS(X)
#undef X
  return false;
}
// Tensor map of one submit's decoded input [S*F*n_in rows][N] (f32): k_stream stages the n_in rows x 240 instants of a
// tile with ONE cp.async.bulk.tensor instruction (a box of {240, n_in}) instead of one bulk copy per row - the per-row
// copies took a per-lane issue loop on the worker warp that was a third of the tile's critical path (profiles/).
// The encoder comes from the driver through the runtime (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}
static int make_input_tmap(CUtensorMap *tm, const void *base, size_t rows, int N, int n_in, bool s16) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return fail(IAMFB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  if ((((size_t)base) & 15) != 0) return fail(IAMFB_ERR_BAD_ARG, "decoded input must be 16-byte aligned");
  const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)N * (s16 ? 2 : 4)};
  const cuuint32_t box[2] = {(cuuint32_t)kStreamTile, (cuuint32_t)n_in};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, s16 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(IAMFB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return IAMFB_OK;
}

static int launch_stream(iamfb_ctx *ctx, const iamfb_plan *p, const FusedArgs &fa, int S) {
  const KernelPlan &kp = p->kp;
  cudaStream_t st = ctx->stream;
  const size_t smem = p->stream_smem;
  bool done = false;
  CUtensorMap tm;
  {
    int r = make_input_tmap(&tm, fa.in[0], (size_t)S * fa.n_frames * kp.el[0].n_in, kp.frame_size, kp.el[0].n_in, false);
    if (r) return r;
  }
#define X(L, T)                                                                                                       \
  if (!done && p->stream_sig == L * 16 + T) {                                                                         \
    CU(cudaFuncSetAttribute(k_stream<L, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
    ScopedKernelTimer tm_(ctx, "k_stream");                                                                           \
    cudaLaunchConfig_t lc_ = {};                                                                                      \
    lc_.gridDim = dim3(S); lc_.blockDim = dim3(kStreamThreads); lc_.dynamicSmemBytes = smem; lc_.stream = st;         \
    cudaLaunchAttribute at_[1];                                                                                       \
    at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                   \
    at_[0].val.programmaticStreamSerializationAllowed = 1;   /* may start under k_resolve (see the kernels) */        \
    lc_.attrs = at_; lc_.numAttrs = 1;                                                                                \
    CU(cudaLaunchKernelEx(&lc_, k_stream<L, T>, kp, fa, tm));                                                         \
    done = true;                                                                                                      \
  }
  IAMFB_STREAM_SIGS(X)
#undef X
  if (!done) return fail(IAMFB_ERR_INTERNAL, "no k_stream variant");
  cudaError_t e_ = cudaGetLastError();
  if (e_ != cudaSuccess) return fail(IAMFB_ERR_CUDA, "launch of k_stream failed: %s", cudaGetErrorString(e_));
  ++ctx->launches;
  return IAMFB_OK;
}

extern "C" void iamfb_plan_destroy(iamfb_plan *p);

extern "C" int iamfb_plan_create(iamfb_ctx *ctx, const iamfb_plan_desc *d, iamfb_plan **out) {
  if (!ctx || !d || !out) return fail(IAMFB_ERR_BAD_ARG, "plan_create: null argument");
  if (d->frame_size <= 0 || d->frame_size > 32768) return fail(IAMFB_ERR_BAD_ARG, "frame_size %d", d->frame_size);
  if (d->n_elements < 1 || d->n_elements > kMaxEl) return fail(IAMFB_ERR_BAD_ARG, "n_elements %d", d->n_elements);
  if (d->target < 0 || d->target >= IAMFB_TARGET_COUNT) return fail(IAMFB_ERR_BAD_ARG, "target %d", d->target);
  if (d->bit_depth != 0 && d->bit_depth != 16 && d->bit_depth != 24 && d->bit_depth != 32)
    return fail(IAMFB_ERR_BAD_ARG, "bit_depth %d", d->bit_depth);
  if (d->arithmetic != IAMFB_ARITH_EXACT && d->arithmetic != IAMFB_ARITH_FMA) return fail(IAMFB_ERR_BAD_ARG, "arithmetic %d", d->arithmetic);
  if (d->in_rate <= 0 || d->out_rate <= 0) return fail(IAMFB_ERR_BAD_ARG, "rates %d -> %d", d->in_rate, d->out_rate);
  CU(cudaSetDevice(ctx->device));
  for (int e = 0; e < d->n_elements; ++e)
    if (d->el[e].n_in < 1 || d->el[e].n_in > IAMFB_MAX_SCENE_CH) return fail(IAMFB_ERR_BAD_ARG, "element %d: n_in %d", e, d->el[e].n_in);

  // elements rendered binaurally through the HRTF front end become 2-channel pass-through elements of the plan built below
  iamfb_plan_desc back = *d;
  iamfb_hrtf_front *hf = nullptr;
  {
    int r = iamfb_hrtf_front_create(d, &back, &hf);
    if (r) return r;
  }
  const iamfb_plan_desc *orig = d;
  d = &back;

  iamfb_plan *p = new iamfb_plan();
  memset(p, 0, sizeof(*p));
  p->ctx = ctx;
  p->desc = *d;
  p->hrtf = hf;
  for (int e = 0; e < d->n_elements; ++e) p->in_rows[e] = orig->el[e].n_in;
  if (hf) {
    bool any = false;
    for (int e = 0; e < d->n_elements; ++e) any |= iamfb_hrtf_needs_demix(hf, e);
    if (any) {
      p->kp_front = new KernelPlan();
      memset(p->kp_front, 0, sizeof(KernelPlan));
      p->kp_front->frame_size = orig->frame_size;
      p->kp_front->n_elements = orig->n_elements;
      p->kp_front->out_channels = k_target_channels[orig->target];
      p->kp_front->overlap = (orig->frame_size / 8) / 2;
      memset(&p->init_state_front, 0, sizeof(p->init_state_front));
      p->init_state_front.lim_j = -1;
      p->init_state_front.lim_start = p->init_state_front.lim_end = -1.f;
      for (int e = 0; e < d->n_elements; ++e) {   // (every element: k_resolve walks them all)
        int r = build_element(*orig, e, *p->kp_front, p->front_tmpl[e], p->init_state_front);
        if (r != IAMFB_OK) { iamfb_plan_destroy(p); return r; }
        p->front_demix[e] = iamfb_hrtf_needs_demix(hf, e);
      }
    }
  }
  KernelPlan &kp = p->kp;
  kp.frame_size = d->frame_size;
  kp.n_elements = d->n_elements;
  kp.out_channels = k_target_channels[d->target];
  kp.overlap = (d->frame_size / 8) / 2;
  kp.resample = d->in_rate != d->out_rate;
  kp.limiter = d->limiter ? 1 : 0;
  kp.hist = kp.limiter ? kLimDelay : 0;
  kp.bit_depth = d->bit_depth;
  kp.loud_gain = d->loudness_gain;
  memset(&p->init_state, 0, sizeof(p->init_state));
  for (int e = 0; e < d->n_elements; ++e) {
    int r = build_element(*d, e, kp, p->tmpl[e], p->init_state);
    if (r != IAMFB_OK) { iamfb_plan_destroy(p); return r; }
  }
  p->init_state.lim_j = -1;
  p->init_state.lim_start = -1.f;
  p->init_state.lim_end = -1.f;
  p->init_state.lim_pad = kp.limiter ? kLimDelay : 0;
  p->init_state.lim_init = 0;

  // recon-gain cross-fade windows (demixer_open demixer.c:502-505 + demixer_set_frame_offset(0) :537-563)
  {
    int wl = d->frame_size / 8, ol = wl / 2;
    std::vector<float> hann(wl > 0 ? wl : 1), sw(ol > 0 ? ol : 1), ew(ol > 0 ? ol : 1);
    for (int i = 0; i < wl; ++i) hann[i] = (0.5 * (1.0 - cos(2.0 * M_PI * (double)i / (double)(wl - 1))));
    for (int j = 0; j < ol; ++j) { sw[j] = hann[j]; ew[j] = hann[j + ol]; }
    int r;
    if ((r = upload(&p->d_start_win, sw.data(), (size_t)ol)) || (r = upload(&p->d_stop_win, ew.data(), (size_t)ol))) { iamfb_plan_destroy(p); return r; }
  }
  {   // qf_to_float(q, 8), fixedp11_5.c:53-55
    float qf[256];
    for (int q = 0; q < 256; ++q) qf[q] = ((float)q / (pow(2.0f, (float)8) - 1.0));
    int r = upload(&p->d_qf, qf, 256);
    if (r) { iamfb_plan_destroy(p); return r; }
  }
  if (kp.resample) {
    std::vector<float> table;
    build_resampler(kp, table, d->in_rate, d->out_rate);
    kp.rs_hist = (int)((kp.rs_filt_len - 1 + 3) & ~3u);
    if (kp.rs_hist < 64) kp.rs_hist = 64;
    if (kp.rs_hist > kMaxRsHist) { iamfb_plan_destroy(p); return fail(IAMFB_ERR_UNIMPLEMENTED, "resampler filter length %u > %d (ratio %d:%d)", kp.rs_filt_len, kMaxRsHist, d->in_rate, d->out_rate); }
    p->sinc_len = (int)table.size();
    int r = upload(&p->d_sinc, table.data(), table.size());
    if (r) { iamfb_plan_destroy(p); return r; }
    if (!kp.rs_direct) {
      const int Nf = (int)kp.rs_filt_len, os = (int)kp.rs_oversample, trow = Nf + 1;
      std::vector<float4> t4((size_t)os * trow);
      for (int o = 0; o < os; ++o)
        for (int j = 0; j < trow; ++j) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < Nf) {
            const float *t = table.data() + 4 + (j + 1) * os - o;
            v = make_float4(t[-2], t[-1], t[0], t[1]);
          }
          t4[(size_t)o * trow + j] = v;
        }
      r = upload(&p->d_tab4, t4.data(), t4.size());
      if (r) { iamfb_plan_destroy(p); return r; }
      // inputs spanned by 128 consecutive outputs: floor(127 * num / den) + 1, plus the filter length
      p->rs_span = (int)((127ull * kp.rs_num) / kp.rs_den + 2 + kp.rs_filt_len + 3) & ~3;
    }
  }
  if (kp.limiter) {
    std::vector<float> acc;
    build_limiter(kp, acc, d->limiter_threshold_db, d->out_rate);   // limiter runs at the requested rate (:3809-3815)
    int r = upload(&p->d_acc, acc.data(), acc.size());
    if (r) { iamfb_plan_destroy(p); return r; }
  }
  p->fused = false;
  {
    const char *env = getenv("IAMFB_FUSED");
    const bool want = !env || atoi(env) != 0;
    const bool dmr0 = kp.el[0].renderer == kRdrDMR;
    bool eligible = want && !kp.resample && (kp.frame_size & 3) == 0 && fused_variant(p->tmpl, kp.n_elements, dmr0) >= 0;
    int nin = 0;
    for (int e = 0; e < kp.n_elements; ++e) {
      nin += kp.el[e].n_in;
      if (kp.el[e].renderer == kRdrDMR && (e > 0 || kp.n_elements > 1)) eligible = false;   // down-mixer inside a two-element mix: multi-kernel path
      if (kp.el[e].n_rec > kp.el[e].n_in) eligible = false;        // reconstructed rows are written back over the staged ones
    }
    if (eligible) {
      const int co = kp.out_channels, H = kp.limiter ? kLimDelay : 0;
      // floats of shared memory per block for a tile of tl samples (layout in k_fused)
      auto smem_floats = [&](int tl) { return (size_t)(nin + 1) * tl + (size_t)(co + 1) * (H + tl) + tl + 2 * ((size_t)tl + kWmPad); };
      // as many blocks (= streams) per SM as possible while a tile still covers >= 240 samples (or the whole frame):
      // 7, 6, 5, 4, 3 blocks of the 227 KB
      int tl = 0, blocks_per_sm = 0;
      for (int blocks = 7; blocks >= 3 && !tl; --blocks) {
        const int budget = (int)((233472 / blocks - 1024 - 64) / 4);
        int tl_max = (budget - 2 * kWmPad - (co + 1) * H) / (nin + 1 + co + 1 + 3);
        if (tl_max > 1024) tl_max = 1024;
        // (frames that are multiples of 8 samples get tiles that are: int16 rows are then staged 16 bytes at a time)
        const int tq = (kp.frame_size & 7) == 0 ? 8 : 4;
        tl_max &= ~(tq - 1);
        if (tl_max < 64) continue;
        const int n_tiles = (kp.frame_size + tl_max - 1) / tl_max;
        int t = (kp.frame_size + n_tiles - 1) / n_tiles;
        t = (t + tq - 1) & ~(tq - 1);
        if (t > tl_max) continue;
        if (t >= 240 || n_tiles == 1 || blocks == 3) { tl = t; blocks_per_sm = blocks; }
      }
      if (tl >= 64) {
        p->fused = true;
        p->fused_tile = tl;
        // few streams per SM (large rings): spread each tile over more, lighter threads (scene-based signatures only)
        p->fused_variant = 0;
        if (kp.n_elements == 1 && kp.el[0].kind == IAMFB_EL_SCENE) {
          p->fused_variant = blocks_per_sm <= 5 ? 1 : 0;   // measured on C3 (3 streams per SM): 2.93 ms vs 3.45 with 4 samples x 64 threads
        }
        p->fused_smem = sizeof(float) * smem_floats(tl);
        // k_stream: one channel-based element through a channel->channel matrix that matches the compile-time table,
        // limiter on, 16-bit output, frames that are whole limiter windows
        p->stream = false;
        {
          const char *senv = getenv("IAMFB_STREAM");
          const bool swant = !senv || atoi(senv) != 0;
          const ElPlan &ep = kp.el[0];
          if (swant && kp.n_elements == 1 && ep.kind == IAMFB_EL_CHANNEL && ep.renderer == kRdrM2M && kp.limiter && kp.bit_depth == 16 &&
              kp.frame_size % kStreamTile == 0 && stream_sig_exists(ep.layout, d->target)) {
            const int idx = m2m_find(ep.layout, d->target);
            bool same = idx >= 0 && k_m2m_index[idx].m == ep.n_rec && k_m2m_index[idx].n == co && ep.n_mat_out == co;
            for (int oc = 0; same && oc < co; ++oc) {
              if (ep.out_slot[oc] != oc) same = false;
              for (int m = 0; same && m < ep.n_rec; ++m) {
                uint32_t bits;
                memcpy(&bits, &ep.mat[oc * ep.n_rec + m], 4);
                if (bits != k_matrix_pool[k_m2m_index[idx].off + m * co + oc]) same = false;
              }
            }
            // an output gain on a channel of the reconstructed layout itself (not on the layers it is derived from):
            // rare, left to k_fused
            for (int m = 0; same && m < ep.n_rec; ++m)
              if ((ep.gain_mask >> ep.rec_ch[m]) & 1u) same = false;
            if (same) {
              for (int c = 0; c < kChCount; ++c) kp.el[0].f_gain[c] = ((ep.gain_mask >> c) & 1u) ? ep.gain[c] : 1.0f;
              for (int c = 0; c < kChCount; ++c)
                kp.el[0].s_row_off[c] = ep.src_row[c] >= 0 ? ep.src_row[c] * kStreamTile * 4 : -1;
              p->stream = true;
              p->stream_sig = ep.layout * 16 + d->target;
              p->stream_smem = sizeof(float) * ((size_t)(ep.n_in + 2 * co + 6) * kStreamTile);
            }
          }
        }
        fill_fused_offsets(kp, tl, nin);
        p->s16_native = (kp.frame_size & 7) == 0;
        // k_pipe: every element either channel-based through a channel->channel matrix or scene-based with a mono
        // channel mapping, the compile-time tables equal to the plan's matrices, frames that are whole limiter windows
        p->pipe = false;
        {
          const char *penv = getenv("IAMFB_PIPE");
          const bool pwant = !penv || atoi(penv) != 0;
          int key[2][2] = {{0, 0}, {0, 0}};
          bool ok = pwant && kp.frame_size % kStreamTile == 0 && (kp.frame_size & 7) == 0;
          for (int e = 0; ok && e < kp.n_elements; ++e) {
            const ElPlan &ep = kp.el[e];
            if (ep.kind == IAMFB_EL_CHANNEL) {
              if (ep.renderer != kRdrM2M) ok = false;
              for (int m = 0; ok && m < ep.n_rec; ++m)
                if ((ep.gain_mask >> ep.rec_ch[m]) & 1u) ok = false;     // output gain on a channel of the layout itself: k_fused
              key[e][0] = ep.layout; key[e][1] = ep.n_rec;
              const int idx = m2m_find(ep.layout, d->target);
              if (idx < 0 || k_m2m_index[idx].m != ep.n_rec || k_m2m_index[idx].n != co || ep.n_mat_out != co) ok = false;
              for (int oc = 0; ok && oc < co; ++oc) {
                if (ep.out_slot[oc] != oc) ok = false;
                for (int m = 0; ok && m < ep.n_rec; ++m) {
                  uint32_t bits;
                  memcpy(&bits, &ep.mat[oc * ep.n_rec + m], 4);
                  if (bits != k_matrix_pool[k_m2m_index[idx].off + m * co + oc]) ok = false;
                }
              }
            } else {
              if (ep.ambi_mode != 0) ok = false;                         // projection matrices are run-time data: k_fused
              key[e][0] = -1; key[e][1] = ep.n_rec;
              const int order = ep.n_rec == 1 ? 0 : ep.n_rec == 4 ? 1 : ep.n_rec == 9 ? 2 : 3;
              int idx = -1;
              for (size_t i = 0; i < sizeof(k_h2m_index) / sizeof(k_h2m_index[0]); ++i)
                if (k_h2m_index[i].order == order && k_h2m_index[i].out == d->target) idx = (int)i;
              if (idx < 0 || k_h2m_index[idx].m != ep.n_rec || k_h2m_index[idx].n != ep.n_mat_out) ok = false;
              for (int oc = 0; ok && oc < co; ++oc) {
                const int n = ep.out_slot[oc];
                if (n >= ep.n_mat_out) ok = false;
                for (int m = 0; ok && n >= 0 && m < ep.n_rec; ++m) {
                  uint32_t bits;
                  memcpy(&bits, &ep.mat[n * ep.n_rec + m], 4);
                  if (bits != k_matrix_pool[k_h2m_index[idx].off + n * ep.n_rec + m]) ok = false;
                }
              }
            }
          }
          const PipeSigInfo *si = ok ? iamfb_pipe_find(key[0][0], key[0][1], kp.n_elements > 1 ? key[1][0] : 0, kp.n_elements > 1 ? key[1][1] : 0, d->target, d->arithmetic == IAMFB_ARITH_FMA) : nullptr;
          if (si) {
            // the second element's rows follow the first's inside a stage: tensor copies land on 128-byte boundaries
            const int n0 = kp.el[0].n_in;
            // (behind the HRTF front end the rows are always float32)
            if (kp.n_elements > 1 && (((n0 * kStreamTile * (p->hrtf ? 4 : 2)) & 127) != 0)) si = nullptr;
          }
          if (si) {
            p->pipe = true;
            p->pipe_sig = si->id;
            // where both exist k_stream keeps the float32 submits (its single float32 stage is the leaner loop: C2 0.231 vs
            // 0.241 ms per submit) and k_pipe takes the int16 ones (two int16 stages instead of a widening pass)
          }
        }
      }
    }
  }
  // k_pipe_rs: one channel-based element through its channel->channel matrix, the interpolating resampler (the 44.1 -> 48 kHz
  // kind: not the integer-ratio "direct" form), frames of a multiple of 8 samples that hold at least one limiter window
  p->rs_pipe = false;
  {
    const char *penv = getenv("IAMFB_PIPE");
    const bool pwant = !penv || atoi(penv) != 0;
    const ElPlan &ep = kp.el[0];
    const int co = kp.out_channels;
    bool ok = pwant && kp.resample && !kp.rs_direct && kp.n_elements == 1 && ep.kind == IAMFB_EL_CHANNEL && ep.renderer == kRdrM2M &&
              (kp.frame_size & 7) == 0 && kp.frame_size >= kStreamTile && kp.rs_filt_len <= 96 && kp.rs_hist >= (int)kp.rs_filt_len - 1 &&
              (kp.rs_hist & 3) == 0 && p->d_tab4;
    for (int m = 0; ok && m < ep.n_rec; ++m)
      if ((ep.gain_mask >> ep.rec_ch[m]) & 1u) ok = false;
    if (ok) {
      const int idx = m2m_find(ep.layout, d->target);
      if (idx < 0 || k_m2m_index[idx].m != ep.n_rec || k_m2m_index[idx].n != co || ep.n_mat_out != co) ok = false;
      for (int oc = 0; ok && oc < co; ++oc) {
        if (ep.out_slot[oc] != oc) ok = false;
        for (int m = 0; ok && m < ep.n_rec; ++m) {
          uint32_t bits;
          memcpy(&bits, &ep.mat[oc * ep.n_rec + m], 4);
          if (bits != k_matrix_pool[k_m2m_index[idx].off + m * co + oc]) ok = false;
        }
      }
    }
    int sig = -1;
    const char *alt = getenv("IAMFB_PIPE_ALT");   // experiment: another thread shape of the signature
    const int want_alt = alt ? atoi(alt) : 0;
#define X(id, L0, N0, T, NW, VEC, MINB) if (ok && ep.layout == L0 && ep.n_rec == N0 && d->target == T && (sig < 0 || id == want_alt)) sig = id;
    IAMFB_PIPE_RS_SIGS(X)
#undef X
    if (sig >= 0) {
      // ring of pre-resample samples: the inputs 240 outputs span + the filter + one input tile rendered ahead
      const int span = (int)((239ull * kp.rs_num) / kp.rs_den) + 2 + (int)kp.rs_filt_len + kStreamTile;
      p->rs_ring = (span + 63) & ~63;
      p->rs_mirror = ((int)kp.rs_filt_len + 3) & ~3;
      // padded tap rows: the FIR walks the inputs of a thread's four outputs once, every output meeting zero taps outside its window
      {
        const int Nf = (int)kp.rs_filt_len, os = (int)kp.rs_oversample;
        p->rs_tab_pad = 4 * (kp.rs_int_adv + 1);
        p->rs_tab_row = (Nf + 2 * p->rs_tab_pad) | 1;
        std::vector<float4> t4((size_t)os * p->rs_tab_row, make_float4(0.f, 0.f, 0.f, 0.f)), src((size_t)os * (Nf + 1));
        int r = cudaMemcpy(src.data(), p->d_tab4, src.size() * sizeof(float4), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 1;
        for (int o = 0; o < os && !r; ++o)
          for (int j = 0; j < Nf; ++j) t4[(size_t)o * p->rs_tab_row + p->rs_tab_pad + j] = src[(size_t)o * (Nf + 1) + j];
        if (r || upload(&p->d_tab4p, t4.data(), t4.size())) { iamfb_plan_destroy(p); return fail(IAMFB_ERR_CUDA, "resampler tap table upload failed"); }
        p->rs_mirror = (Nf + 3 * (kp.rs_int_adv + 1) + 3) & ~3;
        // cubic_coef (resample.c:246-256) per phase: frac = ((phi * oversample) % den) / den in float, interp[2] in double
        std::vector<float4> ci(kp.rs_den);
        for (unsigned phi = 0; phi < kp.rs_den; ++phi) {
          volatile float frac = ((float)((phi * kp.rs_oversample) % kp.rs_den)) / kp.rs_den;
          volatile float t0 = -0.16667f * frac, t1 = 0.16667f * frac, t2 = t1 * frac, t3 = t2 * frac;
          volatile float i0 = t0 + t3;
          volatile float h1 = 0.5f * frac, h2 = h1 * frac, h3 = h2 * frac;
          volatile float i1a = frac + h2;
          volatile float i1 = i1a - h3;
          volatile float u0 = -0.33333f * frac, u1 = u0 + h2;
          volatile float v1 = 0.16667f * frac, v2 = v1 * frac, v3 = v2 * frac;
          volatile float i3 = u1 - v3;
          const float i2 = (float)(1. - (double)i0 - (double)i1 - (double)i3);
          ci[phi] = make_float4(i0, i1, i2, i3);
        }
        if (upload(&p->d_interp4, ci.data(), ci.size())) { iamfb_plan_destroy(p); return fail(IAMFB_ERR_CUDA, "resampler weight table upload failed"); }
      }
      p->rs_pipe = true;
      p->rs_pipe_sig = sig;
      {
        // the split form is the default; IAMFB_RS_SPLIT=0 forces the single kernel k_pipe_rs (test hook)
        const char *sp = getenv("IAMFB_RS_SPLIT");
        p->rs_split = !sp || atoi(sp) != 0;
        // outputs per work item of k_resample_ls: the largest the staging areas allow (<= 36); a launch picks the size
        // at or below it that leaves the least idle work in the last round of its persistent warps.  IAMFB_LS_CHUNK fixes it.
        const char *ck = getenv("IAMFB_LS_CHUNK");
        p->rs_ls_chunk = ck ? (atoi(ck) & ~3) : 36;
        p->rs_ls_chunk_fixed = ck != nullptr;
        if (p->rs_ls_chunk < 8 || p->rs_ls_chunk > 256) p->rs_ls_chunk = 36;
        // the staging areas of a block's warps must fit the SM: smaller work items for long filters / down-sampling ratios
        while (p->rs_ls_chunk > 8 && rs_ls_smem(p, p->rs_ls_chunk, nullptr) > kLsSmemMax) p->rs_ls_chunk -= 4;
        if (rs_ls_smem(p, p->rs_ls_chunk, nullptr) > kLsSmemMax) p->rs_split = false;
      }
      for (int c = 0; c < kChCount; ++c) kp.el[0].f_gain[c] = ((ep.gain_mask >> c) & 1u) ? ep.gain[c] : 1.0f;
    }
  }
  *out = p;
  return IAMFB_OK;
}

// ---- k_pipe dispatch (the instantiations live in iamfb_pipe_g<N>.cu)
static const PipeSigInfo k_pipe_sigs[] = {
#define X(id, L0, N0, L1, N1, T, NW, VEC, MINB) {id, L0, N0, L1, N1, T, NW, VEC, 0},
    IAMFB_PIPE_SIGS(X)
#undef X
#define X(id, L0, N0, L1, N1, T, NW, VEC, MINB) {id, L0, N0, L1, N1, T, NW, VEC, 1},
    IAMFB_PIPE_SIGS_FMA(X)
#undef X
};
const PipeSigInfo *iamfb_pipe_find(int l0, int n0, int l1, int n1, int target, bool fma) {
  if (fma)
    for (const PipeSigInfo &si : k_pipe_sigs)
      if (si.fma && si.l0 == l0 && si.n0 == n0 && si.n1 == n1 && (n1 == 0 || si.l1 == l1) && si.target == target) return &si;
  const char *alt = getenv("IAMFB_PIPE_ALT");   // experiment: the alternative thread shape of a signature (ids >= 13)
  if (alt && atoi(alt))
    for (const PipeSigInfo &si : k_pipe_sigs)
      if (si.id >= 13 && !si.fma && si.l0 == l0 && si.n0 == n0 && si.n1 == n1 && (n1 == 0 || si.l1 == l1) && si.target == target) return &si;
  for (const PipeSigInfo &si : k_pipe_sigs)
    if (!si.fma && si.l0 == l0 && si.n0 == n0 && si.n1 == n1 && (n1 == 0 || si.l1 == l1) && si.target == target) return &si;
  return nullptr;
}
#define G(n) int iamfb_pipe_launch_g##n(iamfb_ctx *, int, bool, const KernelPlan &, const PipeArgs &, int, size_t, const CUtensorMap &, const CUtensorMap &);
G(0) G(1) G(2) G(3) G(4) G(5) G(6) G(7)
#undef G
#define G(n) int iamfb_pipe_launch_r##n(iamfb_ctx *, int, bool, const KernelPlan &, const PipeArgs &, int, size_t, const CUtensorMap &, const CUtensorMap &);
G(0) G(1) G(2) G(3) G(4) G(5) G(6) G(7)
#undef G
int iamfb_pipe_launch(iamfb_ctx *ctx, int sig_id, bool s16, const KernelPlan &kp, const PipeArgs &pa, int S, size_t smem, const CUtensorMap &m0,
                      const CUtensorMap &m1) {
  if (pa.gain_ramp[0] || pa.gain_ramp[kMaxEl - 1] || pa.out_gain_ramp) {   // animated mix gains: the k_pipe<SIG, true> instantiations
    if (sig_id >= kPipeFmaFirstId) return iamfb_pipe_launch_r7(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
    switch (IAMFB_PIPE_GROUP_OF(sig_id)) {
      case 0: return iamfb_pipe_launch_r0(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
      case 1: return iamfb_pipe_launch_r1(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
      case 2: return iamfb_pipe_launch_r2(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
      case 3: return iamfb_pipe_launch_r3(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
      case 4: return iamfb_pipe_launch_r4(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
      case 5: return iamfb_pipe_launch_r5(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
      default: return iamfb_pipe_launch_r6(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
    }
  }
  if (sig_id >= kPipeFmaFirstId) return iamfb_pipe_launch_g7(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);   // IAMFB_ARITH_FMA variants
  switch (IAMFB_PIPE_GROUP_OF(sig_id)) {
    case 0: return iamfb_pipe_launch_g0(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
    case 1: return iamfb_pipe_launch_g1(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
    case 2: return iamfb_pipe_launch_g2(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
    case 3: return iamfb_pipe_launch_g3(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
    case 4: return iamfb_pipe_launch_g4(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
    case 5: return iamfb_pipe_launch_g5(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
    default: return iamfb_pipe_launch_g6(ctx, sig_id, s16, kp, pa, S, smem, m0, m1);
  }
}

static int make_input_tmap(CUtensorMap *tm, const void *base, size_t rows, int N, int n_in, bool s16);
static int launch_pipe(iamfb_ctx *ctx, const iamfb_plan *p, const FusedArgs &fa, int S, bool s16) {
  const KernelPlan &kp = p->kp;
  const PipeSigInfo *si = nullptr;
  for (const PipeSigInfo &k : k_pipe_sigs)
    if (k.id == p->pipe_sig) si = &k;
  if (!si) return fail(IAMFB_ERR_INTERNAL, "no k_pipe signature %d", p->pipe_sig);
  PipeArgs pa;
  memset(&pa, 0, sizeof(pa));
  int nin = 0;
  CUtensorMap tm[2];
  for (int e = 0; e < kp.n_elements; ++e) {
    pa.in[e] = fa.in[e];
    nin += kp.el[e].n_in;
    int r = make_input_tmap(&tm[e], fa.in[e], (size_t)S * fa.n_frames * kp.el[e].n_in, kp.frame_size, kp.el[e].n_in, s16);
    if (r) return r;
  }
  if (kp.n_elements == 1) tm[1] = tm[0];
  pa.frames = fa.frames; pa.start_win = fa.start_win; pa.stop_win = fa.stop_win; pa.submit = fa.submit; pa.state = fa.state;
  pa.acc = fa.acc; pa.hist_y = fa.hist_y; pa.hist_pk = fa.hist_pk; pa.pcm = fa.pcm; pa.stride_bytes = fa.stride_bytes;
  pa.n_frames = fa.n_frames;
  for (int e = 0; e < kp.n_elements; ++e) pa.gain_ramp[e] = fa.gain_ramp[e];
  pa.out_gain_ramp = fa.out_gain_ramp;
  pa.row_bytes = kStreamTile * (s16 ? 2 : 4);
  pa.stage_bytes = (nin * pa.row_bytes + 127) & ~127;
  // active output channels of the signature (rows of the time line): every channel some matrix row of an element lands on
  int ny = 0;
  for (int oc = 0; oc < kp.out_channels; ++oc) {
    bool any = false;
    for (int e = 0; e < kp.n_elements; ++e) {
      const ElPlan &ep = kp.el[e];
      const int n = ep.out_slot[oc];
      for (int m = 0; n >= 0 && m < ep.n_rec; ++m)
        if (ep.mat[n * ep.n_rec + m] != 0.f) any = true;
    }
    ny += any ? 1 : 0;
  }
  const size_t smem = (((size_t)(ny * 2 + 6) * kStreamTile * 4 + 127) & ~(size_t)127) + (size_t)(s16 ? 2 : 1) * pa.stage_bytes;
  const unsigned tpf = (unsigned)(kp.frame_size / kStreamTile);
  pa.neg_zero = -0.0f;
  pa.tpf_magic = tpf <= 1 ? 0u : (unsigned)((0x100000000ull + tpf - 1) / tpf);
  if ((long long)fa.n_frames * tpf >= 65536) return fail(IAMFB_ERR_BAD_ARG, "submit of %d frames is too long for k_pipe", fa.n_frames);
  // staged-row byte offsets for this launch's input format (by IAChannel id; by ambisonics channel for a scene-based element)
  static thread_local KernelPlan kpl;
  kpl = kp;
  for (int e = 0; e < kp.n_elements; ++e) {
    ElPlan &ep = kpl.el[e];
    if (ep.kind == IAMFB_EL_CHANNEL) {
      for (int c = 0; c < kChCount; ++c) ep.s_row_off[c] = ep.src_row[c] >= 0 ? ep.src_row[c] * pa.row_bytes : -1;
    } else {
      for (int m = 0; m < ep.n_rec && m < kChCount; ++m) ep.s_row_off[m] = (int)ep.ambi_map[m] * pa.row_bytes;
    }
  }
  int r = iamfb_pipe_launch(ctx, si->id, s16, kpl, pa, S, smem, tm[0], tm[1]);
  if (r) return r;
  cudaError_t e_ = cudaGetLastError();
  if (e_ != cudaSuccess) return fail(IAMFB_ERR_CUDA, "launch of k_pipe failed: %s", cudaGetErrorString(e_));
  ++ctx->launches;
  return IAMFB_OK;
}

// k_resample_ls: inputs staged per stream for a work item of `chunk` outputs (odd: lanes on different banks) and the
// dynamic shared memory of a block (tap table + cubic weights + one staging area per warp)
static int rs_ls_smem(const iamfb_plan *p, int chunk, int *span_out) {
  const KernelPlan &kp = p->kp;
  const int steps = (int)kp.rs_filt_len + 3 * ((int)kp.rs_int_adv + 1);
  // (a team of warps shares one staging area: the inputs of kLsTeam x chunk outputs)
  const int span = ((int)(((unsigned long long)(kLsTeam * chunk) * kp.rs_num) / kp.rs_den) + 3 + steps) | 1;
  if (span_out) *span_out = span;
  if (kp.rs_den > 2048) return 1 << 30;       // (the per-phase weights would not fit next to the staging areas)
  return (int)((kp.rs_oversample * p->rs_tab_row + kp.rs_den + (kp.rs_den + 3) / 4) * sizeof(float4)) + (kLsWarps / kLsTeam) * 32 * span * 8;
}

// pre: the limiter half only (k_pipe_rs<PRE>) behind k_resample_ls, else the whole pipeline in k_pipe_rs
static int launch_pipe_rs(iamfb_ctx *ctx, const iamfb_plan *p, iamfb_batch *b, const iamfb_io *io, int F, bool s16, bool pre = false) {
  const KernelPlan &kp = p->kp;
  const int co = kp.out_channels;
  PipeRsArgs ra;
  memset(&ra, 0, sizeof(ra));
  PipeArgs &pa = ra.p;
  pa.in[0] = io->in[0];
  if (!pre && (((size_t)io->in[0]) & 15) != 0) return fail(IAMFB_ERR_BAD_ARG, "decoded input must be 16-byte aligned");
  pa.frames = b->d_frames; pa.start_win = p->d_start_win; pa.stop_win = p->d_stop_win; pa.submit = b->d_submit; pa.state = b->d_state;
  pa.acc = p->d_acc;
  pa.hist_y = b->d_tl_b; pa.hist_pk = b->d_pk;
  pa.pcm = io->pcm; pa.stride_bytes = iamfb_plan_out_stride_bytes(p, F);
  pa.n_frames = F;
  pa.row_bytes = kStreamTile * (s16 ? 2 : 4);
  pa.stage_bytes = (kp.el[0].n_in * pa.row_bytes + 127) & ~127;
  pa.neg_zero = -0.0f;
  ra.hist_y_stride = b->cap_b; ra.hist_pk_stride = b->cap_b;
  ra.hist_rs = b->d_tl_a; ra.hist_rs_stride = b->cap_a;
  ra.tab4 = p->d_tab4p; ra.tab_row = p->rs_tab_row; ra.tab_pad = p->rs_tab_pad;
  ra.interp4 = p->d_interp4;
  ra.ring = p->rs_ring; ra.mirror = p->rs_mirror;
  int ny = 0;
  for (int oc = 0; oc < co; ++oc) {
    bool any = false;
    const ElPlan &ep = kp.el[0];
    const int n = ep.out_slot[oc];
    for (int m = 0; n >= 0 && m < ep.n_rec; ++m)
      if (ep.mat[n * ep.n_rec + m] != 0.f) any = true;
    ny += any ? 1 : 0;
  }
  const int np = (ny + 1) / 2;
  int off = (ny * 2 + 6) * kStreamTile * 4;
  ra.off_pkr = off; off += 2 * kStreamTile * 4;
  if (!pre) {
    ra.off_ring = off; off += np * (ra.ring + ra.mirror) * 8;
    off = (off + 15) & ~15;
    ra.off_tab = off; off += (int)(kp.rs_oversample * p->rs_tab_row * sizeof(float4));
    off = (off + 127) & ~127;
    ra.off_stage = off; off += 2 * pa.stage_bytes;
  } else {
    ra.off_ring = ra.off_tab = ra.off_stage = off;    // (not used by the limiter half)
    ra.tl_pre = b->d_tl_b; ra.tl_pre_stride = b->cap_b; ra.tl_pre_off = kp.hist;
  }
  ra.smem_bytes = off;
  static thread_local KernelPlan kpl;
  kpl = kp;
  for (int c = 0; c < kChCount; ++c) kpl.el[0].s_row_off[c] = kp.el[0].src_row[c] >= 0 ? kp.el[0].src_row[c] * pa.row_bytes : -1;
  int r = pre ? iamfb_pipe_rs_lim_launch(ctx, p->rs_pipe_sig, kpl, ra, b->S) : iamfb_pipe_rs_launch(ctx, p->rs_pipe_sig, s16, kpl, ra, b->S);
  if (r) return r;
  cudaError_t e_ = cudaGetLastError();
  if (e_ != cudaSuccess) return fail(IAMFB_ERR_CUDA, "launch of k_pipe_rs failed: %s", cudaGetErrorString(e_));
  ++ctx->launches;
  return IAMFB_OK;
}

// split form of a resampling pipeline, regular streams: k_pipe_prerender -> tl_a, k_resample_ls (one stream per lane) -> tl_b,
// k_pipe_rs<PRE> (limiter, PCM; it also carries the resampler history to the head of tl_a).  The irregular streams of the
// submit follow on the multi-kernel path (gated), exactly as behind k_pipe_rs.
static int launch_rs_split(iamfb_ctx *ctx, const iamfb_plan *p, iamfb_batch *b, const iamfb_io *io, int F, bool s16) {
  const KernelPlan &kp = p->kp;
  const int co = kp.out_channels, S = b->S, N = kp.frame_size;
  {
    PreRenderArgs ra;
    memset(&ra, 0, sizeof(ra));
    PipeArgs &pa = ra.p;
    pa.in[0] = io->in[0];
    if ((((size_t)io->in[0]) & 15) != 0) return fail(IAMFB_ERR_BAD_ARG, "decoded input must be 16-byte aligned");
    pa.frames = b->d_frames; pa.start_win = p->d_start_win; pa.stop_win = p->d_stop_win; pa.submit = b->d_submit; pa.state = b->d_state;
    pa.n_frames = F;
    pa.neg_zero = -0.0f;
    ra.tl = b->d_tl_a; ra.cap = b->cap_a; ra.hist = kp.rs_hist;
    static thread_local KernelPlan kpl;
    kpl = kp;
    const int rowb = N * (s16 ? 2 : 4);      // rows are read where the caller put them: a frame's rows are N samples apart
    for (int c = 0; c < kChCount; ++c) kpl.el[0].s_row_off[c] = kp.el[0].src_row[c] >= 0 ? kp.el[0].src_row[c] * rowb : -1;
    int r = iamfb_pipe_prerender_launch(ctx, p->rs_pipe_sig, s16, kpl, ra, S, F);
    if (r) return r;
    cudaError_t e_ = cudaGetLastError();
    if (e_ != cudaSuccess) return fail(IAMFB_ERR_CUDA, "launch of k_pipe_prerender failed: %s", cudaGetErrorString(e_));
    ++ctx->launches;
  }
  {
    ResampleLsArgs la;
    memset(&la, 0, sizeof(la));
    la.src = b->d_tl_a; la.dst = b->d_tl_b; la.submit = b->d_submit; la.state = b->d_state;
    la.tab4 = p->d_tab4p; la.interp4 = p->d_interp4; la.tab_row = p->rs_tab_row; la.tab_pad = p->rs_tab_pad;
    la.cap_a = b->cap_a; la.cap_b = b->cap_b; la.hist_b = kp.hist; la.n_streams = S; la.co = co;
    const int groups = (S + 31) / 32;
    const int max_out = iamfb_plan_max_out_samples(p, F);
    la.chunk = p->rs_ls_chunk;
    if (!p->rs_ls_chunk_fixed) {
      // every round of the persistent warps lasts one chunk: the size with the fewest (rounds x chunk) wins, ties to the larger
      long long best = -1;
      for (int c = p->rs_ls_chunk; c >= 16 && c >= p->rs_ls_chunk - 16; c -= 4) {
        const long long items = (long long)groups * ((max_out + kLsTeam * c - 1) / (kLsTeam * c)), teams = (long long)ctx->n_sm * (kLsWarps / kLsTeam);
        const long long cost = ((items + teams - 1) / teams) * c;
        if (best < 0 || cost < best) { best = cost; la.chunk = c; }
      }
    }
    la.n_chunks = (max_out + kLsTeam * la.chunk - 1) / (kLsTeam * la.chunk);      // work items: kLsTeam x chunk outputs of 32 streams for a team of warps
    la.neg_zero = -0.0f;
    const int smem_ls = rs_ls_smem(p, la.chunk, &la.span);
    int blocks = ctx->n_sm;      // persistent: one block of kLsWarps warps per SM
    const long long items = (long long)groups * la.n_chunks;
    if ((long long)blocks * (kLsWarps / kLsTeam) > items) blocks = (int)((items + kLsWarps / kLsTeam - 1) / (kLsWarps / kLsTeam));
    int r = iamfb_resample_ls_launch(ctx, kp, la, blocks, smem_ls);
    if (r) return r;
    cudaError_t e_ = cudaGetLastError();
    if (e_ != cudaSuccess) return fail(IAMFB_ERR_CUDA, "launch of k_resample_ls failed: %s", cudaGetErrorString(e_));
    ++ctx->launches;
  }
  return launch_pipe_rs(ctx, p, b, io, F, false, true);
}

extern "C" void iamfb_plan_destroy(iamfb_plan *p) {
  if (!p) return;
  cudaFree(p->d_start_win); cudaFree(p->d_stop_win); cudaFree(p->d_qf); cudaFree(p->d_sinc); cudaFree(p->d_acc); cudaFree(p->d_tab4); cudaFree(p->d_tab4p); cudaFree(p->d_interp4);
  iamfb_hrtf_front_destroy(p->hrtf);
  delete p->kp_front;
  delete p;
}

// every float of [2^-60, 2^60]: bit patterns 0x21800000 .. 0x5d800000
__global__ void __launch_bounds__(256) k_selftest_quotient(float thr, unsigned long long *mismatches) {
  const uint32_t lo = 0x21800000u, hi = 0x5d800000u;
  unsigned long long bad = 0;
  for (uint64_t b = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= hi; b += (uint64_t)gridDim.x * blockDim.x) {
    const float w = __uint_as_float((uint32_t)b);
    bool ok;
    const float q = stream_quot(thr, w, ok);
    const float r = thr / w;
    if (!ok || __float_as_uint(q) != __float_as_uint(r)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}
extern "C" int iamfb_selftest_quotient(iamfb_ctx *ctx, float thr, uint64_t *mismatches) {
  if (!ctx || !mismatches) return fail(IAMFB_ERR_BAD_ARG, "selftest_quotient: null argument");
  if (!(thr >= kQuotThrLo && thr <= kQuotThrHi)) return fail(IAMFB_ERR_BAD_ARG, "selftest_quotient: thr %g outside the served range", thr);
  CU(cudaSetDevice(ctx->device));
  unsigned long long *d = nullptr;
  CU(cudaMalloc(&d, sizeof(*d)));
  CU(cudaMemsetAsync(d, 0, sizeof(*d), ctx->stream));
  k_selftest_quotient<<<148 * 8, 256, 0, ctx->stream>>>(thr, d);
  unsigned long long h = 0;
  cudaError_t e = cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(IAMFB_ERR_CUDA, "selftest_quotient: %s", cudaGetErrorString(e));
  ++ctx->launches;
  *mismatches = h;
  return IAMFB_OK;
}

extern "C" int iamfb_plan_out_channels(const iamfb_plan *p) { return p ? p->kp.out_channels : 0; }

extern "C" int iamfb_plan_kernel_path_fmt(const iamfb_plan *p, int in_format) {
  if (p && p->rs_pipe) return IAMFB_PATH_PIPE;
  if (!p || !p->fused) return IAMFB_PATH_MULTI;
  const bool s16 = in_format == IAMFB_IN_S16 && !p->hrtf;
  if (p->stream && (!s16 || !p->pipe)) return IAMFB_PATH_STREAM;
  return p->pipe ? IAMFB_PATH_PIPE : IAMFB_PATH_FUSED;
}
extern "C" int iamfb_plan_kernel_path(const iamfb_plan *p) { return iamfb_plan_kernel_path_fmt(p, IAMFB_IN_F32); }
extern "C" int iamfb_plan_arithmetic(const iamfb_plan *p) {
  if (!p) return -1;
  return (p->pipe && p->pipe_sig >= kPipeFmaFirstId) ? IAMFB_ARITH_FMA : IAMFB_ARITH_EXACT;
}

extern "C" int iamfb_plan_max_out_samples(const iamfb_plan *p, int n_frames) {
  if (!p) return 0;
  long long in = (long long)n_frames * p->kp.frame_size;
  long long out = in;
  if (p->kp.resample) out = (in * p->kp.rs_den + p->kp.rs_num - 1) / p->kp.rs_num + 2;
  // a flush can add the limiter delay plus the resampler's output latency
  long long flush = kLimDelay + 64 + (p->kp.resample ? (long long)p->kp.rs_filt_len * p->kp.rs_den / p->kp.rs_num : 0);
  return (int)(out > flush ? out : flush);
}

extern "C" size_t iamfb_plan_out_stride_bytes(const iamfb_plan *p, int n_frames) {
  if (!p) return 0;
  size_t bps = p->kp.bit_depth ? p->kp.bit_depth / 8 : 4;
  size_t b = (size_t)iamfb_plan_max_out_samples(p, n_frames) * p->kp.out_channels * bps;
  return (b + 15) & ~(size_t)15;
}

// ---------------------------------------------------------------------------------------------------------------------
// batch
// ---------------------------------------------------------------------------------------------------------------------
static int round4(int x) { return (x + 3) & ~3; }

extern "C" int iamfb_batch_reset(iamfb_batch *b) {
  if (!b) return fail(IAMFB_ERR_BAD_ARG, "null batch");
  iamfb_plan *p = b->plan;
  CU(cudaSetDevice(p->ctx->device));
  std::vector<StreamState> init(b->S, p->init_state);
  CU(cudaMemcpyAsync(b->d_state, init.data(), sizeof(StreamState) * b->S, cudaMemcpyHostToDevice, p->ctx->stream));
  if (b->d_gate) CU(cudaMemsetAsync(b->d_gate, 0, 2 * sizeof(int), p->ctx->stream));
  const int co = p->kp.out_channels;
  if (b->d_tl_a) CU(cudaMemsetAsync(b->d_tl_a, 0, sizeof(float) * (size_t)b->S * co * b->cap_a, p->ctx->stream));
  if (b->d_tl_b) CU(cudaMemsetAsync(b->d_tl_b, 0, sizeof(float) * (size_t)b->S * co * b->cap_b, p->ctx->stream));
  if (b->d_hist_y) {
    CU(cudaMemsetAsync(b->d_hist_y, 0, sizeof(float) * (size_t)b->S * co * kLimDelay, p->ctx->stream));
    CU(cudaMemsetAsync(b->d_hist_pk, 0, sizeof(float) * (size_t)b->S * kLimDelay, p->ctx->stream));
  }
  if (b->d_pk) {
    CU(cudaMemsetAsync(b->d_pk, 0, sizeof(float) * (size_t)b->S * b->cap_b, p->ctx->stream));
    CU(cudaMemsetAsync(b->d_wm, 0, sizeof(float) * (size_t)b->S * b->cap_b, p->ctx->stream));
    CU(cudaMemsetAsync(b->d_gn, 0, sizeof(float) * (size_t)b->S * b->cap_b, p->ctx->stream));
  }
  if (b->hrtf) {
    int r = iamfb_hrtf_batch_reset(p->hrtf, b->hrtf, p->ctx->stream);
    if (r) return r;
  }
  std::vector<StreamState> init_f;
  if (b->d_state_f) {
    init_f.assign(b->S, p->init_state_front);
    CU(cudaMemcpyAsync(b->d_state_f, init_f.data(), sizeof(StreamState) * b->S, cudaMemcpyHostToDevice, p->ctx->stream));
  }
  CU(cudaStreamSynchronize(p->ctx->stream));
  return IAMFB_OK;
}

extern "C" int iamfb_batch_create(iamfb_plan *p, int n_streams, int max_frames, iamfb_batch **out) {
  if (!p || !out || n_streams <= 0 || max_frames <= 0) return fail(IAMFB_ERR_BAD_ARG, "batch_create: bad argument");
  CU(cudaSetDevice(p->ctx->device));
  iamfb_batch *b = new iamfb_batch();
  memset(b, 0, sizeof(*b));
  b->plan = p;
  b->S = n_streams;
  b->Fmax = max_frames;
  const KernelPlan &kp = p->kp;
  const int co = kp.out_channels;
  b->cap_a = kp.resample ? round4(kp.rs_hist + max_frames * kp.frame_size + (int)kp.rs_filt_len) : 0;
  b->cap_b = round4(kp.hist + iamfb_plan_max_out_samples(p, max_frames) + 64);
  b->out_stride = iamfb_plan_out_stride_bytes(p, max_frames);
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void **ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes ? bytes : 16); };
  alloc((void **)&b->d_state, sizeof(StreamState) * n_streams);
  alloc((void **)&b->d_frames, sizeof(FrameRec) * (size_t)n_streams * max_frames);
  alloc((void **)&b->d_submit, sizeof(SubmitRec) * n_streams);
  if (p->rs_pipe) alloc((void **)&b->d_submit_mk, sizeof(SubmitRec) * n_streams);
  if (p->rs_pipe) alloc((void **)&b->d_gate, 2 * sizeof(int));
  if (kp.resample) alloc((void **)&b->d_tl_a, sizeof(float) * (size_t)n_streams * co * b->cap_a);
  if (p->fused) {
    if (kp.limiter) {
      alloc((void **)&b->d_hist_y, sizeof(float) * (size_t)n_streams * co * kLimDelay);
      alloc((void **)&b->d_hist_pk, sizeof(float) * (size_t)n_streams * kLimDelay);
    }
  } else {
    alloc((void **)&b->d_tl_b, sizeof(float) * (size_t)n_streams * co * b->cap_b);
  }
  if (kp.limiter && !p->fused) {
    alloc((void **)&b->d_pk, sizeof(float) * (size_t)n_streams * b->cap_b);
    alloc((void **)&b->d_wm, sizeof(float) * (size_t)n_streams * b->cap_b);
    alloc((void **)&b->d_gn, sizeof(float) * (size_t)n_streams * b->cap_b);
  }
  if (e != cudaSuccess) {
    int r = fail(IAMFB_ERR_ALLOC_FAIL, "device allocation failed: %s", cudaGetErrorString(e));
    iamfb_batch_destroy(b);
    return r;
  }
  if (p->hrtf) {
    int r = iamfb_hrtf_batch_create(p->hrtf, n_streams, max_frames, &b->hrtf);
    if (r) { iamfb_batch_destroy(b); return r; }
  }
  if (p->kp_front) {
    alloc((void **)&b->d_state_f, sizeof(StreamState) * n_streams);
    alloc((void **)&b->d_frames_f, sizeof(FrameRec) * (size_t)n_streams * max_frames);
    alloc((void **)&b->d_submit_f, sizeof(SubmitRec) * n_streams);
    for (int el = 0; el < kp.n_elements; ++el)
      if (p->front_demix[el])
        alloc((void **)&b->d_demixed[el], sizeof(float) * (size_t)n_streams * max_frames * p->kp_front->el[el].n_rec * kp.frame_size);
    if (e != cudaSuccess) {
      int r = fail(IAMFB_ERR_ALLOC_FAIL, "device allocation failed: %s", cudaGetErrorString(e));
      iamfb_batch_destroy(b);
      return r;
    }
  }
  int r = iamfb_batch_reset(b);
  if (r) { iamfb_batch_destroy(b); return r; }
  *out = b;
  return IAMFB_OK;
}

static void free_staging(iamfb_batch *b) {
  for (int e = 0; e < kMaxEl; ++e) {
    cudaFree(b->d_in[e]); cudaFree(b->d_ramp[e]); cudaFree(b->d_in16[e]);
    b->d_in[e] = b->d_ramp[e] = nullptr;
    b->d_in16[e] = nullptr;
  }
  cudaFree(b->d_oramp); cudaFree(b->d_params); cudaFree(b->d_pcm); cudaFree(b->d_counts);
  b->d_oramp = nullptr; b->d_params = nullptr; b->d_pcm = nullptr; b->d_counts = nullptr;
  b->stage_frames = 0;
}

extern "C" void iamfb_batch_destroy(iamfb_batch *b) {
  if (!b) return;
  cudaFree(b->d_state); cudaFree(b->d_frames); cudaFree(b->d_submit); cudaFree(b->d_submit_mk); cudaFree(b->d_gate);
  cudaFree(b->d_tl_a); cudaFree(b->d_tl_b); cudaFree(b->d_pk); cudaFree(b->d_wm); cudaFree(b->d_gn);
  cudaFree(b->d_hist_y); cudaFree(b->d_hist_pk);
  for (int e = 0; e < kMaxEl; ++e) cudaFree(b->d_wide[e]);
  iamfb_hrtf_batch_destroy(b->hrtf);
  cudaFree(b->d_state_f); cudaFree(b->d_frames_f); cudaFree(b->d_submit_f);
  for (int e = 0; e < kMaxEl; ++e) cudaFree(b->d_demixed[e]);
  for (int e = 0; e < kMaxEl; ++e) cudaFree(b->d_seg_ramp[e]);
  cudaFree(b->d_seg_oramp);
  for (int e = 0; e <= kMaxEl; ++e) cudaFree(b->d_segs[e]);
  free_staging(b);
  delete b;
}

// ---------------------------------------------------------------------------------------------------------------------
// kernel sequencing
// ---------------------------------------------------------------------------------------------------------------------
#define LAUNCH_CHECK(name)                                                                                      \
  do {                                                                                                          \
    cudaError_t e_ = cudaGetLastError();                                                                        \
    if (e_ != cudaSuccess) return fail(IAMFB_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
    ++ctx->launches;                                                                                            \
  } while (0)

template <int VEC>
static int launch_render(iamfb_ctx *ctx, int tmpl, const KernelPlan &kp, const RenderArgs &ra, int blocks) {
  cudaStream_t st = ctx->stream;
#define RCASE(ID, LAYOUT, NREC) \
  case ID: { ScopedKernelTimer tm_(ctx, "k_render"); k_render<LAYOUT, NREC, VEC><<<blocks, 128, 0, st>>>(kp, ra); } break;
  switch (tmpl) {
    RCASE(0, 0, 1) RCASE(1, 1, 2) RCASE(2, 2, 6) RCASE(3, 3, 8) RCASE(4, 4, 10) RCASE(5, 5, 8) RCASE(6, 6, 10)
    RCASE(7, 7, 12) RCASE(8, 8, 6)
    RCASE(10, -1, 1) RCASE(11, -1, 4) RCASE(12, -1, 9) RCASE(13, -1, 16)
    default: return fail(IAMFB_ERR_INTERNAL, "no render kernel variant %d", tmpl);
  }
#undef RCASE
  LAUNCH_CHECK("k_render");
  return IAMFB_OK;
}

static int n_subchunks(const KernelPlan &kp, int F, bool flush) {
  if (flush || !kp.limiter || kp.resample) return 1;
  int n = F >= 8 ? 4 : (F >= 4 ? 2 : 1);
  if (n > F) n = F;
  return n;
}

// int16 -> float32 * 2^-15 (the codec glue's scaling, opus/IAMF_opus_decoder.c:133-135); 8 samples per thread
static __global__ void __launch_bounds__(256) k_widen_s16(const int16_t *__restrict__ src, float *__restrict__ dst, size_t n8, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n8) {
    const int4 v = reinterpret_cast<const int4 *>(src)[i];
    const int w[4] = {v.x, v.y, v.z, v.w};
    float o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[2 * k] = (float)(short)(w[k] & 0xffff) / 32768.f;
      o[2 * k + 1] = (float)(short)(w[k] >> 16) / 32768.f;
    }
    reinterpret_cast<float4 *>(dst)[2 * i] = make_float4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<float4 *>(dst)[2 * i + 1] = make_float4(o[4], o[5], o[6], o[7]);
  } else if (i == n8) {
    for (size_t k = n8 * 8; k < n; ++k) dst[k] = (float)src[k] / 32768.f;
  }
}


// the same for the rows of the streams a submit flags irregular only (per = elements per stream, a multiple of 8)
static __global__ void __launch_bounds__(256) k_widen_s16_irregular(const int16_t *__restrict__ src, float *__restrict__ dst, size_t per,
                                                                    const SubmitRec *__restrict__ submit, const int *gate) {
  const int s = blockIdx.y;
  if (*gate == 0 || !submit[s].irregular) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per / 8) return;
  const int4 v = reinterpret_cast<const int4 *>(src + (size_t)s * per)[i];
  const int w[4] = {v.x, v.y, v.z, v.w};
  float o[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    o[2 * k] = (float)(short)(w[k] & 0xffff) / 32768.f;
    o[2 * k + 1] = (float)(short)(w[k] >> 16) / 32768.f;
  }
  float4 *d = reinterpret_cast<float4 *>(dst + (size_t)s * per);
  d[2 * i] = make_float4(o[0], o[1], o[2], o[3]);
  d[2 * i + 1] = make_float4(o[4], o[5], o[6], o[7]);
}

// float32 copy of an int16 submit for the kernels that stage float32 (everything but k_pipe): per element
// [S][F][n_in][N], allocated on first use for the batch's Fmax
static int ensure_wide(iamfb_batch *b) {
  const KernelPlan &kp = b->plan->kp;
  for (int e = 0; e < kp.n_elements; ++e)
    if (!b->d_wide[e]) {
      cudaError_t er = cudaMalloc((void **)&b->d_wide[e], sizeof(float) * (size_t)b->S * b->Fmax * b->plan->in_rows[e] * kp.frame_size);
      if (er != cudaSuccess) return fail(IAMFB_ERR_ALLOC_FAIL, "float32 staging of an int16 submit: %s", cudaGetErrorString(er));
    }
  return IAMFB_OK;
}
static int widen_streams(iamfb_batch *b, const iamfb_io *io, int F, int s_lo, int s_cnt, iamfb_io *wio) {
  iamfb_ctx *ctx = b->plan->ctx;
  const KernelPlan &kp = b->plan->kp;
  int r = ensure_wide(b);
  if (r) return r;
  *wio = *io;
  wio->in_format = IAMFB_IN_F32;
  for (int e = 0; e < kp.n_elements; ++e) {
    const size_t per = (size_t)F * b->plan->in_rows[e] * kp.frame_size, n = (size_t)s_cnt * per, n8 = n / 8, off = (size_t)s_lo * per;
    const int16_t *src = reinterpret_cast<const int16_t *>(io->in[e]) + off;
    float *dst = b->d_wide[e] + off;
    {
      ScopedKernelTimer tm_(ctx, "k_widen_s16");
      if ((off & 7) == 0 && (((size_t)src) & 15) == 0) k_widen_s16<<<(unsigned)((n8 + 1 + 255) / 256), 256, 0, ctx->stream>>>(src, dst, n8, n);
      else k_widen_s16<<<1, 256, 0, ctx->stream>>>(src, dst, 0, n);   // unaligned group start: scalar
    }
    cudaError_t e_ = cudaGetLastError();
    if (e_ != cudaSuccess) return fail(IAMFB_ERR_CUDA, "launch of k_widen_s16 failed: %s", cudaGetErrorString(e_));
    ++ctx->launches;
    wio->in[e] = b->d_wide[e];
  }
  return IAMFB_OK;
}

// ---- k_gain_expand: animated mix gains from their parameter segments (iamfb_gain_ramp) to per-sample gains, with the
// reference's expressions in the reference's precision (mix_gain_bezier_linear / _quad, IAMF_decoder.c:639-664; pow(x, 2)
// of a double is x * x exactly).  One block per (stream, frame); a frame without segments gets the constant of its
// iamfb_frame_params (a constant the reference would skip - exactly 1 or not positive, :1392 - becomes 1.0: exact).
static __global__ void __launch_bounds__(256) k_gain_expand(const iamfb_gain_ramp *__restrict__ segs, const iamfb_frame_params *__restrict__ params,
                                                            float *__restrict__ ramp, int N, int element) {
  const size_t sf = blockIdx.x;
  const iamfb_gain_ramp &r = segs[sf];
  float *g = ramp + sf * (size_t)N;
  if (r.n_segs <= 0) {
    float c = element < 0 ? params[sf].out_gain : params[sf].el[element].mix_gain;
    if (!(c != 1.f && c > 0.f)) c = 1.f;
    for (int k = threadIdx.x; k < N; k += blockDim.x) g[k] = c;
    return;
  }
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    int base = 0, q = 0;
    while (q < r.n_segs && q < IAMFB_MAX_GAIN_SEGS && k >= base + r.seg[q].count) base += r.seg[q++].count;
    float v = 1.f;
    if (q < r.n_segs && q < IAMFB_MAX_GAIN_SEGS) {
      const iamfb_gain_seg &sg = r.seg[q];
      const int i = sg.offset + (k - base);
      const float s = sg.start, e = sg.end, c = sg.control;
      if (sg.type == 0) {
        v = s;
      } else if (sg.type == 1) {
        v = s + (e - s) * i / sg.interval;
      } else {
        const long long alpha = (long long)sg.interval - 2 * (long long)sg.ct;
        float a;
        if (alpha) {
          a = (float)((sqrt((double)sg.ct * (double)sg.ct + (double)(alpha * i)) - (double)sg.ct) / (double)alpha);
        } else {
          a = (float)i;
          a /= (float)(2 * sg.ct);
        }
        v = (float)((double)(s + e - 2 * c) * ((double)a * (double)a) + (double)(2 * a * (c - s)) + (double)s);
      }
    }
    g[k] = v;
  }
}

// ---- k_hrtf_demix: the channels that enter the binaural renderer for a scalable element - the de-mixer's derivation chain,
// output gains and recon-gain cross-fade (demixer_demixing, demixer.c:636-664; exactly the first half of k_render's thread),
// written as float32 [S][F][n_rec][N] in layout order.  Trims do not apply here: the renderer sees the untrimmed frame.
struct DemixArgs {
  const float *in;            // [S][F][n_in][N]
  const iamfb_frame_params *params;
  const FrameRec *frames;     // resolved with the FRONT plan
  const float *start_win, *stop_win;
  float *out;                 // [S][F][NREC][N]
  int e;
};
template <int LAYOUT, int NREC>
static __global__ void __launch_bounds__(128) k_hrtf_demix(const __grid_constant__ KernelPlan plan, DemixArgs a) {
  const int N = plan.frame_size;
  const size_t sf = blockIdx.y;
  const int i0 = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (i0 >= N || a.params[sf].trim_start == 0xFFFF) return;
  const ElPlan &ep = plan.el[a.e];
  const ElFrame &ef = a.frames[sf].el[a.e];
  Vec<4> v[kChCount];
  reconstruct_channels<4, false>(ep, ef, a.in + sf * ep.n_in * N + i0, N, true, 4, v);
  constexpr unsigned char kOrder[9][12] = {
      {13}, {14, 15}, {1, 2, 3, 4, 20, 21}, {1, 2, 3, 4, 20, 21, 22, 23}, {1, 2, 3, 4, 20, 21, 9, 10, 11, 12},
      {1, 2, 3, 4, 5, 6, 7, 8}, {1, 2, 3, 4, 5, 6, 7, 8, 22, 23}, {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12},
      {18, 19, 3, 4, 16, 17}};
#pragma unroll
  for (int m = 0; m < NREC; ++m) {
    Vec<4> x = v[kOrder[LAYOUT][m]];
    if ((ef.rmask >> m) & 1u) {   // dmx_rms cross-fade, demixer.c:461-468
      const float last = ef.rlast[m], cur = ef.rcur[m];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k;
        float st = 0.f, sw = 1.f;
        if (i < plan.overlap) { st = a.stop_win[i]; sw = a.start_win[i]; }
        x.v[k] *= last * st + cur * sw;
      }
    }
    *reinterpret_cast<float4 *>(a.out + (sf * NREC + m) * N + i0) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
  }
}

// per-batch float ramp buffers for device-resident submits with gain segments (host-resident submits use the staging ones)
static int ensure_ramps(iamfb_batch *b) {
  const KernelPlan &kp = b->plan->kp;
  const size_t n = (size_t)b->S * b->Fmax * kp.frame_size;
  for (int e = 0; e <= kp.n_elements; ++e) {
    float **p = e == kp.n_elements ? &b->d_seg_oramp : &b->d_seg_ramp[e];
    if (!*p && cudaMalloc((void **)p, sizeof(float) * n) != cudaSuccess) return fail(IAMFB_ERR_ALLOC_FAIL, "gain ramp buffers");
  }
  return IAMFB_OK;
}

// runs the kernels of one submit (or flush) for the streams [s_lo, s_lo + s_cnt) of the batch; all pointers in `io`,
// `pcm` and `counts` are DEVICE pointers to the whole batch's arrays (stream 0 first)
static int run_pipeline(iamfb_batch *b, const iamfb_io *io, int F, bool flush, void *pcm, int32_t *counts, size_t stride,
                        int s_lo = 0, int s_cnt = -1) {
  iamfb_plan *p = b->plan;
  iamfb_ctx *ctx = p->ctx;
  const KernelPlan &kp = p->kp;
  cudaStream_t st = ctx->stream;
  if (s_cnt < 0) s_cnt = b->S;
  iamfb_io wio, hio, gio;
  if (!flush && (io->gain_segs[0] || io->gain_segs[1] || io->out_gain_segs)) {
    // animated mix gains arrive as parameter segments: evaluated here, on the device, into the ramp arrays the kernels read
    int r = ensure_ramps(b);
    if (r) return r;
    gio = *io;
    const size_t off = (size_t)s_lo * F;
    for (int e = 0; e <= kp.n_elements; ++e) {
      const bool out = e == kp.n_elements;
      const iamfb_gain_ramp *sg = out ? io->out_gain_segs : io->gain_segs[e];
      if (!sg) continue;
      float *dst = out ? b->d_seg_oramp : b->d_seg_ramp[e];
      {
        ScopedKernelTimer tm_(ctx, "k_gain_expand");
        k_gain_expand<<<(unsigned)((size_t)s_cnt * F), 256, 0, st>>>(sg + off, io->params + off, dst + off * kp.frame_size, kp.frame_size, out ? -1 : e);
      }
      cudaError_t e_ = cudaGetLastError();
      if (e_ != cudaSuccess) return fail(IAMFB_ERR_CUDA, "launch of k_gain_expand failed: %s", cudaGetErrorString(e_));
      ++ctx->launches;
      if (out) { gio.out_gain_ramp = dst; gio.out_gain_segs = nullptr; }
      else { gio.gain_ramp[e] = dst; gio.gain_segs[e] = nullptr; }
    }
    io = &gio;
  }
  if (p->hrtf && !flush) {
    const float *demixed[kMaxEl] = {nullptr, nullptr};
    iamfb_io fio;
    if (p->kp_front) {
      // scalable elements in front of the renderer: their frames resolved with the front plan (own de-mixer state), their
      // layout channels de-mixed into float32
      if (io->in_format == IAMFB_IN_S16) {
        int r = widen_streams(b, io, F, s_lo, s_cnt, &fio);
        if (r) return r;
        io = &fio;
      }
      const KernelPlan &kf = *p->kp_front;
      ResolveArgs a;
      memset(&a, 0, sizeof(a));
      a.params = io->params + (size_t)s_lo * F;
      a.state = b->d_state_f + s_lo;
      a.frames = b->d_frames_f + (size_t)s_lo * F;
      a.submit = b->d_submit_f + s_lo;
      a.qf_table = p->d_qf;
      a.n_streams = s_cnt;
      a.n_frames = F;
      a.n_sub = 1;
      a.sub_frame[0] = 0;
      for (int c = 1; c <= kMaxSub; ++c) a.sub_frame[c] = F;
      {
        const int per_block = kResolveThreads / 32;
        ScopedKernelTimer tm_(ctx, "k_resolve");
        k_resolve<<<(s_cnt + per_block - 1) / per_block, kResolveThreads, 0, st>>>(kf, a);
      }
      { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return fail(IAMFB_ERR_CUDA, "launch of k_resolve (front) failed: %s", cudaGetErrorString(e_)); ++ctx->launches; }
      for (int e = 0; e < kf.n_elements; ++e) {
        if (!p->front_demix[e]) continue;
        DemixArgs da;
        da.in = io->in[e] + (size_t)s_lo * F * kf.el[e].n_in * kf.frame_size;
        da.params = io->params + (size_t)s_lo * F;
        da.frames = b->d_frames_f + (size_t)s_lo * F;
        da.start_win = p->d_start_win;
        da.stop_win = p->d_stop_win;
        da.out = b->d_demixed[e] + (size_t)s_lo * F * kf.el[e].n_rec * kf.frame_size;
        da.e = e;
        dim3 grid((kf.frame_size / 4 + 127) / 128, (unsigned)((size_t)s_cnt * F));
        {
          ScopedKernelTimer tm_(ctx, "k_hrtf_demix");
#define DCASE(ID, LAYOUT, NREC) case ID: k_hrtf_demix<LAYOUT, NREC><<<grid, 128, 0, st>>>(kf, da); break;
          switch (p->front_tmpl[e]) {
            DCASE(0, 0, 1) DCASE(1, 1, 2) DCASE(2, 2, 6) DCASE(3, 3, 8) DCASE(4, 4, 10) DCASE(5, 5, 8) DCASE(6, 6, 10) DCASE(7, 7, 12) DCASE(8, 8, 6)
            default: return fail(IAMFB_ERR_INTERNAL, "no de-mixing kernel variant %d", p->front_tmpl[e]);
          }
#undef DCASE
        }
        { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return fail(IAMFB_ERR_CUDA, "launch of k_hrtf_demix failed: %s", cudaGetErrorString(e_)); ++ctx->launches; }
        demixed[e] = b->d_demixed[e];
      }
    }
    // binaural HRTF front end: the elements it renders reach the kernels below as float32 [2][N] frames
    int r = iamfb_hrtf_run(ctx, p->hrtf, b->hrtf, io, F, s_lo, s_cnt, &hio, demixed);
    if (r) return r;
    io = &hio;
  }
  // resampling plans: k_pipe_rs renders the regular streams, the multi-kernel path below the irregular ones
  const bool rs_native = p->rs_pipe && !flush && !io->gain_ramp[0] && !io->out_gain_ramp;
  const bool rs_split = rs_native && p->rs_split;
  const bool s16_in = !flush && io->in_format == IAMFB_IN_S16 &&
                      ((p->fused && p->s16_native && !(p->stream && !p->pipe)) || rs_native);
  if (!flush && io->in_format == IAMFB_IN_S16 && !s16_in) {
    // kernels that stage float32 get a widened copy (x / 32768 is exact)
    int r = widen_streams(b, io, F, s_lo, s_cnt, &wio);
    if (r) return r;
    io = &wio;
  }
  const int S = s_cnt, N = kp.frame_size, co = kp.out_channels;
  if (!p->fused && (s_lo != 0 || s_cnt != b->S)) return fail(IAMFB_ERR_INTERNAL, "stream ranges need the fused path");
  const int n_sub = p->fused ? 1 : n_subchunks(kp, F, flush);
  int sub_frame[kMaxSub + 1];
  for (int c = 0; c <= kMaxSub; ++c) sub_frame[c] = flush ? 0 : (c >= n_sub ? F : (int)((long long)F * c / n_sub));
  const size_t fF = flush ? 0 : (size_t)F;   // frames per stream in the per-frame arrays of this call

  // K0
  {
    ResolveArgs a;
    a.params = flush ? nullptr : io->params + (size_t)s_lo * fF;
    a.state = b->d_state + s_lo;
    a.frames = b->d_frames + (size_t)s_lo * fF;
    a.submit = b->d_submit + s_lo;
    a.submit_mk = rs_native ? b->d_submit_mk : nullptr;
    a.gate = rs_native ? b->d_gate + (b->submit_seq & 1) : nullptr;
    a.gate_next = rs_native ? b->d_gate + ((b->submit_seq + 1) & 1) : nullptr;
    a.out_counts = counts ? counts + (size_t)s_lo * (flush ? 1 : fF) : nullptr;
    a.qf_table = p->d_qf;
    a.n_streams = S;
    a.n_frames = flush ? 0 : F;
    a.flush = flush ? 1 : 0;
    a.n_sub = n_sub;
    for (int c = 0; c <= kMaxSub; ++c) a.sub_frame[c] = sub_frame[c];
    {
      const int per_block = kResolveThreads / 32;   // one warp per stream
      ScopedKernelTimer tm_(ctx, "k_resolve");
      k_resolve<<<(S + per_block - 1) / per_block, kResolveThreads, 0, st>>>(kp, a);
    }
    LAUNCH_CHECK("k_resolve");
  }
  if (p->fused) {
    FusedArgs fa;
    memset(&fa, 0, sizeof(fa));
    if (!flush)
      for (int e = 0; e < kp.n_elements; ++e) {
        fa.in[e] = s16_in ? reinterpret_cast<const float *>(reinterpret_cast<const int16_t *>(io->in[e]) + (size_t)s_lo * fF * kp.el[e].n_in * N)
                          : io->in[e] + (size_t)s_lo * fF * kp.el[e].n_in * N;
        fa.gain_ramp[e] = io->gain_ramp[e] ? io->gain_ramp[e] + (size_t)s_lo * fF * N : nullptr;
      }
    fa.out_gain_ramp = (flush || !io->out_gain_ramp) ? nullptr : io->out_gain_ramp + (size_t)s_lo * fF * N;
    fa.frames = b->d_frames + (size_t)s_lo * fF;
    fa.start_win = p->d_start_win;
    fa.stop_win = p->d_stop_win;
    fa.submit = b->d_submit + s_lo;
    fa.state = b->d_state + s_lo;
    fa.acc = p->d_acc;
    fa.hist_y = b->d_hist_y ? b->d_hist_y + (size_t)s_lo * co * kLimDelay : nullptr;
    fa.hist_pk = b->d_hist_pk ? b->d_hist_pk + (size_t)s_lo * kLimDelay : nullptr;
    fa.pcm = (char *)pcm + (size_t)s_lo * stride;
    fa.stride_bytes = stride;
    fa.n_frames = flush ? 0 : F;
    fa.flush = flush ? 1 : 0;
    fa.tile = p->fused_tile;
    fa.only_irregular = 0;
    fa.in_s16 = s16_in ? 1 : 0;
    const bool ramps = !flush && (io->gain_ramp[0] || io->gain_ramp[1] || io->out_gain_ramp);   // (io is null in a flush)
    // (k_stream has no animated-gain form: those submits take k_pipe, which has)
    const bool stream_first = p->stream && !s16_in && !(ramps && p->pipe);
    if (p->pipe && !stream_first && !flush) {
      // untrimmed streams: the pipelined kernel; the rest (flagged by k_resolve): k_fused
      int r = launch_pipe(ctx, p, fa, S, s16_in);
      if (r) return r;
      fa.only_irregular = 1;
    } else if (p->stream && !flush && !io->gain_ramp[0] && !io->out_gain_ramp) {
      // untrimmed streams: the register-resident pipelined kernel; the rest (flagged by k_resolve): k_fused
      int r = launch_stream(ctx, p, fa, S);
      if (r) return r;
      fa.only_irregular = 1;
    }
    return launch_fused(ctx, p, fa, S);
  }
  const SubmitRec *mk_submit = b->d_submit;      // the records the kernels below work from
  // streams per grid row: behind k_pipe_rs only the (rare) irregular streams are left - a small grid strides over the batch
  const int gy = rs_native ? (S < 32 ? S : 32) : S;
  const int *gate = rs_native ? b->d_gate + (b->submit_seq & 1) : nullptr;
  if (rs_native) ++b->submit_seq;
  if (rs_native) {
    mk_submit = b->d_submit_mk;                    // regular streams: nothing left to do
    iamfb_io pio = *io;
    pio.pcm = pcm;
    int r = rs_split ? launch_rs_split(ctx, p, b, &pio, F, s16_in) : launch_pipe_rs(ctx, p, b, &pio, F, s16_in);
    if (r) return r;
    if (s16_in) {
      // the irregular streams' rows as float32 for k_render (regular streams are skipped)
      r = ensure_wide(b);
      if (r) return r;
      wio = *io;
      wio.in_format = IAMFB_IN_F32;
      for (int e = 0; e < kp.n_elements; ++e) {
        const size_t per = (size_t)F * kp.el[e].n_in * N;
        {
          ScopedKernelTimer tm_(ctx, "k_widen_s16");
          k_widen_s16_irregular<<<dim3((unsigned)((per / 8 + 255) / 256), S), 256, 0, st>>>(reinterpret_cast<const int16_t *>(io->in[e]), b->d_wide[e], per, b->d_submit, gate);
        }
        LAUNCH_CHECK("k_widen_s16_irregular");
        wio.in[e] = b->d_wide[e];
      }
      io = &wio;
    }
  }
  float *tl_first = kp.resample ? b->d_tl_a : b->d_tl_b;
  const int cap_first = kp.resample ? b->cap_a : b->cap_b;
  const int hist_first = kp.resample ? kp.rs_hist : kp.hist;

  if (flush) {
    // the resampler is fed filt_len/2 zeros, the limiter the resampler tail followed by 240 zeros
    if (kp.resample)
      CU(cudaMemset2DAsync(b->d_tl_a + kp.rs_hist, sizeof(float) * b->cap_a, 0, sizeof(float) * (kp.rs_filt_len / 2), (size_t)S * co, st));
    const size_t zlen = (size_t)iamfb_plan_max_out_samples(p, 0);
    CU(cudaMemset2DAsync(b->d_tl_b + kp.hist, sizeof(float) * b->cap_b, 0, sizeof(float) * zlen, (size_t)S * co, st));
    if (b->d_pk) CU(cudaMemset2DAsync(b->d_pk + kp.hist, sizeof(float) * b->cap_b, 0, sizeof(float) * zlen, (size_t)S, st));
  }

  auto render = [&](int f_lo, int nf) -> int {
    const bool vec4 = (N % 4) == 0;
    const int vec = vec4 ? 4 : 1;
    const int tiles = (N + 128 * vec - 1) / (128 * vec);
    for (int e = 0; e < kp.n_elements; ++e) {
      RenderArgs ra;
      ra.in = io->in[e];
      ra.frames = b->d_frames;
      ra.gain_ramp = io->gain_ramp[e];
      ra.out_gain_ramp = io->out_gain_ramp;
      ra.start_win = p->d_start_win;
      ra.stop_win = p->d_stop_win;
      ra.tl = tl_first;
      ra.pk = (!kp.resample && kp.limiter) ? b->d_pk : nullptr;
      ra.cap = cap_first;
      ra.hist = hist_first;
      ra.n_frames = F;
      ra.e = e;
      ra.f_lo = f_lo;
      ra.nf = nf;
      ra.first = e == 0;
      ra.last = e == kp.n_elements - 1;
      ra.tiles_per_frame = tiles;
      ra.only_irregular = rs_native ? b->d_submit : nullptr;
      ra.n_blocks = S * nf * tiles;
      ra.gate = gate;
      int blocks = rs_native ? (ra.n_blocks < 2048 ? ra.n_blocks : 2048) : ra.n_blocks;
      int r = vec4 ? launch_render<4>(ctx, p->tmpl[e], kp, ra, blocks) : launch_render<1>(ctx, p->tmpl[e], kp, ra, blocks);
      if (r) return r;
    }
    return IAMFB_OK;
  };
  auto output = [&](int sub, int max_len) -> int {
    OutputArgs o;
    o.tl = b->d_tl_b; o.gn = kp.limiter ? b->d_gn : nullptr; o.submit = mk_submit; o.pcm = pcm;
    o.stride_bytes = stride; o.cap = b->cap_b; o.hist = kp.hist; o.sub = sub; o.n_streams = S; o.gate = gate;
    {
      ScopedKernelTimer tm_(ctx, "k_output");
      dim3 grid((max_len / 4 + 1 + 255) / 256, gy);
      switch (kp.bit_depth) {
        case 16: k_output<16><<<grid, 256, 0, st>>>(kp, o); break;
        case 24: k_output<24><<<grid, 256, 0, st>>>(kp, o); break;
        case 32: k_output<32><<<grid, 256, 0, st>>>(kp, o); break;
        default: k_output<0><<<grid, 256, 0, st>>>(kp, o); break;
      }
    }
    LAUNCH_CHECK("k_output");
    return IAMFB_OK;
  };

  const int max_out = flush ? iamfb_plan_max_out_samples(p, 0) : iamfb_plan_max_out_samples(p, F);
  if (n_sub == 1) {
    if (!flush) { int r = render(0, F); if (r) return r; }
    if (kp.resample) {
      ResampleArgs a;
      a.src = b->d_tl_a;
      a.dst = b->d_tl_b;
      a.pk = b->d_pk;
      a.submit = mk_submit;
      a.state = b->d_state;
      a.sinc = p->d_sinc;
      a.cap_a = b->cap_a;
      a.cap_b = b->cap_b;
      a.hist_b = kp.hist;
      a.max_out = max_out;
      a.flush = flush ? 1 : 0;
      a.n_streams = S;
      a.gate = gate;
      dim3 grid((max_out + 127) / 128, gy);
      const size_t smem2 = p->d_tab4 ? sizeof(float4) * kp.rs_oversample * (kp.rs_filt_len + 1) + sizeof(float) * 2 * p->rs_span : 0;
      if (p->d_tab4 && smem2 <= 200 * 1024) {
        Resample2Args a2;
        a2.r = a;
        a2.tab4 = p->d_tab4;
        a2.span = p->rs_span;
        CU(cudaFuncSetAttribute(k_resample_interp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        { ScopedKernelTimer tm_(ctx, "k_resample"); k_resample_interp<<<grid, 128, smem2, st>>>(kp, a2); }
      } else {
        size_t smem = p->sinc_len <= 12 * 1024 ? sizeof(float) * p->sinc_len : 0;
        { ScopedKernelTimer tm_(ctx, "k_resample"); k_resample<<<grid, 128, smem, st>>>(kp, a); }
      }
      LAUNCH_CHECK("k_resample");
      if (!flush) {
        CarryArgs c;
        c.tl = b->d_tl_a; c.submit = mk_submit; c.rows = co; c.cap = b->cap_a; c.hist = kp.rs_hist; c.use_in_len = 1; c.gate = gate;
        { ScopedKernelTimer tm_(ctx, "k_carry"); k_carry<<<dim3(co, S), 256, 0, st>>>(c); }
        LAUNCH_CHECK("k_carry");
      }
    }
  }
  if (kp.limiter) {
    const size_t scan_smem = (kp.lim_jr + 4) <= kScanAccSmem ? sizeof(float) * (kp.lim_jr + 4) : 0;
    if (scan_smem) CU(cudaFuncSetAttribute(k_limiter_scan<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem));
    cudaStream_t scan_st = n_sub > 1 ? ctx->aux : st;
    for (int c = 0; c < n_sub; ++c) {
      const int sub_len = n_sub == 1 ? max_out : (sub_frame[c + 1] - sub_frame[c]) * N;
      if (n_sub > 1) { int r = render(sub_frame[c], sub_frame[c + 1] - sub_frame[c]); if (r) return r; }
      WmaxArgs w;
      w.pk = b->d_pk; w.wm = b->d_wm; w.submit = mk_submit; w.cap = b->cap_b; w.hist = kp.hist; w.flush = flush; w.sub = c; w.n_streams = S; w.gate = gate;
      { ScopedKernelTimer tm_(ctx, "k_window_max"); k_window_max<<<dim3((sub_len + kWmTile - 1) / kWmTile, gy), 256, 0, st>>>(kp, w); }
      LAUNCH_CHECK("k_window_max");
      if (n_sub > 1) {
        CU(cudaEventRecord(ctx->ev_w[c], st));
        CU(cudaStreamWaitEvent(scan_st, ctx->ev_w[c], 0));
      }
      ScanArgs sa;
      sa.wm = b->d_wm; sa.gn = b->d_gn; sa.state = b->d_state; sa.submit = mk_submit; sa.acc = p->d_acc;
      sa.cap = b->cap_b; sa.hist = kp.hist; sa.n_streams = S; sa.max_len = sub_len; sa.sub = c; sa.gate = gate;
      {
        ScopedKernelTimer tm_(ctx, "k_limiter_scan", scan_st);
        if (scan_smem) k_limiter_scan<true><<<(S + 31) / 32, kScanThreads, scan_smem, scan_st>>>(kp, sa);
        else k_limiter_scan<false><<<(S + 31) / 32, kScanThreads, 0, scan_st>>>(kp, sa);
      }
      LAUNCH_CHECK("k_limiter_scan");
      if (n_sub > 1) {
        CU(cudaEventRecord(ctx->ev_s[c], scan_st));
        if (c > 0) {   // quantise the previous sub-chunk while this one is being scanned
          CU(cudaStreamWaitEvent(st, ctx->ev_s[c - 1], 0));
          int r = output(c - 1, (sub_frame[c] - sub_frame[c - 1]) * N);
          if (r) return r;
        }
      }
    }
    if (n_sub > 1) {
      CU(cudaStreamWaitEvent(st, ctx->ev_s[n_sub - 1], 0));
      int r = output(n_sub - 1, (sub_frame[n_sub] - sub_frame[n_sub - 1]) * N);
      if (r) return r;
    } else {
      int r = output(0, max_out);
      if (r) return r;
    }
    if (!flush) {
      CarryArgs c;
      c.tl = b->d_tl_b; c.submit = mk_submit; c.rows = co; c.cap = b->cap_b; c.hist = kLimDelay; c.use_in_len = 0; c.gate = gate;
      { ScopedKernelTimer tm_(ctx, "k_carry"); k_carry<<<dim3(co, S), 256, 0, st>>>(c); }
      LAUNCH_CHECK("k_carry");
      c.tl = b->d_pk; c.rows = 1;
      { ScopedKernelTimer tm_(ctx, "k_carry"); k_carry<<<dim3(1, S), 256, 0, st>>>(c); }
      LAUNCH_CHECK("k_carry");
    }
  } else {
    int r = output(0, max_out);
    if (r) return r;
  }
  return IAMFB_OK;
}

extern "C" int iamfb_batch_submit_device(iamfb_batch *b, const iamfb_io *io, int n_frames) {
  if (!b || !io) return fail(IAMFB_ERR_BAD_ARG, "submit: null argument");
  if (n_frames <= 0 || n_frames > b->Fmax) return fail(IAMFB_ERR_BAD_ARG, "submit: %d frames (batch sized for %d)", n_frames, b->Fmax);
  for (int e = 0; e < b->plan->kp.n_elements; ++e)
    if (!io->in[e]) return fail(IAMFB_ERR_BAD_ARG, "submit: input of element %d is null", e);
  if (!io->params || !io->pcm) return fail(IAMFB_ERR_BAD_ARG, "submit: params / pcm is null");
  if (io->in_format != IAMFB_IN_F32 && io->in_format != IAMFB_IN_S16) return fail(IAMFB_ERR_BAD_ARG, "submit: in_format %d", io->in_format);
  CU(cudaSetDevice(b->plan->ctx->device));
  return run_pipeline(b, io, n_frames, false, io->pcm, io->out_counts, iamfb_plan_out_stride_bytes(b->plan, n_frames));
}

extern "C" int iamfb_batch_flush_device(iamfb_batch *b, void *pcm, int32_t *counts) {
  if (!b || !pcm) return fail(IAMFB_ERR_BAD_ARG, "flush: null argument");
  CU(cudaSetDevice(b->plan->ctx->device));
  if (!b->plan->kp.limiter && !b->plan->kp.resample) {   // iamf_delay_buffer_handle returns 0 (:3259-3260)
    if (counts) CU(cudaMemsetAsync(counts, 0, sizeof(int32_t) * b->S, b->plan->ctx->stream));
    return IAMFB_OK;
  }
  return run_pipeline(b, nullptr, 0, true, pcm, counts, iamfb_plan_out_stride_bytes(b->plan, 1));
}

static int ensure_staging(iamfb_batch *b, int F) {
  if ((size_t)F <= b->stage_frames) return IAMFB_OK;
  free_staging(b);
  const KernelPlan &kp = b->plan->kp;
  const size_t S = b->S, N = kp.frame_size;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void **ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes ? bytes : 16); };
  for (int el = 0; el < kp.n_elements; ++el) {
    alloc((void **)&b->d_in[el], sizeof(float) * S * F * b->plan->in_rows[el] * N);
    alloc((void **)&b->d_in16[el], sizeof(int16_t) * S * F * b->plan->in_rows[el] * N);
    alloc((void **)&b->d_ramp[el], sizeof(float) * S * F * N);
  }
  alloc((void **)&b->d_oramp, sizeof(float) * S * F * N);
  alloc((void **)&b->d_params, sizeof(iamfb_frame_params) * S * F);
  alloc((void **)&b->d_pcm, S * iamfb_plan_out_stride_bytes(b->plan, F));
  alloc((void **)&b->d_counts, sizeof(int32_t) * S * (F > 1 ? F : 1));
  if (e != cudaSuccess) { free_staging(b); return fail(IAMFB_ERR_ALLOC_FAIL, "staging allocation failed: %s", cudaGetErrorString(e)); }
  b->stage_frames = F;
  return IAMFB_OK;
}

// Host-resident submit.  hooks == nullptr: the caller's buffers are complete when the call is made.  With hooks the caller
// produces and consumes them group by group while the device works on the neighbouring groups: fill(s_lo, s_cnt, io) is
// called right before the inputs of the streams [s_lo, s_lo + s_cnt) are uploaded (it writes their slices of in / params /
// ramps and may set or clear the ramp pointers of `io` for this group), drain(s_lo, s_cnt) once their PCM and counts are
// back in host memory.
static int submit_host(iamfb_batch *b, const iamfb_io *io_in, int F, const iamfb_chunk_hooks *hooks) {
  if (!b || !io_in) return fail(IAMFB_ERR_BAD_ARG, "submit: null argument");
  if (F <= 0 || F > b->Fmax) return fail(IAMFB_ERR_BAD_ARG, "submit: %d frames (batch sized for %d)", F, b->Fmax);
  if (!io_in->params || !io_in->pcm) return fail(IAMFB_ERR_BAD_ARG, "submit: params / pcm is null");
  if (io_in->in_format != IAMFB_IN_F32 && io_in->in_format != IAMFB_IN_S16) return fail(IAMFB_ERR_BAD_ARG, "submit: in_format %d", io_in->in_format);
  iamfb_plan *p = b->plan;
  iamfb_ctx *ctx = p->ctx;
  CU(cudaSetDevice(ctx->device));
  int r = ensure_staging(b, F);
  if (r) return r;
  const KernelPlan &kp = p->kp;
  cudaStream_t st = ctx->stream;
  const size_t S = b->S, N = kp.frame_size;
  const bool s16 = io_in->in_format == IAMFB_IN_S16;
  for (int e = 0; e < kp.n_elements; ++e)
    if (!io_in->in[e]) return fail(IAMFB_ERR_BAD_ARG, "submit: input of element %d is null", e);
  iamfb_io io = *io_in;
  const size_t stride = iamfb_plan_out_stride_bytes(p, F);
  // groups of streams flow through upload -> kernels -> download on three streams; the fused path can run any stream
  // range, the multi-kernel path runs the batch as one group
  int n_chunks = 1;
  if (p->fused && S >= 64) {
    n_chunks = 8;   // measured on configs[1]: 2 -> 35.6k, 4 -> 38.9k, 8 -> 41.1k audio-s/s (the last group's download is the exposed tail)
  }
  // the staging buffers are reused by every submit: uploads must not start before the previous submit's kernels are done
  CU(cudaEventRecord(ctx->ev_free, st));
  CU(cudaStreamWaitEvent(ctx->h2d, ctx->ev_free, 0));
  int drained = 0;
  auto drain_to = [&](int upto) -> int {   // groups [drained, upto): wait for their download, hand them to the caller
    for (; drained < upto; ++drained) {
      const size_t s_lo = S * drained / n_chunks, s_hi = S * (drained + 1) / n_chunks;
      if (s_hi == s_lo) continue;
      CU(cudaEventSynchronize(ctx->ev_back[drained]));
      hooks->drain(hooks->user, (int)s_lo, (int)(s_hi - s_lo));
    }
    return IAMFB_OK;
  };
  for (int c = 0; c < n_chunks; ++c) {
    const size_t s_lo = S * c / n_chunks, s_hi = S * (c + 1) / n_chunks, cnt = s_hi - s_lo;
    if (!cnt) continue;
    if (hooks) {
      io = *io_in;
      hooks->fill(hooks->user, (int)s_lo, (int)cnt, &io);
    }
    iamfb_io dio;
    memset(&dio, 0, sizeof(dio));
    for (int e = 0; e < kp.n_elements; ++e) dio.in[e] = s16 ? reinterpret_cast<const float *>(b->d_in16[e]) : b->d_in[e];
    dio.in_format = io.in_format;
    dio.params = b->d_params;
    dio.pcm = b->d_pcm;
    dio.out_counts = b->d_counts;
    // animated gains (rare): this group's slices
    for (int e = 0; e < kp.n_elements; ++e)
      if (io.gain_ramp[e]) {
        CU(cudaMemcpyAsync(b->d_ramp[e] + s_lo * F * N, io.gain_ramp[e] + s_lo * F * N, sizeof(float) * cnt * F * N, cudaMemcpyHostToDevice, ctx->h2d));
        dio.gain_ramp[e] = b->d_ramp[e];
      }
    if (io.out_gain_ramp) {
      CU(cudaMemcpyAsync(b->d_oramp + s_lo * F * N, io.out_gain_ramp + s_lo * F * N, sizeof(float) * cnt * F * N, cudaMemcpyHostToDevice, ctx->h2d));
      dio.out_gain_ramp = b->d_oramp;
    }
    // ... or their parameter segments (272 bytes per frame), expanded on the device by run_pipeline
    for (int e = 0; e <= kp.n_elements; ++e) {
      const iamfb_gain_ramp *sg = e == kp.n_elements ? io.out_gain_segs : io.gain_segs[e];
      if (!sg) continue;
      if (!b->d_segs[e] && cudaMalloc((void **)&b->d_segs[e], sizeof(iamfb_gain_ramp) * S * b->Fmax) != cudaSuccess)
        return fail(IAMFB_ERR_ALLOC_FAIL, "gain segment staging");
      CU(cudaMemcpyAsync(b->d_segs[e] + s_lo * F, sg + s_lo * F, sizeof(iamfb_gain_ramp) * cnt * F, cudaMemcpyHostToDevice, ctx->h2d));
      if (e == kp.n_elements) dio.out_gain_segs = b->d_segs[e];
      else dio.gain_segs[e] = b->d_segs[e];
    }
    CU(cudaMemcpyAsync(b->d_params + s_lo * F, io.params + s_lo * F, sizeof(iamfb_frame_params) * cnt * F, cudaMemcpyHostToDevice, ctx->h2d));
    for (int e = 0; e < kp.n_elements; ++e) {
      const size_t per = (size_t)F * p->in_rows[e] * N;
      if (s16)
        CU(cudaMemcpyAsync(b->d_in16[e] + s_lo * per, (const int16_t *)io.in[e] + s_lo * per, sizeof(int16_t) * cnt * per, cudaMemcpyHostToDevice, ctx->h2d));
      else
        CU(cudaMemcpyAsync(b->d_in[e] + s_lo * per, io.in[e] + s_lo * per, sizeof(float) * cnt * per, cudaMemcpyHostToDevice, ctx->h2d));
    }
    CU(cudaEventRecord(ctx->ev_up[c], ctx->h2d));
    CU(cudaStreamWaitEvent(st, ctx->ev_up[c], 0));
    r = run_pipeline(b, &dio, F, false, b->d_pcm, b->d_counts, stride, (int)s_lo, (int)cnt);
    if (r) return r;
    CU(cudaEventRecord(ctx->ev_done[c], st));
    CU(cudaStreamWaitEvent(ctx->d2h, ctx->ev_done[c], 0));
    CU(cudaMemcpyAsync((char *)io.pcm + s_lo * stride, b->d_pcm + s_lo * stride, cnt * stride, cudaMemcpyDeviceToHost, ctx->d2h));
    if (io.out_counts)
      CU(cudaMemcpyAsync(io.out_counts + s_lo * F, b->d_counts + s_lo * F, sizeof(int32_t) * cnt * F, cudaMemcpyDeviceToHost, ctx->d2h));
    if (hooks) {
      CU(cudaEventRecord(ctx->ev_back[c], ctx->d2h));
      // two groups stay in flight behind the one being filled next
      if (c >= 2) { r = drain_to(c - 1); if (r) return r; }
    }
  }
  if (hooks) { r = drain_to(n_chunks); if (r) return r; }
  CU(cudaStreamSynchronize(ctx->d2h));
  CU(cudaStreamSynchronize(st));
  return IAMFB_OK;
}

extern "C" int iamfb_batch_submit_host(iamfb_batch *b, const iamfb_io *io, int F) { return submit_host(b, io, F, nullptr); }
extern "C" int iamfb_batch_submit_host_hooks(iamfb_batch *b, const iamfb_io *io, int F, const iamfb_chunk_hooks *hooks) {
  if (!hooks || !hooks->fill || !hooks->drain) return fail(IAMFB_ERR_BAD_ARG, "submit: null hooks");
  return submit_host(b, io, F, hooks);
}

extern "C" int iamfb_batch_flush_host(iamfb_batch *b, void *pcm, int32_t *counts) {
  if (!b || !pcm) return fail(IAMFB_ERR_BAD_ARG, "flush: null argument");
  iamfb_plan *p = b->plan;
  CU(cudaSetDevice(p->ctx->device));
  int r = ensure_staging(b, 1);
  if (r) return r;
  cudaStream_t st = p->ctx->stream;
  const size_t stride = iamfb_plan_out_stride_bytes(p, 1);
  r = iamfb_batch_flush_device(b, b->d_pcm, b->d_counts);
  if (r) return r;
  CU(cudaMemcpyAsync(pcm, b->d_pcm, (size_t)b->S * stride, cudaMemcpyDeviceToHost, st));
  if (counts) CU(cudaMemcpyAsync(counts, b->d_counts, sizeof(int32_t) * b->S, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return IAMFB_OK;
}
