// iamfb_hrtf.cu - host side of the binaural (HRTF) front end: which elements take it, their Toeplitz tables, the
// per-batch buffers, and the launches per submit: k_hrtf_index once, then per element k_hrtf_gemm - its RAW variant straight
// from the decoded int16 frames when they reach the renderer untouched, else k_hrtf_prep (gains, projection, float32
// input -> Q20 limb planes) in front of it.
//
// A plan with HRTF elements is the front end below followed by an ordinary plan in which every such element has become a
// 2-channel pass-through element (the rendered [2][N] frame takes the place of the decoded frame, exactly where the
// reference's iamf_stream_render hands sout to the rest of iamf_decoder_internal_decode, IAMF_decoder.c:2565-2612):
// trimming, mix gains, mixing, resampling, loudness, limiter and quantisation are the same kernels as for any other plan.
#include "iamfb_internal.h"

#include <cstdlib>

#include "iamfb_hrtf.cuh"
#include "iamfb_hrir.inc"

using namespace iamfb;

extern "C" int iamfb_get_hrir(int kind, int index, int16_t *taps) {
  if (!taps) return IAMFB_ERR_BAD_ARG;
  if (kind == IAMFB_EL_CHANNEL) {
    if (index < 1 || index >= IAMFB_CH_COUNT) return IAMFB_ERR_BAD_ARG;
    memcpy(taps, k_hrir_q15 + (size_t)index * 2 * k_hrir_taps, sizeof(int16_t) * 2 * k_hrir_taps);
  } else {
    if (index < 0 || index >= IAMFB_MAX_SCENE_CH) return IAMFB_ERR_BAD_ARG;
    memcpy(taps, k_hrir_q15 + k_hrir_amb_off + (size_t)index * 2 * k_hrir_taps, sizeof(int16_t) * 2 * k_hrir_taps);
  }
  return IAMFB_OK;
}
extern "C" int iamfb_hrir_taps(void) { return k_hrir_taps; }

struct HrtfEl {
  bool on;
  int C, n_in, mode, proj_cols;
  int row[IAMFB_MAX_SCENE_CH];
  float gain[IAMFB_MAX_SCENE_CH];
  bool plain;               // rows pass through untouched: 16-bit submits travel as two limbs
  bool demix;               // the layout's channels are de-mixed first (scalable layers / recon gain): the renderer reads them
                            // from a float32 buffer the caller of iamfb_hrtf_run provides
  float proj[IAMFB_MAX_SCENE_CH * IAMFB_MAX_SCENE_CH];
  uint8_t *d_tab;
};
struct iamfb_hrtf_front {
  int n_elements, N;
  HrtfEl el[kMaxEl];
};
struct iamfb_hrtf_batch {
  int S, Fmax, NB, NT, NBP;
  uint8_t *d_planes[kMaxEl];
  int *d_hist[kMaxEl][2];
  float *d_bin[kMaxEl];
  int *d_np;
  short *d_fos, *d_sof;
  unsigned seq;
};

void iamfb_hrtf_front_destroy(iamfb_hrtf_front *h) {
  if (!h) return;
  for (int e = 0; e < kMaxEl; ++e) cudaFree(h->el[e].d_tab);
  delete h;
}

// Builds the front end for the elements of `d` that ask for binaural HRTF rendering toward the binaural target and rewrites
// those elements of `back` (a copy of d) as 2-channel pass-through elements.  *out stays null when no element takes it.
int iamfb_hrtf_front_create(const iamfb_plan_desc *d, iamfb_plan_desc *back, iamfb_hrtf_front **out) {
  *out = nullptr;
  bool any = false;
  for (int e = 0; e < d->n_elements && e < kMaxEl; ++e)
    if (d->el[e].binaural_hrtf && d->target == IAMFB_TARGET_BINAURAL && !(d->el[e].kind == IAMFB_EL_CHANNEL && d->el[e].layout == IAMFB_LAYOUT_BINAURAL))
      any = true;
  if (!any) return IAMFB_OK;
  if (d->in_rate != 48000) return fail(IAMFB_ERR_UNIMPLEMENTED, "binaural HRTF rendering: the HRIR set is sampled at 48 kHz (stream rate %d)", d->in_rate);
  if (d->frame_size % 16 != 0) return fail(IAMFB_ERR_UNIMPLEMENTED, "binaural HRTF rendering needs frames of a multiple of 16 samples (%d)", d->frame_size);
  iamfb_hrtf_front *h = new iamfb_hrtf_front();
  memset(h, 0, sizeof(*h));
  h->n_elements = d->n_elements;
  h->N = d->frame_size;
  for (int e = 0; e < d->n_elements; ++e) {
    const iamfb_element_desc &de = d->el[e];
    HrtfEl &he = h->el[e];
    if (!de.binaural_hrtf || (de.kind == IAMFB_EL_CHANNEL && de.layout == IAMFB_LAYOUT_BINAURAL)) continue;
    he.on = true;
    he.n_in = de.n_in;
    he.plain = true;
    std::vector<int16_t> taps;
    if (de.kind == IAMFB_EL_CHANNEL) {
      int32_t chs[IAMFB_MAX_LAYOUT_CH];
      const int n = iamfb_layout_channels(de.layout, chs);
      if (n <= 0) { iamfb_hrtf_front_destroy(h); return fail(IAMFB_ERR_BAD_ARG, "binaural HRTF rendering: layout %d", de.layout); }
      if (de.n_in < 1 || de.n_in > IAMFB_MAX_LAYOUT_CH) { iamfb_hrtf_front_destroy(h); return fail(IAMFB_ERR_BAD_ARG, "element %d: n_in %d", e, de.n_in); }
      he.C = n;
      he.mode = 0;
      taps.resize((size_t)n * 2 * k_hrir_taps);
      for (int m = 0; m < n; ++m) {
        int row = -1;
        for (int r = 0; r < de.n_in; ++r)
          if (de.chs_in[r] == chs[m]) row = r;
        // a channel the layers do not carry is derived by the de-mixer (demixer.c:127-378) from per-frame parameters
        if (row < 0 || de.recon_present) he.demix = true;
        he.row[m] = row;
        he.gain[m] = 1.0f;
        for (int g = 0; g < de.n_out_gain && g < IAMFB_MAX_LAYOUT_CH; ++g)
          if (de.out_gain_ch[g] == chs[m]) { he.gain[m] = de.out_gain[g]; he.plain = false; }
        iamfb_get_hrir(IAMFB_EL_CHANNEL, chs[m], &taps[(size_t)m * 2 * k_hrir_taps]);
      }
      if (he.demix) {   // rows = the layout's channels in order, gains already applied by the de-mixer
        for (int m = 0; m < n; ++m) { he.row[m] = m; he.gain[m] = 1.0f; }
        he.plain = false;
      }
    } else {
      const int n = de.ambi_channels;
      if (!(n == 1 || n == 4 || n == 9 || n == 16) || de.n_in < 1 || de.n_in > IAMFB_MAX_SCENE_CH) {
        iamfb_hrtf_front_destroy(h);
        return fail(IAMFB_ERR_BAD_ARG, "binaural HRTF rendering: %d ambisonics channels", n);
      }
      he.C = n;
      taps.resize((size_t)n * 2 * k_hrir_taps);
      for (int m = 0; m < n; ++m) {
        iamfb_get_hrir(IAMFB_EL_SCENE, m, &taps[(size_t)m * 2 * k_hrir_taps]);
        he.gain[m] = 1.0f;
      }
      if (de.ambi_mode == 0) {
        he.mode = 0;
        for (int m = 0; m < n; ++m) {
          if (de.ambi_map[m] >= de.n_in) { iamfb_hrtf_front_destroy(h); return fail(IAMFB_ERR_BAD_ARG, "element %d: ambisonics mapping", e); }
          he.row[m] = de.ambi_map[m];
        }
      } else {
        he.mode = 1;
        he.plain = false;
        he.proj_cols = de.ambi_cols;
        if (de.ambi_cols < 1 || de.ambi_cols > IAMFB_MAX_SCENE_CH || de.ambi_cols > de.n_in) {
          iamfb_hrtf_front_destroy(h);
          return fail(IAMFB_ERR_BAD_ARG, "element %d: projection columns", e);
        }
        for (int l = 0; l < de.ambi_cols; ++l)
          for (int m = 0; m < n; ++m) he.proj[l * n + m] = de.ambi_matrix[l * n + m];
      }
    }
    std::vector<uint8_t> tab((size_t)he.C * kHrHLimbs * kHrTabBytes);
    for (int c = 0; c < he.C; ++c) hrtf_build_table(&taps[(size_t)c * 2 * k_hrir_taps], &tab[(size_t)c * kHrHLimbs * kHrTabBytes]);
    if (cudaMalloc((void **)&he.d_tab, tab.size()) != cudaSuccess ||
        cudaMemcpy(he.d_tab, tab.data(), tab.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      iamfb_hrtf_front_destroy(h);
      return fail(IAMFB_ERR_CUDA, "binaural HRTF rendering: table upload failed");
    }
    // what the rest of the pipeline sees: the rendered [2][N] frame as a 2-channel element passed straight through
    iamfb_element_desc &be = back->el[e];
    memset(&be, 0, sizeof(be));
    be.kind = IAMFB_EL_CHANNEL;
    be.n_in = 2;
    be.layout = IAMFB_LAYOUT_STEREO;
    be.chs_in[0] = IAMFB_CH_L2;
    be.chs_in[1] = IAMFB_CH_R2;
    be.default_mode = be.default_w_idx = -1;
    be.first_layer_layout = IAMFB_LAYOUT_STEREO;
  }
  *out = h;
  return IAMFB_OK;
}

bool iamfb_hrtf_needs_demix(const iamfb_hrtf_front *h, int e) { return h && h->el[e].on && h->el[e].demix; }

void iamfb_hrtf_batch_destroy(iamfb_hrtf_batch *b) {
  if (!b) return;
  for (int e = 0; e < kMaxEl; ++e) {
    cudaFree(b->d_planes[e]); cudaFree(b->d_hist[e][0]); cudaFree(b->d_hist[e][1]); cudaFree(b->d_bin[e]);
  }
  cudaFree(b->d_np); cudaFree(b->d_fos); cudaFree(b->d_sof);
  delete b;
}

int iamfb_hrtf_batch_reset(const iamfb_hrtf_front *h, iamfb_hrtf_batch *b, cudaStream_t st) {
  for (int e = 0; e < h->n_elements; ++e)
    if (h->el[e].on)
      for (int k = 0; k < 2; ++k) CU(cudaMemsetAsync(b->d_hist[e][k], 0, sizeof(int) * (size_t)b->S * h->el[e].C * kHrHist, st));
  b->seq = 0;
  return IAMFB_OK;
}

int iamfb_hrtf_batch_create(const iamfb_hrtf_front *h, int S, int Fmax, iamfb_hrtf_batch **out) {
  iamfb_hrtf_batch *b = new iamfb_hrtf_batch();
  memset(b, 0, sizeof(*b));
  b->S = S;
  b->Fmax = Fmax;
  const int T = Fmax * h->N;
  int nb = (T + kHrBlock - 1) / kHrBlock;
  nb = (nb + 63) & ~63;   // (the epilogue splits a tile into four column groups of whole 16-column pieces)
  if (nb > kHrMaxNB) nb = kHrMaxNB;
  b->NB = nb;
  b->NT = (T + nb * kHrBlock - 1) / (nb * kHrBlock);
  b->NBP = 4 + b->NT * nb;
  cudaError_t er = cudaSuccess;
  auto alloc = [&](void **p, size_t bytes) { if (er == cudaSuccess) er = cudaMalloc(p, bytes ? bytes : 16); };
  for (int e = 0; e < h->n_elements; ++e) {
    if (!h->el[e].on) continue;
    const size_t C = h->el[e].C;
    alloc((void **)&b->d_hist[e][0], sizeof(int) * (size_t)S * C * kHrHist);
    alloc((void **)&b->d_hist[e][1], sizeof(int) * (size_t)S * C * kHrHist);
    alloc((void **)&b->d_bin[e], sizeof(float) * (size_t)S * Fmax * 2 * h->N);
  }
  alloc((void **)&b->d_np, sizeof(int) * S);
  alloc((void **)&b->d_fos, sizeof(short) * (size_t)S * Fmax);
  alloc((void **)&b->d_sof, sizeof(short) * (size_t)S * Fmax);
  if (er != cudaSuccess) {
    iamfb_hrtf_batch_destroy(b);
    return fail(IAMFB_ERR_ALLOC_FAIL, "binaural HRTF rendering: device allocation failed: %s", cudaGetErrorString(er));
  }
  *out = b;
  return IAMFB_OK;
}

#define HR_LAUNCH_CHECK(name)                                                                                   \
  do {                                                                                                          \
    cudaError_t e_ = cudaGetLastError();                                                                        \
    if (e_ != cudaSuccess) return fail(IAMFB_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
    ++ctx->launches;                                                                                            \
  } while (0)

// Renders the HRTF elements of the streams [s_lo, s_lo + s_cnt) for a submit of F frames.  io holds DEVICE pointers with the
// caller's shapes; out_io receives what the rest of the pipeline takes (the binaural frames in place of those elements).
int iamfb_hrtf_run(iamfb_ctx *ctx, const iamfb_hrtf_front *h, iamfb_hrtf_batch *b, const iamfb_io *io, int F, int s_lo, int s_cnt,
                   iamfb_io *out_io, const float *const *demixed) {
  cudaStream_t st = ctx->stream;
  const int N = h->N;
  const bool s16_io = io->in_format == IAMFB_IN_S16;
  *out_io = *io;
  out_io->in_format = IAMFB_IN_F32;
  for (int e = 0; e < h->n_elements; ++e)
    if (!h->el[e].on && s16_io)
      return fail(IAMFB_ERR_UNIMPLEMENTED, "a mix of HRTF-rendered and matrix-rendered elements takes float32 input");
  // tiles for THIS submit's length (the planes are sized for Fmax)
  const int T = F * N;
  int nb = (T + kHrBlock - 1) / kHrBlock;
  nb = (nb + 63) & ~63;   // (the epilogue splits a tile into four column groups of whole 16-column pieces)
  if (nb > b->NB) nb = b->NB;
  const int nt = (T + nb * kHrBlock - 1) / (nb * kHrBlock);
  const int nbp = b->NBP;
  {
    HrtfIndexArgs ia;
    ia.params = io->params + (size_t)s_lo * F;
    ia.n_present = b->d_np + s_lo;
    ia.frame_of_slot = b->d_fos + (size_t)s_lo * F;
    ia.slot_of_frame = b->d_sof + (size_t)s_lo * F;
    ia.S = s_cnt; ia.F = F; ia.N = N;
    { ScopedKernelTimer tm_(ctx, "k_hrtf_index"); k_hrtf_index<<<(s_cnt + 127) / 128, 128, 0, st>>>(ia); }
    HR_LAUNCH_CHECK("k_hrtf_index");
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  // (stream groups of a host-resident submit run one after the other on the same stream: the history parity advances once
  // per submit, with the first group)
  if (s_lo == 0) ++b->seq;
  const int par = (int)(b->seq & 1u);
  for (int e = 0; e < h->n_elements; ++e) {
    const HrtfEl &he = h->el[e];
    if (!he.on) continue;
    const int C = he.C;
    const bool dm = he.demix;
    if (dm && !(demixed && demixed[e])) return fail(IAMFB_ERR_INTERNAL, "binaural HRTF rendering: element %d arrives without its de-mixed channels", e);
    const bool s16 = s16_io && !dm;
    const int n_in = dm ? C : he.n_in;
    const void *in_e = dm ? (const void *)demixed[e] : (const void *)io->in[e];
    const int NL = (s16 && he.plain) ? 2 : 3;
    const size_t in_per = (size_t)F * n_in * N;
    // 16-bit PCM that reaches the renderer untouched: the contraction kernel makes its limb rows itself (no prep pass)
    const bool raw = s16 && he.plain;
    if (!raw && !b->d_planes[e]) {   // limb planes: only submits that need the prep pass pay for them
      if (cudaMalloc((void **)&b->d_planes[e], (size_t)b->S * C * kHrMaxXLimbs * 4 * nbp * 16) != cudaSuccess)
        return fail(IAMFB_ERR_ALLOC_FAIL, "binaural HRTF rendering: limb planes");
    }
    uint8_t *planes = raw ? nullptr : b->d_planes[e] + (size_t)s_lo * C * NL * 4 * nbp * 16;
    if (!raw) {
      HrtfPrepArgs pa;
      memset(&pa, 0, sizeof(pa));
      pa.in = s16 ? (const void *)(reinterpret_cast<const int16_t *>(in_e) + (size_t)s_lo * in_per)
                  : (const void *)(reinterpret_cast<const float *>(in_e) + (size_t)s_lo * in_per);
      pa.planes = planes;
      pa.hist_in = b->d_hist[e][par ^ 1] + (size_t)s_lo * C * kHrHist;
      pa.hist_out = b->d_hist[e][par] + (size_t)s_lo * C * kHrHist;
      pa.n_present = b->d_np + s_lo;
      pa.frame_of_slot = b->d_fos + (size_t)s_lo * F;
      pa.C = C; pa.n_in = n_in; pa.NL = NL; pa.NBP = nbp; pa.F = F; pa.N = N;
      pa.mode = he.mode;
      for (int m = 0; m < C; ++m) { pa.row[m] = he.row[m]; pa.gain[m] = he.gain[m]; }
      pa.proj_cols = he.proj_cols;
      memcpy(pa.proj, he.proj, sizeof(pa.proj));
      const int groups = (kHrHist + T) / 16;
      dim3 grid((groups + 255) / 256, s_cnt * C);
      {
        ScopedKernelTimer tm_(ctx, "k_hrtf_prep");
        if (s16) k_hrtf_prep<true><<<grid, 256, 0, st>>>(pa);
        else k_hrtf_prep<false><<<grid, 256, 0, st>>>(pa);
      }
      HR_LAUNCH_CHECK("k_hrtf_prep");
    }
    {
      HrtfGemmArgs ga;
      ga.tab = he.d_tab;
      ga.planes = planes;
      ga.out = b->d_bin[e] + (size_t)s_lo * F * 2 * N;
      ga.n_present = b->d_np + s_lo;
      ga.frame_of_slot = b->d_fos + (size_t)s_lo * F;
      ga.S = s_cnt; ga.C = C; ga.NL = NL; ga.NB = nb; ga.NT = nt; ga.NBP = nbp; ga.F = F; ga.N = N;
      ga.x_shift = NL == 2 ? 15 : 20;
      ga.raw_in = reinterpret_cast<const int16_t *>(in_e) + (size_t)s_lo * in_per;
      ga.n_in = n_in;
      for (int m = 0; m < C; ++m) ga.row[m] = he.row[m];
      ga.hist_in = b->d_hist[e][par ^ 1] + (size_t)s_lo * C * kHrHist;
      ga.hist_out = b->d_hist[e][par] + (size_t)s_lo * C * kHrHist;
      const int smem = kHrStages * hrtf_stage_bytes(nb, NL) + (raw ? 2 * hrtf_raw_bytes(nb) : 0);
      const int tiles = s_cnt * nt, grid = tiles < sms ? tiles : sms;
      if (raw) {
        CU(cudaFuncSetAttribute(k_hrtf_gemm<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ScopedKernelTimer tm_(ctx, "k_hrtf_gemm");
        k_hrtf_gemm<2, true><<<grid, kHrThreads + kHrConvThreads, smem, st>>>(ga);
      } else if (NL == 2) {
        CU(cudaFuncSetAttribute(k_hrtf_gemm<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ScopedKernelTimer tm_(ctx, "k_hrtf_gemm");
        k_hrtf_gemm<2, false><<<grid, kHrThreads, smem, st>>>(ga);
      } else {
        CU(cudaFuncSetAttribute(k_hrtf_gemm<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ScopedKernelTimer tm_(ctx, "k_hrtf_gemm");
        k_hrtf_gemm<3, false><<<grid, kHrThreads, smem, st>>>(ga);
      }
      HR_LAUNCH_CHECK("k_hrtf_gemm");
    }
    out_io->in[e] = b->d_bin[e];
  }
  return IAMFB_OK;
}
