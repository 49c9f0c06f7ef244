// CLIPSeg-ensemble fusion step downstream of the UNet logits, fused into two kernels.
//
// Reference (all on the host / ATen, one pass per alpha and image -- 100 alphas x N images x ~8 tensor ops + a numpy/cv2 round trip):
//   eval_CLIPseg.py:885-888   clip_up = F.interpolate(clip_logits, size=unet.shape[2:], mode='bilinear', align_corners=False)
//   eval_CLIPseg.py:656-724   search_best_alpha: for alpha in linspace(0.1, 10, 100): pred = argmax(clip_up + alpha*unet);
//                             cv2.resize(pred, label size, INTER_NEAREST); ConfusionMatrix.update(label, pred); mIoU
//   eval_CLIPseg.py:726-748   ConfusionMatrix.update / compute
//   eval_CLIPseg.py:901-912, predict_CLIPseg.py:519-526   final mask = uint8(argmax(clip_up + best_alpha*unet)) resized INTER_NEAREST
// Here ONE kernel per image evaluates every alpha for every label pixel (bilinear CLIP sample computed once, reused by all
// alphas) and accumulates the [n_alpha][C][C] confusion counts; a one-block kernel turns them into mIoU per alpha and the first
// maximiser.  Arithmetic mirrors oracle/ensemble_oracle.py operation by operation (explicit round-to-nearest mul/add, no FMA
// contraction), so the confusion counts are bit-identical to the oracle's.
#include "common.cuh"

constexpr int ENS_MAXC = 4;

struct EnsGeom { int hc, wc, H, W, Ho, Wo, C; };

// cv2 resizeNN: source index of destination index x  (fx = n_dst / n_src in double, ifx = 1 / fx)
__device__ __forceinline__ int nn_index(int x, int n_src, int n_dst) {
  const double fx = (double)n_dst / (double)n_src, ifx = 1.0 / fx;
  const int s = (int)floor((double)x * ifx);
  return s < n_src - 1 ? s : n_src - 1;
}
struct Lerp { int i0, i1; float l0, l1; };
// ATen area_pixel_compute_source_index, align_corners = False, float32
__device__ __forceinline__ Lerp lerp_axis(int dst, int n_in, int n_out) {
  const float scale = (float)n_in / (float)n_out;
  float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  if (src < 0.f) src = 0.f;
  Lerp r;
  r.i0 = (int)floorf(src); if (r.i0 > n_in - 1) r.i0 = n_in - 1;
  r.i1 = r.i0 + 1 < n_in ? r.i0 + 1 : n_in - 1;
  r.l1 = __fsub_rn(src, (float)r.i0); r.l0 = __fsub_rn(1.f, r.l1);
  return r;
}
// bilinear sample of clip[c] (NCHW plane) at pred pixel (sy, sx), and the UNet logit, for every class
__device__ __forceinline__ void sample_logits(const float* __restrict__ clip, const float* __restrict__ unet, const EnsGeom& g, int sy, int sx,
                                              float (&cu)[ENS_MAXC], float (&un)[ENS_MAXC]) {
  const Lerp ly = lerp_axis(sy, g.hc, g.H), lx = lerp_axis(sx, g.wc, g.W);
#pragma unroll
  for (int c = 0; c < ENS_MAXC; ++c) {
    if (c < g.C) {
      const float* p = clip + (size_t)c * g.hc * g.wc;
      const float top = __fadd_rn(__fmul_rn(p[(size_t)ly.i0 * g.wc + lx.i0], lx.l0), __fmul_rn(p[(size_t)ly.i0 * g.wc + lx.i1], lx.l1));
      const float bot = __fadd_rn(__fmul_rn(p[(size_t)ly.i1 * g.wc + lx.i0], lx.l0), __fmul_rn(p[(size_t)ly.i1 * g.wc + lx.i1], lx.l1));
      cu[c] = __fadd_rn(__fmul_rn(top, ly.l0), __fmul_rn(bot, ly.l1));
      un[c] = unet[((size_t)c * g.H + sy) * g.W + sx];
    }
  }
}
__device__ __forceinline__ int fused_argmax(const float (&cu)[ENS_MAXC], const float (&un)[ENS_MAXC], float alpha, int C) {
  int best = 0; float bv = __fadd_rn(cu[0], __fmul_rn(alpha, un[0]));
#pragma unroll
  for (int c = 1; c < ENS_MAXC; ++c)
    if (c < C) { const float v = __fadd_rn(cu[c], __fmul_rn(alpha, un[c])); if (v > bv) { bv = v; best = c; } }   // ties -> lowest class (torch.argmax)
  return best;
}

// confusion[a][label][pred] += ...   for one image; label pixels at (Ho, Wo) = label size
__global__ void __launch_bounds__(256) k_ens_confusion(const float* __restrict__ clip, const float* __restrict__ unet, const unsigned char* __restrict__ label,
                                                       EnsGeom g, const double* __restrict__ alphas, int n_alpha, unsigned long long* __restrict__ conf) { egm_pdl_enter();
  extern __shared__ unsigned int s_cnt[];                 // [n_alpha][C*C]
  const int CC = g.C * g.C;
  for (int i = threadIdx.x; i < n_alpha * CC; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  const long long total = (long long)g.Ho * g.Wo;
  const int lane = threadIdx.x & 31;
  // whole warps iterate together (ballots below): pad the trip count to a multiple of the warp size
  const long long padded = (total + 31) / 32 * 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < padded; i += (long long)gridDim.x * blockDim.x) {
    const bool in = i < total;
    int lab = 255; float cu[ENS_MAXC], un[ENS_MAXC];
    if (in) {
      const int yl = (int)(i / g.Wo), xl = (int)(i - (long long)yl * g.Wo);
      const int sy = (g.Ho == g.H) ? yl : nn_index(yl, g.H, g.Ho), sx = (g.Wo == g.W) ? xl : nn_index(xl, g.W, g.Wo);
      lab = label[i];
      sample_logits(clip, unet, g, sy, sx, cu, un);
    }
    const bool counted = in && lab < g.C;
    // warp-aggregated counting: label masks once per pixel batch, one ballot per class and alpha; lane k < C*C owns cell k
    unsigned labMask = 0;                                   // lanes whose label equals this lane's row (lane / C)
#pragma unroll
    for (int c = 0; c < ENS_MAXC; ++c) {
      const unsigned m = __ballot_sync(0xffffffffu, counted && lab == c);
      if (lane / g.C == c) labMask = m;
    }
    if (__ballot_sync(0xffffffffu, counted) == 0) continue;
    for (int a = 0; a < n_alpha; ++a) {
      const int pred = counted ? fused_argmax(cu, un, (float)alphas[a], g.C) : -1;
      unsigned predMask = 0;                                // lanes whose prediction equals this lane's column (lane % C)
#pragma unroll
      for (int c = 0; c < ENS_MAXC; ++c) {
        if (c < g.C) { const unsigned m = __ballot_sync(0xffffffffu, pred == c); if (lane % g.C == c) predMask = m; }
      }
      const unsigned hit = labMask & predMask;
      if (lane < CC && hit) atomicAdd(&s_cnt[a * CC + lane], (unsigned)__popc(hit));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_alpha * CC; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(conf + i, (unsigned long long)s_cnt[i]);
}

// mIoU per alpha (float32, ConfusionMatrix.compute) and the FIRST alpha whose mIoU is strictly larger than everything before it
__global__ void k_ens_best(const unsigned long long* __restrict__ conf, const double* __restrict__ alphas, int n_alpha, int C, float* __restrict__ miou,
                           double* __restrict__ best) { egm_pdl_enter();
  for (int a = threadIdx.x; a < n_alpha; a += blockDim.x) {
    const unsigned long long* m = conf + (size_t)a * C * C;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) {
      float row = 0.f, col = 0.f;
      for (int k = 0; k < C; ++k) { row += (float)m[c * C + k]; col += (float)m[k * C + c]; }
      const float diag = (float)m[c * C + c];
      float denom = row + col - diag;
      if (denom == 0.f) denom = 1.f;
      acc += diag / denom;
    }
    miou[a] = acc / (float)C;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ba = 0.0; float bm = 0.f;                       // search_best_alpha starts from alpha 0.0 / mIoU 0.0
    for (int a = 0; a < n_alpha; ++a)
      if (miou[a] > bm) { bm = miou[a]; ba = alphas[a]; }
    best[0] = ba; best[1] = (double)bm;
  }
}

// final mask at (Ho, Wo): uint8(argmax(clip_up + alpha*unet)) resized INTER_NEAREST
__global__ void k_ens_predict(const float* __restrict__ clip, const float* __restrict__ unet, EnsGeom g, float alpha, unsigned char* __restrict__ out) { egm_pdl_enter();
  const long long total = (long long)g.Ho * g.Wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int yo = (int)(i / g.Wo), xo = (int)(i - (long long)yo * g.Wo);
    const int sy = (g.Ho == g.H) ? yo : nn_index(yo, g.H, g.Ho), sx = (g.Wo == g.W) ? xo : nn_index(xo, g.W, g.Wo);
    float cu[ENS_MAXC], un[ENS_MAXC];
    sample_logits(clip, unet, g, sy, sx, cu, un);
    out[i] = (unsigned char)fused_argmax(cu, un, alpha, g.C);
  }
}

extern "C" int egm_ensemble_confusion(const float* clip_logits, int hc, int wc, const float* unet_logits, int H, int W, const unsigned char* label, int Hl,
                                      int Wl, int num_classes, const double* alphas, int n_alpha, unsigned long long* confusion, void* stream) {
  EGM_REQUIRE(num_classes >= 2 && num_classes <= ENS_MAXC, EGM_E_SHAPE, "ensemble: 2..%d classes", ENS_MAXC);
  EGM_REQUIRE(n_alpha >= 1 && n_alpha <= 1024, EGM_E_SHAPE, "ensemble: 1..1024 alphas");
  EGM_REQUIRE(hc > 0 && wc > 0 && H > 0 && W > 0, EGM_E_SHAPE, "ensemble: empty logits");
  if ((long long)Hl * Wl == 0) return EGM_OK;
  EnsGeom g{hc, wc, H, W, Hl, Wl, num_classes};
  const size_t sm = (size_t)n_alpha * num_classes * num_classes * sizeof(unsigned int);
  egm_launch(k_ens_confusion, egm_grid_for((long long)Hl * Wl, 256, 4), 256, sm, (cudaStream_t)stream, clip_logits, unet_logits, label, g, alphas, n_alpha, confusion);
  EGM_LAUNCH_CHECK("ensemble_confusion"); return EGM_OK;
}
extern "C" int egm_ensemble_best_alpha(const unsigned long long* confusion, const double* alphas, int n_alpha, int num_classes, float* miou, double* best,
                                       void* stream) {
  EGM_REQUIRE(num_classes >= 2 && num_classes <= ENS_MAXC && n_alpha >= 1, EGM_E_SHAPE, "ensemble_best_alpha: bad shape");
  egm_launch(k_ens_best, 1, 128, 0, (cudaStream_t)stream, confusion, alphas, n_alpha, num_classes, miou, best);
  EGM_LAUNCH_CHECK("ensemble_best_alpha"); return EGM_OK;
}
extern "C" int egm_ensemble_predict(const float* clip_logits, int hc, int wc, const float* unet_logits, int H, int W, int num_classes, float alpha,
                                    unsigned char* mask, int Ho, int Wo, void* stream) {
  EGM_REQUIRE(num_classes >= 2 && num_classes <= ENS_MAXC, EGM_E_SHAPE, "ensemble: 2..%d classes", ENS_MAXC);
  if ((long long)Ho * Wo == 0) return EGM_OK;
  EnsGeom g{hc, wc, H, W, Ho, Wo, num_classes};
  egm_launch(k_ens_predict, egm_grid_for((long long)Ho * Wo, 256, 8), 256, 0, (cudaStream_t)stream, clip_logits, unet_logits, g, alpha, mask);
  EGM_LAUNCH_CHECK("ensemble_predict"); return EGM_OK;
}
