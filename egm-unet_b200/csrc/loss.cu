// Fused loss: criterion(dice=True) of train_utils/train_and_eval.py:7-19 =
//   CE(weight, ignore_index) + multiclass Dice + laplace_loss + lap_loss + sobel_loss
// (train_utils/dice_coefficient_loss.py:22-108), forward and backward in two passes over the
// fp32 NCHW logits, plus the fused SGD(momentum, weight-decay) update (train.py:113-118).
// Quirks reproduced: lap/sobel use ONLY sample 0's target, raw 255s included,
// broadcast against every sample; Dice is averaged per (sample, class) over non-ignored pixels.
#include "common.cuh"

#define EGM_MAXC 16
// accumulator layout (double): [0] ce_num [1] ce_den [2] lap4 [3] lap8 [4] sobel [5] out-of-range labels, then per (n,c): inter, psum, tsum
#define ACC_HDR 8

__device__ __forceinline__ float tgt0(const long long* t0, int H, int W, int h, int w) {
  return (h >= 0 && h < H && w >= 0 && w < W) ? (float)t0[(long long)h * W + w] : 0.f;
}
__device__ __forceinline__ float x0at(const float* x0, int H, int W, int h, int w) {
  return (h >= 0 && h < H && w >= 0 && w < W) ? x0[(long long)h * W + w] : 0.f;
}
__device__ __forceinline__ int sgn(float v) { return (v > 0.f) - (v < 0.f); }

// the four stencil responses at (h, w): lap4(x0), lap8(x0)-lap8(t0), sobx(x0)-sobx(t0), soby(x0)-soby(t0)
__device__ __forceinline__ void stencils(const float* x0, const long long* t0, int H, int W, int h, int w, float (&f)[4]) {
  float a[3][3], b[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < 3; ++s) { a[r][s] = x0at(x0, H, W, h + r - 1, w + s - 1); b[r][s] = tgt0(t0, H, W, h + r - 1, w + s - 1); }
  f[0] = a[0][1] + a[1][0] + a[1][2] + a[2][1] - 4.f * a[1][1];
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < 3; ++s) { sa += a[r][s]; sb += b[r][s]; }
  f[1] = (9.f * a[1][1] - sa) - (9.f * b[1][1] - sb);
  float sxa = (a[0][0] - a[0][2]) + 2.f * (a[1][0] - a[1][2]) + (a[2][0] - a[2][2]);
  float sxb = (b[0][0] - b[0][2]) + 2.f * (b[1][0] - b[1][2]) + (b[2][0] - b[2][2]);
  float sya = (a[0][0] + 2.f * a[0][1] + a[0][2]) - (a[2][0] + 2.f * a[2][1] + a[2][2]);
  float syb = (b[0][0] + 2.f * b[0][1] + b[0][2]) - (b[2][0] + 2.f * b[2][1] + b[2][2]);
  f[2] = sxa - sxb; f[3] = sya - syb;
}

__global__ void __launch_bounds__(256) k_loss_pass1(const float* __restrict__ logits, const long long* __restrict__ target, const float* __restrict__ weight,
                                                    int N, int C, int H, int W, int ignore_index, double* __restrict__ acc, unsigned char* __restrict__ smap) { egm_pdl_enter();
  __shared__ float red[32];
  const int n = blockIdx.y;
  const long long HW = (long long)H * W;
  const float* lg = logits + (long long)n * C * HW;
  const long long* tg = target + (long long)n * HW;
  float ce_num = 0.f, ce_den = 0.f, s4 = 0.f, s8 = 0.f, ss = 0.f, nbad = 0.f;
  float inter[EGM_MAXC], psum[EGM_MAXC], tsum[EGM_MAXC];
#pragma unroll
  for (int c = 0; c < EGM_MAXC; ++c) { inter[c] = 0.f; psum[c] = 0.f; tsum[c] = 0.f; }
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    const int h = (int)((unsigned)p / (unsigned)W), w = (int)((unsigned)p - (unsigned)h * (unsigned)W);   // H*W < 2^31 (checked by the launcher)
    long long t = tg[p];
    float z[EGM_MAXC], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < EGM_MAXC; ++c) if (c < C) { z[c] = lg[(long long)c * HW + p]; mx = fmaxf(mx, z[c]); }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < EGM_MAXC; ++c) if (c < C) { z[c] = expf(z[c] - mx); se += z[c]; }
    // a label that is neither ignore_index nor a class id makes the reference raise (F.cross_entropy / one_hot device assert); here such
    // pixels are dropped from BOTH the CE and the Dice sums, never index class_weight, and are counted in loss_out[6]
    const bool bad = t != ignore_index && (t < 0 || t >= C);
    nbad += bad ? 1.f : 0.f;
    if (t != ignore_index && !bad) {
      float inv = 1.f / se;
#pragma unroll
      for (int c = 0; c < EGM_MAXC; ++c) if (c < C) {
        float pc = z[c] * inv; psum[c] += pc;
        if (t == c) { inter[c] += pc; tsum[c] += 1.f; float wt = weight ? weight[c] : 1.f; ce_num += -wt * logf(fmaxf(pc, 1e-38f)); ce_den += wt; }
      }
    }
    float f[4]; stencils(lg, target, H, W, h, w, f);     // channel 0 of sample n vs target of sample 0
    s4 += fabsf(f[0]); s8 += fabsf(f[1]); ss += fabsf(f[2]) + fabsf(f[3]);
    smap[(long long)n * HW + p] = (unsigned char)((sgn(f[0]) + 1) | ((sgn(f[1]) + 1) << 2) | ((sgn(f[2]) + 1) << 4) | ((sgn(f[3]) + 1) << 6));
  }
  float v;
  v = block_sum(ce_num, red); if (threadIdx.x == 0) atomicAdd(acc + 0, (double)v);
  v = block_sum(ce_den, red); if (threadIdx.x == 0) atomicAdd(acc + 1, (double)v);
  v = block_sum(s4, red); if (threadIdx.x == 0) atomicAdd(acc + 2, (double)v);
  v = block_sum(s8, red); if (threadIdx.x == 0) atomicAdd(acc + 3, (double)v);
  v = block_sum(ss, red); if (threadIdx.x == 0) atomicAdd(acc + 4, (double)v);
  v = block_sum(nbad, red); if (threadIdx.x == 0 && v != 0.f) atomicAdd(acc + 5, (double)v);
#pragma unroll
  for (int c = 0; c < EGM_MAXC; ++c) if (c < C) {
    double* a = acc + ACC_HDR + ((long long)n * C + c) * 3;
    v = block_sum(inter[c], red); if (threadIdx.x == 0) atomicAdd(a + 0, (double)v);
    v = block_sum(psum[c], red); if (threadIdx.x == 0) atomicAdd(a + 1, (double)v);
    v = block_sum(tsum[c], red); if (threadIdx.x == 0) atomicAdd(a + 2, (double)v);
  }
}

// out[0] = total, out[1..5] = ce, dice, laplace, lap, sobel, out[6] = number of out-of-range labels (0 for valid input)
__global__ void k_loss_finalize(const double* __restrict__ acc, int N, int C, double NHW, int with_dice, float* __restrict__ out) { egm_pdl_enter();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double eps = 1e-6;
  double ce = acc[0] / acc[1];
  double d = 0.0;
  for (int i = 0; i < N * C; ++i) {
    const double* a = acc + ACC_HDR + (long long)i * 3;
    double inter = a[0], sets = a[1] + a[2];
    if (sets == 0.0) sets = 2.0 * inter;
    d += (2.0 * inter + eps) / (sets + eps);
  }
  double dice = 1.0 - d / (double)(N * C);
  double l4 = acc[2] / NHW, l8 = acc[3] / NHW, sb = acc[4] / NHW;
  out[0] = with_dice ? (float)(ce + dice + l4 + l8 + sb) : (float)ce;
  out[1] = (float)ce; out[2] = (float)dice; out[3] = (float)l4; out[4] = (float)l8; out[5] = (float)sb; out[6] = (float)acc[5];
}

__global__ void __launch_bounds__(256) k_loss_pass2(const float* __restrict__ logits, const long long* __restrict__ target, const float* __restrict__ weight,
                                                    int N, int C, int H, int W, int ignore_index, const double* __restrict__ acc,
                                                    const unsigned char* __restrict__ smap, float gscale, int with_dice, float* __restrict__ dlogits) { egm_pdl_enter();
  const int n = blockIdx.y;
  const long long HW = (long long)H * W;
  const float* lg = logits + (long long)n * C * HW;
  const long long* tg = target + (long long)n * HW;
  const unsigned char* sm = smap + (long long)n * HW;
  float* dl = dlogits + (long long)n * C * HW;
  const float inv_den = (float)(1.0 / acc[1]);
  const float inv_nhw = (float)(1.0 / ((double)N * (double)HW));
  const float eps = 1e-6f;
  float dA[EGM_MAXC], dB[EGM_MAXC];   // d(dice_nc)/dp = dA*[t==c] - dB   (for valid pixels)
#pragma unroll
  for (int c = 0; c < EGM_MAXC; ++c) if (c < C) {
    const double* a = acc + ACC_HDR + ((long long)n * C + c) * 3;
    double inter = a[0], sets = a[1] + a[2];
    if (sets == 0.0) { dA[c] = 0.f; dB[c] = 0.f; }
    else { double den = sets + eps; dA[c] = (float)(2.0 / den); dB[c] = (float)((2.0 * inter + eps) / (den * den)); }
  }
  const float dice_w = with_dice ? -1.f / (float)(N * C) : 0.f;
  // transposed stencil taps: contribution of the response at p = q - (r-1, s-1) to x0[q] is k[r][s]
  const float K4[3][3] = {{0, 1, 0}, {1, -4, 1}, {0, 1, 0}};
  const float K8[3][3] = {{-1, -1, -1}, {-1, 8, -1}, {-1, -1, -1}};
  const float KX[3][3] = {{1, 0, -1}, {2, 0, -2}, {1, 0, -1}};
  const float KY[3][3] = {{1, 2, 1}, {0, 0, 0}, {-1, -2, -1}};
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    const int h = (int)((unsigned)p / (unsigned)W), w = (int)((unsigned)p - (unsigned)h * (unsigned)W);   // H*W < 2^31 (checked by the launcher)
    long long t = tg[p];
    float z[EGM_MAXC], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < EGM_MAXC; ++c) if (c < C) { z[c] = lg[(long long)c * HW + p]; mx = fmaxf(mx, z[c]); }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < EGM_MAXC; ++c) if (c < C) { z[c] = expf(z[c] - mx); se += z[c]; }
    float g[EGM_MAXC];
    if (t != ignore_index && t >= 0 && t < C) {          // same validity rule as pass 1
      float inv = 1.f / se, dot = 0.f, wt = weight ? weight[(int)t] : 1.f;
      float dp[EGM_MAXC];
#pragma unroll
      for (int c = 0; c < EGM_MAXC; ++c) if (c < C) { z[c] *= inv; dp[c] = dice_w * ((t == c ? dA[c] : 0.f) - dB[c]); dot += z[c] * dp[c]; }
#pragma unroll
      for (int c = 0; c < EGM_MAXC; ++c) if (c < C) g[c] = z[c] * (dp[c] - dot) + wt * inv_den * (z[c] - (t == c ? 1.f : 0.f));
    } else {
#pragma unroll
      for (int c = 0; c < EGM_MAXC; ++c) if (c < C) g[c] = 0.f;
    }
    float st = 0.f;
    unsigned char bb[3][3];                            // unconditional (clamped) loads first, masks applied afterwards
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int ph = h - (r - 1), pw = w - (s - 1);
        ph = ph < 0 ? 0 : (ph >= H ? H - 1 : ph); pw = pw < 0 ? 0 : (pw >= W ? W - 1 : pw);
        bb[r][s] = sm[(long long)ph * W + pw];
      }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int ph = h - (r - 1), pw = w - (s - 1);
        if (ph < 0 || ph >= H || pw < 0 || pw >= W) continue;
        unsigned char b = bb[r][s];
        st += K4[r][s] * (float)((int)(b & 3) - 1) + K8[r][s] * (float)((int)((b >> 2) & 3) - 1)
            + KX[r][s] * (float)((int)((b >> 4) & 3) - 1) + KY[r][s] * (float)((int)((b >> 6) & 3) - 1);
      }
    if (with_dice) g[0] += st * inv_nhw;
#pragma unroll
    for (int c = 0; c < EGM_MAXC; ++c) if (c < C) dl[(long long)c * HW + p] = gscale * g[c];
  }
}

extern "C" long long egm_loss_workspace_bytes(int N, int C, int H, int W) {
  long long accb = (long long)(ACC_HDR + 3LL * N * C) * sizeof(double);
  accb = (accb + 255) / 256 * 256;
  return accb + (long long)N * H * W;
}
// ------------------------------------------------------------------ shared-memory tiled passes (the ones launched)
// The flat passes above fetch the 3x3 neighbourhoods of x0 (channel-0 logits), of sample 0's int64 target and of the sign map through L1
// -- 18 + 9 scattered loads per pixel, 0.45 ms at cfg2 for 88 MB of algorithmic traffic.  Here a block walks 64x4 pixel tiles: the
// (66 x 6) halo of x0 / target-0 (pass 1) or of the sign map (pass 2) is staged in shared memory once, everything else is one coalesced
// load per pixel.  Same arithmetic, same accumulator layout.
constexpr int LT_W = 64, LT_H = 4, LT_HW = LT_W + 2, LT_HH = LT_H + 2;
// The per-class register arrays are sized by the template parameter (2 / 4 / 16 classes): with the fixed 16 the passes needed 125-140
// registers per thread and ran at 12-24 % occupancy (ncu, profiles/membound_r2_ncu.txt) for a 2-class problem.
#pragma push_macro("EGM_MAXC")
#undef EGM_MAXC
#define EGM_MAXC MAXC

template <int MAXC>
__global__ void __launch_bounds__(256) k_loss_pass1_t(const float* __restrict__ logits, const long long* __restrict__ target, const float* __restrict__ weight,
                                                      int N, int C, int H, int W, int ignore_index, double* __restrict__ acc, unsigned char* __restrict__ smap) { egm_pdl_enter();
  __shared__ float red[32];
  __shared__ float sx[LT_HH][LT_HW], st0[LT_HH][LT_HW];
  const int n = blockIdx.z;
  const long long HW = (long long)H * W;
  const float* lg = logits + (long long)n * C * HW;
  const long long* tg = target + (long long)n * HW;
  const int tx = threadIdx.x & (LT_W - 1), ty = threadIdx.x / LT_W;
  const int w0 = blockIdx.x * LT_W;
  float ce_num = 0.f, ce_den = 0.f, s4 = 0.f, s8 = 0.f, ss = 0.f, nbad = 0.f;
  float inter[EGM_MAXC], psum[EGM_MAXC], tsum[EGM_MAXC];
#pragma unroll
  for (int c = 0; c < EGM_MAXC; ++c) { inter[c] = 0.f; psum[c] = 0.f; tsum[c] = 0.f; }
  const int tilesY = (H + LT_H - 1) / LT_H;
  for (int tyb = blockIdx.y; tyb < tilesY; tyb += gridDim.y) {
    const int h0 = tyb * LT_H;
    __syncthreads();                                   // previous tile fully consumed
    for (int i = threadIdx.x; i < LT_HH * LT_HW; i += 256) {
      const int r = i / LT_HW, cidx = i - r * LT_HW, hh = h0 - 1 + r, ww = w0 - 1 + cidx;
      const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
      sx[r][cidx] = ok ? lg[(long long)hh * W + ww] : 0.f;                       // channel 0 of sample n
      st0[r][cidx] = ok ? (float)target[(long long)hh * W + ww] : 0.f;           // target of sample 0 (reference quirk), raw 255s included
    }
    __syncthreads();
    const int h = h0 + ty, w = w0 + tx;
    if (h < H && w < W) {
      const long long p = (long long)h * W + w;
      const long long t = tg[p];
      float z[EGM_MAXC], mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < EGM_MAXC; ++c) if (c < C) { z[c] = c == 0 ? sx[ty + 1][tx + 1] : lg[(long long)c * HW + p]; mx = fmaxf(mx, z[c]); }
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < EGM_MAXC; ++c) if (c < C) { z[c] = expf(z[c] - mx); se += z[c]; }
      const bool bad = t != ignore_index && (t < 0 || t >= C);
      nbad += bad ? 1.f : 0.f;
      if (t != ignore_index && !bad) {
        float inv = 1.f / se;
#pragma unroll
        for (int c = 0; c < EGM_MAXC; ++c) if (c < C) {
          float pc = z[c] * inv; psum[c] += pc;
          if (t == c) { inter[c] += pc; tsum[c] += 1.f; float wt = weight ? weight[c] : 1.f; ce_num += -wt * logf(fmaxf(pc, 1e-38f)); ce_den += wt; }
        }
      }
      float a[3][3], b[3][3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) { a[r][s] = sx[ty + r][tx + s]; b[r][s] = st0[ty + r][tx + s]; }
      float f[4];
      f[0] = a[0][1] + a[1][0] + a[1][2] + a[2][1] - 4.f * a[1][1];
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) { sa += a[r][s]; sb += b[r][s]; }
      f[1] = (9.f * a[1][1] - sa) - (9.f * b[1][1] - sb);
      const float sxa = (a[0][0] - a[0][2]) + 2.f * (a[1][0] - a[1][2]) + (a[2][0] - a[2][2]);
      const float sxb = (b[0][0] - b[0][2]) + 2.f * (b[1][0] - b[1][2]) + (b[2][0] - b[2][2]);
      const float sya = (a[0][0] + 2.f * a[0][1] + a[0][2]) - (a[2][0] + 2.f * a[2][1] + a[2][2]);
      const float syb = (b[0][0] + 2.f * b[0][1] + b[0][2]) - (b[2][0] + 2.f * b[2][1] + b[2][2]);
      f[2] = sxa - sxb; f[3] = sya - syb;
      s4 += fabsf(f[0]); s8 += fabsf(f[1]); ss += fabsf(f[2]) + fabsf(f[3]);
      smap[(long long)n * HW + p] = (unsigned char)((sgn(f[0]) + 1) | ((sgn(f[1]) + 1) << 2) | ((sgn(f[2]) + 1) << 4) | ((sgn(f[3]) + 1) << 6));
    }
  }
  float v;
  v = block_sum(ce_num, red); if (threadIdx.x == 0) atomicAdd(acc + 0, (double)v);
  v = block_sum(ce_den, red); if (threadIdx.x == 0) atomicAdd(acc + 1, (double)v);
  v = block_sum(s4, red); if (threadIdx.x == 0) atomicAdd(acc + 2, (double)v);
  v = block_sum(s8, red); if (threadIdx.x == 0) atomicAdd(acc + 3, (double)v);
  v = block_sum(ss, red); if (threadIdx.x == 0) atomicAdd(acc + 4, (double)v);
  v = block_sum(nbad, red); if (threadIdx.x == 0 && v != 0.f) atomicAdd(acc + 5, (double)v);
#pragma unroll
  for (int c = 0; c < EGM_MAXC; ++c) if (c < C) {
    double* a = acc + ACC_HDR + ((long long)n * C + c) * 3;
    v = block_sum(inter[c], red); if (threadIdx.x == 0) atomicAdd(a + 0, (double)v);
    v = block_sum(psum[c], red); if (threadIdx.x == 0) atomicAdd(a + 1, (double)v);
    v = block_sum(tsum[c], red); if (threadIdx.x == 0) atomicAdd(a + 2, (double)v);
  }
}

template <int MAXC>
__global__ void __launch_bounds__(256) k_loss_pass2_t(const float* __restrict__ logits, const long long* __restrict__ target, const float* __restrict__ weight,
                                                      int N, int C, int H, int W, int ignore_index, const double* __restrict__ acc,
                                                      const unsigned char* __restrict__ smap, float gscale, int with_dice, float* __restrict__ dlogits) { egm_pdl_enter();
  __shared__ unsigned char sb[LT_HH][LT_HW + 2];
  const int n = blockIdx.z;
  const long long HW = (long long)H * W;
  const float* lg = logits + (long long)n * C * HW;
  const long long* tg = target + (long long)n * HW;
  const unsigned char* sm = smap + (long long)n * HW;
  float* dl = dlogits + (long long)n * C * HW;
  const int tx = threadIdx.x & (LT_W - 1), ty = threadIdx.x / LT_W;
  const int w0 = blockIdx.x * LT_W;
  const float inv_den = (float)(1.0 / acc[1]);
  const float inv_nhw = (float)(1.0 / ((double)N * (double)HW));
  const float eps = 1e-6f;
  float dA[EGM_MAXC], dB[EGM_MAXC];   // d(dice_nc)/dp = dA*[t==c] - dB   (for valid pixels)
#pragma unroll
  for (int c = 0; c < EGM_MAXC; ++c) if (c < C) {
    const double* a = acc + ACC_HDR + ((long long)n * C + c) * 3;
    double inter = a[0], sets = a[1] + a[2];
    if (sets == 0.0) { dA[c] = 0.f; dB[c] = 0.f; }
    else { double den = sets + eps; dA[c] = (float)(2.0 / den); dB[c] = (float)((2.0 * inter + eps) / (den * den)); }
  }
  const float dice_w = with_dice ? -1.f / (float)(N * C) : 0.f;
  // transposed stencil taps: contribution of the response at p = q - (r-1, s-1) to x0[q] is k[r][s]
  const float K4[3][3] = {{0, 1, 0}, {1, -4, 1}, {0, 1, 0}};
  const float K8[3][3] = {{-1, -1, -1}, {-1, 8, -1}, {-1, -1, -1}};
  const float KX[3][3] = {{1, 0, -1}, {2, 0, -2}, {1, 0, -1}};
  const float KY[3][3] = {{1, 2, 1}, {0, 0, 0}, {-1, -2, -1}};
  const int tilesY = (H + LT_H - 1) / LT_H;
  for (int tyb = blockIdx.y; tyb < tilesY; tyb += gridDim.y) {
    const int h0 = tyb * LT_H;
    __syncthreads();
    for (int i = threadIdx.x; i < LT_HH * LT_HW; i += 256) {
      const int r = i / LT_HW, cidx = i - r * LT_HW, hh = h0 - 1 + r, ww = w0 - 1 + cidx;
      // code 0x55 = all four signs zero: what an out-of-image response contributes
      sb[r][cidx] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? sm[(long long)hh * W + ww] : (unsigned char)0x55;
    }
    __syncthreads();
    const int h = h0 + ty, w = w0 + tx;
    if (h < H && w < W) {
      const long long p = (long long)h * W + w;
      const long long t = tg[p];
      float z[EGM_MAXC], mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < EGM_MAXC; ++c) if (c < C) { z[c] = lg[(long long)c * HW + p]; mx = fmaxf(mx, z[c]); }
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < EGM_MAXC; ++c) if (c < C) { z[c] = expf(z[c] - mx); se += z[c]; }
      float g[EGM_MAXC];
      if (t != ignore_index && t >= 0 && t < C) {          // same validity rule as pass 1
        float inv = 1.f / se, dot = 0.f, wt = weight ? weight[(int)t] : 1.f;
        float dp[EGM_MAXC];
#pragma unroll
        for (int c = 0; c < EGM_MAXC; ++c) if (c < C) { z[c] *= inv; dp[c] = dice_w * ((t == c ? dA[c] : 0.f) - dB[c]); dot += z[c] * dp[c]; }
#pragma unroll
        for (int c = 0; c < EGM_MAXC; ++c) if (c < C) g[c] = z[c] * (dp[c] - dot) + wt * inv_den * (z[c] - (t == c ? 1.f : 0.f));
      } else {
#pragma unroll
        for (int c = 0; c < EGM_MAXC; ++c) if (c < C) g[c] = 0.f;
      }
      float st = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const unsigned char b = sb[ty + 2 - r][tx + 2 - s];     // response at (h - (r-1), w - (s-1))
          st += K4[r][s] * (float)((int)(b & 3) - 1) + K8[r][s] * (float)((int)((b >> 2) & 3) - 1)
              + KX[r][s] * (float)((int)((b >> 4) & 3) - 1) + KY[r][s] * (float)((int)((b >> 6) & 3) - 1);
        }
      if (with_dice) g[0] += st * inv_nhw;
#pragma unroll
      for (int c = 0; c < EGM_MAXC; ++c) if (c < C) dl[(long long)c * HW + p] = gscale * g[c];
    }
  }
}

#pragma pop_macro("EGM_MAXC")

// logits fp32 NCHW, target int64 [N,H,W]; loss_out[8] (total, ce, dice, laplace, lap, sobel, #out-of-range labels, unused); dlogits may be null (forward only).
extern "C" int egm_loss_fwd_bwd(const float* logits, const long long* target, const float* class_weight, int N, int C, int H, int W, int ignore_index,
                                int with_dice, float grad_scale, float* loss_out, float* dlogits, void* workspace, long long workspace_bytes, void* stream) {
  EGM_REQUIRE(C >= 1 && C <= EGM_MAXC, EGM_E_SHAPE, "loss: num_classes %d > %d", C, EGM_MAXC);
  EGM_REQUIRE(N >= 1 && N <= 65535, EGM_E_SHAPE, "loss: batch %d", N);
  EGM_REQUIRE((long long)H * W < (1ll << 31), EGM_E_SHAPE, "loss: H*W must be < 2^31");
  EGM_REQUIRE(workspace_bytes >= egm_loss_workspace_bytes(N, C, H, W), EGM_E_BADARG, "loss: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  long long accb = (long long)(ACC_HDR + 3LL * N * C) * sizeof(double);
  accb = (accb + 255) / 256 * 256;
  double* acc = (double*)workspace;
  unsigned char* smap = (unsigned char*)workspace + accb;
  cudaMemsetAsync(acc, 0, (size_t)accb, st);
  long long HW = (long long)H * W;
  static int flat = -1;
  if (flat < 0) { const char* e = getenv("EGM_LOSS_FLAT"); flat = (e && e[0] == '1') ? 1 : 0; }
  if (flat) {                                            // round-1 kernels (diagnostic switch)
    int bx = (int)((HW + 255) / 256); int cap = egm_num_sms() * 8 / N + 1; if (bx > cap) bx = cap; if (bx < 1) bx = 1;
    dim3 grid(bx, N);
    egm_launch(k_loss_pass1, grid, 256, 0, st, logits, target, class_weight, N, C, H, W, ignore_index, acc, smap);
    egm_launch(k_loss_finalize, 1, 32, 0, st, acc, N, C, (double)N * (double)HW, with_dice, loss_out);
    if (dlogits) egm_launch(k_loss_pass2, grid, 256, 0, st, logits, target, class_weight, N, C, H, W, ignore_index, acc, smap, grad_scale, with_dice, dlogits);
    EGM_LAUNCH_CHECK("loss_fwd_bwd"); return EGM_OK;
  }
  const int tilesX = (W + LT_W - 1) / LT_W, tilesY = (H + LT_H - 1) / LT_H;
  long long gy = (long long)egm_num_sms() * 8 / ((long long)N * tilesX); if (gy < 1) gy = 1; if (gy > tilesY) gy = tilesY;
  dim3 grid(tilesX, (unsigned)gy, N);
#define EGM_LOSS_LAUNCH(MC)                                                                                                                    \
  {                                                                                                                                            \
    egm_launch(k_loss_pass1_t<MC>, grid, 256, 0, st, logits, target, class_weight, N, C, H, W, ignore_index, acc, smap);                               \
    egm_launch(k_loss_finalize, 1, 32, 0, st, acc, N, C, (double)N * (double)HW, with_dice, loss_out);                                                 \
    if (dlogits) egm_launch(k_loss_pass2_t<MC>, grid, 256, 0, st, logits, target, class_weight, N, C, H, W, ignore_index, acc, smap, grad_scale, with_dice, dlogits); \
  }
  if (C <= 2) EGM_LOSS_LAUNCH(2) else if (C <= 4) EGM_LOSS_LAUNCH(4) else EGM_LOSS_LAUNCH(16)
#undef EGM_LOSS_LAUNCH
  EGM_LAUNCH_CHECK("loss_fwd_bwd"); return EGM_OK;
}

// ------------------------------------------------------------------ fused SGD (torch.optim.SGD semantics, dampening 0, no nesterov)
// hp (device or host-mapped): hp[0]=lr hp[1]=momentum hp[2]=weight_decay hp[3]=grad_scale (e.g. 1/world_size)
__global__ void k_sgd(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n, const float* __restrict__ hp) { egm_pdl_enter();
  const float lr = hp[0], mom = hp[1], wd = hp[2], gs = hp[3];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float w = p[i];
    float d = fmaf(wd, w, g[i] * gs);
    float b = fmaf(mom, buf[i], d);
    buf[i] = b;
    p[i] = w - lr * b;
  }
}
extern "C" int egm_sgd_step(float* params, const float* grads, float* momentum_buf, long long n, const float* hyper_dev, void* stream) {
  if (n == 0) return EGM_OK;
  egm_launch(k_sgd, egm_grid_for(n, 256), 256, 0, (cudaStream_t)stream, params, grads, momentum_buf, n, hyper_dev);
  EGM_LAUNCH_CHECK("sgd_step"); return EGM_OK;
}

// ------------------------------------------------------------------ eval metrics: argmax + confusion matrix + per-sample dice sums
// mat[n_cls*n_cls] int64 += bincount(n_cls*t + argmax) over valid (0 <= t < n_cls) pixels (distributed_utils.py:81-91)
// dice_acc[N][C][3] double += (inter, pred_sum, tgt_sum) over t != ignore_index pixels of one-hot argmax vs one-hot target (:135-144)
__global__ void k_eval_metrics(const float* __restrict__ logits, const long long* __restrict__ target, int N, int C, long long HW, int ignore_index,
                               unsigned long long* __restrict__ mat, double* __restrict__ dice_acc) { egm_pdl_enter();
  extern __shared__ unsigned int sh[];     // [C*C] confusion + [C*3] dice
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < C * C + C * 3; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const float* lg = logits + (long long)n * C * HW;
  const long long* tg = target + (long long)n * HW;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    float best = lg[p]; int am = 0;
    for (int c = 1; c < C; ++c) { float v = lg[(long long)c * HW + p]; if (v > best) { best = v; am = c; } }
    long long t = tg[p];
    if (t >= 0 && t < C) atomicAdd(&sh[(int)t * C + am], 1u);
    if (t != ignore_index) {
      atomicAdd(&sh[C * C + am * 3 + 1], 1u);
      if (t >= 0 && t < C) { atomicAdd(&sh[C * C + (int)t * 3 + 2], 1u); if (t == am) atomicAdd(&sh[C * C + am * 3 + 0], 1u); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) if (sh[i]) atomicAdd(mat + i, (unsigned long long)sh[i]);
  for (int i = threadIdx.x; i < C * 3; i += blockDim.x) if (sh[C * C + i]) atomicAdd(dice_acc + (long long)n * C * 3 + i, (double)sh[C * C + i]);
}
extern "C" int egm_eval_metrics(const float* logits, const long long* target, int N, int C, int H, int W, int ignore_index, long long* confmat,
                                double* dice_acc, void* stream) {
  EGM_REQUIRE(C >= 1 && C <= 64 && N >= 1 && N <= 65535, EGM_E_SHAPE, "eval_metrics: bad N/C");
  long long HW = (long long)H * W;
  if (HW == 0) return EGM_OK;
  int bx = (int)((HW + 255) / 256); int cap = egm_num_sms() * 8 / N + 1; if (bx > cap) bx = cap;
  egm_launch(k_eval_metrics, dim3(bx, N), 256, (C * C + C * 3) * sizeof(unsigned int), (cudaStream_t)stream, logits, target, N, C, HW, ignore_index,
                                                                                                    (unsigned long long*)confmat, dice_acc);
  EGM_LAUNCH_CHECK("eval_metrics"); return EGM_OK;
}
