// CUDA-core implicit-GEMM convolution (forward / dgrad via repacked weights / wgrad), NHWC,
// stride 1, "same" padding (pad = dil*(k-1)/2), any kernel size, dilation and group count.
// This is the fp32 check-mode path for every conv and the bf16 path for the small / odd-channel
// convs of the EGM blocks; the DoubleConv 3x3 layers run on the
// tcgen05 kernel in conv_tc.cu when dtype == bf16.
//
// Weight layouts (fp32, produced by egm_pack_conv_weight from the reference's [Cout][Cin_g][kh][kw]):
//   wf [taps][Cin_g][Cout]   forward:  y[m,co] = sum_{t,ci} x[m+off(t), g*Cin_g+ci] * wf[t][ci][co]
//   wd [taps][Cout_g][Cin]   dgrad  :  the same kernel run on dy with flipped taps and swapped roles
#include "common.cuh"

struct ConvGeom {
  int N, H, W, Cin_g, Cout_g, groups, kh, kw, dil, pad;
  long long xcs, xco, ycs, yco;
};

template <typename T, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) k_conv_tiled(const T* __restrict__ x, const float* __restrict__ wf, const float* __restrict__ bias,
                                                    T* __restrict__ y, ConvGeom g, int accumulate) { egm_pdl_enter();
  constexpr int BK = 16;
  constexpr int NJ = BM / 16;
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int grp = blockIdx.z;
  const long long M = (long long)g.N * g.H * g.W;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Ktot = g.kh * g.kw * g.Cin_g;
  const int Cout = g.Cout_g * g.groups;

  // A-load mapping: kk = tid % 16, rows mrow + 16*j
  const int kk = tid & 15, mrow = tid >> 4;
  int hj[NJ], wj[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    long long m = m0 + mrow + 16 * j;
    if (m < M) { long long r = m % ((long long)g.H * g.W); hj[j] = (int)(r / g.W); wj[j] = (int)(r - (long long)hj[j] * g.W); }
    else { hj[j] = -1000000; wj[j] = 0; }
  }
  const int tm = tid / (BN / TN), tn = tid % (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const T* xg = x + g.xco + (long long)grp * g.Cin_g;
  for (int k0 = 0; k0 < Ktot; k0 += BK) {
    {  // A tile
      int k = k0 + kk;
      bool kv = k < Ktot;
      int tap = kv ? k / g.Cin_g : 0, ci = k - tap * g.Cin_g;
      int r = tap / g.kw, s = tap - r * g.kw;
      int dh = r * g.dil - g.pad, dw = s * g.dil - g.pad;
      long long doff = (long long)dh * g.W + dw;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        int hh = hj[j] + dh, ww = wj[j] + dw;
        float v = 0.f;
        if (kv && hh >= 0 && hh < g.H && ww >= 0 && ww < g.W) v = ldf(xg + (m0 + mrow + 16 * j + doff) * g.xcs + ci);
        As[kk][mrow + 16 * j] = v;
      }
    }
    for (int idx = tid; idx < BK * BN; idx += 256) {  // B tile
      int k2 = idx / BN, nn = idx - k2 * BN;
      int k = k0 + k2, col = n0 + nn;
      Bs[k2][nn] = (k < Ktot && col < g.Cout_g) ? wf[(long long)k * Cout + grp * g.Cout_g + col] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < BK; ++q) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[q][tm * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[q][tn * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    long long m = m0 + tm * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int col = n0 + tn * TN + j;
      if (col >= g.Cout_g) continue;
      int co = grp * g.Cout_g + col;
      T* yp = y + m * g.ycs + g.yco + co;
      float v = acc[i][j] + (bias ? bias[co] : 0.f);
      if (accumulate) v += ldf(yp);
      stf(yp, v);
    }
  }
}

template <typename T>
static int launch_conv_tiled(const T* x, const float* wf, const float* bias, T* y, const ConvGeom& g, int accumulate, cudaStream_t st) {
  long long M = (long long)g.N * g.H * g.W;
  if (M == 0) return EGM_OK;
  int cg = g.Cout_g;
  if (cg > 32) {
    dim3 grid(cdiv(M, 64), cdiv(cg, 64), g.groups);
    egm_launch(k_conv_tiled<T, 64, 64, 4, 4>, grid, 256, 0, st, x, wf, bias, y, g, accumulate);
  } else if (cg > 16) {
    dim3 grid(cdiv(M, 64), 1, g.groups);
    egm_launch(k_conv_tiled<T, 64, 32, 4, 2>, grid, 256, 0, st, x, wf, bias, y, g, accumulate);
  } else if (cg > 8) {
    dim3 grid(cdiv(M, 128), 1, g.groups);
    egm_launch(k_conv_tiled<T, 128, 16, 4, 2>, grid, 256, 0, st, x, wf, bias, y, g, accumulate);
  } else {
    dim3 grid(cdiv(M, 128), 1, g.groups);
    egm_launch(k_conv_tiled<T, 128, 8, 4, 1>, grid, 256, 0, st, x, wf, bias, y, g, accumulate);
  }
  return egm_check_launch("conv_tiled");
}

extern "C" int egm_conv2d_direct(const void* x, long long x_cstride, long long x_coff, const float* w_packed, const float* bias, void* y,
                                 long long y_cstride, long long y_coff, int accumulate, int dtype, int N, int H, int W, int Cin, int Cout,
                                 int kh, int kw, int dil, int groups, void* stream) {
  EGM_REQUIRE(groups >= 1 && Cin % groups == 0 && Cout % groups == 0, EGM_E_SHAPE, "conv2d: bad groups %d for %d->%d", groups, Cin, Cout);
  EGM_REQUIRE((kh & 1) && (kw & 1) && kh == kw, EGM_E_SHAPE, "conv2d: only odd square kernels (got %dx%d)", kh, kw);
  EGM_REQUIRE(groups <= 65535, EGM_E_SHAPE, "conv2d: too many groups");
  ConvGeom g{N, H, W, Cin / groups, Cout / groups, groups, kh, kw, dil, dil * (kh - 1) / 2, x_cstride, x_coff, y_cstride, y_coff};
  EGM_DISPATCH_DTYPE(dtype, return launch_conv_tiled<T>((const T*)x, w_packed, bias, (T*)y, g, accumulate, (cudaStream_t)stream));
  return EGM_OK;
}

// ------------------------------------------------------------------ weight gradient
// dwf[t][ci][co] (+)= sum_m x[m+off(t), g*Cin_g+ci] * dy[m, g*Cout_g+co]; split over m with fp32 atomics.
template <typename T, int BN, int TN>
__global__ void __launch_bounds__(256) k_conv_wgrad_tiled(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dwf, ConvGeom g,
                                                          int ktiles, int splits, long long m_per_split) { egm_pdl_enter();
  constexpr int BKT = 64, BMC = 16, TK = 4;
  __shared__ float As[BMC][BKT + 1];
  __shared__ float Bs[BMC][BN];
  const int tid = threadIdx.x;
  const int grp = blockIdx.z;
  const int kt = blockIdx.x % ktiles, sp = blockIdx.x / ktiles;
  const int k0 = kt * BKT, n0 = blockIdx.y * BN;
  const long long M = (long long)g.N * g.H * g.W;
  const long long HW = (long long)g.H * g.W;
  const int Ktot = g.kh * g.kw * g.Cin_g;
  const int Cout = g.Cout_g * g.groups;
  long long mb = (long long)sp * m_per_split, me = mb + m_per_split; if (me > M) me = M;

  // A-load mapping: kk = tid % 64 (fixed k per thread), rows tid/64 + 4*j
  const int kk = tid & 63, mr = tid >> 6;
  const int k = k0 + kk;
  const bool kv = k < Ktot;
  const int tap = kv ? k / g.Cin_g : 0, ci = k - tap * g.Cin_g;
  const int r = tap / g.kw, s = tap - r * g.kw;
  const int dh = r * g.dil - g.pad, dw = s * g.dil - g.pad;
  const long long doff = (long long)dh * g.W + dw;
  const T* xg = x + g.xco + (long long)grp * g.Cin_g + ci;
  const T* dyg = dy + g.yco + (long long)grp * g.Cout_g;

  const int tk = tid / (BN / TN), tn = tid % (BN / TN);   // BN/TN == 16 -> tk in 0..15
  float acc[TK][TN];
#pragma unroll
  for (int i = 0; i < TK; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // (h, w) of this thread's four rows, advanced incrementally (no per-element division)
  int hq[4], wq[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    long long m = mb + mr + 4 * j;
    long long rr = m % HW; hq[j] = (int)(rr / g.W); wq[j] = (int)(rr - (long long)hq[j] * g.W);
  }
  for (long long mc = mb; mc < me; mc += BMC) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long m = mc + mr + 4 * j;
      float v = 0.f;
      if (kv && m < me) {
        int hh = hq[j] + dh, ww = wq[j] + dw;
        if (hh >= 0 && hh < g.H && ww >= 0 && ww < g.W) v = ldf(xg + (m + doff) * g.xcs);
      }
      As[mr + 4 * j][kk] = v;
      wq[j] += BMC;
      while (wq[j] >= g.W) { wq[j] -= g.W; hq[j] += 1; }
      while (hq[j] >= g.H) hq[j] -= g.H;
    }
    for (int idx = tid; idx < BMC * BN; idx += 256) {
      int mm = idx / BN, nn = idx - mm * BN;
      long long m = mc + mm; int col = n0 + nn;
      Bs[mm][nn] = (m < me && col < g.Cout_g) ? ldf(dyg + m * g.ycs + col) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < BMC; ++q) {
      float a[TK], b[TN];
#pragma unroll
      for (int i = 0; i < TK; ++i) a[i] = As[q][tk * TK + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[q][tn * TN + j];
#pragma unroll
      for (int i = 0; i < TK; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TK; ++i) {
    int kr = k0 + tk * TK + i;
    if (kr >= Ktot) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int col = n0 + tn * TN + j;
      if (col >= g.Cout_g) continue;
      atomicAdd(dwf + (long long)kr * Cout + grp * g.Cout_g + col, acc[i][j]);
    }
  }
}

// x: the conv's forward input; dy: gradient of its output. dwf must hold taps*Cin_g*Cout floats (zeroed here).
extern "C" int egm_conv2d_wgrad_direct(const void* x, long long x_cstride, long long x_coff, const void* dy, long long dy_cstride, long long dy_coff,
                                       float* dw_packed, int dtype, int N, int H, int W, int Cin, int Cout, int kh, int kw, int dil, int groups,
                                       void* stream) {
  EGM_REQUIRE(groups >= 1 && Cin % groups == 0 && Cout % groups == 0, EGM_E_SHAPE, "wgrad: bad groups");
  EGM_REQUIRE((kh & 1) && kh == kw, EGM_E_SHAPE, "wgrad: only odd square kernels");
  cudaStream_t st = (cudaStream_t)stream;
  ConvGeom g{N, H, W, Cin / groups, Cout / groups, groups, kh, kw, dil, dil * (kh - 1) / 2, x_cstride, x_coff, dy_cstride, dy_coff};
  long long M = (long long)N * H * W;
  int Ktot = kh * kw * g.Cin_g;
  cudaMemsetAsync(dw_packed, 0, sizeof(float) * (size_t)Ktot * Cout, st);
  if (M == 0) return EGM_OK;
  int ktiles = cdiv(Ktot, 64);
  int cg = g.Cout_g;
  int BN = cg > 32 ? 64 : (cg > 16 ? 32 : 16);
  int ntiles = cdiv(cg, BN);
  long long base = (long long)ktiles * ntiles * groups;
  long long want = ((long long)egm_num_sms() * 4 + base - 1) / base;
  long long maxs = (M + 255) / 256;
  int splits = (int)(want < 1 ? 1 : (want > maxs ? maxs : want));
  long long mps = ((M + splits - 1) / splits + 15) / 16 * 16;
  splits = (int)((M + mps - 1) / mps);
  dim3 grid(ktiles * splits, ntiles, groups);
  EGM_DISPATCH_DTYPE(dtype, {
    if (BN == 64) egm_launch(k_conv_wgrad_tiled<T, 64, 4>, grid, 256, 0, st, (const T*)x, (const T*)dy, dw_packed, g, ktiles, splits, mps);
    else if (BN == 32) egm_launch(k_conv_wgrad_tiled<T, 32, 2>, grid, 256, 0, st, (const T*)x, (const T*)dy, dw_packed, g, ktiles, splits, mps);
    else egm_launch(k_conv_wgrad_tiled<T, 16, 1>, grid, 256, 0, st, (const T*)x, (const T*)dy, dw_packed, g, ktiles, splits, mps);
  });
  EGM_LAUNCH_CHECK("conv_wgrad_tiled"); return EGM_OK;
}

// ------------------------------------------------------------------ weight packing / unpacking
// w [Cout][Cin_g][kh][kw] fp32 (reference layout) -> wf [taps][Cin_g][Cout], wd [taps][Cout_g][Cin] (flipped taps); either may be null.
__global__ void k_pack_w(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wd, int Cout, int Cin_g, int kh, int kw, int groups) { egm_pdl_enter();
  int taps = kh * kw, Cout_g = Cout / groups, Cin = Cin_g * groups;
  long long total = (long long)Cout * Cin_g * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i % taps); long long q = i / taps; int ci = (int)(q % Cin_g); int co = (int)(q / Cin_g);
    float v = w[i];
    if (wf) wf[((long long)t * Cin_g + ci) * Cout + co] = v;
    if (wd) { int gq = co / Cout_g, col = co - gq * Cout_g; int tf = taps - 1 - t;
      wd[((long long)tf * Cout_g + col) * Cin + gq * Cin_g + ci] = v; }
  }
}
extern "C" int egm_pack_conv_weight(const float* w, float* wf, float* wd, int Cout, int Cin_g, int kh, int kw, int groups, void* stream) {
  long long total = (long long)Cout * Cin_g * kh * kw;
  if (total == 0) return EGM_OK;
  egm_launch(k_pack_w, egm_grid_for(total, 256), 256, 0, (cudaStream_t)stream, w, wf, wd, Cout, Cin_g, kh, kw, groups);
  EGM_LAUNCH_CHECK("pack_conv_weight"); return EGM_OK;
}
// dwf [taps][Cin_g][Cout] -> dw [Cout][Cin_g][kh][kw]  (dw = beta*dw + dwf^T)
__global__ void k_unpack_dw(const float* __restrict__ dwf, float* __restrict__ dw, int Cout, int Cin_g, int taps, float beta) { egm_pdl_enter();
  long long total = (long long)Cout * Cin_g * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i % taps); long long q = i / taps; int ci = (int)(q % Cin_g); int co = (int)(q / Cin_g);
    float v = dwf[((long long)t * Cin_g + ci) * Cout + co];
    dw[i] = beta != 0.f ? beta * dw[i] + v : v;
  }
}
extern "C" int egm_unpack_conv_wgrad(const float* dw_packed, float* dw, int Cout, int Cin_g, int kh, int kw, float beta, void* stream) {
  long long total = (long long)Cout * Cin_g * kh * kw;
  if (total == 0) return EGM_OK;
  egm_launch(k_unpack_dw, egm_grid_for(total, 256), 256, 0, (cudaStream_t)stream, dw_packed, dw, Cout, Cin_g, kh * kw, beta);
  EGM_LAUNCH_CHECK("unpack_conv_wgrad"); return EGM_OK;
}

// Embed a [Co][Ci][ks][ks] kernel into the centre of a [Co][Ci][kb][kb] one (big (+)= small), or crop back (small = centre(big)).
__global__ void k_embed(float* big, float* small_, long long CoCi, int kb, int ks, int mode, int accumulate) { egm_pdl_enter();
  long long total = CoCi * ks * ks; int o = (kb - ks) / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int s = (int)(i % ks); long long q = i / ks; int r = (int)(q % ks); long long cc = q / ks;
    long long bi = (cc * kb + r + o) * kb + s + o;
    if (mode == 0) big[bi] = (accumulate ? big[bi] : 0.f) + small_[i];
    else small_[i] = big[bi];
  }
}
extern "C" int egm_kernel_embed(float* big, float* small_, long long CoCi, int kb, int ks, int mode, int accumulate, void* stream) {
  long long total = CoCi * ks * ks;
  if (total == 0) return EGM_OK;
  egm_launch(k_embed, egm_grid_for(total, 256), 256, 0, (cudaStream_t)stream, big, small_, CoCi, kb, ks, mode, accumulate);
  EGM_LAUNCH_CHECK("kernel_embed"); return EGM_OK;
}

// EdgeAwareFeatureEnhancer (src/EGM-UNet.py:872-886): conv1x1(x - AvgPool2d(3, 1, 1)(x)).  AvgPool2d with count_include_pad divides by 9
// everywhere, i.e. x - avg3(x) IS a zero-padded depthwise 3x3 filter k = delta - 1/9, so high-pass + 1x1 conv == ONE dense 3x3 conv with
// W3[co][ci][t] = W1[co][ci] * k[t].  That conv runs on the tcgen05 halo kernel (the smem-staged halo tile feeds the MMAs directly, BN
// statistics in its epilogue); the high-pass tensor and both of its HBM round trips (forward and transposed backward) disappear.
//   mode 0: w3 <- w1.  round_bf16: the off-centre weight is wn = bf16(-w1/9) and the centre is -8*wn (exact in bf16), so the bf16 operand
//           is EXACTLY W1' (x - avg3 x) with W1' = -9*wn within bf16 rounding of W1 -- a constant input still maps to exactly zero.
//   mode 1: dw1 <- sum_t k[t] * dw3[co][ci][t]   (gradient of the composition)
__device__ __forceinline__ float hp_tap(float w1, int t, int round_bf16) {
  float wn = -w1 * (1.f / 9.f);
  if (round_bf16) wn = __bfloat162float(__float2bfloat16_rn(wn));
  return t == 4 ? -8.f * wn : wn;
}
__global__ void k_highpass_compose(float* __restrict__ w1, float* __restrict__ w3, long long CoCi, int mode, int round_bf16) { egm_pdl_enter();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < CoCi; i += (long long)gridDim.x * blockDim.x) {
    if (mode == 0) {
      const float w = w1[i];
#pragma unroll
      for (int t = 0; t < 9; ++t) w3[i * 9 + t] = hp_tap(w, t, round_bf16);
    } else {
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) s += w3[i * 9 + t] * (t == 4 ? 8.f / 9.f : -1.f / 9.f);
      w1[i] = s;
    }
  }
}
extern "C" int egm_highpass_compose(float* w1, float* w3, long long CoCi, int mode, int round_bf16, void* stream) {
  if (CoCi == 0) return EGM_OK;
  egm_launch(k_highpass_compose, egm_grid_for(CoCi, 256), 256, 0, (cudaStream_t)stream, w1, w3, CoCi, mode, round_bf16);
  EGM_LAUNCH_CHECK("highpass_compose"); return EGM_OK;
}

// Lift a (possibly grouped, possibly thin) conv weight to a dense zero-padded one so it can run on the tcgen05 path:
//   mode 0: wp[CoutP][CinP][taps] (zeroed here) <- w[Cout][Cin_g][taps] placed on the block diagonal (ci = group*Cin_g + cil)
//   mode 1: w <- the same entries read back out of wp (gradient extraction)
__global__ void k_weight_lift(float* __restrict__ w, float* __restrict__ wp, int Cout, int Cin_g, int groups, int taps, int CinP, int mode) { egm_pdl_enter();
  const int Cout_g = Cout / groups;
  long long total = (long long)Cout * Cin_g * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i % taps); long long q = i / taps; int cil = (int)(q % Cin_g); int co = (int)(q / Cin_g);
    long long j = ((long long)co * CinP + (co / Cout_g) * Cin_g + cil) * taps + t;
    if (mode == 0) wp[j] = w[i]; else w[i] = wp[j];
  }
}
extern "C" int egm_conv_weight_lift(float* w, float* wp, int Cout, int Cin_g, int groups, int taps, int CoutP, int CinP, int mode, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGM_REQUIRE(CoutP >= Cout && CinP >= Cin_g * groups && Cout % groups == 0, EGM_E_SHAPE, "weight_lift: bad padded shape");
  if (mode == 0) cudaMemsetAsync(wp, 0, sizeof(float) * (size_t)CoutP * CinP * taps, st);
  long long total = (long long)Cout * Cin_g * taps;
  if (total == 0) return EGM_OK;
  egm_launch(k_weight_lift, egm_grid_for(total, 256), 256, 0, st, w, wp, Cout, Cin_g, groups, taps, CinP, mode);
  EGM_LAUNCH_CHECK("conv_weight_lift"); return EGM_OK;
}
