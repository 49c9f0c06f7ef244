// Down-/up-sampling kernels of the encoder-decoder skeleton (HBM-bound, vectorised NHWC):
//   MaxPool2d(2, stride 2)                 src/EGM-UNet.py:905-912, src/unet.py:21-26
//   Upsample(x2, bilinear, align_corners) + F.pad + cat([skip, up])   src/EGM-UNet.py:927-949
// The up-sample kernel writes straight into the concatenated tensor the following conv reads.
#include "common.cuh"

// ------------------------------------------------------------------ max pool 2x2 s2
template <typename T, int V>
__global__ void k_maxpool2_fwd(const T* __restrict__ x, long long xcs, long long xco, T* __restrict__ y, int N, int H, int W, int CV) { egm_pdl_enter();
  const int Ho = H / 2, Wo = W / 2, C = CV * V;
  long long total = (long long)N * Ho * Wo * CV;
  const NhwcIndexer ix(CV, Wo, Ho, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const Nhwc4 e = ix(i); const int cv = e.cv, wo = e.w, ho = e.h, n = e.n; const long long p = e.p;
    const T* b = x + (((long long)n * H + 2 * ho) * W + 2 * wo) * xcs + xco + cv * V;
    FVec<V> a = ldv<V>(b), c1 = ldv<V>(b + xcs), c2 = ldv<V>(b + (long long)W * xcs), c3 = ldv<V>(b + (long long)W * xcs + xcs), o;
#pragma unroll
    for (int j = 0; j < V; ++j) o.v[j] = fmaxf(fmaxf(a.v[j], c1.v[j]), fmaxf(c2.v[j], c3.v[j]));
    stv<V>(y + p * C + cv * V, o);
  }
}
// gradient goes to the FIRST maximum in (h, w) scan order (ATen tie rule)
template <typename T, int V>
__global__ void k_maxpool2_bwd(const T* __restrict__ x, long long xcs, long long xco, const T* __restrict__ dy, T* __restrict__ dx, long long dcs, long long dco,
                               int N, int H, int W, int CV, int accumulate) { egm_pdl_enter();
  const int Ho = H / 2, Wo = W / 2, C = CV * V;
  long long total = (long long)N * Ho * Wo * CV;
  const NhwcIndexer ix(CV, Wo, Ho, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const Nhwc4 e = ix(i); const int cv = e.cv, wo = e.w, ho = e.h, n = e.n; const long long p = e.p;
    const long long pix = ((long long)n * H + 2 * ho) * W + 2 * wo;
    const long long pofs[4] = {0, 1, (long long)W, (long long)W + 1};
    FVec<V> v[4], g = ldv<V>(dy + p * C + cv * V);
#pragma unroll
    for (int t = 0; t < 4; ++t) v[t] = ldv<V>(x + (pix + pofs[t]) * xcs + xco + cv * V);
    int arg[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float m = v[0].v[j]; int a = 0;
#pragma unroll
      for (int t = 1; t < 4; ++t) if (v[t].v[j] > m) { m = v[t].v[j]; a = t; }
      arg[j] = a;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      FVec<V> o;
      T* dp = dx + (pix + pofs[t]) * dcs + dco + cv * V;
      if (accumulate) o = ldv<V>(dp);
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] = (accumulate ? o.v[j] : 0.f) + (arg[j] == t ? g.v[j] : 0.f);
      stv<V>(dp, o);
    }
  }
}
// x may be a channel-strided view (element (pixel, c) at x[pixel * x_cstride + x_coff + c]): the first skip connection of the U lives
// inside the Up level's concat buffer (egm_upsample_concat_fwd with skip == NULL) and is pooled from there.
extern "C" int egm_maxpool2x2_fwd_view(const void* x, long long x_cstride, long long x_coff, void* y, int dtype, int N, int H, int W, int C, void* stream) {
  long long total = (long long)N * (H / 2) * (W / 2) * C;
  if (total == 0) return EGM_OK;
  int v = egm_pick_vec(C, x_cstride, x_coff);
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_maxpool2_fwd<T, V>, egm_grid_for(total / V, 256), 256, 0, (cudaStream_t)stream,
      (const T*)x, x_cstride, x_coff, (T*)y, N, H, W, C / V))));
  EGM_LAUNCH_CHECK("maxpool2x2_fwd"); return EGM_OK;
}
extern "C" int egm_maxpool2x2_fwd(const void* x, void* y, int dtype, int N, int H, int W, int C, void* stream) {
  return egm_maxpool2x2_fwd_view(x, C, 0, y, dtype, N, H, W, C, stream);
}
// dx may be a channel-strided view as well (the gradient of the concat buffer); accumulate = 0 needs a dense dx.
extern "C" int egm_maxpool2x2_bwd_view(const void* x, long long x_cstride, long long x_coff, const void* dy, void* dx, long long dx_cstride, long long dx_coff,
                                       int accumulate, int dtype, int N, int H, int W, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  size_t es = dtype == EGM_F32 ? 4 : 2;
  EGM_REQUIRE(accumulate || (dx_cstride == C && dx_coff == 0), EGM_E_BADARG, "maxpool2x2_bwd: a strided dx must be accumulated into");
  if (!accumulate && ((H & 1) || (W & 1))) cudaMemsetAsync(dx, 0, (size_t)N * H * W * C * es, st);   // dropped trailing row/col gets no gradient
  long long total = (long long)N * (H / 2) * (W / 2) * C;
  if (total == 0) return EGM_OK;
  int v = egm_pick_vec(C, x_cstride, x_coff), v2 = egm_pick_vec(C, dx_cstride, dx_coff); if (v2 < v) v = v2;
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_maxpool2_bwd<T, V>, egm_grid_for(total / V, 256), 256, 0, st,
      (const T*)x, x_cstride, x_coff, (const T*)dy, (T*)dx, dx_cstride, dx_coff, N, H, W, C / V, accumulate))));
  EGM_LAUNCH_CHECK("maxpool2x2_bwd"); return EGM_OK;
}
extern "C" int egm_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int accumulate, int dtype, int N, int H, int W, int C, void* stream) {
  return egm_maxpool2x2_bwd_view(x, C, 0, dy, dx, C, 0, accumulate, dtype, N, H, W, C, stream);
}

// ------------------------------------------------------------------ bilinear x2 (align_corners=True) + pad + concat
struct UpGeom { int N, Hl, Wl, H, W, Cs, Cu, padt, padl; float sh, sw; };

__device__ __forceinline__ void up_src(int o, int in, float scale, int& i0, int& ip, float& l0, float& l1) {
  float src = scale * (float)o;               // ATen area_pixel_compute_source_index, align_corners=True
  i0 = (int)src; ip = (i0 < in - 1) ? 1 : 0; l1 = src - (float)i0; l0 = 1.f - l1;
}
template <typename T, int V>
__global__ void k_upcat_fwd(const T* __restrict__ skip, const T* __restrict__ low, T* __restrict__ out, UpGeom g) { egm_pdl_enter();
  // skip == nullptr: the skip half of `out` was written in place by its producer; only the up-sampled channels are produced here
  const int C = g.Cs + g.Cu, c_begin = skip ? 0 : g.Cs, CV = (C - c_begin) / V;
  long long total = (long long)g.N * g.H * g.W * CV;
  const NhwcIndexer ix(CV, g.W, g.H, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const Nhwc4 e = ix(i); const int c = c_begin + e.cv * V, w = e.w, h = e.h, n = e.n; const long long p = e.p;
    FVec<V> o;
    if (c < g.Cs) o = ldv<V>(skip + p * g.Cs + c);
    else {
      int cu = c - g.Cs, uh = h - g.padt, uw = w - g.padl;
      if (uh < 0 || uh >= 2 * g.Hl || uw < 0 || uw >= 2 * g.Wl) {
#pragma unroll
        for (int j = 0; j < V; ++j) o.v[j] = 0.f;
      } else {
        int h0, hp, w0, wp; float a0, a1, b0, b1;
        up_src(uh, g.Hl, g.sh, h0, hp, a0, a1); up_src(uw, g.Wl, g.sw, w0, wp, b0, b1);
        const T* r0 = low + (((long long)n * g.Hl + h0) * g.Wl + w0) * g.Cu + cu;
        const T* r1 = r0 + (long long)hp * g.Wl * g.Cu;
        FVec<V> v00 = ldv<V>(r0), v01 = ldv<V>(r0 + (long long)wp * g.Cu), v10 = ldv<V>(r1), v11 = ldv<V>(r1 + (long long)wp * g.Cu);
#pragma unroll
        for (int j = 0; j < V; ++j) o.v[j] = a0 * (b0 * v00.v[j] + b1 * v01.v[j]) + a1 * (b0 * v10.v[j] + b1 * v11.v[j]);
      }
    }
    stv<V>(out + p * C + c, o);
  }
}
// gather form of the transpose: each low-res pixel sums the (<= 6x6) high-res pixels that read it
template <typename T, int V>
__global__ void k_upcat_bwd_low(const T* __restrict__ dcat, T* __restrict__ dlow, UpGeom g) { egm_pdl_enter();
  const int C = g.Cs + g.Cu, CV = g.Cu / V;
  long long total = (long long)g.N * g.Hl * g.Wl * CV;
  const NhwcIndexer ix(CV, g.Wl, g.Hl, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const Nhwc4 e = ix(i); const int cu = e.cv * V, wl = e.w, hl = e.h, n = e.n; const long long p = e.p;
    float wh[6], ww[6];
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      int uh = 2 * hl - 2 + t, uw = 2 * wl - 2 + t; wh[t] = 0.f; ww[t] = 0.f;
      if (uh >= 0 && uh < 2 * g.Hl) { int i0, ip; float l0, l1; up_src(uh, g.Hl, g.sh, i0, ip, l0, l1); if (i0 == hl) wh[t] += l0; if (i0 + ip == hl) wh[t] += l1; }
      if (uw >= 0 && uw < 2 * g.Wl) { int i0, ip; float l0, l1; up_src(uw, g.Wl, g.sw, i0, ip, l0, l1); if (i0 == wl) ww[t] += l0; if (i0 + ip == wl) ww[t] += l1; }
    }
    FVec<V> acc;
#pragma unroll
    for (int j = 0; j < V; ++j) acc.v[j] = 0.f;
    // rows with a non-zero weight only (uniform-ish skip), but inside a row all 6 candidate loads are issued unconditionally on
    // clamped coordinates and masked through the weight: 6 loads in flight instead of 6 dependent, predicated ones
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      int h = 2 * hl - 2 + a + g.padt;
      if (wh[a] == 0.f || h < 0 || h >= g.H) continue;
      const T* rowp = dcat + (((long long)n * g.H + h) * g.W) * C + g.Cs + cu;
      FVec<V> d[6]; float f[6];
#pragma unroll
      for (int b = 0; b < 6; ++b) {
        int w = 2 * wl - 2 + b + g.padl;
        const bool ok = w >= 0 && w < g.W;
        f[b] = ok ? wh[a] * ww[b] : 0.f;
        w = w < 0 ? 0 : (w >= g.W ? g.W - 1 : w);
        d[b] = ldv<V>(rowp + (long long)w * C);
      }
#pragma unroll
      for (int b = 0; b < 6; ++b)
#pragma unroll
        for (int j = 0; j < V; ++j) acc.v[j] = fmaf(f[b], d[b].v[j], acc.v[j]);
    }
    stv<V>(dlow + p * g.Cu + cu, acc);
  }
}
static int up_geom(UpGeom& g, int N, int Hl, int Wl, int H, int W, int Cs, int Cu) {
  int dy = H - 2 * Hl, dx = W - 2 * Wl;
  EGM_REQUIRE(dy >= 0 && dx >= 0, EGM_E_SHAPE, "upsample_concat: skip %dx%d smaller than 2x low %dx%d", H, W, Hl, Wl);
  g = UpGeom{N, Hl, Wl, H, W, Cs, Cu, dy / 2, dx / 2,
             2 * Hl > 1 ? (float)(Hl - 1) / (float)(2 * Hl - 1) : 0.f, 2 * Wl > 1 ? (float)(Wl - 1) / (float)(2 * Wl - 1) : 0.f};
  return EGM_OK;
}
// out[N,H,W,Cs+Cu] = cat([skip[N,H,W,Cs], pad(upsample2x(low[N,Hl,Wl,Cu]))]);  skip == NULL: out[..., :Cs] already holds the skip
// (its producer wrote it in place -- the concat is virtual) and only out[..., Cs:] is written
extern "C" int egm_upsample_concat_fwd(const void* skip, const void* low, void* out, int dtype, int N, int Hl, int Wl, int H, int W, int Cs, int Cu, void* stream) {
  UpGeom g; int e = up_geom(g, N, Hl, Wl, H, W, Cs, Cu); if (e) return e;
  long long total = (long long)N * H * W * (skip ? Cs + Cu : Cu);
  if (total == 0) return EGM_OK;
  int v = egm_pick_vec(Cs); int v2 = egm_pick_vec(Cu); if (v2 < v) v = v2;
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_upcat_fwd<T, V>, egm_grid_for(total / V, 256), 256, 0, (cudaStream_t)stream, (const T*)skip, (const T*)low, (T*)out, g))));
  EGM_LAUNCH_CHECK("upsample_concat_fwd"); return EGM_OK;
}
// dlow[N,Hl,Wl,Cu] = transpose of the bilinear part applied to dcat[..., Cs:]; the skip part is a plain slice (egm_copy_slice).
extern "C" int egm_upsample_concat_bwd_low(const void* dcat, void* dlow, int dtype, int N, int Hl, int Wl, int H, int W, int Cs, int Cu, void* stream) {
  UpGeom g; int e = up_geom(g, N, Hl, Wl, H, W, Cs, Cu); if (e) return e;
  long long total = (long long)N * Hl * Wl * Cu;
  if (total == 0) return EGM_OK;
  int v = egm_pick_vec(Cu, Cs + Cu, Cs);
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_upcat_bwd_low<T, V>, egm_grid_for(total / V, 256), 256, 0, (cudaStream_t)stream, (const T*)dcat, (T*)dlow, g))));
  EGM_LAUNCH_CHECK("upsample_concat_bwd_low"); return EGM_OK;
}

// ------------------------------------------------------------------ ConvTranspose2d(k=2, s=2) as a per-pixel GEMM + pixel shuffle
// (UNet(bilinear=False): src/unet.py:35-37).  The GEMM is a 1x1 conv with the re-laid-out weight
//   wT[(a*2+b)*Cout + co][ci] = W[ci][co][a][b]        (mode 0: W -> wT,  mode 1: wT -> W, used for the gradient)
// and the kernels below scatter / gather its [N,Hl,Wl,4*Cout] output into the up-sampled half of the concat tensor.
__global__ void k_deconv_wpack(float* __restrict__ w, float* __restrict__ wt, int Cin, int Cout, int mode) { egm_pdl_enter();
  long long total = (long long)Cin * Cout * 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int ab = (int)(i & 3); long long q = i >> 2; int co = (int)(q % Cout); int ci = (int)(q / Cout);
    long long j = ((long long)ab * Cout + co) * Cin + ci;
    if (mode == 0) wt[j] = w[i]; else w[i] = wt[j];
  }
}
extern "C" int egm_deconv_weight_pack(float* w, float* wt, int Cin, int Cout, int mode, void* stream) {
  long long total = (long long)Cin * Cout * 4;
  if (total == 0) return EGM_OK;
  egm_launch(k_deconv_wpack, egm_grid_for(total, 256), 256, 0, (cudaStream_t)stream, w, wt, Cin, Cout, mode);
  EGM_LAUNCH_CHECK("deconv_weight_pack"); return EGM_OK;
}
// out[N,H,W,Cs+Cu] = cat([skip, pad(pixel_shuffle(z[N,Hl,Wl,4*Cu]))])
template <typename T, int V>
__global__ void k_shufcat_fwd(const T* __restrict__ skip, const T* __restrict__ z, T* __restrict__ out, UpGeom g) { egm_pdl_enter();
  const int C = g.Cs + g.Cu, CV = C / V;
  long long total = (long long)g.N * g.H * g.W * CV;
  const NhwcIndexer ix(CV, g.W, g.H, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const Nhwc4 e = ix(i); const int c = e.cv * V, w = e.w, h = e.h, n = e.n; const long long p = e.p;
    FVec<V> o;
    if (c < g.Cs) o = ldv<V>(skip + p * g.Cs + c);
    else {
      int uh = h - g.padt, uw = w - g.padl;
      if (uh < 0 || uh >= 2 * g.Hl || uw < 0 || uw >= 2 * g.Wl) {
#pragma unroll
        for (int j = 0; j < V; ++j) o.v[j] = 0.f;
      } else {
        int ab = (uh & 1) * 2 + (uw & 1);
        o = ldv<V>(z + ((((long long)n * g.Hl + (uh >> 1)) * g.Wl + (uw >> 1)) * 4 + ab) * g.Cu + (c - g.Cs));
      }
    }
    stv<V>(out + p * C + c, o);
  }
}
// dz[N,Hl,Wl,4*Cu] gathered from dcat[..., Cs:]
template <typename T, int V>
__global__ void k_shufcat_bwd(const T* __restrict__ dcat, T* __restrict__ dz, UpGeom g) { egm_pdl_enter();
  const int C = g.Cs + g.Cu, CV = g.Cu / V;
  long long total = (long long)g.N * g.Hl * g.Wl * 4 * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cu = (int)(i % CV) * V; long long p = i / CV; int ab = (int)(p & 3); long long pl = p >> 2;
    int wl = (int)(pl % g.Wl); long long q = pl / g.Wl; int hl = (int)(q % g.Hl); int n = (int)(q / g.Hl);
    int h = 2 * hl + (ab >> 1) + g.padt, w = 2 * wl + (ab & 1) + g.padl;
    FVec<V> o = ldv<V>(dcat + (((long long)n * g.H + h) * g.W + w) * C + g.Cs + cu);
    stv<V>(dz + p * g.Cu + cu, o);
  }
}
extern "C" int egm_pixel_shuffle_concat_fwd(const void* skip, const void* z, void* out, int dtype, int N, int Hl, int Wl, int H, int W, int Cs, int Cu, void* stream) {
  UpGeom g; int e = up_geom(g, N, Hl, Wl, H, W, Cs, Cu); if (e) return e;
  long long total = (long long)N * H * W * (skip ? Cs + Cu : Cu);
  if (total == 0) return EGM_OK;
  int v = egm_pick_vec(Cs); int v2 = egm_pick_vec(Cu); if (v2 < v) v = v2;
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_shufcat_fwd<T, V>, egm_grid_for(total / V, 256), 256, 0, (cudaStream_t)stream, (const T*)skip, (const T*)z, (T*)out, g))));
  EGM_LAUNCH_CHECK("pixel_shuffle_concat_fwd"); return EGM_OK;
}
extern "C" int egm_pixel_shuffle_concat_bwd(const void* dcat, void* dz, int dtype, int N, int Hl, int Wl, int H, int W, int Cs, int Cu, void* stream) {
  UpGeom g; int e = up_geom(g, N, Hl, Wl, H, W, Cs, Cu); if (e) return e;
  long long total = (long long)N * Hl * Wl * 4 * Cu;
  if (total == 0) return EGM_OK;
  int v = egm_pick_vec(Cu, Cs + Cu, Cs);
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_shufcat_bwd<T, V>, egm_grid_for(total / V, 256), 256, 0, (cudaStream_t)stream, (const T*)dcat, (T*)dz, g))));
  EGM_LAUNCH_CHECK("pixel_shuffle_concat_bwd"); return EGM_OK;
}
