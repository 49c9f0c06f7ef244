// Elementwise / stencil / small-reduction kernels of the EGM blocks (all HBM- or latency-bound):
//   EdgeAwareFeatureEnhancer high-pass           src/EGM-UNet.py:872-886
//   FusionConv spatial + channel attention       src/EGM-UNet.py:1171-1236
//   EdgeEnhancedGRFB tail (target enhancer)      src/EGM-UNet.py:1289-1323
//   RecursiveGatedAttention gating               src/EGM-UNet.py:458-547
#include "common.cuh"

// ------------------------------------------------------------------ out (+)= in - avgpool3x3(in)   (zero pad, /9; self-adjoint)
template <typename T, int V>
__global__ void k_highpass3(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int CV, int accumulate) { egm_pdl_enter();
  const int C = CV * V; long long total = (long long)N * H * W * CV;
  const NhwcIndexer ix(CV, W, H, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const Nhwc4 e = ix(i); const int c = e.cv * V, w = e.w, h = e.h; const long long p = e.p;
    // all 9 neighbour loads are issued unconditionally (out-of-image neighbours are clamped onto the centre row / column and
    // masked afterwards): predicated loads were consumed one by one -- ncu showed 9 separate long-scoreboard stalls per vector
    const bool vr[3] = {h > 0, true, h < H - 1}, vc[3] = {w > 0, true, w < W - 1};
    const long long ro[3] = {vr[0] ? -(long long)W : 0, 0, vr[2] ? (long long)W : 0}, cofs[3] = {vc[0] ? -1 : 0, 0, vc[2] ? 1 : 0};
    FVec<V> t[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) t[a][b] = ldv<V>(in + (p + ro[a] + cofs[b]) * C + c);
    FVec<V> s, ctr = t[1][1];
#pragma unroll
    for (int j = 0; j < V; ++j) s.v[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const float m = (vr[a] && vc[b]) ? 1.f : 0.f;
#pragma unroll
        for (int j = 0; j < V; ++j) s.v[j] += m * t[a][b].v[j];
      }
    FVec<V> o;
    if (accumulate) o = ldv<V>(out + p * C + c);
#pragma unroll
    for (int j = 0; j < V; ++j) o.v[j] = (accumulate ? o.v[j] : 0.f) + ctr.v[j] - s.v[j] * (1.f / 9.f);
    stv<V>(out + p * C + c, o);
  }
}
int egm_highpass3_walk_launch(const void* in, void* out, int accumulate, int dtype, int N, int H, int W, int C, cudaStream_t st);   // mca_fused.cu
extern "C" int egm_highpass3(const void* in, void* out, int accumulate, int dtype, int N, int H, int W, int C, void* stream) {
  long long total = (long long)N * H * W * C;
  if (total == 0) return EGM_OK;
  // C % 64 == 0 (the GRFB-level enhancers): shared-memory-staged row walk; thin tensors (the C/8 enhancer inside branch_edge): gather kernel
  static int walk = -1;
  if (walk < 0) { const char* e = getenv("EGM_NO_STENCIL_WALK"); walk = (e && e[0] == '1') ? 0 : 1; }
  if (walk && egm_highpass3_walk_launch(in, out, accumulate, dtype, N, H, W, C, (cudaStream_t)stream)) { EGM_LAUNCH_CHECK("highpass3(walk)"); return EGM_OK; }
  int v = egm_pick_vec(C);
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_highpass3<T, V>, egm_grid_for(total / V, 256), 256, 0, (cudaStream_t)stream, (const T*)in, (T*)out, N, H, W, C / V, accumulate))));
  EGM_LAUNCH_CHECK("highpass3"); return EGM_OK;
}

// ------------------------------------------------------------------ per-pixel dot over channels: out[m] = sum_c a*b*(cvec[n][c])
// G threads (power of two <= 32) cooperate on one pixel.
template <typename T, int V>
__global__ void __launch_bounds__(256) k_pixel_dot(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ cvec, float* __restrict__ out,
                                                   long long M, long long HW, int C, int G) { egm_pdl_enter();
  const int CV = C / V;
  const int lane = threadIdx.x % G;
  const long long gpb = blockDim.x / G;
  const long long nIter = (M + gpb * gridDim.x - 1) / (gpb * gridDim.x);
  for (long long it = 0; it < nIter; ++it) {
    long long m = (it * gridDim.x + blockIdx.x) * gpb + threadIdx.x / G;
    float s = 0.f;
    if (m < M) {
      const float* cv = cvec ? cvec + (m / HW) * C : nullptr;
      for (int q = lane; q < CV; q += G) {
        FVec<V> x = ldv<V>(a + m * C + q * V), y = ldv<V>(b + m * C + q * V);
#pragma unroll
        for (int j = 0; j < V; ++j) s += x.v[j] * y.v[j] * (cv ? cv[q * V + j] : 1.f);
      }
    }
    for (int o = G >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (m < M && lane == 0) out[m] = s;
  }
}
extern "C" int egm_pixel_dot(const void* a, const void* b, const float* cvec, float* out, int dtype, int N, long long HW, int C, void* stream) {
  long long M = (long long)N * HW;
  if (M == 0) return EGM_OK;
  int v = egm_pick_vec(C); int CV = C / v; int G = 1; while (G < CV && G < 32) G <<= 1;
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_pixel_dot<T, V>, egm_grid_for(M * G, 256), 256, 0, (cudaStream_t)stream, (const T*)a, (const T*)b, cvec, out, M, HW, C, G))));
  EGM_LAUNCH_CHECK("pixel_dot"); return EGM_OK;
}

// ------------------------------------------------------------------ per-(sample, channel) dot over pixels: out[n][c] = sum_p a*b*pvec[p]
template <typename T, int V>
__global__ void k_sample_chan_dot(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ pvec, float* __restrict__ out, long long HW, int C) { egm_pdl_enter();
  extern __shared__ float sm[];
  const int CV = C / V, rpi = blockDim.x / CV, cv = threadIdx.x % CV, r = threadIdx.x / CV, n = blockIdx.y;
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  const long long base = (long long)n * HW;
  for (long long p = (long long)blockIdx.x * rpi + r; p < HW; p += (long long)gridDim.x * rpi) {
    FVec<V> x = ldv<V>(a + (base + p) * C + cv * V);
    float f = pvec ? pvec[base + p] : 1.f;
    if (b) { FVec<V> y = ldv<V>(b + (base + p) * C + cv * V);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += x.v[j] * y.v[j] * f; }
    else {
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += x.v[j] * f; }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) sm[(size_t)r * C + cv * V + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < rpi; ++q) s += sm[(size_t)q * C + c];
    atomicAdd(out + (long long)n * C + c, s);
  }
}
extern "C" int egm_sample_chan_dot(const void* a, const void* b, const float* pvec, float* out, int dtype, int N, long long HW, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N * C, st);
  if ((long long)N * HW == 0) return EGM_OK;
  int v = egm_pick_vec(C); int CV = C / v; EGM_REQUIRE(CV <= 256, EGM_E_SHAPE, "sample_chan_dot: C too large");
  int rpi = 256 / CV; int threads = CV * rpi;
  long long bx = (HW + (long long)rpi * 16 - 1) / ((long long)rpi * 16); long long cap = egm_num_sms() * 4 / N + 1; if (bx > cap) bx = cap; if (bx < 1) bx = 1;
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_sample_chan_dot<T, V>, dim3((unsigned)bx, N), threads, (size_t)rpi * C * sizeof(float), st, 
      (const T*)a, (const T*)b, pvec, out, HW, C))));
  EGM_LAUNCH_CHECK("sample_chan_dot"); return EGM_OK;
}

// ------------------------------------------------------------------ pixel gates:  phi(g[m]) with g a T tensor of Gc channels
//   mode 0: sigmoid(g[m][0])          (RGA gate map)
//   mode 1: 1 + mean_j sigmoid(g[m][j])   (GRFB target enhancer, Gc = 3)
template <typename T>
__device__ __forceinline__ float pixel_phi(const T* g, long long m, int Gc, int mode) {
  if (mode == 0) return sigmoidf_(ldf(g + m * Gc));
  float s = 0.f;
  for (int j = 0; j < Gc; ++j) s += sigmoidf_(ldf(g + m * Gc + j));
  return 1.f + s / (float)Gc;
}
template <typename T, int V>
__global__ void k_mul_pixel_gate(const T* __restrict__ a, const T* __restrict__ g, T* __restrict__ y, long long M, int CV, int Gc, int mode) { egm_pdl_enter();
  const int C = CV * V; long long total = M * CV;
  const RowIndexer rix(CV, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long m; int c; rix(i, m, c); c *= V;
    float f = pixel_phi(g, m, Gc, mode);
    FVec<V> x = ldv<V>(a + m * C + c);
#pragma unroll
    for (int j = 0; j < V; ++j) x.v[j] *= f;
    stv<V>(y + m * C + c, x);
  }
}
extern "C" int egm_mul_pixel_gate(const void* a, const void* g, void* y, int dtype, long long M, int C, int Gc, int mode, void* stream) {
  if (M * C == 0) return EGM_OK;
  int v = egm_pick_vec(C);
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_mul_pixel_gate<T, V>, egm_grid_for(M * (C / V), 256), 256, 0, (cudaStream_t)stream, (const T*)a, (const T*)g, (T*)y, M, C / V, Gc, mode))));
  EGM_LAUNCH_CHECK("mul_pixel_gate"); return EGM_OK;
}
// dg[m][j] = dot[m] * d phi / d g_j
template <typename T>
__global__ void k_pixel_gate_bwd(const float* __restrict__ dot, const T* __restrict__ g, T* __restrict__ dg, long long M, int Gc, int mode) { egm_pdl_enter();
  long long total = M * Gc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long m = i / Gc;
    float s = sigmoidf_(ldf(g + i));
    float d = dot[m] * s * (1.f - s);
    if (mode == 1) d /= (float)Gc;
    stf(dg + i, d);
  }
}
extern "C" int egm_pixel_gate_bwd(const float* dot, const void* g, void* dg, int dtype, long long M, int Gc, int mode, void* stream) {
  if (M * Gc == 0) return EGM_OK;
  EGM_DISPATCH_DTYPE(dtype, (egm_launch(k_pixel_gate_bwd<T>, egm_grid_for(M * Gc, 256), 256, 0, (cudaStream_t)stream, dot, (const T*)g, (T*)dg, M, Gc, mode)));
  EGM_LAUNCH_CHECK("pixel_gate_bwd"); return EGM_OK;
}

// ------------------------------------------------------------------ unary elementwise (flat): gelu fwd/bwd, scale by a device scalar
//   op 0: y = gelu(x) (exact erf)      op 1: y = dy * gelu'(x)   (a = dy, b = x)       op 2: y = a * (*scalar)
template <typename T, int V>
__global__ void k_unary(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ scalar, T* __restrict__ y, long long nv, int op) { egm_pdl_enter();
  const float sc = (op == 2) ? *scalar : 1.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    FVec<V> x = ldv<V>(a + i * V), o, z;
    if (op == 1) z = ldv<V>(b + i * V);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (op == 0) o.v[j] = 0.5f * x.v[j] * (1.f + erff(x.v[j] * 0.70710678118654752f));
      else if (op == 1) { float t = z.v[j]; o.v[j] = x.v[j] * (0.5f * (1.f + erff(t * 0.70710678118654752f)) + t * 0.3989422804014327f * __expf(-0.5f * t * t)); }
      else o.v[j] = x.v[j] * sc;
    }
    stv<V>(y + i * V, o);
  }
}
extern "C" int egm_unary(const void* a, const void* b, const float* scalar_dev, void* y, int dtype, long long n, int op, void* stream) {
  if (n == 0) return EGM_OK;
  int v = egm_pick_vec(n);
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_unary<T, V>, egm_grid_for(n / V, 256), 256, 0, (cudaStream_t)stream, (const T*)a, (const T*)b, scalar_dev, (T*)y, n / V, op))));
  EGM_LAUNCH_CHECK("unary"); return EGM_OK;
}
// out[0] = sum_i a[i]*b[i]
template <typename T>
__global__ void k_dot_all(const T* __restrict__ a, const T* __restrict__ b, long long n, float* out) { egm_pdl_enter();
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += ldf(a + i) * ldf(b + i);
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
extern "C" int egm_dot_all(const void* a, const void* b, float* out, int dtype, long long n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(float), st);
  if (n == 0) return EGM_OK;
  EGM_DISPATCH_DTYPE(dtype, (egm_launch(k_dot_all<T>, egm_grid_for(n, 256, 2), 256, 0, st, (const T*)a, (const T*)b, n, out)));
  EGM_LAUNCH_CHECK("dot_all"); return EGM_OK;
}

// ------------------------------------------------------------------ FusionConv spatial attention
// mm[m] = (mean_c s, max_c s) fp32; amax[m] = first arg-max channel
template <typename T>
__global__ void k_chan_meanmax(const T* __restrict__ s, float* __restrict__ mm, unsigned char* __restrict__ amax, long long M, int C) { egm_pdl_enter();
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    float sum = 0.f, mx = -INFINITY; int am = 0;
    for (int c = 0; c < C; ++c) { float v = ldf(s + m * C + c); sum += v; if (v > mx) { mx = v; am = c; } }
    mm[m * 2] = sum / (float)C; mm[m * 2 + 1] = mx; amax[m] = (unsigned char)am;
  }
}
extern "C" int egm_chan_meanmax(const void* s, float* mm, unsigned char* amax, int dtype, long long M, int C, void* stream) {
  EGM_REQUIRE(C <= 256, EGM_E_SHAPE, "chan_meanmax: C > 256");
  if (M == 0) return EGM_OK;
  EGM_DISPATCH_DTYPE(dtype, (egm_launch(k_chan_meanmax<T>, egm_grid_for(M, 128), 128, 0, (cudaStream_t)stream, (const T*)s, mm, amax, M, C)));
  EGM_LAUNCH_CHECK("chan_meanmax"); return EGM_OK;
}
// sa[m] = sigmoid(conv7x7(mm; w[1][2][7][7], pad 3, no bias))
__global__ void k_sa_conv_fwd(const float* __restrict__ mm, const float* __restrict__ w, float* __restrict__ sa, int N, int H, int W) { egm_pdl_enter();
  __shared__ float ws[98];
  for (int i = threadIdx.x; i < 98; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  long long M = (long long)N * H * W;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    int wq = (int)(m % W); int h = (int)((m / W) % H);
    float acc = 0.f;
    for (int r = 0; r < 7; ++r) { int hh = h + r - 3; if (hh < 0 || hh >= H) continue;
      for (int s = 0; s < 7; ++s) { int ww = wq + s - 3; if (ww < 0 || ww >= W) continue;
        const float* p = mm + (m + (long long)(r - 3) * W + (s - 3)) * 2;
        acc += ws[r * 7 + s] * p[0] + ws[49 + r * 7 + s] * p[1]; } }
    sa[m] = sigmoidf_(acc);
  }
}
extern "C" int egm_sa_conv_fwd(const float* mm, const float* w, float* sa, int N, int H, int W, void* stream) {
  long long M = (long long)N * H * W; if (M == 0) return EGM_OK;
  egm_launch(k_sa_conv_fwd, egm_grid_for(M, 128), 128, 0, (cudaStream_t)stream, mm, w, sa, N, H, W);
  EGM_LAUNCH_CHECK("sa_conv_fwd"); return EGM_OK;
}
// dpre = dsa * sa (1-sa);  dmm[q][ch] = sum_{r,s} w[ch][r][s] dpre[q-(r-3,s-3)];  dw[ch][r][s] = sum_p dpre[p] mm[p+(r-3,s-3)][ch]
__global__ void k_sa_conv_bwd_dmm(const float* __restrict__ dsa, const float* __restrict__ sa, const float* __restrict__ w, float* __restrict__ dmm, int N, int H, int W) { egm_pdl_enter();
  __shared__ float ws[98];
  for (int i = threadIdx.x; i < 98; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  long long M = (long long)N * H * W;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    int wq = (int)(m % W); int h = (int)((m / W) % H);
    float a0 = 0.f, a1 = 0.f;
    for (int r = 0; r < 7; ++r) { int hh = h - (r - 3); if (hh < 0 || hh >= H) continue;
      for (int s = 0; s < 7; ++s) { int ww = wq - (s - 3); if (ww < 0 || ww >= W) continue;
        long long p = m - (long long)(r - 3) * W - (s - 3);
        float sg = sa[p], d = dsa[p] * sg * (1.f - sg);
        a0 += ws[r * 7 + s] * d; a1 += ws[49 + r * 7 + s] * d; } }
    dmm[m * 2] = a0; dmm[m * 2 + 1] = a1;
  }
}
__global__ void __launch_bounds__(256) k_sa_conv_bwd_dw(const float* __restrict__ dsa, const float* __restrict__ sa, const float* __restrict__ mm,
                                                        float* __restrict__ dw, int N, int H, int W) { egm_pdl_enter();
  __shared__ float red[32];
  const int ch = blockIdx.y / 7, r = blockIdx.y % 7;
  long long M = (long long)N * H * W;
  float acc[7];
#pragma unroll
  for (int s = 0; s < 7; ++s) acc[s] = 0.f;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    int wq = (int)(m % W); int h = (int)((m / W) % H);
    int hh = h + r - 3; if (hh < 0 || hh >= H) continue;
    float sg = sa[m], d = dsa[m] * sg * (1.f - sg);
#pragma unroll
    for (int s = 0; s < 7; ++s) { int ww = wq + s - 3; if (ww < 0 || ww >= W) continue;
      acc[s] += d * mm[(m + (long long)(r - 3) * W + (s - 3)) * 2 + ch]; }
  }
#pragma unroll
  for (int s = 0; s < 7; ++s) { float v = block_sum(acc[s], red); if (threadIdx.x == 0) atomicAdd(dw + ch * 49 + r * 7 + s, v); }
}
extern "C" int egm_sa_conv_bwd(const float* dsa, const float* sa, const float* mm, const float* w, float* dmm, float* dw, int N, int H, int W, void* stream) {
  // dmm == NULL or dw == NULL skips that half: the input gradient is on the backward dependency chain, the 98-element weight gradient
  // is not (the engine launches it on the weight-gradient lane)
  cudaStream_t st = (cudaStream_t)stream;
  if (dw) cudaMemsetAsync(dw, 0, sizeof(float) * 98, st);
  long long M = (long long)N * H * W; if (M == 0) return EGM_OK;
  if (dmm) egm_launch(k_sa_conv_bwd_dmm, egm_grid_for(M, 128), 128, 0, st, dsa, sa, w, dmm, N, H, W);
  int bx = egm_grid_for(M, 256, 2) / 14 + 1;
  if (dw) egm_launch(k_sa_conv_bwd_dw, dim3(bx, 14), 256, 0, st, dsa, sa, mm, dw, N, H, W);
  EGM_LAUNCH_CHECK("sa_conv_bwd"); return EGM_OK;
}

// ------------------------------------------------------------------ FusionConv channel attention
__device__ __forceinline__ unsigned int f2ord(float f) { unsigned int u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned int u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
// global avg / max (+ FIRST arg-max pixel) per (n, c).  keys: packed (ordered value << 32 | ~pixel) combined with 64-bit atomicMax.
template <typename T, int V>
__global__ void k_gap_gmp(const T* __restrict__ f, float* __restrict__ sum, unsigned long long* __restrict__ keys, long long HW, int C) { egm_pdl_enter();
  const int CV = C / V, rpi = blockDim.x / CV, cv = threadIdx.x % CV, r = threadIdx.x / CV, n = blockIdx.y;
  float acc[V], mx[V]; unsigned int am[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { acc[j] = 0.f; mx[j] = -INFINITY; am[j] = 0; }
  const long long base = (long long)n * HW;
#pragma unroll 4
  for (long long p = (long long)blockIdx.x * rpi + r; p < HW; p += (long long)gridDim.x * rpi) {
    FVec<V> x = ldv<V>(f + (base + p) * C + cv * V);
#pragma unroll
    for (int j = 0; j < V; ++j) { acc[j] += x.v[j]; if (x.v[j] > mx[j]) { mx[j] = x.v[j]; am[j] = (unsigned int)p; } }
  }
  // block-level combine first: 128 rows x same channel hammering one address serialised the whole kernel at L2 (ncu: 0.1 TB/s)
  __shared__ float s_acc[256 * V];
  __shared__ unsigned long long s_key[256 * V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    s_acc[(r * CV + cv) * V + j] = acc[j];
    s_key[(r * CV + cv) * V + j] = mx[j] > -INFINITY ? (((unsigned long long)f2ord(mx[j]) << 32) | (unsigned long long)(0xffffffffu - am[j])) : 0ull;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f; unsigned long long k = 0ull;
    for (int rr = 0; rr < rpi; ++rr) { a += s_acc[rr * C + c]; unsigned long long kk = s_key[rr * C + c]; k = kk > k ? kk : k; }
    atomicAdd(sum + (long long)n * C + c, a);
    if (k) atomicMax(keys + (long long)n * C + c, k);
  }
}
__global__ void k_gap_gmp_fin(const float* sum, const unsigned long long* keys, float* avg, float* mx, int* arg, long long NC, float inv_hw) { egm_pdl_enter();
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= NC) return;
  unsigned long long k = keys[i];
  avg[i] = sum[i] * inv_hw; mx[i] = ord2f((unsigned int)(k >> 32)); arg[i] = (int)(0xffffffffu - (unsigned int)(k & 0xffffffffu));
}
// scratch: N*C floats + N*C u64 (16-byte aligned)
extern "C" int egm_gap_gmp(const void* f, float* avg, float* mx, int* arg, void* scratch, int dtype, int N, long long HW, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  long long NC = (long long)N * C;
  unsigned long long* keys = (unsigned long long*)scratch; float* sum = (float*)(keys + NC);
  cudaMemsetAsync(scratch, 0, (size_t)NC * 12, st);
  if (NC * HW == 0) return EGM_OK;
  int v = egm_pick_vec(C); int CV = C / v; EGM_REQUIRE(CV <= 256, EGM_E_SHAPE, "gap_gmp: C too large");
  int rpi = 256 / CV; int threads = CV * rpi;
  long long bx = (HW + (long long)rpi * 16 - 1) / ((long long)rpi * 16); long long cap = egm_num_sms() * 4 / N + 1; if (bx > cap) bx = cap; if (bx < 1) bx = 1;
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_gap_gmp<T, V>, dim3((unsigned)bx, N), threads, 0, st, (const T*)f, sum, keys, HW, C))));
  egm_launch(k_gap_gmp_fin, cdiv(NC, 128), 128, 0, st, sum, keys, avg, mx, arg, NC, 1.f / (float)HW);
  EGM_LAUNCH_CHECK("gap_gmp"); return EGM_OK;
}
// ca[n][c] = sigmoid(W2 relu(W0 avg) + W2 relu(W0 max));  hid[2][N][Cr] keeps the pre-ReLU hidden activations.  One block per sample.
__global__ void k_ca_mlp_fwd(const float* __restrict__ avg, const float* __restrict__ mx, const float* __restrict__ w0, const float* __restrict__ w2,
                             float* __restrict__ ca, float* __restrict__ hid, int N, int C, int Cr) { egm_pdl_enter();
  extern __shared__ float sh[];   // ha[Cr] hm[Cr]
  const int n = blockIdx.x;
  for (int j = threadIdx.x; j < Cr; j += blockDim.x) {
    float a = 0.f, m = 0.f;
    for (int c = 0; c < C; ++c) { float w = w0[j * C + c]; a += w * avg[n * C + c]; m += w * mx[n * C + c]; }
    hid[(0 * N + n) * Cr + j] = a; hid[(1 * N + n) * Cr + j] = m; sh[j] = fmaxf(a, 0.f); sh[Cr + j] = fmaxf(m, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float o = 0.f;
    for (int j = 0; j < Cr; ++j) o += w2[c * Cr + j] * (sh[j] + sh[Cr + j]);
    ca[n * C + c] = sigmoidf_(o);
  }
}
extern "C" int egm_ca_mlp_fwd(const float* avg, const float* mx, const float* w0, const float* w2, float* ca, float* hid, int N, int C, int Cr, void* stream) {
  egm_launch(k_ca_mlp_fwd, N, 128, 2 * Cr * sizeof(float), (cudaStream_t)stream, avg, mx, w0, w2, ca, hid, N, C, Cr);
  EGM_LAUNCH_CHECK("ca_mlp_fwd"); return EGM_OK;
}
// single block: parameter grads need sums over samples
__global__ void __launch_bounds__(256) k_ca_mlp_bwd(const float* __restrict__ dca, const float* __restrict__ ca, const float* __restrict__ avg,
                                                    const float* __restrict__ mx, const float* __restrict__ hid, const float* __restrict__ w0,
                                                    const float* __restrict__ w2, float* __restrict__ dw0, float* __restrict__ dw2,
                                                    float* __restrict__ davg, float* __restrict__ dmx, int N, int C, int Cr) { egm_pdl_enter();
  extern __shared__ float sh[];   // dpre[N*C] | dha[N*Cr] | dhm[N*Cr]
  float* dpre = sh; float* dha = sh + N * C; float* dhm = dha + N * Cr;
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) { float s = ca[i]; dpre[i] = dca[i] * s * (1.f - s); }
  __syncthreads();
  for (int i = threadIdx.x; i < N * Cr; i += blockDim.x) {
    int n = i / Cr, j = i - n * Cr; float t = 0.f;
    for (int c = 0; c < C; ++c) t += w2[c * Cr + j] * dpre[n * C + c];
    dha[i] = hid[(0 * N + n) * Cr + j] > 0.f ? t : 0.f; dhm[i] = hid[(1 * N + n) * Cr + j] > 0.f ? t : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * Cr; i += blockDim.x) {       // dw2[c][j]
    int c = i / Cr, j = i - c * Cr; float t = 0.f;
    for (int n = 0; n < N; ++n) t += dpre[n * C + c] * (fmaxf(hid[(0 * N + n) * Cr + j], 0.f) + fmaxf(hid[(1 * N + n) * Cr + j], 0.f));
    dw2[i] = t;
  }
  for (int i = threadIdx.x; i < Cr * C; i += blockDim.x) {       // dw0[j][c]
    int j = i / C, c = i - j * C; float t = 0.f;
    for (int n = 0; n < N; ++n) t += dha[n * Cr + j] * avg[n * C + c] + dhm[n * Cr + j] * mx[n * C + c];
    dw0[i] = t;
  }
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) {
    int n = i / C, c = i - n * C; float a = 0.f, m = 0.f;
    for (int j = 0; j < Cr; ++j) { float w = w0[j * C + c]; a += w * dha[n * Cr + j]; m += w * dhm[n * Cr + j]; }
    davg[i] = a; dmx[i] = m;
  }
}
extern "C" int egm_ca_mlp_bwd(const float* dca, const float* ca, const float* avg, const float* mx, const float* hid, const float* w0, const float* w2,
                              float* dw0, float* dw2, float* davg, float* dmx, int N, int C, int Cr, void* stream) {
  size_t smb = (size_t)(N * C + 2 * N * Cr) * sizeof(float);
  EGM_REQUIRE(smb <= 48 * 1024, EGM_E_SHAPE, "ca_mlp_bwd: N*C too large for one block");
  egm_launch(k_ca_mlp_bwd, 1, 256, smb, (cudaStream_t)stream, dca, ca, avg, mx, hid, w0, w2, dw0, dw2, davg, dmx, N, C, Cr);
  EGM_LAUNCH_CHECK("ca_mlp_bwd"); return EGM_OK;
}

// t = f + s * sa[m] * ca[n][c]
template <typename T, int V>
__global__ void k_fuse_mix_fwd(const T* __restrict__ f, const T* __restrict__ s, const float* __restrict__ sa, const float* __restrict__ ca, T* __restrict__ t,
                               long long M, long long HW, int CV) { egm_pdl_enter();
  const int C = CV * V; long long total = M * CV;
  const RowIndexer rix(CV, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long m; int c; rix(i, m, c); c *= V;
    FVec<V> a = ldv<V>(f + m * C + c), b = ldv<V>(s + m * C + c), cc = ldv<V>(ca + (m / HW) * C + c);
    float g = sa[m];
#pragma unroll
    for (int j = 0; j < V; ++j) a.v[j] = fmaf(b.v[j] * g, cc.v[j], a.v[j]);
    stv<V>(t + m * C + c, a);
  }
}
extern "C" int egm_fuse_mix_fwd(const void* f, const void* s, const float* sa, const float* ca, void* t, int dtype, int N, long long HW, int C, void* stream) {
  long long M = (long long)N * HW; if (M * C == 0) return EGM_OK;
  int v = egm_pick_vec(C);
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_fuse_mix_fwd<T, V>, egm_grid_for(M * (C / V), 256), 256, 0, (cudaStream_t)stream, (const T*)f, (const T*)s, sa, ca, (T*)t, M, HW, C / V))));
  EGM_LAUNCH_CHECK("fuse_mix_fwd"); return EGM_OK;
}
// ds = dt*sa*ca + dmm[m][0]/C + [c == amax[m]] * dmm[m][1]
template <typename T, int V>
__global__ void k_fuse_mix_bwd_s(const T* __restrict__ dt, const float* __restrict__ sa, const float* __restrict__ ca, const float* __restrict__ dmm,
                                 const unsigned char* __restrict__ amax, T* __restrict__ ds, long long M, long long HW, int CV) { egm_pdl_enter();
  const int C = CV * V; long long total = M * CV;
  const RowIndexer rix(CV, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long m; int c; rix(i, m, c); c *= V;
    FVec<V> a = ldv<V>(dt + m * C + c), cc = ldv<V>(ca + (m / HW) * C + c);
    float g = sa[m], d0 = dmm[m * 2] / (float)C, d1 = dmm[m * 2 + 1]; int am = amax[m];
#pragma unroll
    for (int j = 0; j < V; ++j) a.v[j] = a.v[j] * g * cc.v[j] + d0 + ((c + j) == am ? d1 : 0.f);
    stv<V>(ds + m * C + c, a);
  }
}
extern "C" int egm_fuse_mix_bwd_s(const void* dt, const float* sa, const float* ca, const float* dmm, const unsigned char* amax, void* ds, int dtype, int N,
                                  long long HW, int C, void* stream) {
  long long M = (long long)N * HW; if (M * C == 0) return EGM_OK;
  int v = egm_pick_vec(C);
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_fuse_mix_bwd_s<T, V>, egm_grid_for(M * (C / V), 256), 256, 0, (cudaStream_t)stream, (const T*)dt, sa, ca, dmm, amax, (T*)ds, M, HW, C / V))));
  EGM_LAUNCH_CHECK("fuse_mix_bwd_s"); return EGM_OK;
}
// df (+)= dt + davg[n][c]/HW + [p == arg[n][c]] * dmx[n][c]
template <typename T, int V>
__global__ void k_fuse_df_finish(T* __restrict__ df, const T* __restrict__ dt, const float* __restrict__ davg, const float* __restrict__ dmx,
                                 const int* __restrict__ arg, int accumulate, long long M, long long HW, int CV) { egm_pdl_enter();
  const int C = CV * V; long long total = M * CV; const float ih = 1.f / (float)HW;
  const RowIndexer rix(CV, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long m; int c; rix(i, m, c); c *= V; long long n = m / HW; int p = (int)(m - n * HW);
    FVec<V> a, b = ldv<V>(dt + m * C + c), da = ldv<V>(davg + n * C + c), dm = ldv<V>(dmx + n * C + c);
    if (accumulate) a = ldv<V>(df + m * C + c);
#pragma unroll
    for (int j = 0; j < V; ++j) a.v[j] = (accumulate ? a.v[j] : 0.f) + b.v[j] + da.v[j] * ih + (arg[n * C + c + j] == p ? dm.v[j] : 0.f);
    stv<V>(df + m * C + c, a);
  }
}
extern "C" int egm_fuse_df_finish(void* df, const void* dt, const float* davg, const float* dmx, const int* arg, int accumulate, int dtype, int N, long long HW, int C, void* stream) {
  long long M = (long long)N * HW; if (M * C == 0) return EGM_OK;
  int v = egm_pick_vec(C);
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_fuse_df_finish<T, V>, egm_grid_for(M * (C / V), 256), 256, 0, (cudaStream_t)stream, (T*)df, (const T*)dt, davg, dmx, arg, accumulate, M, HW, C / V))));
  EGM_LAUNCH_CHECK("fuse_df_finish"); return EGM_OK;
}
