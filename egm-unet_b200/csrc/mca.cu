// MCALayer (src/EGM-UNet.py:686-791) with its three MCAGates (:836-869) -- forward and backward.
// Eager PyTorch runs ~30 full-tensor passes, 4 permute copies and an FFT here;
// this file does:  stats (1 read) -> gates (tiny) -> apply (1 read + 1 write, 5x5 halo from L1/L2).
// frequency_enhancement is the identity scaled by 1.1, so the blend is
//   y = 0.51*u + 0.2*(max3x3 u - min3x3 u) + 0.2*avg3x3((u - avg3x3 u)^2) + 0.1*shuffle4(u),   u = x*(g_c+g_h+g_w)/3
#include "common.cuh"
#include <stdlib.h>

// Per-axis vectors are packed [h: N*H | w: N*W | c: N*C], each segment padded to 4 entries so the
// channel segment stays 16-byte aligned for vector loads.
static inline long long al4(long long x) { return (x + 3) & ~3LL; }
static inline long long mca_ow(int N, int H) { return al4((long long)N * H); }
static inline long long mca_oc(int N, int H, int W) { return mca_ow(N, H) + al4((long long)N * W); }
static inline long long mca_total(int N, int H, int W, int C) { return mca_oc(N, H, W) + al4((long long)N * C); }

// ------------------------------------------------------------------ three-axis sums of x (or of a*b)
// rowS[N*H][K], colS[N*W][K], chS[N*C][K] (double, zeroed by the caller side of this file).
// K=2: (sum x, sum x^2);  K=1 with b != null: sum a*b.
template <typename T, int V, int K>
__global__ void k_axis_sums(const T* __restrict__ a, const T* __restrict__ b, int H, int W, int C, double* __restrict__ rowS, double* __restrict__ colS,
                            double* __restrict__ chS) { egm_pdl_enter();
  extern __shared__ float sm[];           // col[W][K] | ch[rows][K][C] | red[32]
  const int CV = C / V;
  const int rows = blockDim.x / CV;
  const int cv = threadIdx.x % CV, rr = threadIdx.x / CV;
  float* col = sm; float* chp = sm + (size_t)W * K; float* red = chp + (size_t)rows * K * C;
  const int nh = blockIdx.x, n = nh / H;
  for (int i = threadIdx.x; i < W * K; i += blockDim.x) col[i] = 0.f;
  __syncthreads();
  const T* ap = a + (long long)nh * W * C;
  const T* bp = b ? b + (long long)nh * W * C : nullptr;
  float tot[K], chacc[K][V];
#pragma unroll
  for (int k = 0; k < K; ++k) { tot[k] = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) chacc[k][j] = 0.f; }
  const bool lanes_pow2 = CV <= 32 && (CV & (CV - 1)) == 0;      // CV threads of one pixel sit in one warp, aligned
  for (int w0 = 0; w0 < W; w0 += rows) {                          // uniform trip count (shuffles below need the full warp)
    const int w = w0 + rr;
    const bool valid = w < W;
    FVec<V> x;
#pragma unroll
    for (int j = 0; j < V; ++j) x.v[j] = 0.f;
    if (valid) {
      x = ldv<V>(ap + (long long)w * C + cv * V);
      if (bp) { FVec<V> y = ldv<V>(bp + (long long)w * C + cv * V);
#pragma unroll
        for (int j = 0; j < V; ++j) x.v[j] *= y.v[j]; }
    }
    float s[K];
#pragma unroll
    for (int k = 0; k < K; ++k) s[k] = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) { s[0] += x.v[j]; chacc[0][j] += x.v[j]; if (K == 2) { s[K - 1] += x.v[j] * x.v[j]; chacc[K - 1][j] += x.v[j] * x.v[j]; } }
#pragma unroll
    for (int k = 0; k < K; ++k) tot[k] += s[k];
    if (lanes_pow2) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        for (int o = CV >> 1; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if (valid && cv == 0) atomicAdd(&col[w * K + k], s[k]);
      }
    } else if (valid) {
#pragma unroll
      for (int k = 0; k < K; ++k) atomicAdd(&col[w * K + k], s[k]);
    }
  }
  if (rr < rows) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < V; ++j) chp[((size_t)rr * K + k) * C + cv * V + j] = chacc[k][j];
  }
#pragma unroll
  for (int k = 0; k < K; ++k) { float v = block_sum(tot[k], red); if (threadIdx.x == 0) rowS[(long long)nh * K + k] = (double)v; }
  __syncthreads();
  for (int i = threadIdx.x; i < W * K; i += blockDim.x) atomicAdd(colS + (long long)n * W * K + i, (double)col[i]);
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) {
    int k = i / C, c = i - k * C; float s = 0.f;
    for (int q = 0; q < rows; ++q) s += chp[((size_t)q * K + k) * C + c];
    atomicAdd(chS + ((long long)n * C + c) * K + k, (double)s);
  }
}
// Banded variant (the production shapes): the kernel above launches one block per image ROW and ends every block with W*K + C*K
// fp64 global atomics -- 2.3 M atomics for the 240^2 x 64 map -- and funnels the per-pixel sums through contended shared-memory
// atomics; it ran at 0.2-1.2 TB/s (137 us for the 29 MB 60^2 x 256 map).  Here a block owns a band of TH rows of one image, every
// thread keeps ONE 16-byte channel vector position (threads = PX pixels x CV vectors, 256 % CV == 0) and walks the band with fixed
// pixel slots, so the channel sums (chacc) and the column sums (colacc, one slot per pass over the row) live in registers for the
// whole band and the row sums are reduced once per row inside the warp; global fp64 atomics drop by the band height.
constexpr int AS_MAXP = 16;     // passes over one row: ceil(W / PX)
constexpr int AS_MAXTH = 8;
template <typename T, int K>
__global__ void __launch_bounds__(256) k_axis_sums_band(const T* __restrict__ a, const T* __restrict__ b, int H, int W, int C, int TH, int bands,
                                                        double* __restrict__ rowS, double* __restrict__ colS, double* __restrict__ chS) { egm_pdl_enter();
  constexpr int VE = 16 / (int)sizeof(T);
  extern __shared__ float sm[];           // col[W][K] | rowp[AS_MAXTH][8 warps][K] | chp[PX][K][C]
  const int CV = C / VE, PX = 256 / CV;
  const int cg = threadIdx.x % CV, px = threadIdx.x / CV, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* col = sm; float* rowp = col + (size_t)W * K; float* chp = rowp + AS_MAXTH * 8 * K;
  const int n = blockIdx.x / bands, h0 = (blockIdx.x - n * bands) * TH, h1 = min(h0 + TH, H);
  for (int i = threadIdx.x; i < W * K; i += 256) col[i] = 0.f;
  float chacc[K][VE], colacc[AS_MAXP][K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int j = 0; j < VE; ++j) chacc[k][j] = 0.f;
#pragma unroll
    for (int i = 0; i < AS_MAXP; ++i) colacc[i][k] = 0.f;
  }
  const int np = (W + PX - 1) / PX;
  for (int h = h0; h < h1; ++h) {
    const long long base = ((long long)n * H + h) * W * C + cg * VE;
    float rs[K];
#pragma unroll
    for (int k = 0; k < K; ++k) rs[k] = 0.f;
#pragma unroll
    for (int i = 0; i < AS_MAXP; ++i) {
      const int w = i * PX + px;
      if (i < np && w < W) {
        FVec<VE> x = ldv<VE>(a + base + (long long)w * C);
        if (b) { const FVec<VE> y = ldv<VE>(b + base + (long long)w * C);
#pragma unroll
          for (int j = 0; j < VE; ++j) x.v[j] *= y.v[j]; }
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < VE; ++j) {
          s0 += x.v[j]; chacc[0][j] += x.v[j];
          if (K == 2) { const float q = x.v[j] * x.v[j]; s1 += q; chacc[K - 1][j] += q; }
        }
        colacc[i][0] += s0; rs[0] += s0;
        if (K == 2) { colacc[i][K - 1] += s1; rs[K - 1] += s1; }
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) { const float v = warp_sum(rs[k]); if (lane == 0) rowp[((h - h0) * 8 + wid) * K + k] = v; }
  }
  __syncthreads();                                      // col zeroed, rowp complete
  // column sums: the CV threads of one pixel are consecutive lanes (CV <= 32: one aligned lane group; CV == 64: two warps)
  const int grp = CV < 32 ? CV : 32;
#pragma unroll
  for (int i = 0; i < AS_MAXP; ++i) {
    if (i < np) {                                       // block-uniform
      const int w = i * PX + px;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float v = colacc[i][k];
        for (int o = grp >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((cg & (grp - 1)) == 0 && w < W) { if (CV > 32) atomicAdd(&col[w * K + k], v); else col[w * K + k] = v; }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < VE; ++j) chp[((size_t)px * K + k) * C + cg * VE + j] = chacc[k][j];
  __syncthreads();
  for (int i = threadIdx.x; i < (h1 - h0) * K; i += 256) {
    const int hr = i / K, k = i - hr * K; float v = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += rowp[(hr * 8 + q) * K + k];
    rowS[((long long)n * H + h0 + hr) * K + k] = (double)v;
  }
  for (int i = threadIdx.x; i < W * K; i += 256) atomicAdd(colS + (long long)n * W * K + i, (double)col[i]);
  for (int i = threadIdx.x; i < K * C; i += 256) {
    const int k = i / C, c = i - k * C; float v = 0.f;
    for (int q = 0; q < PX; ++q) v += chp[((size_t)q * K + k) * C + c];
    atomicAdd(chS + ((long long)n * C + c) * K + k, (double)v);
  }
}
template <typename T, int K>
static bool launch_axis_sums_band(const T* a, const T* b, int N, int H, int W, int C, double* rowS, double* colS, double* chS, cudaStream_t st) {
  constexpr int VE = 16 / (int)sizeof(T);
  if (C % VE != 0 || ((uintptr_t)a & 15) || ((uintptr_t)b & 15)) return false;
  const int CV = C / VE;
  if (CV < 1 || CV > 64 || 256 % CV != 0) return false;
  const int PX = 256 / CV;
  if ((W + PX - 1) / PX > AS_MAXP) return false;
  const size_t smb = ((size_t)W * K + AS_MAXTH * 8 * K + (size_t)PX * K * C) * sizeof(float);
  if (smb > 48 * 1024) return false;
  long long th = (long long)N * H / (4LL * egm_num_sms()); if (th < 1) th = 1; if (th > AS_MAXTH) th = AS_MAXTH;
  const int bands = cdiv(H, th);
  egm_launch(k_axis_sums_band<T, K>, N * bands, 256, smb, st, a, b, H, W, C, (int)th, bands, rowS, colS, chS);
  return true;
}

template <typename T, int K>
static int launch_axis_sums(const T* a, const T* b, int N, int H, int W, int C, double* sums, cudaStream_t st) {
  double* rowS = sums; double* colS = sums + mca_ow(N, H) * K; double* chS = sums + mca_oc(N, H, W) * K;
  cudaMemsetAsync(sums, 0, sizeof(double) * K * (size_t)mca_total(N, H, W, C), st);
  if ((long long)N * H * W * C == 0) return EGM_OK;
  {
    static int band_off = -1;                           // EGM_MCA_SUMS_V1=1: the row-per-block kernel everywhere (A/B measurements)
    if (band_off < 0) { const char* e = getenv("EGM_MCA_SUMS_V1"); band_off = (e && e[0] == '1') ? 1 : 0; }
    if (!band_off && launch_axis_sums_band<T, K>(a, b, N, H, W, C, rowS, colS, chS, st)) return egm_check_launch("mca_axis_sums(band)");
  }
  int v = egm_pick_vec(C); if (v > 4) v = 4;
  int CV = C / v; EGM_REQUIRE(CV <= 512, EGM_E_SHAPE, "mca: C=%d too large", C);
  int rows = 256 / CV; if (rows < 1) rows = 1; int threads = CV * rows;
  size_t smb = ((size_t)W * K + (size_t)rows * K * C + 32) * sizeof(float);
  EGM_REQUIRE(smb <= 200 * 1024, EGM_E_SHAPE, "mca: row too wide for shared memory (W=%d C=%d)", W, C);
  EGM_DISPATCH_VEC(v, {
    if (V <= 4) {
      auto kern = k_axis_sums<T, (V <= 4 ? V : 4), K>;
      if (smb > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
      egm_launch(kern, N * H, threads, smb, st, a, b, H, W, C, rowS, colS, chS);
    }
  });
  return egm_check_launch("mca_axis_sums");
}
// sums layout (double): rowS[N*H][2] | colS[N*W][2] | chS[N*C][2]
// number of entries of one per-axis vector set (gates / avg / std / coef: floats; sums: doubles x2)
extern "C" long long egm_mca_vec_len(int N, int H, int W, int C) { return mca_total(N, H, W, C); }
extern "C" long long egm_mca_vec_off_w(int N, int H) { return mca_ow(N, H); }
extern "C" long long egm_mca_vec_off_c(int N, int H, int W) { return mca_oc(N, H, W); }
extern "C" int egm_mca_stats(const void* x, int dtype, int N, int H, int W, int C, double* sums, void* stream) {
  EGM_DISPATCH_DTYPE(dtype, return (launch_axis_sums<T, 2>((const T*)x, nullptr, N, H, W, C, sums, (cudaStream_t)stream)));
  return EGM_OK;
}
// K=1 layout: rowS[N*H] | colS[N*W] | chS[N*C]  of sum(a*b)
extern "C" int egm_mca_prod_sums(const void* a, const void* b, int dtype, int N, int H, int W, int C, double* sums, void* stream) {
  EGM_DISPATCH_DTYPE(dtype, return (launch_axis_sums<T, 1>((const T*)a, (const T*)b, N, H, W, C, sums, (cudaStream_t)stream)));
  return EGM_OK;
}

// ------------------------------------------------------------------ gates
// Per gate (axis length L, n_el elements behind each statistic): pre_i = (0.5+sig(w0))*avg_i + (0.5+sig(w1))*std_i,
// gate_i = sigmoid(sum_j k[j] * pre_{i+j-pad}).   stats out: avg, std (for backward).
struct GateDesc { const double* S; int L; double n_el; const float* w2; const float* kw; int ks; float* gate; float* avg; float* stdv; };

__device__ __forceinline__ void gate_stat(const GateDesc& d, int n, int i, float& avg, float& sd) {
  const double* s = d.S + ((long long)n * d.L + i) * 2;
  double m = s[0] / d.n_el, var = (s[1] - s[0] * m) / (d.n_el - 1.0);
  avg = (float)m; sd = (float)sqrt(var > 0.0 ? var : 0.0);
}
__global__ void k_mca_gates(GateDesc g0, GateDesc g1, GateDesc g2, int N) { egm_pdl_enter();
  GateDesc d = blockIdx.y == 0 ? g0 : (blockIdx.y == 1 ? g1 : g2);
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * d.L) return;
  int n = idx / d.L, i = idx - n * d.L;
  float a0 = 0.5f + sigmoidf_(d.w2[0]), a1 = 0.5f + sigmoidf_(d.w2[1]);
  int pad = (d.ks - 1) / 2; float acc = 0.f;
  for (int j = 0; j < d.ks; ++j) {
    int q = i + j - pad; if (q < 0 || q >= d.L) continue;
    float av, sd; gate_stat(d, n, q, av, sd);
    acc += d.kw[j] * (a0 * av + a1 * sd);
  }
  float av, sd; gate_stat(d, n, i, av, sd);
  d.avg[idx] = av; d.stdv[idx] = sd; d.gate[idx] = sigmoidf_(acc);
}
// gates/avg/std: fp32 arrays laid out [h: N*H | w: N*W | c: N*C]
extern "C" int egm_mca_gates(const double* sums, int N, int H, int W, int C, const float* w_h, const float* k_h, int ks_h, const float* w_w,
                             const float* k_w, int ks_w, const float* w_c, const float* k_c, int ks_c, float* gates, float* avg, float* stdv, void* stream) {
  long long oh = 0, ow = mca_ow(N, H), oc = mca_oc(N, H, W);
  GateDesc gh{sums, H, (double)W * C, w_h, k_h, ks_h, gates + oh, avg + oh, stdv + oh};
  GateDesc gw{sums + 2 * ow, W, (double)H * C, w_w, k_w, ks_w, gates + ow, avg + ow, stdv + ow};
  GateDesc gc{sums + 2 * oc, C, (double)H * W, w_c, k_c, ks_c, gates + oc, avg + oc, stdv + oc};
  int mx = H > W ? H : W; if (C > mx) mx = C;
  egm_launch(k_mca_gates, dim3(cdiv((long long)N * mx, 128), 3), 128, 0, (cudaStream_t)stream, gh, gw, gc, N);
  EGM_LAUNCH_CHECK("mca_gates"); return EGM_OK;
}

// ------------------------------------------------------------------ apply
struct McaGeom { int N, H, W, C; const float* gh; const float* gw; const float* gc; };

template <typename T, int V>
__device__ __forceinline__ FVec<V> mca_u(const T* x, const McaGeom& g, int n, int h, int w, int c, const FVec<V>& gcv) {
  FVec<V> a = ldv<V>(x + (((long long)n * g.H + h) * g.W + w) * g.C + c);
  float s = g.gh[n * g.H + h] + g.gw[n * g.W + w];
#pragma unroll
  for (int j = 0; j < V; ++j) a.v[j] = a.v[j] * ((gcv.v[j] + s) * (1.f / 3.f));
  return a;
}
// Three streaming passes (each 16-byte vectorised, 9-point stencils served by L1/L2) instead of one 25-point gather per output:
//   U : u  = x * (g_c + g_h + g_w) / 3                                  (1 read, 1 write)
//   D : d2 = (u - avg3x3(u))^2                                          (9 cached reads, 1 write)
//   O : y  = 0.51 u + 0.2 (max3x3 u - min3x3 u) + 0.2 avg3x3(d2) + 0.1 shuffle4(u)  (+ arg-max/min byte map)
template <typename T, int V>
__global__ void k_mca_u(const T* __restrict__ x, T* __restrict__ u, McaGeom g) { egm_pdl_enter();
  const int CV = g.C / V;
  long long total = (long long)g.N * g.H * g.W * CV;
  const NhwcIndexer ix(CV, g.W, g.H, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const Nhwc4 e = ix(i); const int c = e.cv * V, w = e.w, h = e.h, n = e.n; const long long p = e.p;
    FVec<V> gcv = ldv<V>(g.gc + n * g.C + c);
    FVec<V> o = mca_u<T, V>(x, g, n, h, w, c, gcv);
    stv<V>(u + p * g.C + c, o);
  }
}
template <typename T, int V>
__global__ void k_mca_d2(const T* __restrict__ u, T* __restrict__ d2, int N, int H, int W, int CV) { egm_pdl_enter();
  const int C = CV * V; long long total = (long long)N * H * W * CV;
  const NhwcIndexer ix(CV, W, H, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const Nhwc4 e = ix(i); const int c = e.cv * V, w = e.w, h = e.h; const long long p = e.p;
    // unconditional neighbour loads (clamped + masked): all 9 in flight at once instead of 9 dependent round trips
    const bool vr[3] = {h > 0, true, h < H - 1}, vc[3] = {w > 0, true, w < W - 1};
    const long long ro[3] = {vr[0] ? -(long long)W : 0, 0, vr[2] ? (long long)W : 0}, cofs[3] = {vc[0] ? -1 : 0, 0, vc[2] ? 1 : 0};
    FVec<V> t[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) t[a][b] = ldv<V>(u + (p + ro[a] + cofs[b]) * C + c);
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
#pragma unroll
    for (int j = 0; j < V; ++j) { float d = ctr.v[j] - s.v[j] * (1.f / 9.f); o.v[j] = d * d; }
    stv<V>(d2 + p * C + c, o);
  }
}
template <typename T, int V>
__global__ void k_mca_out(const T* __restrict__ u, const T* __restrict__ d2, T* __restrict__ y, unsigned char* __restrict__ idx, int N, int H, int W, int CV) { egm_pdl_enter();
  const int C = CV * V; long long total = (long long)N * H * W * CV;
  const NhwcIndexer ix(CV, W, H, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const Nhwc4 e = ix(i); const int c = e.cv * V, w = e.w, h = e.h; const long long p = e.p;
    float mx[V], mn[V], var[V], uc[V]; int amx[V], amn[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { mx[j] = -INFINITY; mn[j] = INFINITY; var[j] = 0.f; uc[j] = 0.f; amx[j] = 4; amn[j] = 4; }
    const bool vr[3] = {h > 0, true, h < H - 1}, vc[3] = {w > 0, true, w < W - 1};
    const long long ro[3] = {vr[0] ? -(long long)W : 0, 0, vr[2] ? (long long)W : 0}, cofs[3] = {vc[0] ? -1 : 0, 0, vc[2] ? 1 : 0};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      FVec<V> tu[3], td[3];                              // one row of neighbours at a time: 6 unconditional loads in flight
#pragma unroll
      for (int b = 0; b < 3; ++b) { const long long off = (p + ro[a] + cofs[b]) * C + c; tu[b] = ldv<V>(u + off); td[b] = ldv<V>(d2 + off); }
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        if (!(vr[a] && vc[b])) continue;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          if (tu[b].v[j] > mx[j]) { mx[j] = tu[b].v[j]; amx[j] = a * 3 + b; }
          if (tu[b].v[j] < mn[j]) { mn[j] = tu[b].v[j]; amn[j] = a * 3 + b; }
          var[j] += td[b].v[j];
          if (a == 1 && b == 1) uc[j] = tu[b].v[j];
        }
      }
    }
    FVec<V> o;
    const T* up = u + p * C;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      int cc = c + j, src = (cc & 3) * (C >> 2) + (cc >> 2);          // channel shuffle (groups = 4)
      o.v[j] = 0.51f * uc[j] + 0.2f * (mx[j] - mn[j]) + 0.2f * var[j] * (1.f / 9.f) + 0.1f * ldf(up + src);
    }
    stv<V>(y + p * C + c, o);
    if (idx) {
      unsigned char codes[V];
#pragma unroll
      for (int j = 0; j < V; ++j) codes[j] = (unsigned char)(amx[j] | (amn[j] << 4));
      if constexpr (V == 8) *reinterpret_cast<uint2*>(idx + p * C + c) = *reinterpret_cast<uint2*>(codes);
      else *reinterpret_cast<unsigned int*>(idx + p * C + c) = *reinterpret_cast<unsigned int*>(codes);
    }
  }
}
// u_scratch, d2_scratch: two tensors shaped like x (caller-owned); u_scratch holds u = x_out of the reference on return
extern "C" int egm_mca_apply(const void* x, const float* gates, void* y, unsigned char* argidx, void* u_scratch, void* d2_scratch, int dtype, int N, int H,
                             int W, int C, void* stream) {
  EGM_REQUIRE(C % 4 == 0, EGM_E_SHAPE, "mca: C %% 4 != 0");
  long long total = (long long)N * H * W * C;
  if (total == 0) return EGM_OK;
  McaGeom g{N, H, W, C, gates, gates + mca_ow(N, H), gates + mca_oc(N, H, W)};
  cudaStream_t st = (cudaStream_t)stream;
  int v = egm_pick_vec(C); if (v > 8) v = 8; if (v < 4) v = 4;
  EGM_DISPATCH_DTYPE(dtype, {
    if (v == 8) {
      egm_launch(k_mca_u<T, 8>, egm_grid_for(total / 8, 256), 256, 0, st, (const T*)x, (T*)u_scratch, g);
      egm_launch(k_mca_d2<T, 8>, egm_grid_for(total / 8, 256), 256, 0, st, (const T*)u_scratch, (T*)d2_scratch, N, H, W, C / 8);
      egm_launch(k_mca_out<T, 8>, egm_grid_for(total / 8, 256), 256, 0, st, (const T*)u_scratch, (const T*)d2_scratch, (T*)y, argidx, N, H, W, C / 8);
    } else {
      egm_launch(k_mca_u<T, 4>, egm_grid_for(total / 4, 256), 256, 0, st, (const T*)x, (T*)u_scratch, g);
      egm_launch(k_mca_d2<T, 4>, egm_grid_for(total / 4, 256), 256, 0, st, (const T*)u_scratch, (T*)d2_scratch, N, H, W, C / 4);
      egm_launch(k_mca_out<T, 4>, egm_grid_for(total / 4, 256), 256, 0, st, (const T*)u_scratch, (const T*)d2_scratch, (T*)y, argidx, N, H, W, C / 4);
    }
  });
  EGM_LAUNCH_CHECK("mca_apply"); return EGM_OK;
}

// ------------------------------------------------------------------ backward
// pass E:  E[q] = 2*(u[q]-avg3(u)[q]) * (0.2/9) * sum_{p in N(q)} dy[p]
template <typename T, int V>
__global__ void __launch_bounds__(256) k_mca_bwd_e(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ E, McaGeom g) { egm_pdl_enter();
  const int CV = g.C / V;
  long long total = (long long)g.N * g.H * g.W * CV;
  const NhwcIndexer ix(CV, g.W, g.H, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const Nhwc4 e = ix(i); const int c = e.cv * V, w = e.w, h = e.h, n = e.n; const long long p = e.p;
    FVec<V> gcv = ldv<V>(g.gc + n * g.C + c), su, sd, uc;
#pragma unroll
    for (int j = 0; j < V; ++j) { su.v[j] = 0.f; sd.v[j] = 0.f; uc.v[j] = 0.f; }
    const bool vr[3] = {h > 0, true, h < g.H - 1}, vc[3] = {w > 0, true, w < g.W - 1};
    const int hr[3] = {vr[0] ? h - 1 : h, h, vr[2] ? h + 1 : h}, wc[3] = {vc[0] ? w - 1 : w, w, vc[2] ? w + 1 : w};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      FVec<V> tx[3], td[3];                              // one row of neighbours at a time, loads unconditional (clamped + masked)
      const float gh = g.gh[n * g.H + hr[a]];
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const long long off = (((long long)n * g.H + hr[a]) * g.W + wc[b]) * g.C + c;
        tx[b] = ldv<V>(x + off); td[b] = ldv<V>(dy + off);
      }
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        if (!(vr[a] && vc[b])) continue;
        const float sgate = gh + g.gw[n * g.W + wc[b]];
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float uu = tx[b].v[j] * ((gcv.v[j] + sgate) * (1.f / 3.f));
          su.v[j] += uu; sd.v[j] += td[b].v[j]; if (a == 1 && b == 1) uc.v[j] = uu;
        }
      }
    }
    FVec<V> o;
#pragma unroll
    for (int j = 0; j < V; ++j) o.v[j] = 2.f * (uc.v[j] - su.v[j] * (1.f / 9.f)) * (0.2f / 9.f) * sd.v[j];
    stv<V>(E + p * g.C + c, o);
  }
}
// pass U: du[q] = 0.51 dy[q] + 0.1 unshuffle(dy)[q] + 0.2 (sum_p dy[p][argmax_p==q] - sum_p dy[p][argmin_p==q]) + E[q] - avg3(E)[q]
template <typename T, int V>
__global__ void __launch_bounds__(256) k_mca_bwd_du(const T* __restrict__ dy, const unsigned char* __restrict__ idx, const T* __restrict__ E,
                                                    T* __restrict__ du, McaGeom g) { egm_pdl_enter();
  const int CV = g.C / V;
  long long total = (long long)g.N * g.H * g.W * CV;
  const NhwcIndexer ix(CV, g.W, g.H, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const Nhwc4 e = ix(i); const int c = e.cv * V, w = e.w, h = e.h, n = e.n; const long long p = e.p;
    FVec<V> o, se;
#pragma unroll
    for (int j = 0; j < V; ++j) { o.v[j] = 0.f; se.v[j] = 0.f; }
    const bool vr[3] = {h > 0, true, h < g.H - 1}, vc[3] = {w > 0, true, w < g.W - 1};
    const int hr[3] = {vr[0] ? h - 1 : h, h, vr[2] ? h + 1 : h}, wc[3] = {vc[0] ? w - 1 : w, w, vc[2] ? w + 1 : w};
#pragma unroll
    for (int a = -1; a <= 1; ++a) {
      FVec<V> dv[3], ev[3]; unsigned char codes[3][V];   // one row of neighbours at a time, loads unconditional (clamped + masked)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const long long off = (((long long)n * g.H + hr[a + 1]) * g.W + wc[b]) * g.C + c;
        dv[b] = ldv<V>(dy + off); ev[b] = ldv<V>(E + off);
        if constexpr (V == 8) *reinterpret_cast<uint2*>(codes[b]) = *reinterpret_cast<const uint2*>(idx + off);
        else if constexpr (V == 4) *reinterpret_cast<unsigned int*>(codes[b]) = *reinterpret_cast<const unsigned int*>(idx + off);
        else { for (int j = 0; j < V; ++j) codes[b][j] = idx[off + j]; }
      }
#pragma unroll
      for (int b = -1; b <= 1; ++b) {
        if (!(vr[a + 1] && vc[b + 1])) continue;
        const int want = (1 - a) * 3 + (1 - b);          // position of (h,w) inside the window centred at (hh,ww)
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const unsigned char code = codes[b + 1][j];
          const float d = dv[b + 1].v[j];
          float r = ((code & 15) == want ? d : 0.f) - ((code >> 4) == want ? d : 0.f);
          o.v[j] += 0.2f * r; se.v[j] += ev[b + 1].v[j];
          if (a == 0 && b == 0) o.v[j] += 0.51f * d + ev[b + 1].v[j];
        }
      }
    }
    const T* dyp = dy + p * g.C;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      int cc = c + j, q4 = g.C >> 2; int a4 = cc / q4, b4 = cc - a4 * q4;       // cc = a4*(C/4) + b4  ->  read by shuffled channel b4*4 + a4
      o.v[j] += 0.1f * ldf(dyp + b4 * 4 + a4) - se.v[j] * (1.f / 9.f);
    }
    stv<V>(du + p * g.C + c, o);
  }
}
extern "C" int egm_mca_bwd_du(const void* x, const float* gates, const void* dy, const unsigned char* argidx, void* E_scratch, void* du, int dtype, int N,
                              int H, int W, int C, void* stream) {
  long long total = (long long)N * H * W * C;
  if (total == 0) return EGM_OK;
  McaGeom g{N, H, W, C, gates, gates + mca_ow(N, H), gates + mca_oc(N, H, W)};
  cudaStream_t st = (cudaStream_t)stream;
  EGM_DISPATCH_DTYPE(dtype, {
    if (C % 8 == 0) {
      egm_launch(k_mca_bwd_e<T, 8>, egm_grid_for(total / 8, 256, 16), 256, 0, st, (const T*)x, (const T*)dy, (T*)E_scratch, g);
      egm_launch(k_mca_bwd_du<T, 8>, egm_grid_for(total / 8, 256, 16), 256, 0, st, (const T*)dy, argidx, (const T*)E_scratch, (T*)du, g);
    } else {
      egm_launch(k_mca_bwd_e<T, 4>, egm_grid_for(total / 4, 256, 16), 256, 0, st, (const T*)x, (const T*)dy, (T*)E_scratch, g);
      egm_launch(k_mca_bwd_du<T, 4>, egm_grid_for(total / 4, 256, 16), 256, 0, st, (const T*)dy, argidx, (const T*)E_scratch, (T*)du, g);
    }
  });
  EGM_LAUNCH_CHECK("mca_bwd_du"); return EGM_OK;
}

// gate backward: from dG (sum du*x/3 per row/col/channel) -> coefficients a, b of the statistic gradients and the gate parameter grads.
//   dpre_i = dG_i * g_i (1-g_i);   dout_q = sum_j k[j] dpre_{q-j+pad};   davg = (0.5+sig w0) dout, dstd = (0.5+sig w1) dout
//   a_i = davg/n - dstd*avg/((n-1) std),  b_i = dstd/((n-1) std)    (d std / d x = (x-avg)/((n-1) std))
struct GateBwd { const double* dG; int L; double n_el; const float* w2; const float* kw; int ks; const float* gate; const float* avg; const float* stdv;
                 float* a; float* b; float* dw2; float* dkw; };
__global__ void __launch_bounds__(256) k_mca_gates_bwd(GateBwd g0, GateBwd g1, GateBwd g2, int N) { egm_pdl_enter();
  __shared__ float red[32];
  GateBwd d = blockIdx.x == 0 ? g0 : (blockIdx.x == 1 ? g1 : g2);
  const int pad = (d.ks - 1) / 2, tot = N * d.L;
  const float s0 = sigmoidf_(d.w2[0]), s1 = sigmoidf_(d.w2[1]);
  const float a0 = 0.5f + s0, a1 = 0.5f + s1;
  float gw0 = 0.f, gw1 = 0.f, gk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gk[j] = 0.f;
  for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) {
    int n = idx / d.L, i = idx - n * d.L;
    // dout_i = sum_j k[j] * dpre_{i - j + pad}
    float dout = 0.f;
    for (int j = 0; j < d.ks; ++j) {
      int q = i - j + pad; if (q < 0 || q >= d.L) continue;
      float gq = d.gate[n * d.L + q];
      dout += d.kw[j] * ((float)d.dG[n * d.L + q] * (1.f / 3.f)) * gq * (1.f - gq);
    }
    float av = d.avg[idx], sd = d.stdv[idx];
    float davg = a0 * dout, dstd = a1 * dout;
    gw0 += dout * av * s0 * (1.f - s0); gw1 += dout * sd * s1 * (1.f - s1);
    float inv = sd > 0.f ? 1.f / ((float)(d.n_el - 1.0) * sd) : 0.f;
    d.a[idx] = davg / (float)d.n_el - dstd * av * inv;
    d.b[idx] = dstd * inv;
    // dk[j] = sum_i dpre_i * pre_{i + j - pad}
    float gi = d.gate[idx]; float dpre = ((float)d.dG[idx] * (1.f / 3.f)) * gi * (1.f - gi);
    for (int j = 0; j < d.ks; ++j) {
      int q = i + j - pad; if (q < 0 || q >= d.L) continue;
      gk[j] += dpre * (a0 * d.avg[n * d.L + q] + a1 * d.stdv[n * d.L + q]);
    }
  }
  float v = block_sum(gw0, red); if (threadIdx.x == 0) d.dw2[0] = v;
  v = block_sum(gw1, red); if (threadIdx.x == 0) d.dw2[1] = v;
  for (int j = 0; j < d.ks; ++j) { v = block_sum(gk[j], red); if (threadIdx.x == 0) d.dkw[j] = v; }
}
// dG: output of egm_mca_prod_sums(du, x) (UNscaled: the 1/3 of u = x*s/3 is applied here). coef_a/coef_b laid out like gates.
extern "C" int egm_mca_gates_bwd(const double* dG, int N, int H, int W, int C, const float* gates, const float* avg, const float* stdv, const float* w_h,
                                 const float* k_h, int ks_h, const float* w_w, const float* k_w, int ks_w, const float* w_c, const float* k_c, int ks_c,
                                 float* coef_a, float* coef_b, float* dw_h, float* dk_h, float* dw_w, float* dk_w, float* dw_c, float* dk_c, void* stream) {
  EGM_REQUIRE(ks_h <= 8 && ks_w <= 8 && ks_c <= 8, EGM_E_SHAPE, "mca: gate kernel > 8");
  long long oh = 0, ow = mca_ow(N, H), oc = mca_oc(N, H, W);
  GateBwd gh{dG + oh, H, (double)W * C, w_h, k_h, ks_h, gates + oh, avg + oh, stdv + oh, coef_a + oh, coef_b + oh, dw_h, dk_h};
  GateBwd gw{dG + ow, W, (double)H * C, w_w, k_w, ks_w, gates + ow, avg + ow, stdv + ow, coef_a + ow, coef_b + ow, dw_w, dk_w};
  GateBwd gc{dG + oc, C, (double)H * W, w_c, k_c, ks_c, gates + oc, avg + oc, stdv + oc, coef_a + oc, coef_b + oc, dw_c, dk_c};
  egm_launch(k_mca_gates_bwd, 3, 256, 0, (cudaStream_t)stream, gh, gw, gc, N);
  EGM_LAUNCH_CHECK("mca_gates_bwd"); return EGM_OK;
}

// final: dx = du * s/3 + (a_h + a_w + a_c) + (b_h + b_w + b_c) * x
template <typename T, int V>
__global__ void k_mca_bwd_dx(const T* __restrict__ du, const T* __restrict__ x, const float* __restrict__ gates, const float* __restrict__ ca,
                             const float* __restrict__ cb, T* __restrict__ dx, int N, int H, int W, int CV, long long ow, long long oc) { egm_pdl_enter();
  const int C = CV * V; long long total = (long long)N * H * W * CV;
  const NhwcIndexer ix(CV, W, H, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const Nhwc4 e = ix(i); const int c = e.cv * V, w = e.w, h = e.h, n = e.n; const long long p = e.p;
    FVec<V> d = ldv<V>(du + p * C + c), xv = ldv<V>(x + p * C + c), gc = ldv<V>(gates + oc + n * C + c), ac = ldv<V>(ca + oc + n * C + c),
            bc = ldv<V>(cb + oc + n * C + c), o;
    float gs = gates[n * H + h] + gates[ow + n * W + w], as = ca[n * H + h] + ca[ow + n * W + w], bs = cb[n * H + h] + cb[ow + n * W + w];
#pragma unroll
    for (int j = 0; j < V; ++j) o.v[j] = d.v[j] * ((gc.v[j] + gs) * (1.f / 3.f)) + (ac.v[j] + as) + (bc.v[j] + bs) * xv.v[j];
    stv<V>(dx + p * C + c, o);
  }
}
extern "C" int egm_mca_bwd_dx(const void* du, const void* x, const float* gates, const float* coef_a, const float* coef_b, void* dx, int dtype, int N,
                              int H, int W, int C, void* stream) {
  long long total = (long long)N * H * W * C;
  if (total == 0) return EGM_OK;
  int v = egm_pick_vec(C); if (v > 4) v = 4;
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_mca_bwd_dx<T, (V > 4 ? 4 : V)>, egm_grid_for(total / v, 256), 256, 0, (cudaStream_t)stream, 
      (const T*)du, (const T*)x, gates, coef_a, coef_b, (T*)dx, N, H, W, C / v, mca_ow(N, H), mca_oc(N, H, W)))));
  EGM_LAUNCH_CHECK("mca_bwd_dx"); return EGM_OK;
}
