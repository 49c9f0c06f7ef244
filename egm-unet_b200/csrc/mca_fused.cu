// MCALayer (src/EGM-UNet.py:686-791), fused apply and fused backward: ONE pass over x per direction.
//
//   u   = x * (g_c + g_h + g_w) / 3                                   (the three MCAGates, :755-768)
//   y   = 0.51 u + 0.2 (max3x3 u - min3x3 u) + 0.2 avg3x3((u - avg3x3 u)^2) + 0.1 shuffle4(u)      (:774-790, FFT branch == 1.1 u)
//
// Round 1 ran three stencil passes (u, d^2, out) that each fetched 9 neighbours per output through L1 (1.3-1.7 TB/s) and stored u and
// d^2 in bf16.  Here a CTA owns a column band of 28 output pixels x 64 channels and walks DOWN the image: per input row every thread
// (one pixel column, 4 channels) loads its 8 bytes of x once (prefetched 3 rows ahead), exchanges u / d^2 with its two horizontal
// neighbours through a shared-memory ring, and keeps the vertical 3-row windows (row sums, row max/min with first-occurrence argument)
// in registers.  The 5x5 receptive field of y costs one 8-byte global load, ~12 16-byte shared-memory accesses and ONE barrier per
// row; u and d^2 never leave the SM and stay fp32.  HBM traffic = read x + write y (+ 1 byte/element arg map in training) -- the
// algorithmic minimum of SURVEY.md s8 a4.
//
// Backward (same walk): from x, dy and the arg map, du = 0.51 dy + 0.1 unshuffle(dy) + 0.2 (routed max - routed min) + E - avg3x3(E),
// E = 2 (u - avg3x3 u) (0.2/9) box3x3(dy); the gate gradients need global sums of du*x, taken by egm_mca_prod_sums afterwards.
#include "common.cuh"

// Tunables (tools/build_mca_variants.sh + tools/mca_variants.py measured them on the cfg2 shapes, profiles/mca_variants_r2.txt):
// 32 channels x 32 columns per CTA = 256 threads, >= 2 CTAs per SM was the fastest of the five configurations tried -- these kernels
// are instruction-issue bound (~130 instructions per element in the forward, the arg-tracking 3x3 max/min being the largest part), so
// what matters is how many independent barrier groups an SM can interleave, not bytes in flight.
#ifndef MF_CC
#define MF_CC 32
#endif
#ifndef MF_MINB_FWD
#define MF_MINB_FWD 2
#endif
#ifndef MF_MINB_BWD
#define MF_MINB_BWD 2
#endif
namespace mf {
constexpr int CC = MF_CC;         // channels per CTA
constexpr int V = 4;              // channels per thread
constexpr int CV = CC / V;        // 16 channel vectors
constexpr int TX = 32;            // pixel columns per CTA (incl. 2 + 2 halo)
constexpr int HALO = 2;
constexpr int TW = TX - 2 * HALO; // 28 output columns
constexpr int THREADS = TX * CV;  // 512
constexpr int ROW = TX * CC;      // floats per shared-memory row
constexpr int PF = 3;             // rows of x in flight per thread
}  // namespace mf

static inline long long mf_al4(long long x) { return (x + 3) & ~3LL; }

struct McaFusedParams {
  int N, H, W, C, TH;
  const float* gh; const float* gw; const float* gc;     // gate vectors [N*H], [N*W], [N*C]
};

template <typename T> struct RawV4;
template <> struct RawV4<float> { float4 v; };
template <> struct RawV4<__nv_bfloat16> { uint2 v; };
__device__ __forceinline__ void raw_zero(RawV4<float>& r) { r.v = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void raw_zero(RawV4<__nv_bfloat16>& r) { r.v = make_uint2(0u, 0u); }
__device__ __forceinline__ void raw_load(RawV4<float>& r, const float* p) { r.v = *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void raw_load(RawV4<__nv_bfloat16>& r, const __nv_bfloat16* p) { r.v = *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ void raw_get(const RawV4<float>& r, float (&f)[4]) { f[0] = r.v.x; f[1] = r.v.y; f[2] = r.v.z; f[3] = r.v.w; }
__device__ __forceinline__ void raw_get(const RawV4<__nv_bfloat16>& r, float (&f)[4]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.v);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ float tof(float v) { return v; }
__device__ __forceinline__ float tof(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void ld4(const float* p, float (&f)[4]) { float4 v = *reinterpret_cast<const float4*>(p); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
__device__ __forceinline__ void st4(float* p, const float (&f)[4]) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }

// ===================================================================================== forward
// smem: U[5][TX][CC] (ring over rows: r, r-1, r-2, r-3 are read while r+1 is written) | D[2][TX][CC]
template <typename T>
__global__ void __launch_bounds__(mf::THREADS, MF_MINB_FWD) k_mca_fwd(const T* __restrict__ x, T* __restrict__ y, unsigned char* __restrict__ idx, McaFusedParams g) { egm_pdl_enter();
  using namespace mf;
  extern __shared__ float sm[];
  float* U = sm;
  float* D = sm + 5 * ROW;
  const int tid = threadIdx.x, cv = tid & (CV - 1), tx = tid / CV;
  const int chunks = g.C / CC;
  const int n = blockIdx.z / chunks, c0 = (blockIdx.z - n * chunks) * CC + cv * V;
  const int wx = blockIdx.x * TW - HALO + tx;
  const int R0 = blockIdx.y * g.TH, R1 = min(R0 + g.TH, g.H);
  const bool colv = wx >= 0 && wx < g.W, colL = wx >= 1 && wx - 1 < g.W, colR = wx + 1 >= 0 && wx + 1 < g.W;
  const bool outcol = tx >= HALO && tx < TX - HALO && wx < g.W;
  const int own = tx * CC + cv * V;
  const int left = (tx > 0 ? own - CC : own), right = (tx < TX - 1 ? own + CC : own);
  float gcv[V], gcs[V];
  const int q4 = g.C >> 2;
#pragma unroll
  for (int j = 0; j < V; ++j) { gcv[j] = g.gc[n * g.C + c0 + j]; gcs[j] = g.gc[n * g.C + j * q4 + (c0 >> 2)]; }   // shuffle4: out ch c reads ch (c&3)*(C/4) + (c>>2)
  const float gwv = colv ? g.gw[n * g.W + wx] : 0.f;
  const long long img = (long long)n * g.H * g.W;
  const int wxc = colv ? wx : 0;

  // prefetch queue
  RawV4<T> xq[PF]; float ghq[PF];
  const int rstart = R0 - 2, rend = R1 + 2;
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    const int rr = rstart + i; const bool ok = rr >= 0 && rr < g.H && rr <= R1 + 1;
    raw_zero(xq[i]); ghq[i] = 0.f;
    if (ok) { ghq[i] = g.gh[n * g.H + rr]; if (colv) raw_load(xq[i], x + (img + (long long)rr * g.W + wxc) * g.C + c0); }
  }
  float hsu1[V], hsu2[V], hsd3[V], hsd4[V], mx3[V], mx4[V], mn3[V], mn4[V], d2p[V];
  unsigned ax3 = 0, ax4 = 0;
  T sq[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { hsu1[j] = hsu2[j] = hsd3[j] = hsd4[j] = d2p[j] = 0.f; mx3[j] = mx4[j] = -INFINITY; mn3[j] = mn4[j] = INFINITY; sq[j] = T(0.f); }
  int s0 = 0, s1 = 4, s2 = 3, s3 = 2, s4 = 1;          // ring slots of rows r, r-1, r-2, r-3 and the free one

  for (int r = rstart; r <= rend; ++r) {
    // ---- (1) u[r] from the prefetched row; queue the load of row r + PF
    float u[V];
    {
      float xv[V]; raw_get(xq[0], xv);
      const bool rowv = r >= 0 && r < g.H;
      const float s = (ghq[0] + gwv) * (1.f / 3.f);
#pragma unroll
      for (int j = 0; j < V; ++j) u[j] = (rowv && colv) ? xv[j] * fmaf(gcv[j], 1.f / 3.f, s) : 0.f;
#pragma unroll
      for (int i = 0; i < PF - 1; ++i) { xq[i] = xq[i + 1]; ghq[i] = ghq[i + 1]; }
      const int rr = r + PF; const bool ok = rr >= 0 && rr < g.H && rr <= R1 + 1;
      raw_zero(xq[PF - 1]); ghq[PF - 1] = 0.f;
      if (ok) { ghq[PF - 1] = g.gh[n * g.H + rr]; if (colv) raw_load(xq[PF - 1], x + (img + (long long)rr * g.W + wxc) * g.C + c0); }
    }
    st4(U + s0 * ROW + own, u);
    st4(D + (r & 1) * ROW + own, d2p);                 // d2[r-2], computed in the previous iteration
    __syncthreads();
    // ---- (2) horizontal sum of u[r]; d2[r-1] = (u - avg3x3 u)^2
    float hsu0[V], d2n[V];
    {
      float a[V], b[V], c1[V];
      ld4(U + s0 * ROW + left, a); ld4(U + s0 * ROW + right, b); ld4(U + s1 * ROW + own, c1);
      const bool v1 = r - 1 >= 0 && r - 1 < g.H && colv;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        hsu0[j] = (tx > 0 ? a[j] : 0.f) + u[j] + (tx < TX - 1 ? b[j] : 0.f);
        const float d = c1[j] - (hsu2[j] + hsu1[j] + hsu0[j]) * (1.f / 9.f);
        d2n[j] = v1 ? d * d : 0.f;
      }
    }
    // ---- (3) horizontal sum of d2[r-2]
    float hsd2[V];
    {
      float a[V], b[V];
      ld4(D + (r & 1) * ROW + left, a); ld4(D + (r & 1) * ROW + right, b);
#pragma unroll
      for (int j = 0; j < V; ++j) hsd2[j] = (tx > 0 ? a[j] : 0.f) + d2p[j] + (tx < TX - 1 ? b[j] : 0.f);
    }
    // ---- (4) row extremes of row r-2 (first occurrence wins: strict comparisons in column order)
    float mx2[V], mn2[V]; unsigned ax2 = 0;
    {
      float a[V], b[V], c[V];
      ld4(U + s2 * ROW + left, a); ld4(U + s2 * ROW + own, b); ld4(U + s2 * ROW + right, c);
      const bool rv = r - 2 >= 0 && r - 2 < g.H;
      const bool vl = rv && colL && tx > 0, vc = rv && colv, vr = rv && colR && tx < TX - 1;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float m = vl ? a[j] : -INFINITY; unsigned am = 0;
        const float bm = vc ? b[j] : -INFINITY, cm = vr ? c[j] : -INFINITY;
        if (bm > m) { m = bm; am = 1; }
        if (cm > m) { m = cm; am = 2; }
        float q = vl ? a[j] : INFINITY; unsigned aq = 0;
        const float bq = vc ? b[j] : INFINITY, cq = vr ? c[j] : INFINITY;
        if (bq < q) { q = bq; aq = 1; }
        if (cq < q) { q = cq; aq = 2; }
        mx2[j] = m; mn2[j] = q; ax2 |= (am | (aq << 2)) << (4 * j);
      }
    }
    // ---- (5) output row o = r - 3
    const int o = r - 3;
    if (o >= R0 && o < R1 && outcol) {
      float uo[V]; ld4(U + s3 * ROW + own, uo);
      const float so = (g.gh[n * g.H + o] + gwv) * (1.f / 3.f);
      FVec<V> out; unsigned codes = 0;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float M = mx4[j]; unsigned aM = (ax4 >> (4 * j)) & 3u;
        if (mx3[j] > M) { M = mx3[j]; aM = 3u + ((ax3 >> (4 * j)) & 3u); }
        if (mx2[j] > M) { M = mx2[j]; aM = 6u + ((ax2 >> (4 * j)) & 3u); }
        float Q = mn4[j]; unsigned aQ = (ax4 >> (4 * j + 2)) & 3u;
        if (mn3[j] < Q) { Q = mn3[j]; aQ = 3u + ((ax3 >> (4 * j + 2)) & 3u); }
        if (mn2[j] < Q) { Q = mn2[j]; aQ = 6u + ((ax2 >> (4 * j + 2)) & 3u); }
        const float var = (hsd4[j] + hsd3[j] + hsd2[j]) * (1.f / 9.f);
        const float shuf = tof(sq[j]) * fmaf(gcs[j], 1.f / 3.f, so);
        out.v[j] = 0.51f * uo[j] + 0.2f * (M - Q) + 0.2f * var + 0.1f * shuf;
        codes |= (aM | (aQ << 4)) << (8 * j);
      }
      const long long e = (img + (long long)o * g.W + wx) * g.C + c0;
      stv<V>(y + e, out);
      if (idx) *reinterpret_cast<unsigned*>(idx + e) = codes;
    }
    // ---- rotate the windows; fetch the shuffle sources of the next output row
#pragma unroll
    for (int j = 0; j < V; ++j) {
      hsu2[j] = hsu1[j]; hsu1[j] = hsu0[j]; hsd4[j] = hsd3[j]; hsd3[j] = hsd2[j];
      mx4[j] = mx3[j]; mx3[j] = mx2[j]; mn4[j] = mn3[j]; mn3[j] = mn2[j]; d2p[j] = d2n[j];
    }
    ax4 = ax3; ax3 = ax2;
    { const int t = s4; s4 = s3; s3 = s2; s2 = s1; s1 = s0; s0 = t; }
    const int on = o + 1;
    if (on >= R0 && on < R1 && outcol) {
      const T* xs = x + (img + (long long)on * g.W + wx) * g.C + (c0 >> 2);
#pragma unroll
      for (int j = 0; j < V; ++j) sq[j] = xs[j * q4];
    }
  }
}

// pick the row-band height: fewest wasted SM-waves, then fewest halo rows
static int mf_pick_th(int N, int H, int W, int C) {
  const long long per = (long long)cdiv(W, mf::TW) * N * (C / mf::CC);
  const int sms = egm_num_sms();
  double best = 1e30; int bestTH = H;
  for (int bands = 1; bands <= H; ++bands) {
    const int th = cdiv(H, bands);
    if (th < 8 && bands > 1) break;
    const long long ctas = per * cdiv(H, th);
    const double waves = (double)cdiv(ctas, sms);
    const double cost = waves * (th + 4);                // rows walked per SM
    if (cost < best - 1e-9) { best = cost; bestTH = th; }
  }
  return bestTH;
}

extern "C" int egm_mca_fused_supported(int C) { return (C % mf::CC) == 0 ? 1 : 0; }

// y = MCALayer blend of x with the gate vectors `gates` ([h: N*H | w: N*W | c: N*C], segments padded to 4 floats as egm_mca_vec_off_*);
// argidx (nullable): 1 byte per element, arg-max position | arg-min position << 4 inside the 3x3 window (row-major), for the backward.
extern "C" int egm_mca_fwd(const void* x, const float* gates, void* y, unsigned char* argidx, int dtype, int N, int H, int W, int C, void* stream) {
  EGM_REQUIRE(egm_mca_fused_supported(C), EGM_E_SHAPE, "mca_fwd: C %% 64 != 0 (use egm_mca_apply)");
  if ((long long)N * H * W == 0) return EGM_OK;
  McaFusedParams g{N, H, W, C, mf_pick_th(N, H, W, C), gates, gates + mf_al4((long long)N * H), gates + mf_al4((long long)N * H) + mf_al4((long long)N * W)};
  dim3 grid(cdiv(W, mf::TW), cdiv(H, g.TH), N * (C / mf::CC));
  EGM_REQUIRE(grid.z <= 65535 && grid.y <= 65535, EGM_E_SHAPE, "mca_fwd: grid too large");
  const size_t smb = (size_t)7 * mf::ROW * sizeof(float);
  EGM_DISPATCH_DTYPE(dtype, {
    static bool attr[64] = {};
    egm_ensure_smem(k_mca_fwd<T>, (int)smb, attr);
    egm_launch(k_mca_fwd<T>, grid, mf::THREADS, smb, (cudaStream_t)stream, (const T*)x, (T*)y, argidx, g);
  });
  EGM_LAUNCH_CHECK("mca_fwd"); return EGM_OK;
}

// ===================================================================================== backward
// smem rings over rows (slot = row index mod depth):
//   U  [3][TX][CC] fp32   u rows r, r-1                       (row r+1 is written while r, r-1 are read)
//   E  [2][TX][CC] fp32   E rows r-2 (written this iteration), r-3
//   DY [6][TX][CC] fp32   dy rows r .. r-4                    (the routed terms of output row r-3 read rows r-4 .. r-2)
//   IX [6][TX][CV] u32    arg codes of the same rows
template <typename T>
__global__ void __launch_bounds__(mf::THREADS, MF_MINB_BWD) k_mca_bwd(const T* __restrict__ x, const T* __restrict__ dy, const unsigned char* __restrict__ idx,
                                                           T* __restrict__ du, McaFusedParams g) { egm_pdl_enter();
  using namespace mf;
  extern __shared__ float sm[];
  float* U = sm;
  float* E = U + 3 * ROW;
  float* DY = E + 2 * ROW;
  unsigned* IX = reinterpret_cast<unsigned*>(DY + 6 * ROW);
  const int tid = threadIdx.x, cv = tid & (CV - 1), tx = tid / CV;
  const int chunks = g.C / CC;
  const int n = blockIdx.z / chunks, c0 = (blockIdx.z - n * chunks) * CC + cv * V;
  const int wx = blockIdx.x * TW - HALO + tx;
  const int R0 = blockIdx.y * g.TH, R1 = min(R0 + g.TH, g.H);
  const bool colv = wx >= 0 && wx < g.W;
  const bool outcol = tx >= HALO && tx < TX - HALO && wx < g.W;
  const int own = tx * CC + cv * V, ownx = tx * CV + cv;
  const int left = (tx > 0 ? own - CC : own), right = (tx < TX - 1 ? own + CC : own);
  const int leftx = (tx > 0 ? ownx - CV : ownx), rightx = (tx < TX - 1 ? ownx + CV : ownx);
  const bool hasL = tx > 0 && wx >= 1 && wx - 1 < g.W, hasR = tx < TX - 1 && wx + 1 >= 0 && wx + 1 < g.W;
  float gcv[V];
#pragma unroll
  for (int j = 0; j < V; ++j) gcv[j] = g.gc[n * g.C + c0 + j];
  const float gwv = colv ? g.gw[n * g.W + wx] : 0.f;
  const long long img = (long long)n * g.H * g.W;
  const int wxc = colv ? wx : 0;
  const int q4 = g.C >> 2;
  const int a4 = c0 / q4, b4 = c0 - a4 * q4;            // unshuffle: du ch c = a4*(C/4)+b4 receives 0.1 * dy ch b4*4 + a4

  RawV4<T> xq[PF], dq[PF]; unsigned iq[PF]; float ghq[PF];
  const int rstart = R0 - 2, rend = R1 + 2;
  auto fetch = [&](int rr, RawV4<T>& xr, RawV4<T>& dr, unsigned& ir, float& gr) {
    raw_zero(xr); raw_zero(dr); ir = 0x44444444u; gr = 0.f;
    if (rr >= 0 && rr < g.H && rr <= R1 + 1) {
      gr = g.gh[n * g.H + rr];
      if (colv) {
        const long long e = (img + (long long)rr * g.W + wxc) * g.C + c0;
        raw_load(xr, x + e); raw_load(dr, dy + e); ir = *reinterpret_cast<const unsigned*>(idx + e);
      }
    }
  };
#pragma unroll
  for (int i = 0; i < PF; ++i) fetch(rstart + i, xq[i], dq[i], iq[i], ghq[i]);
  float hsu1[V], hsu2[V], hsy1[V], hsy2[V], hse3[V], hse4[V], ep[V];
  T uq[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { hsu1[j] = hsu2[j] = hsy1[j] = hsy2[j] = hse3[j] = hse4[j] = ep[j] = 0.f; uq[j] = T(0.f); }
  int k = 0;                                            // r - rstart

  for (int r = rstart; r <= rend; ++r, ++k) {
    const int su0 = k % 3, su1 = (k + 2) % 3;           // U slots of rows r, r-1
    const int sy0 = k % 6;                              // DY / IX slot of row r
    float u[V], dyr[V];
    {
      float xv[V]; raw_get(xq[0], xv); raw_get(dq[0], dyr);
      const bool rowv = r >= 0 && r < g.H;
      const float s = (ghq[0] + gwv) * (1.f / 3.f);
#pragma unroll
      for (int j = 0; j < V; ++j) { u[j] = (rowv && colv) ? xv[j] * fmaf(gcv[j], 1.f / 3.f, s) : 0.f; if (!(rowv && colv)) dyr[j] = 0.f; }
      const unsigned ir = iq[0];
#pragma unroll
      for (int i = 0; i < PF - 1; ++i) { xq[i] = xq[i + 1]; dq[i] = dq[i + 1]; iq[i] = iq[i + 1]; ghq[i] = ghq[i + 1]; }
      fetch(r + PF, xq[PF - 1], dq[PF - 1], iq[PF - 1], ghq[PF - 1]);
      st4(U + su0 * ROW + own, u);
      st4(DY + sy0 * ROW + own, dyr);
      IX[sy0 * (TX * CV) + ownx] = ir;
      st4(E + (k & 1) * ROW + own, ep);               // E[r-2], computed in the previous iteration
    }
    __syncthreads();
    // ---- hs_u[r], hs_dy[r]; E[r-1] = 2 (u - avg3 u)(0.2/9) box3(dy) at row r-1
    float hsu0[V], hsy0[V], en[V];
    {
      float a[V], b[V], c1[V], ya[V], yb[V];
      ld4(U + su0 * ROW + left, a); ld4(U + su0 * ROW + right, b); ld4(U + su1 * ROW + own, c1);
      ld4(DY + sy0 * ROW + left, ya); ld4(DY + sy0 * ROW + right, yb);
      const bool v1 = r - 1 >= 0 && r - 1 < g.H && colv;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        hsu0[j] = (tx > 0 ? a[j] : 0.f) + u[j] + (tx < TX - 1 ? b[j] : 0.f);
        hsy0[j] = (tx > 0 ? ya[j] : 0.f) + dyr[j] + (tx < TX - 1 ? yb[j] : 0.f);
        const float d = c1[j] - (hsu2[j] + hsu1[j] + hsu0[j]) * (1.f / 9.f);
        en[j] = v1 ? 2.f * d * (0.2f / 9.f) * (hsy2[j] + hsy1[j] + hsy0[j]) : 0.f;
      }
    }
    // ---- hs_E[r-2]
    float hse2[V];
    {
      float a[V], b[V];
      ld4(E + (k & 1) * ROW + left, a); ld4(E + (k & 1) * ROW + right, b);
#pragma unroll
      for (int j = 0; j < V; ++j) hse2[j] = (tx > 0 ? a[j] : 0.f) + ep[j] + (tx < TX - 1 ? b[j] : 0.f);
    }
    // ---- output row q = r - 3
    const int q = r - 3;
    if (q >= R0 && q < R1 && outcol) {
      float acc[V], eq[V], dq0[V];
      ld4(E + ((k + 1) & 1) * ROW + own, eq);           // E[r-3] (written in the previous iteration)
      ld4(DY + ((k + 3) % 6) * ROW + own, dq0);         // dy[r-3]
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = 0.51f * dq0[j] + 0.1f * tof(uq[j]) + eq[j] - (hse4[j] + hse3[j] + hse2[j]) * (1.f / 9.f);
      // routed max / min gradients: windows centred at p = (q-1+a, wx-1+b); q sits at position (2-a)*3 + (2-b) of window p
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int pr = q - 1 + a;
        if (pr < 0 || pr >= g.H) continue;
        const int sl = (k + 2 + a) % 6;                 // slot of row r-4+a
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          if ((b == 0 && !hasL) || (b == 2 && !hasR)) continue;
          const int off = b == 0 ? left : (b == 1 ? own : right), offx = b == 0 ? leftx : (b == 1 ? ownx : rightx);
          float dv[V]; ld4(DY + sl * ROW + off, dv);
          const unsigned code = IX[sl * (TX * CV) + offx];
          const unsigned want = (unsigned)((2 - a) * 3 + (2 - b));
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const unsigned cj = (code >> (8 * j)) & 0xffu;
            acc[j] += 0.2f * (((cj & 15u) == want ? dv[j] : 0.f) - ((cj >> 4) == want ? dv[j] : 0.f));
          }
        }
      }
      FVec<V> o;
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] = acc[j];
      stv<V>(du + (img + (long long)q * g.W + wx) * g.C + c0, o);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      hsu2[j] = hsu1[j]; hsu1[j] = hsu0[j]; hsy2[j] = hsy1[j]; hsy1[j] = hsy0[j]; hse4[j] = hse3[j]; hse3[j] = hse2[j]; ep[j] = en[j];
    }
    const int qn = q + 1;
    if (qn >= R0 && qn < R1 && outcol) {
      const T* ys = dy + (img + (long long)qn * g.W + wx) * g.C + a4;
#pragma unroll
      for (int j = 0; j < V; ++j) uq[j] = ys[(b4 + j) * 4];
    }
  }
}

// du = dL/du of the MCALayer blend (see the header comment); x, dy, du: [N,H,W,C]; argidx from egm_mca_fwd.
extern "C" int egm_mca_bwd(const void* x, const float* gates, const void* dy, const unsigned char* argidx, void* du, int dtype, int N, int H, int W, int C,
                           void* stream) {
  EGM_REQUIRE(egm_mca_fused_supported(C), EGM_E_SHAPE, "mca_bwd: C %% 64 != 0 (use egm_mca_bwd_du)");
  EGM_REQUIRE(argidx != nullptr, EGM_E_BADARG, "mca_bwd: needs the arg map written by egm_mca_fwd");
  if ((long long)N * H * W == 0) return EGM_OK;
  McaFusedParams g{N, H, W, C, mf_pick_th(N, H, W, C), gates, gates + mf_al4((long long)N * H), gates + mf_al4((long long)N * H) + mf_al4((long long)N * W)};
  dim3 grid(cdiv(W, mf::TW), cdiv(H, g.TH), N * (C / mf::CC));
  EGM_REQUIRE(grid.z <= 65535 && grid.y <= 65535, EGM_E_SHAPE, "mca_bwd: grid too large");
  const size_t smb = (size_t)11 * mf::ROW * sizeof(float) + (size_t)6 * mf::TX * mf::CV * sizeof(unsigned);
  EGM_DISPATCH_DTYPE(dtype, {
    static bool attr[64] = {};
    egm_ensure_smem(k_mca_bwd<T>, (int)smb, attr);
    egm_launch(k_mca_bwd<T>, grid, mf::THREADS, smb, (cudaStream_t)stream, (const T*)x, (const T*)dy, argidx, (T*)du, g);
  });
  EGM_LAUNCH_CHECK("mca_bwd"); return EGM_OK;
}

// ===================================================================================== 3x3 high-pass (EdgeAwareFeatureEnhancer)
// out (+)= in - avgpool3x3(in)  (zero pad, /9; self-adjoint, so the backward is the same kernel with accumulate): src/EGM-UNet.py:875,883.
// Same row walk as above with a 1-pixel halo: one 8-byte global load per thread and row, the horizontal 3-sums through a double-buffered
// shared-memory row, the vertical window in registers.  Round 1's gather kernel pulled 9 neighbours per output through L1 (26 % of HBM).
namespace hp {
constexpr int TX = 32, HALO = 1, TW = TX - 2 * HALO, PF = 4;
}
template <typename T>
__global__ void __launch_bounds__(mf::THREADS, 2) k_highpass3_walk(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int C, int TH,
                                                                  int accumulate) { egm_pdl_enter();
  using mf::CC; using mf::V; using mf::CV; using mf::ROW;
  using namespace hp;
  extern __shared__ float sm[];
  const int tid = threadIdx.x, cv = tid & (CV - 1), tx = tid / CV;
  const int chunks = C / CC;
  const int n = blockIdx.z / chunks, c0 = (blockIdx.z - n * chunks) * CC + cv * V;
  const int wx = blockIdx.x * TW - HALO + tx;
  const int R0 = blockIdx.y * TH, R1 = min(R0 + TH, H);
  const bool colv = wx >= 0 && wx < W;
  const bool outcol = tx >= HALO && tx < TX - HALO && wx < W;
  const int own = tx * CC + cv * V;
  const int left = (tx > 0 ? own - CC : own), right = (tx < TX - 1 ? own + CC : own);
  const long long img = (long long)n * H * W;
  const int wxc = colv ? wx : 0;
  RawV4<T> xq[PF];
  const int rstart = R0 - 1, rend = R1;
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    const int rr = rstart + i;
    raw_zero(xq[i]);
    if (rr >= 0 && rr < H && rr <= R1 && colv) raw_load(xq[i], in + (img + (long long)rr * W + wxc) * C + c0);
  }
  float hs1[V], hs2[V], xp[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { hs1[j] = hs2[j] = xp[j] = 0.f; }
  for (int r = rstart; r <= rend; ++r) {
    float xv[V]; raw_get(xq[0], xv);                  // zero outside the image / the band's needs
#pragma unroll
    for (int i = 0; i < PF - 1; ++i) xq[i] = xq[i + 1];
    {
      const int rr = r + PF;
      raw_zero(xq[PF - 1]);
      if (rr >= 0 && rr < H && rr <= R1 && colv) raw_load(xq[PF - 1], in + (img + (long long)rr * W + wxc) * C + c0);
    }
    const int o = r - 1;
    const bool doout = o >= R0 && o < R1 && outcol;
    FVec<V> acc;
#pragma unroll
    for (int j = 0; j < V; ++j) acc.v[j] = 0.f;
    const long long e = (img + (long long)(doout ? o : 0) * W + wxc) * C + c0;
    if (doout && accumulate) acc = ldv<V>(out + e);
    float* S = sm + (r & 1) * ROW;
    st4(S + own, xv);
    __syncthreads();
    float a[V], b[V], hs0[V];
    ld4(S + left, a); ld4(S + right, b);
#pragma unroll
    for (int j = 0; j < V; ++j) hs0[j] = (tx > 0 ? a[j] : 0.f) + xv[j] + (tx < TX - 1 ? b[j] : 0.f);
    if (doout) {
#pragma unroll
      for (int j = 0; j < V; ++j) acc.v[j] += xp[j] - (hs2[j] + hs1[j] + hs0[j]) * (1.f / 9.f);
      stv<V>(out + e, acc);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) { hs2[j] = hs1[j]; hs1[j] = hs0[j]; xp[j] = xv[j]; }
  }
}
static int hp_pick_th(int N, int H, int W, int C) {
  const long long per = (long long)cdiv(W, hp::TW) * N * (C / mf::CC);
  const int slots = egm_num_sms() * 2;
  double best = 1e30; int bestTH = H;
  for (int bands = 1; bands <= H; ++bands) {
    const int th = cdiv(H, bands);
    if (th < 8 && bands > 1) break;
    const double cost = (double)cdiv(per * cdiv(H, th), slots) * (th + 2);
    if (cost < best - 1e-9) { best = cost; bestTH = th; }
  }
  return bestTH;
}
// returns 1 if the walking kernel took the launch, 0 if the shape is not covered (caller falls back to the gather kernel)
int egm_highpass3_walk_launch(const void* in, void* out, int accumulate, int dtype, int N, int H, int W, int C, cudaStream_t st) {
  if (C % mf::CC != 0 || (long long)N * (C / mf::CC) > 65535) return 0;
  const int TH = hp_pick_th(N, H, W, C);
  dim3 grid(cdiv(W, hp::TW), cdiv(H, TH), N * (C / mf::CC));
  if (grid.y > 65535) return 0;
  const size_t smb = (size_t)2 * mf::ROW * sizeof(float);
  if (dtype == EGM_F32) egm_launch(k_highpass3_walk<float>, grid, mf::THREADS, smb, st, (const float*)in, (float*)out, N, H, W, C, TH, accumulate);
  else if (dtype == EGM_BF16) egm_launch(k_highpass3_walk<__nv_bfloat16>, grid, mf::THREADS, smb, st, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, N, H, W, C, TH, accumulate);
  else return 0;
  return 1;
}
