// Layout conversion, slice copies, small elementwise helpers (all HBM-bound, vectorised NHWC).
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

// ------------------------------------------------------------------ error state
static thread_local char g_err[512] = "";
void egm_set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
int egm_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { egm_set_error("%s: %s", what, cudaGetErrorString(e)); return EGM_E_LAUNCH; }
  return EGM_OK;
}
extern "C" const char* egm_last_error(void) { return g_err; }
// ------------------------------------------------------------------ programmatic dependent launch switch (see common.cuh)
static int g_pdl = -1;
int egm_launch_overlap_enabled() {
  if (g_pdl < 0) { const char* e = getenv("EGM_PDL"); g_pdl = (e && e[0] == '1') ? 1 : 0; }    // measured slower on cfg2 (see common.cuh): off unless asked for
  return g_pdl;
}
extern "C" int egm_set_launch_overlap(int enabled) { int prev = egm_launch_overlap_enabled(); g_pdl = enabled ? 1 : 0; return prev; }
extern "C" int egm_abi_version(void) { return EGM_ABI_VERSION; }
extern "C" int egm_device_check(void) {
  int dev = 0; cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    egm_set_error("no CUDA device"); return EGM_E_ARCH;
  }
  if (p.major != 10) { egm_set_error("libegm_b200 needs compute capability 10.x (sm_100a), found %d.%d", p.major, p.minor); return EGM_E_ARCH; }
  return EGM_OK;
}

// ------------------------------------------------------------------ NCHW fp32 <-> NHWC T
template <typename T>
__global__ void k_nchw_to_nhwc(const float* __restrict__ x, T* __restrict__ y, long long NHW, int C, long long HW) { egm_pdl_enter();
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < NHW; p += (long long)gridDim.x * blockDim.x) {
    long long n = p / HW, r = p - n * HW;
    const float* xp = x + n * C * HW + r;
    T* yp = y + p * C;
    for (int c = 0; c < C; ++c) stf(yp + c, xp[(long long)c * HW]);
  }
}
template <typename T>
__global__ void k_nhwc_to_nchw(const T* __restrict__ x, float* __restrict__ y, long long NHW, int C, long long HW) { egm_pdl_enter();
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < NHW; p += (long long)gridDim.x * blockDim.x) {
    long long n = p / HW, r = p - n * HW;
    const T* xp = x + p * C;
    float* yp = y + n * C * HW + r;
    for (int c = 0; c < C; ++c) yp[(long long)c * HW] = ldf(xp + c);
  }
}
extern "C" int egm_nchw_to_nhwc(const float* x, void* y, int dtype, int N, int C, int H, int W, void* stream) {
  long long HW = (long long)H * W, NHW = HW * N;
  if (NHW == 0) return EGM_OK;
  EGM_DISPATCH_DTYPE(dtype, (egm_launch(k_nchw_to_nhwc<T>, egm_grid_for(NHW, 256), 256, 0, (cudaStream_t)stream, x, (T*)y, NHW, C, HW)));
  EGM_LAUNCH_CHECK("nchw_to_nhwc"); return EGM_OK;
}
extern "C" int egm_nhwc_to_nchw(const void* x, float* y, int dtype, int N, int C, int H, int W, void* stream) {
  long long HW = (long long)H * W, NHW = HW * N;
  if (NHW == 0) return EGM_OK;
  EGM_DISPATCH_DTYPE(dtype, (egm_launch(k_nhwc_to_nchw<T>, egm_grid_for(NHW, 256), 256, 0, (cudaStream_t)stream, (const T*)x, y, NHW, C, HW)));
  EGM_LAUNCH_CHECK("nhwc_to_nchw"); return EGM_OK;
}

// ------------------------------------------------------------------ channel-slice copy (concat / split)
template <typename T, int V>
__global__ void k_copy_slice(const T* __restrict__ src, T* __restrict__ dst, long long M, int CV, long long scs, long long sco,
                             long long dcs, long long dco, int accumulate) { egm_pdl_enter();
  long long total = M * CV;
  const RowIndexer rix(CV, total);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long m; int c; rix(i, m, c); c *= V;
    FVec<V> a = ldv<V>(src + m * scs + sco + c);
    T* dp = dst + m * dcs + dco + c;
    if (accumulate) { FVec<V> b = ldv<V>(dp);
#pragma unroll
      for (int j = 0; j < V; ++j) a.v[j] += b.v[j]; }
    stv<V>(dp, a);
  }
}
extern "C" int egm_copy_slice(const void* src, void* dst, int dtype, long long M, int C, long long s_cstride, long long s_coff,
                              long long d_cstride, long long d_coff, int accumulate, void* stream) {
  if (M * C == 0) return EGM_OK;
  int v = egm_pick_vec(C, s_cstride, s_coff); int v2 = egm_pick_vec(C, d_cstride, d_coff); if (v2 < v) v = v2;
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_copy_slice<T, V>, egm_grid_for(M * (C / V), 256), 256, 0, (cudaStream_t)stream, 
      (const T*)src, (T*)dst, M, C / V, s_cstride, s_coff, d_cstride, d_coff, accumulate))));
  EGM_LAUNCH_CHECK("copy_slice"); return EGM_OK;
}

// ------------------------------------------------------------------ dst = alpha*dst + beta*src  (flat)
template <typename T, int V>
__global__ void k_axpby(T* __restrict__ dst, const T* __restrict__ src, long long nv, float alpha, float beta) { egm_pdl_enter();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    FVec<V> a = ldv<V>(dst + i * V), b = ldv<V>(src + i * V);
#pragma unroll
    for (int j = 0; j < V; ++j) a.v[j] = alpha * a.v[j] + beta * b.v[j];
    stv<V>(dst + i * V, a);
  }
}
extern "C" int egm_axpby(void* dst, const void* src, int dtype, long long n, float alpha, float beta, void* stream) {
  if (n == 0) return EGM_OK;
  int v = egm_pick_vec(n);
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_axpby<T, V>, egm_grid_for(n / V, 256), 256, 0, (cudaStream_t)stream, (T*)dst, (const T*)src, n / V, alpha, beta))));
  EGM_LAUNCH_CHECK("axpby"); return EGM_OK;
}

extern "C" int egm_memset_zero(void* p, long long bytes, void* stream) {
  if (bytes == 0) return EGM_OK;
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) { egm_set_error("memset: %s", cudaGetErrorString(e)); return EGM_E_LAUNCH; }
  return EGM_OK;
}

// fp32 -> T cast (flat), used for staging parameters
template <typename T>
__global__ void k_cast_from_f32(const float* __restrict__ s, T* __restrict__ d, long long n) { egm_pdl_enter();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) stf(d + i, s[i]);
}
extern "C" int egm_cast_from_f32(const float* src, void* dst, int dtype, long long n, void* stream) {
  if (n == 0) return EGM_OK;
  EGM_DISPATCH_DTYPE(dtype, (egm_launch(k_cast_from_f32<T>, egm_grid_for(n, 256), 256, 0, (cudaStream_t)stream, src, (T*)dst, n)));
  EGM_LAUNCH_CHECK("cast_from_f32"); return EGM_OK;
}
template <typename T>
__global__ void k_cast_to_f32(const T* __restrict__ s, float* __restrict__ d, long long n, int accumulate) { egm_pdl_enter();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = (accumulate ? d[i] : 0.f) + ldf(s + i);
}
extern "C" int egm_cast_to_f32(const void* src, float* dst, int dtype, long long n, int accumulate, void* stream) {
  if (n == 0) return EGM_OK;
  EGM_DISPATCH_DTYPE(dtype, (egm_launch(k_cast_to_f32<T>, egm_grid_for(n, 256), 256, 0, (cudaStream_t)stream, (const T*)src, dst, n, accumulate)));
  EGM_LAUNCH_CHECK("cast_to_f32"); return EGM_OK;
}

// out[r][c] = w[r][c] * scale[r]: per-output-channel scale of an inference BatchNorm folded into the conv weight [Cout][Cin*kh*kw]
__global__ void k_scale_rows(const float* __restrict__ w, const float* __restrict__ scale, float* __restrict__ out, long long total, long long cols) { egm_pdl_enter();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) out[i] = w[i] * scale[i / cols];
}
extern "C" int egm_scale_rows(const float* w, const float* scale, float* out, int rows, long long cols, void* stream) {
  const long long total = (long long)rows * cols;
  if (total == 0) return EGM_OK;
  egm_launch(k_scale_rows, egm_grid_for(total, 256), 256, 0, (cudaStream_t)stream, w, scale, out, total, cols);
  EGM_LAUNCH_CHECK("scale_rows"); return EGM_OK;
}
