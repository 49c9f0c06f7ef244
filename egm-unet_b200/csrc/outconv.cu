// OutConv (src/EGM-UNet.py:952-956 == src/unet.py:54-58): Conv2d(C -> num_classes, kernel 1) + bias on the last decoder map, fused with
// the layout change at the model boundary: reads the NHWC activation once and writes fp32 NCHW logits (what the criterion and
// `output['out']` consumers see); backward reads fp32 NCHW dlogits and produces dL/dy (NHWC), dL/dW and dL/db in one pass.
//
// Round 1 ran this as a zero-padded 32 -> 16 tcgen05 conv + NHWC->NCHW conversion (and NCHW->NHWC + wgrad + dgrad + channel_sum on
// the way back); that path rounded the logits AND the incoming dlogits to bf16 -- the first and the last rounding of the whole step,
// and the one that decides near-tie argmax pixels.  Here both stay fp32.  HBM-bound: C*2 + K*4 bytes per pixel.
#include "common.cuh"

constexpr int OC_MAXK = 4;

template <typename T, int C, int K>
__global__ void __launch_bounds__(256) k_outconv_fwd(const T* __restrict__ y, const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ logits,
                                                     long long HW, long long total) { egm_pdl_enter();
  __shared__ float sw[K * C + K];
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < K) sw[K * C + threadIdx.x] = b ? b[threadIdx.x] : 0.f;
  __syncthreads();
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = sw[K * C + k];
    const T* yp = y + p * C;
#pragma unroll
    for (int c = 0; c < C; c += 8) {
      FVec<8> v = ldv<8>(yp + c);
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = fmaf(v.v[j], sw[k * C + c + j], acc[k]);
    }
    const long long n = p / HW, q = p - n * HW;
#pragma unroll
    for (int k = 0; k < K; ++k) logits[(n * K + k) * HW + q] = acc[k];
  }
}

// dy[p][c] = sum_k dl[k][p] w[k][c];  dw[k][c] += sum_p dl[k][p] y[p][c];  db[k] += sum_p dl[k][p]     (dw / db zeroed by the launcher)
template <typename T, int C, int K>
__global__ void __launch_bounds__(256) k_outconv_bwd(const T* __restrict__ y, const float* __restrict__ w, const float* __restrict__ dl, T* __restrict__ dy,
                                                     float* __restrict__ dw, float* __restrict__ db, long long HW, long long total) { egm_pdl_enter();
  __shared__ float sw[K * C];
  __shared__ float red[8][K * C + K];
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  float aw[K][C], ab[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { ab[k] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) aw[k][c] = 0.f; }
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const long long n = p / HW, q = p - n * HW;
    float d[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { d[k] = dl[(n * K + k) * HW + q]; ab[k] += d[k]; }
    const T* yp = y + p * C;
    T* gp = dy ? dy + p * C : nullptr;
#pragma unroll
    for (int c = 0; c < C; c += 8) {
      FVec<8> v = ldv<8>(yp + c), g;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) { s = fmaf(d[k], sw[k * C + c + j], s); aw[k][c + j] = fmaf(d[k], v.v[j], aw[k][c + j]); }
        g.v[j] = s;
      }
      if (gp) stv<8>(gp + c, g);
    }
  }
  // block reduction: warp shuffles, then one row per warp in shared memory, then one atomic per value and block
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int c = 0; c < C; ++c) { const float s = warp_sum(aw[k][c]); if (lane == 0) red[wid][k * C + c] = s; }
    const float s = warp_sum(ab[k]); if (lane == 0) red[wid][K * C + k] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * C + K; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += red[r][i];
    if (i < K * C) atomicAdd(dw + i, s); else if (db) atomicAdd(db + (i - K * C), s);
  }
}

// the backward keeps K*C weight-gradient accumulators in registers: K*C <= 128
extern "C" int egm_outconv_supported(int C, int K) { return ((C == 32 && K >= 1 && K <= OC_MAXK) || (C == 64 && K >= 1 && K <= 2)) ? 1 : 0; }

#define EGM_OC_DISPATCH(C, K, ...)                                                              \
  do {                                                                                          \
    if (C == 32 && K == 1) { constexpr int CC = 32, KK = 1; __VA_ARGS__; }                      \
    else if (C == 32 && K == 2) { constexpr int CC = 32, KK = 2; __VA_ARGS__; }                 \
    else if (C == 32 && K == 3) { constexpr int CC = 32, KK = 3; __VA_ARGS__; }                 \
    else if (C == 32 && K == 4) { constexpr int CC = 32, KK = 4; __VA_ARGS__; }                 \
    else if (C == 64 && K == 1) { constexpr int CC = 64, KK = 1; __VA_ARGS__; }                 \
    else { constexpr int CC = 64, KK = 2; __VA_ARGS__; }                                        \
  } while (0)

// y [N,H,W,C] (dtype) -> logits [N,K,H,W] fp32;  w [K][C] fp32 (the nn.Conv2d weight [K,C,1,1]), bias [K] or NULL
extern "C" int egm_outconv_fwd(const void* y, const float* w, const float* bias, float* logits, int dtype, int N, long long HW, int C, int K, void* stream) {
  EGM_REQUIRE(egm_outconv_supported(C, K), EGM_E_SHAPE, "outconv: C=%d K=%d unsupported", C, K);
  const long long total = (long long)N * HW;
  if (total == 0) return EGM_OK;
  const int grid = egm_grid_for(total, 256);
  EGM_DISPATCH_DTYPE(dtype, EGM_OC_DISPATCH(C, K, (egm_launch(k_outconv_fwd<T, CC, KK>, grid, 256, 0, (cudaStream_t)stream, (const T*)y, w, bias, logits, HW, total))));
  EGM_LAUNCH_CHECK("outconv_fwd"); return EGM_OK;
}
// dlogits [N,K,H,W] fp32 -> dy [N,H,W,C] (dtype; NULL to skip), dw [K][C] fp32, dbias [K] fp32 (NULL to skip); dw / dbias are overwritten
extern "C" int egm_outconv_bwd(const void* y, const float* w, const float* dlogits, void* dy, float* dw, float* dbias, int dtype, int N, long long HW, int C,
                               int K, void* stream) {
  EGM_REQUIRE(egm_outconv_supported(C, K), EGM_E_SHAPE, "outconv: C=%d K=%d unsupported", C, K);
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(dw, 0, sizeof(float) * K * C, st);
  if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * K, st);
  const long long total = (long long)N * HW;
  if (total == 0) return EGM_OK;
  const int grid = egm_grid_for(total, 256, 4);
  EGM_DISPATCH_DTYPE(dtype, EGM_OC_DISPATCH(C, K, (egm_launch(k_outconv_bwd<T, CC, KK>, grid, 256, 0, st, (const T*)y, w, dlogits, (T*)dy, dw, dbias, HW, total))));
  EGM_LAUNCH_CHECK("outconv_bwd"); return EGM_OK;
}
