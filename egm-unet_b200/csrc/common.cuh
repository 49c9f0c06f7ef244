// Common device/host helpers for libegm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/egm_b200.h"

// ---------------------------------------------------------------- error plumbing
void egm_set_error(const char* fmt, ...);
int egm_check_launch(const char* what);

#define EGM_REQUIRE(cond, code, ...)                 \
  do {                                               \
    if (!(cond)) {                                   \
      egm_set_error(__VA_ARGS__);                    \
      return (code);                                 \
    }                                                \
  } while (0)

#define EGM_LAUNCH_CHECK(what)                       \
  do {                                               \
    int _e = egm_check_launch(what);                 \
    if (_e) return _e;                               \
  } while (0)

static inline int egm_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: set it once per (kernel, device), not once per process
template <typename K>
static inline void egm_ensure_smem(K kernel, int bytes, bool (&done)[64]) {
  int dev = 0; cudaGetDevice(&dev); dev &= 63;
  if (!done[dev]) { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); done[dev] = true; }
}
// ---------------------------------------------------------------- launches (optionally with programmatic dependent launch, PDL)
// Every kernel of the library is launched through egm_launch() and starts with egm_pdl_enter().  With the launch attribute
// cudaLaunchAttributeProgrammaticStreamSerialization a kernel lets the NEXT grid of the stream be scheduled once all of its own CTAs
// have started (griddepcontrol.launch_dependents) and waits for the PREVIOUS grid's completion and memory visibility before its
// first global access (griddepcontrol.wait); because every kernel executes the wait, ordering stays transitive and results are
// identical.  MEASURED on cfg2 inside the captured step graph (profiles/step_variants_r2.txt): 24.56 ms with the early trigger,
// 24.08 ms with the wait only, 24.08 ms with plain stream order -- graph replay already leaves no gap between dependent kernel
// nodes, and early-resident successor CTAs only take resources from the running grid's tail.  So the attribute is OFF by default
// (both instructions are no-ops then); EGM_PDL=1 / egm_set_launch_overlap(1) / bench.py --pdl switch it on for A/B runs.
#ifndef EGM_PDL_MODE
#define EGM_PDL_MODE 2      // 2: early trigger + wait; 1: wait only (successor scheduled when this grid's CTAs exit); 0: no PDL instructions
#endif
__device__ __forceinline__ void egm_pdl_enter() {
#if EGM_PDL_MODE >= 2
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
#if EGM_PDL_MODE >= 1
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
int egm_launch_overlap_enabled();
template <typename... P, typename... A>
static inline void egm_launch(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = egm_launch_overlap_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- dtype dispatch
// Storage type T in {float, __nv_bfloat16}; arithmetic is always fp32.
#define EGM_DISPATCH_DTYPE(dtype, ...)                                        \
  do {                                                                        \
    if ((dtype) == EGM_F32) { using T = float; __VA_ARGS__; }                 \
    else if ((dtype) == EGM_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }   \
    else { egm_set_error("bad dtype %d", (int)(dtype)); return EGM_E_BADARG; }\
  } while (0)

__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// Vector of V channels (V in {1,2,4,8}) loaded with the widest aligned access.
template <int V> struct FVec { float v[V]; };

template <int V>
__device__ __forceinline__ FVec<V> ldv(const float* p) {
  FVec<V> r;
  if constexpr (V == 8) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  } else if constexpr (V == 4) {
    float4 a = *reinterpret_cast<const float4*>(p);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  } else if constexpr (V == 2) {
    float2 a = *reinterpret_cast<const float2*>(p);
    r.v[0] = a.x; r.v[1] = a.y;
  } else {
    r.v[0] = *p;
  }
  return r;
}
template <int V>
__device__ __forceinline__ FVec<V> ldv(const __nv_bfloat16* p) {
  FVec<V> r;
  if constexpr (V == 8) {
    uint4 a = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y; }
  } else if constexpr (V == 4) {
    uint2 a = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
    for (int i = 0; i < 2; ++i) { float2 f = __bfloat1622float2(h[i]); r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y; }
  } else if constexpr (V == 2) {
    __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(p);
    float2 f = __bfloat1622float2(h); r.v[0] = f.x; r.v[1] = f.y;
  } else {
    r.v[0] = __bfloat162float(*p);
  }
  return r;
}
template <int V>
__device__ __forceinline__ void stv(float* p, const FVec<V>& r) {
  if constexpr (V == 8) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
  } else if constexpr (V == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else if constexpr (V == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(r.v[0], r.v[1]);
  } else {
    *p = r.v[0];
  }
}
template <int V>
__device__ __forceinline__ void stv(__nv_bfloat16* p, const FVec<V>& r) {
  if constexpr (V == 8) {
    uint4 a; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&a);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = a;
  } else if constexpr (V == 4) {
    uint2 a; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&a);
#pragma unroll
    for (int i = 0; i < 2; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
    *reinterpret_cast<uint2*>(p) = a;
  } else if constexpr (V == 2) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(r.v[0], r.v[1]);
  } else {
    *p = __float2bfloat16_rn(r.v[0]);
  }
}

// Widest vector width usable for a tensor with C channels whose rows start at
// multiples of `cstride` elements from an (at least 16-byte aligned) base.
static inline int egm_pick_vec(long long C, long long cstride = 0, long long coff = 0) {
  if (cstride == 0) cstride = C;
  if (C % 8 == 0 && cstride % 8 == 0 && coff % 8 == 0) return 8;
  if (C % 4 == 0 && cstride % 4 == 0 && coff % 4 == 0) return 4;
  if (C % 2 == 0 && cstride % 2 == 0 && coff % 2 == 0) return 2;
  return 1;
}

#define EGM_DISPATCH_VEC(vec, ...)                                  \
  do {                                                              \
    if ((vec) == 8) { constexpr int V = 8; __VA_ARGS__; }           \
    else if ((vec) == 4) { constexpr int V = 4; __VA_ARGS__; }      \
    else if ((vec) == 2) { constexpr int V = 2; __VA_ARGS__; }      \
    else { constexpr int V = 1; __VA_ARGS__; }                      \
  } while (0)

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum; `red` is >= 32 floats of shared memory; result valid in all threads
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) { r = warp_sum(r); if (lane == 0) red[0] = r; }
  __syncthreads();
  return red[0];
}

// ------------------------------------------------------------------ index decoding without 64-bit divisions
// The grid-stride elementwise / stencil kernels decode a flat vector index into (channel vector, w, h, n).  Three 64-bit
// divisions by runtime values cost ~300 instructions per element -- more than the kernels' own work (ncu: these kernels
// ran at 1.1 TB/s, issue-bound).  Division by a runtime constant through a 33-bit reciprocal (Granlund-Montgomery round-up
// form, exact for dividends < 2^32) is 3 instructions; index spaces >= 2^31 fall back to the 64-bit path.
struct FastDiv {
  unsigned d, m, s;
  __device__ __forceinline__ explicit FastDiv(int dd) {
    d = (unsigned)dd; s = 32 - __clz((int)(d - 1)); if (d <= 1) s = 0;
    m = (unsigned)((((unsigned long long)1 << 32) * (((unsigned long long)1 << s) - d)) / d + 1);
  }
  __device__ __forceinline__ unsigned div(unsigned x) const { return (unsigned)(((unsigned long long)__umulhi(x, m) + x) >> s); }
};
struct Nhwc4 { int cv, w, h, n; long long p; };
// i = ((n*H + h)*W + w)*CV + cv
struct NhwcIndexer {
  FastDiv dCV, dW, dH; int CV, W, H; bool big;
  __device__ __forceinline__ NhwcIndexer(int CV_, int W_, int H_, long long total)
      : dCV(CV_), dW(W_), dH(H_), CV(CV_), W(W_), H(H_), big(total > 0x7fffffffLL) {}
  __device__ __forceinline__ Nhwc4 operator()(long long i) const {
    Nhwc4 r;
    if (!big) {
      const unsigned x = (unsigned)i, p = dCV.div(x), q = dW.div(p), n = dH.div(q);
      r.cv = (int)(x - p * (unsigned)CV); r.w = (int)(p - q * (unsigned)W); r.h = (int)(q - n * (unsigned)H); r.n = (int)n; r.p = p;
    } else {
      r.cv = (int)(i % CV); r.p = i / CV; r.w = (int)(r.p % W); const long long q = r.p / W; r.h = (int)(q % H); r.n = (int)(q / H);
    }
    return r;
  }
};
// i = m*CV + cv
struct RowIndexer {
  FastDiv dCV; int CV; bool big;
  __device__ __forceinline__ RowIndexer(int CV_, long long total) : dCV(CV_), CV(CV_), big(total > 0x7fffffffLL) {}
  __device__ __forceinline__ void operator()(long long i, long long& m, int& cv) const {
    if (!big) { const unsigned x = (unsigned)i, q = dCV.div(x); m = q; cv = (int)(x - q * (unsigned)CV); }
    else { m = i / CV; cv = (int)(i - m * CV); }
  }
};

static inline int egm_grid_for(long long work_items, int threads, int per_sm = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)egm_num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
