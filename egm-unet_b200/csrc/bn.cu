// BatchNorm2d (training statistics / apply / backward) fused with the activation that follows it.
// Reference semantics: nn.BatchNorm2d as used by DoubleConv (src/EGM-UNet.py:44-55), BasicConv
// (:958-975, momentum 0.01) and EdgeAwareFeatureEnhancer (:872-886); see SURVEY.md App. A.
// All kernels are HBM-bound: one pass per tensor, 16-byte vector accesses, per-channel fp64 atomics.
#include "common.cuh"
#include <stdlib.h>

// act: 0 none, 1 relu, 2 sigmoid.   mode: 0 plain, 1 edge gate (y = s*aux + aux), 2 residual (y = relu(alpha*aux + bn))
struct BnArgs {
  const float* scale; const float* shift; const float* mean; const float* rstd; const float* coef;
  int act, mode; float alpha;
};

// Finalize fused into the APPLY kernels (statistics from a conv epilogue): sums != nullptr makes every thread derive the scale / shift of
// its own channel vector from the batch sums (the same fp64 formulas as k_bn_finalize; a few hundred cycles, hidden behind the first
// loads) instead of reading vectors a separate ~7 us single-warp launch would have produced; the first C/V threads of block 0 publish
// scale / shift / mean / rstd for the backward pass and update the running statistics.
struct BnFin {
  const double* sums; double M;
  const float* gamma; const float* beta; float* rmean; float* rvar; long long* nbt; float momentum, eps;
  float* scale; float* shift; float* mean; float* rstd;
};
template <int V>
__device__ __forceinline__ void bn_fin_vec(const BnFin& f, const BnArgs& a, int c, int C, bool publish, FVec<V>& sc, FVec<V>& sh) {
  if (!f.sums) { sc = ldv<V>(a.scale + c); sh = ldv<V>(a.shift + c); return; }
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const double mean = f.sums[c + j] / f.M; double var = f.sums[C + c + j] / f.M - mean * mean; if (var < 0) var = 0;
    const double rstd = 1.0 / sqrt(var + (double)f.eps);
    const float g = f.gamma[c + j];
    sc.v[j] = (float)(g * rstd); sh.v[j] = (float)(f.beta[c + j] - mean * g * rstd);
    if (publish) {
      if (f.rmean) {
        const double unb = f.M > 1 ? var * f.M / (f.M - 1) : var;
        f.rmean[c + j] = (float)((1.0 - f.momentum) * f.rmean[c + j] + f.momentum * mean);
        f.rvar[c + j] = (float)((1.0 - f.momentum) * f.rvar[c + j] + f.momentum * unb);
      }
      f.scale[c + j] = sc.v[j]; f.shift[c + j] = sh.v[j]; f.mean[c + j] = (float)mean; f.rstd[c + j] = (float)rstd;
    }
  }
  if (publish && c == 0 && f.nbt) *f.nbt += 1;
}

// Finalize fused into the reduction kernels: the LAST block to add its partial sums (ticket counter behind the sums) turns them
// into scale/shift/mean/rstd + running statistics (forward) or the backward coefficients + dgamma/dbeta (backward).  Saves one
// dependent ~3 us launch per BN layer and pass (148 per train step).
struct BnTail {
  unsigned int* counter;                 // nullptr: no fused finalize
  int backward;
  double M;
  const float* gamma; const float* beta; float* rmean; float* rvar; long long* nbt; float momentum, eps;
  float* scale; float* shift; float* mean; float* rstd;      // forward outputs (backward: rstd is an input)
  float* coef; float* dgamma; float* dbeta;                  // backward outputs
};
__device__ __forceinline__ void bn_tail_run(const BnTail& tl, const double* sums, int C, int tid, int nthr, bool named) {
  if (!tl.counter) return;
  __shared__ int s_last;
  __threadfence();                                           // my atomics are visible device-wide before the ticket is taken
  if (named) asm volatile("bar.sync 1, %0;" ::"r"(nthr)); else __syncthreads();
  if (tid == 0) s_last = (atomicAdd(tl.counter, 1u) == gridDim.x - 1) ? 1 : 0;
  if (named) asm volatile("bar.sync 1, %0;" ::"r"(nthr)); else __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (!tl.backward && tid == 0 && tl.nbt) *tl.nbt += 1;
  for (int c = tid; c < C; c += nthr) {
    const double s0 = __ldcg(sums + c), s1 = __ldcg(sums + C + c);
    if (!tl.backward) {                                      // == k_bn_finalize, training branch
      const double mean = s0 / tl.M; double var = s1 / tl.M - mean * mean; if (var < 0) var = 0;
      if (tl.rmean) {
        const double unb = tl.M > 1 ? var * tl.M / (tl.M - 1) : var;
        tl.rmean[c] = (float)((1.0 - tl.momentum) * tl.rmean[c] + tl.momentum * mean);
        tl.rvar[c] = (float)((1.0 - tl.momentum) * tl.rvar[c] + tl.momentum * unb);
      }
      const double rstd = 1.0 / sqrt(var + (double)tl.eps);
      tl.scale[c] = (float)(tl.gamma[c] * rstd); tl.shift[c] = (float)(tl.beta[c] - mean * tl.gamma[c] * rstd);
      tl.mean[c] = (float)mean; tl.rstd[c] = (float)rstd;
    } else {                                                 // == k_bn_bwd_finalize, training branch
      tl.coef[c] = tl.gamma[c] * tl.rstd[c];
      tl.coef[C + c] = (float)(s0 / tl.M);
      tl.coef[2 * C + c] = (float)(s1 / tl.M);
      if (tl.dgamma) tl.dgamma[c] = (float)s1;
      if (tl.dbeta) tl.dbeta[c] = (float)s0;
    }
  }
}

// ---------------------------------------------------------------- per-channel reduction skeleton
// blockDim.x = CV * rpi (CV = C/V channel vectors, rpi rows per iteration). Each thread owns one
// channel vector; partials are combined through shared memory, then one fp64 atomic per channel.
template <int V, int K, typename F>
__device__ __forceinline__ void chan_reduce(F f, long long M, int C, double* __restrict__ out, float* smem) {
  const int CV = C / V;
  const int rpi = blockDim.x / CV;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV;
  float acc[K][V];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < V; ++j) acc[k][j] = 0.f;
  f.prepare(cv * V);                                    // per-channel constants live in registers for the whole loop
  if (r < rpi) {
#pragma unroll 4
    for (long long m = (long long)blockIdx.x * rpi + r; m < M; m += (long long)gridDim.x * rpi) f(m, cv * V, acc);
  }
  // smem layout [rpi][K][C]
  if (r < rpi) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < V; ++j) smem[((size_t)r * K + k) * C + cv * V + j] = acc[k][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) {
    float s = 0.f;
    for (int rr = 0; rr < rpi; ++rr) s += smem[(size_t)rr * K * C + i];
    atomicAdd(out + i, (double)s);
  }
}
static inline int reduce_threads(int C, int V) { int CV = C / V; int rpi = 256 / CV; if (rpi < 1) rpi = 1; return CV * rpi; }
static inline int reduce_blocks(long long M, int threads, int C, int V) {
  int rpi = threads / (C / V);
  long long b = (M + (long long)rpi * 8 - 1) / ((long long)rpi * 8);
  long long cap = (long long)egm_num_sms() * 8;
  if (b > cap) b = cap; if (b < 1) b = 1; return (int)b;
}

// ---------------------------------------------------------------- forward statistics
template <typename T, int V>
struct StatsF {
  const T* x; long long cs, co;
  __device__ void prepare(int) {}
  __device__ void operator()(long long m, int c, float (&acc)[2][V]) const {
    FVec<V> a = ldv<V>(x + m * cs + co + c);
#pragma unroll
    for (int j = 0; j < V; ++j) { acc[0][j] += a.v[j]; acc[1][j] += a.v[j] * a.v[j]; }
  }
};
template <typename T, int V>
__global__ void k_bn_stats(StatsF<T, V> f, long long M, int C, double* out, BnTail tl) { egm_pdl_enter();
  extern __shared__ float smem[];
  chan_reduce<V, 2>(f, M, C, out, smem);
  bn_tail_run(tl, out, C, threadIdx.x, blockDim.x, false);
}
static bool bn_stats_stream_launch(const void* x, int dtype, long long M, int C, long long cstride, long long coff, double* sums, const BnTail& tl, cudaStream_t st);
struct BnFin;
static bool bn_fwd_stream_launch(const void* z, long long zcs, long long zco, const BnArgs& a, const BnFin& fin, const void* aux, void* y, long long ycs,
                                 long long yco, int dtype, long long M, int C, cudaStream_t st);
static int bn_stats_impl(const void* x, int dtype, long long M, int C, long long cstride, long long coff, double* sums, BnTail tl, cudaStream_t st) {
  EGM_REQUIRE(C >= 1 && C <= 2048, EGM_E_SHAPE, "bn_stats: C=%d unsupported", C);
  cudaMemsetAsync(sums, 0, sizeof(double) * (2 * C + (tl.counter ? 1 : 0)), st);
  if (M == 0) return EGM_OK;
  if (bn_stats_stream_launch(x, dtype, M, C, cstride, coff, sums, tl, st)) { EGM_LAUNCH_CHECK("bn_stats(stream)"); return EGM_OK; }
  int v = egm_pick_vec(C, cstride, coff);
  int threads = reduce_threads(C, v); size_t sm = (size_t)(threads / (C / v)) * 2 * C * sizeof(float);
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_bn_stats<T, V>, reduce_blocks(M, threads, C, v), threads, sm, st, 
      StatsF<T, V>{(const T*)x, cstride, coff}, M, C, sums, tl))));
  EGM_LAUNCH_CHECK("bn_stats"); return EGM_OK;
}
extern "C" int egm_bn_stats(const void* x, int dtype, long long M, int C, long long cstride, long long coff, double* sums, void* stream) {
  return bn_stats_impl(x, dtype, M, C, cstride, coff, sums, BnTail{}, (cudaStream_t)stream);
}
// training-mode statistics + finalize in one launch; `sums` holds 2*C + 1 doubles (the last one is the block ticket counter)
extern "C" int egm_bn_stats_finalize(const void* x, int dtype, long long M, int C, long long cstride, long long coff, double* sums, const float* gamma,
                                     const float* beta, float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                                     float eps, float* scale, float* shift, float* mean, float* rstd, void* stream) {
  EGM_REQUIRE(M > 0, EGM_E_SHAPE, "bn_stats_finalize: empty batch");
  BnTail tl{};
  tl.counter = (unsigned int*)(sums + 2 * C); tl.backward = 0; tl.M = (double)M;
  tl.gamma = gamma; tl.beta = beta; tl.rmean = running_mean; tl.rvar = running_var; tl.nbt = num_batches_tracked; tl.momentum = momentum; tl.eps = eps;
  tl.scale = scale; tl.shift = shift; tl.mean = mean; tl.rstd = rstd;
  return bn_stats_impl(x, dtype, M, C, cstride, coff, sums, tl, (cudaStream_t)stream);
}

// sums [2][C] -> scale/shift/mean/rstd (+ running-stat update).  training=0: use the running stats.
__global__ void k_bn_finalize(const double* __restrict__ sums, double M, const float* __restrict__ gamma, const float* __restrict__ beta,
                              float* running_mean, float* running_var, long long* nbt, float momentum, float eps, int training, int C,
                              float* scale, float* shift, float* mean_o, float* rstd_o) { egm_pdl_enter();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && nbt) *nbt += 1;
  if (c >= C) return;
  double mean, var;
  if (training) {
    mean = sums[c] / M; var = sums[C + c] / M - mean * mean; if (var < 0) var = 0;
    if (running_mean) {
      double unb = M > 1 ? var * M / (M - 1) : var;
      running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
      running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
    }
  } else { mean = running_mean[c]; var = running_var[c]; }
  double rstd = 1.0 / sqrt(var + (double)eps);
  float sc = (float)(gamma[c] * rstd);
  scale[c] = sc; shift[c] = (float)(beta[c] - mean * gamma[c] * rstd);
  if (mean_o) mean_o[c] = (float)mean;
  if (rstd_o) rstd_o[c] = (float)rstd;
}
extern "C" int egm_bn_finalize(const double* sums, long long M, const float* gamma, const float* beta, float* running_mean, float* running_var,
                               long long* num_batches_tracked, float momentum, float eps, int training, int C,
                               float* scale, float* shift, float* mean, float* rstd, void* stream) {
  egm_launch(k_bn_finalize, cdiv(C, 128), 128, 0, (cudaStream_t)stream, sums, (double)M, gamma, beta, running_mean, running_var, num_batches_tracked,
                                                                momentum, eps, training, C, scale, shift, mean, rstd);
  EGM_LAUNCH_CHECK("bn_finalize"); return EGM_OK;
}

// ---------------------------------------------------------------- apply (+activation / gate / residual)
template <typename T, int V>
__global__ void k_bn_act_fwd(const T* __restrict__ z, long long zcs, long long zco, BnArgs a, const T* __restrict__ aux, T* __restrict__ y,
                             long long ycs, long long yco, long long M, int CV, BnFin fin) { egm_pdl_enter();
  // blockDim.x = CV * rpb: every thread keeps ONE channel vector, so scale/shift are loaded (or derived from the batch sums) once
  const int rpb = blockDim.x / CV, cv = threadIdx.x % CV, r = threadIdx.x / CV, c = cv * V;
  FVec<V> sc, sh;
  bn_fin_vec<V>(fin, a, c, CV * V, blockIdx.x == 0 && r == 0, sc, sh);
#pragma unroll 2
  for (long long m = (long long)blockIdx.x * rpb + r; m < M; m += (long long)gridDim.x * rpb) {
    FVec<V> zv = ldv<V>(z + m * zcs + zco + c), xv, o;
    if (a.mode != 0) xv = ldv<V>(aux + m * (long long)(CV * V) + c);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float t = fmaf(zv.v[j], sc.v[j], sh.v[j]);
      if (a.mode == 0) o.v[j] = a.act == 1 ? fmaxf(t, 0.f) : (a.act == 2 ? sigmoidf_(t) : t);
      else if (a.mode == 1) { float s = sigmoidf_(t); o.v[j] = fmaf(s, xv.v[j], xv.v[j]); }
      else o.v[j] = fmaxf(fmaf(a.alpha, xv.v[j], t), 0.f);
    }
    stv<V>(y + m * ycs + yco + c, o);
  }
}
static inline int ew_blocks(long long M, int threads, int C, int V) {
  int rpb = threads / (C / V);
  long long b = (M + (long long)rpb * 4 - 1) / ((long long)rpb * 4);
  long long cap = (long long)egm_num_sms() * 8;
  if (b > cap) b = cap; if (b < 1) b = 1; return (int)b;
}
static int bn_act_fwd_impl(const void* z, long long z_cstride, long long z_coff, const BnArgs& a, const BnFin& fin, const void* aux, void* y,
                           long long y_cstride, long long y_coff, int dtype, long long M, int C, cudaStream_t st) {
  if (M * C == 0) return EGM_OK;
  EGM_REQUIRE(a.mode == 0 || aux, EGM_E_BADARG, "bn_act_fwd: mode %d needs aux", a.mode);
  int v = egm_pick_vec(C, z_cstride, z_coff), v2 = egm_pick_vec(C, y_cstride, y_coff); if (v2 < v) v = v2;
  if (bn_fwd_stream_launch(z, z_cstride, z_coff, a, fin, aux, y, y_cstride, y_coff, dtype, M, C, st)) { EGM_LAUNCH_CHECK("bn_act_fwd(stream)"); return EGM_OK; }
  const int threads = reduce_threads(C, v);
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_bn_act_fwd<T, V>, ew_blocks(M, threads, C, V), threads, 0, st,
      (const T*)z, z_cstride, z_coff, a, (const T*)aux, (T*)y, y_cstride, y_coff, M, C / V, fin))));
  EGM_LAUNCH_CHECK("bn_act_fwd"); return EGM_OK;
}
extern "C" int egm_bn_act_fwd(const void* z, long long z_cstride, long long z_coff, const float* scale, const float* shift, int act, int mode,
                              const void* aux, float alpha, void* y, long long y_cstride, long long y_coff, int dtype, long long M, int C, void* stream) {
  BnArgs a{scale, shift, nullptr, nullptr, nullptr, act, mode, alpha};
  return bn_act_fwd_impl(z, z_cstride, z_coff, a, BnFin{}, aux, y, y_cstride, y_coff, dtype, M, C, (cudaStream_t)stream);
}
// egm_bn_finalize (training) + egm_bn_act_fwd in ONE launch: `sums` are the batch sums a conv epilogue produced (egm_conv2d_tc_ex);
// scale / shift / mean / rstd are still written (the backward pass reads them) and the running statistics are updated.
extern "C" int egm_bn_finalize_act_fwd(const double* sums, long long Mstat, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                       long long* num_batches_tracked, float momentum, float eps, float* scale, float* shift, float* mean, float* rstd,
                                       const void* z, long long z_cstride, long long z_coff, int act, int mode, const void* aux, float alpha, void* y,
                                       long long y_cstride, long long y_coff, int dtype, long long M, int C, void* stream) {
  EGM_REQUIRE(sums && Mstat > 0 && M > 0, EGM_E_BADARG, "bn_finalize_act_fwd: needs batch sums over a non-empty batch");
  BnArgs a{scale, shift, nullptr, nullptr, nullptr, act, mode, alpha};
  BnFin fin{sums, (double)Mstat, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, scale, shift, mean, rstd};
  return bn_act_fwd_impl(z, z_cstride, z_coff, a, fin, aux, y, y_cstride, y_coff, dtype, M, C, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- backward
// g = dL/d(bn output) given dy = dL/d(block output); also the gradient routed to aux.
__device__ __forceinline__ void bn_local_grad(const BnArgs& a, float dy, float z, float sc, float sh, float aux, float& g, float& daux) {
  float t = fmaf(z, sc, sh);
  daux = 0.f;
  if (a.mode == 0) {
    if (a.act == 1) g = t > 0.f ? dy : 0.f;
    else if (a.act == 2) { float s = sigmoidf_(t); g = dy * s * (1.f - s); }
    else g = dy;
  } else if (a.mode == 1) {
    float s = sigmoidf_(t); g = dy * aux * s * (1.f - s); daux = dy * (1.f + s);
  } else {
    float m = fmaf(a.alpha, aux, t) > 0.f ? dy : 0.f; g = m; daux = a.alpha * m;
  }
}
#include "bn_stream.cuh"

template <typename T, int V>
struct BwdRedF {
  const T* dy; long long dcs, dco; const T* z; const T* aux; BnArgs a; int C;
  FVec<V> sc, sh, mu, rs;
  __device__ void prepare(int c) { sc = ldv<V>(a.scale + c); sh = ldv<V>(a.shift + c); mu = ldv<V>(a.mean + c); rs = ldv<V>(a.rstd + c); }
  __device__ void operator()(long long m, int c, float (&acc)[2][V]) const {
    FVec<V> d = ldv<V>(dy + m * dcs + dco + c), zv = ldv<V>(z + m * (long long)C + c), xv;
    if (a.mode != 0) xv = ldv<V>(aux + m * (long long)C + c);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float g, da; bn_local_grad(a, d.v[j], zv.v[j], sc.v[j], sh.v[j], a.mode ? xv.v[j] : 0.f, g, da);
      acc[0][j] += g; acc[1][j] += g * (zv.v[j] - mu.v[j]) * rs.v[j];
    }
  }
};
template <typename T, int V>
__global__ void k_bn_bwd_reduce(BwdRedF<T, V> f, long long M, int C, double* out, BnTail tl) { egm_pdl_enter();
  extern __shared__ float smem[];
  chan_reduce<V, 2>(f, M, C, out, smem);
  bn_tail_run(tl, out, C, threadIdx.x, blockDim.x, false);
}
static int bn_bwd_reduce_impl(const void* dy, long long dy_cstride, long long dy_coff, const void* z, const float* scale, const float* shift,
                              const float* mean, const float* rstd, int act, int mode, const void* aux, float alpha, int dtype, long long M, int C,
                              double* sums, BnTail tl, cudaStream_t st) {
  cudaMemsetAsync(sums, 0, sizeof(double) * (2 * C + (tl.counter ? 1 : 0)), st);
  if (M == 0) return EGM_OK;
  BnArgs a{scale, shift, mean, rstd, nullptr, act, mode, alpha};
  if (bn_stream_eligible(mode, M, C, dy_cstride, dy_coff)) {
    const size_t tail = 2 * bs::CONSUMERS * bs::V * sizeof(float);
    EGM_DISPATCH_DTYPE(dtype, {
      if (mode == 0) {
        const size_t smb = bs::ring_bytes<2>(tail);
        static bool attr[64] = {}; egm_ensure_smem(k_bn_bwd_reduce_stream<T, false>, (int)smb, attr);
        egm_launch(k_bn_bwd_reduce_stream<T, false>, bn_stream_grid(M * C, sizeof(T)), bs::THREADS, smb, st, (const T*)dy, (const T*)z, nullptr, a, M * C, C, sums, tl);
      } else {
        const size_t smb = bs::ring_bytes<3>(tail);
        static bool attr[64] = {}; egm_ensure_smem(k_bn_bwd_reduce_stream<T, true>, (int)smb, attr);
        egm_launch(k_bn_bwd_reduce_stream<T, true>, bn_stream_grid(M * C, sizeof(T)), bs::THREADS, smb, st, (const T*)dy, (const T*)z, (const T*)aux, a, M * C, C, sums, tl);
      }
    });
    EGM_LAUNCH_CHECK("bn_act_bwd_reduce(stream)"); return EGM_OK;
  }
  int v = egm_pick_vec(C, dy_cstride, dy_coff); if (v > 4) v = 4;      // 4-wide: fewer live registers -> more warps in flight
  int threads = reduce_threads(C, v); size_t sm = (size_t)(threads / (C / v)) * 2 * C * sizeof(float);
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_bn_bwd_reduce<T, V>, reduce_blocks(M, threads, C, v), threads, sm, st, 
      BwdRedF<T, V>{(const T*)dy, dy_cstride, dy_coff, (const T*)z, (const T*)aux, a, C, {}, {}, {}, {}}, M, C, sums, tl))));
  EGM_LAUNCH_CHECK("bn_act_bwd_reduce"); return EGM_OK;
}
extern "C" int egm_bn_act_bwd_reduce(const void* dy, long long dy_cstride, long long dy_coff, const void* z, const float* scale, const float* shift,
                                     const float* mean, const float* rstd, int act, int mode, const void* aux, float alpha, int dtype, long long M, int C,
                                     double* sums, void* stream) {
  return bn_bwd_reduce_impl(dy, dy_cstride, dy_coff, z, scale, shift, mean, rstd, act, mode, aux, alpha, dtype, M, C, sums, BnTail{}, (cudaStream_t)stream);
}
// training-mode backward reduction + finalize (coef[3][C], dgamma, dbeta) in one launch; `sums` holds 2*C + 1 doubles
extern "C" int egm_bn_act_bwd_reduce_finalize(const void* dy, long long dy_cstride, long long dy_coff, const void* z, const float* scale, const float* shift,
                                              const float* mean, const float* rstd, int act, int mode, const void* aux, float alpha, int dtype,
                                              long long M, int C, double* sums, const float* gamma, float* coef, float* dgamma, float* dbeta,
                                              void* stream) {
  EGM_REQUIRE(M > 0, EGM_E_SHAPE, "bn_act_bwd_reduce_finalize: empty batch");
  BnTail tl{};
  tl.counter = (unsigned int*)(sums + 2 * C); tl.backward = 1; tl.M = (double)M;
  tl.gamma = gamma; tl.rstd = const_cast<float*>(rstd); tl.coef = coef; tl.dgamma = dgamma; tl.dbeta = dbeta;
  return bn_bwd_reduce_impl(dy, dy_cstride, dy_coff, z, scale, shift, mean, rstd, act, mode, aux, alpha, dtype, M, C, sums, tl, (cudaStream_t)stream);
}

// sums -> coef[0][c] = gamma*rstd, coef[1][c] = sum(g)/M, coef[2][c] = sum(g*xhat)/M ; dgamma, dbeta
__global__ void k_bn_bwd_finalize(const double* __restrict__ sums, double M, const float* __restrict__ gamma, const float* __restrict__ rstd, int C,
                                  float* coef, float* dgamma, float* dbeta, int training) { egm_pdl_enter();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double sg = sums[c], sgx = sums[C + c];
  coef[c] = gamma[c] * rstd[c];
  coef[C + c] = training ? (float)(sg / M) : 0.f;
  coef[2 * C + c] = training ? (float)(sgx / M) : 0.f;
  if (dgamma) dgamma[c] = (float)sgx;
  if (dbeta) dbeta[c] = (float)sg;
}
extern "C" int egm_bn_bwd_finalize(const double* sums, long long M, const float* gamma, const float* rstd, int C, float* coef, float* dgamma,
                                   float* dbeta, int training, void* stream) {
  egm_launch(k_bn_bwd_finalize, cdiv(C, 128), 128, 0, (cudaStream_t)stream, sums, (double)M, gamma, rstd, C, coef, dgamma, dbeta, training);
  EGM_LAUNCH_CHECK("bn_bwd_finalize"); return EGM_OK;
}

template <typename T, int V>
__global__ void k_bn_bwd_apply(const T* __restrict__ dy, long long dcs, long long dco, const T* __restrict__ z, const T* __restrict__ aux, BnArgs a,
                               T* __restrict__ dz, T* __restrict__ daux, int daux_acc, long long M, int CV) { egm_pdl_enter();
  const int C = CV * V;
  const int rpb = blockDim.x / CV, cv = threadIdx.x % CV, r = threadIdx.x / CV, c = cv * V;
  const FVec<V> sc = ldv<V>(a.scale + c), sh = ldv<V>(a.shift + c), mu = ldv<V>(a.mean + c), rs = ldv<V>(a.rstd + c),
                k0 = ldv<V>(a.coef + c), k1 = ldv<V>(a.coef + C + c), k2 = ldv<V>(a.coef + 2 * C + c);
#pragma unroll 2
  for (long long m = (long long)blockIdx.x * rpb + r; m < M; m += (long long)gridDim.x * rpb) {
    FVec<V> d = ldv<V>(dy + m * dcs + dco + c), zv = ldv<V>(z + m * (long long)C + c), xv, o, oa;
    if (a.mode != 0) xv = ldv<V>(aux + m * (long long)C + c);
    if (a.mode != 0 && daux && daux_acc) oa = ldv<V>(daux + m * (long long)C + c);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float g, da; bn_local_grad(a, d.v[j], zv.v[j], sc.v[j], sh.v[j], a.mode ? xv.v[j] : 0.f, g, da);
      float xh = (zv.v[j] - mu.v[j]) * rs.v[j];
      o.v[j] = k0.v[j] * (g - k1.v[j] - xh * k2.v[j]);
      if (a.mode != 0) oa.v[j] = (daux_acc ? oa.v[j] : 0.f) + da;
    }
    stv<V>(dz + m * (long long)C + c, o);
    if (a.mode != 0 && daux) stv<V>(daux + m * (long long)C + c, oa);
  }
}
extern "C" int egm_bn_act_bwd_apply(const void* dy, long long dy_cstride, long long dy_coff, const void* z, const float* scale, const float* shift,
                                    const float* mean, const float* rstd, const float* coef, int act, int mode, const void* aux, float alpha,
                                    void* dz, void* daux, int daux_accumulate, int dtype, long long M, int C, void* stream) {
  if (M * C == 0) return EGM_OK;
  BnArgs a{scale, shift, mean, rstd, coef, act, mode, alpha};
  if (mode == 0 && bn_stream_eligible(mode, M, C, dy_cstride, dy_coff)) {   // the aux modes are ALU-bound in the consumer warps (measured slower)
    EGM_DISPATCH_DTYPE(dtype, {
      if (mode == 0) {
        const size_t smb = bs::ring_bytes<2>(0);
        static bool attr[64] = {}; egm_ensure_smem(k_bn_bwd_apply_stream<T, false>, (int)smb, attr);
        egm_launch(k_bn_bwd_apply_stream<T, false>, bn_stream_grid(M * C, sizeof(T)), bs::THREADS, smb, (cudaStream_t)stream, 
            (const T*)dy, (const T*)z, nullptr, a, (T*)dz, nullptr, 0, M * C, C);
      } else {
        const size_t smb = bs::ring_bytes<3>(0);
        static bool attr[64] = {}; egm_ensure_smem(k_bn_bwd_apply_stream<T, true>, (int)smb, attr);
        egm_launch(k_bn_bwd_apply_stream<T, true>, bn_stream_grid(M * C, sizeof(T)), bs::THREADS, smb, (cudaStream_t)stream, 
            (const T*)dy, (const T*)z, (const T*)aux, a, (T*)dz, (T*)daux, daux_accumulate, M * C, C);
      }
    });
    EGM_LAUNCH_CHECK("bn_act_bwd_apply(stream)"); return EGM_OK;
  }
  int v = egm_pick_vec(C, dy_cstride, dy_coff); if (v > 4) v = 4;
  const int threads = reduce_threads(C, v);
  EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_bn_bwd_apply<T, V>, ew_blocks(M, threads, C, V), threads, 0, (cudaStream_t)stream, 
      (const T*)dy, dy_cstride, dy_coff, (const T*)z, (const T*)aux, a, (T*)dz, (T*)daux, daux_accumulate, M, C / V))));
  EGM_LAUNCH_CHECK("bn_act_bwd_apply"); return EGM_OK;
}

// ---------------------------------------------------------------- per-channel sum of a tensor (conv bias gradient)
template <typename T, int V>
struct SumF {
  const T* x; long long cs, co;
  __device__ void prepare(int) {}
  __device__ void operator()(long long m, int c, float (&acc)[1][V]) const {
    FVec<V> a = ldv<V>(x + m * cs + co + c);
#pragma unroll
    for (int j = 0; j < V; ++j) acc[0][j] += a.v[j];
  }
};
template <typename T, int V>
__global__ void k_chan_sum(SumF<T, V> f, long long M, int C, double* out) { egm_pdl_enter();
  extern __shared__ float smem[];
  chan_reduce<V, 1>(f, M, C, out, smem);
}
__global__ void k_d2f(const double* s, float* d, int n) { egm_pdl_enter(); int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) d[i] = (float)s[i]; }
extern "C" int egm_channel_sum(const void* x, int dtype, long long M, int C, long long cstride, long long coff, double* scratch, float* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, st);
  if (M > 0 && bn_stats_stream_launch(x, dtype, M, C, cstride, coff, scratch, BnTail{}, st)) {
    // dense power-of-two tensors: the streaming statistics kernel (its sum-of-squares half of `scratch` is simply unused)
  } else if (M > 0) {
    int v = egm_pick_vec(C, cstride, coff);
    int threads = reduce_threads(C, v); size_t sm = (size_t)(threads / (C / v)) * C * sizeof(float);
    EGM_DISPATCH_DTYPE(dtype, EGM_DISPATCH_VEC(v, (egm_launch(k_chan_sum<T, V>, reduce_blocks(M, threads, C, v), threads, sm, st, 
        SumF<T, V>{(const T*)x, cstride, coff}, M, C, scratch))));
  }
  egm_launch(k_d2f, cdiv(C, 128), 128, 0, st, scratch, out, C);
  EGM_LAUNCH_CHECK("channel_sum"); return EGM_OK;
}

// ---------------------------------------------------------------- streaming launches declared above
static bool bn_stats_stream_launch(const void* x, int dtype, long long M, int C, long long cstride, long long coff, double* sums, const BnTail& tl, cudaStream_t st) {
  if (!bn_stream_eligible(0, M, C, cstride, coff) || (dtype != EGM_F32 && dtype != EGM_BF16)) return false;
  const size_t smb = bs::ring_bytes<1>(2 * bs::CONSUMERS * bs::V * sizeof(float));
  if (dtype == EGM_F32) {
    static bool attr[64] = {}; egm_ensure_smem(k_bn_stats_stream<float>, (int)smb, attr);
    egm_launch(k_bn_stats_stream<float>, bn_stream_grid(M * C, 4), bs::THREADS, smb, st, (const float*)x, M * C, C, sums, tl);
  } else {
    static bool attr[64] = {}; egm_ensure_smem(k_bn_stats_stream<__nv_bfloat16>, (int)smb, attr);
    egm_launch(k_bn_stats_stream<__nv_bfloat16>, bn_stream_grid(M * C, 2), bs::THREADS, smb, st, (const __nv_bfloat16*)x, M * C, C, sums, tl);
  }
  return true;
}
static bool bn_fwd_stream_launch(const void* z, long long zcs, long long zco, const BnArgs& a, const BnFin& fin, const void* aux, void* y, long long ycs,
                                 long long yco, int dtype, long long M, int C, cudaStream_t st) {
  if (!bn_stream_eligible(a.mode, M, C, zcs, zco) || ycs != C || yco != 0 || (dtype != EGM_F32 && dtype != EGM_BF16)) return false;
  if (a.mode == 1) return false;        // sigmoid gate: ALU-bound in the consumer warps, the grid-stride kernel is faster (measured)
  EGM_DISPATCH_DTYPE(dtype, {
    if (a.mode == 0) {
      const size_t smb = bs::ring_bytes<1>(0);
      static bool attr[64] = {}; egm_ensure_smem(k_bn_act_fwd_stream<T, false>, (int)smb, attr);
      egm_launch(k_bn_act_fwd_stream<T, false>, bn_stream_grid(M * C, sizeof(T)), bs::THREADS, smb, st, (const T*)z, nullptr, a, (T*)y, M * C, C, fin);
    } else {
      const size_t smb = bs::ring_bytes<2>(0);
      static bool attr[64] = {}; egm_ensure_smem(k_bn_act_fwd_stream<T, true>, (int)smb, attr);
      egm_launch(k_bn_act_fwd_stream<T, true>, bn_stream_grid(M * C, sizeof(T)), bs::THREADS, smb, st, (const T*)z, (const T*)aux, a, (T*)y, M * C, C, fin);
    }
  });
  return true;
}
