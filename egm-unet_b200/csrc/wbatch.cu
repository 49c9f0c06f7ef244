// Table-driven weight preparation / weight-gradient extraction for ALL tcgen05 convs of a step in one launch each.
//
// A train step has ~100 convs on the tensor-core path.  Each used to cost 2-6 tiny launches per pass (lift to a dense
// zero-padded weight, fold / embed for the FusionConv rewrites, bf16 pack for forward and dgrad, bias pad; and the inverse on
// the gradient side) -- ~450 launches of 2-4 us that do no useful HBM work.  The host builds one job table per model
// (egm-unet_b200/engine.py: WeightPlan); `egm_weight_prep_batch` fills every packed operand from the fp32 master parameters
// at the start of the step and `egm_wgrad_unpack_batch` scatters every packed fp32 gradient into the flat gradient buffer at
// the end of backward.
//
// Replaces (per conv): egm_conv_weight_lift + egm_kernel_embed/egm_copy_slice (fold) + egm_pack_conv_weight_tc (+ bias pad)
// and egm_unpack_conv_wgrad + egm_conv_weight_lift(mode 1) + egm_kernel_embed(mode 1)/egm_copy_slice.
#include "common.cuh"

// One conv.  Mirrored word for word by WeightPlan (20 x int64).
struct EgmWJob {
  const float* src[3];      // kind 0/1: src[0];  kind 2 (embed): 7x7, 5x5, 3x3 kernels
  const float* bsrc[3];     // biases summed into bpad (nullable)
  __nv_bfloat16* wf;        // [taps][CoutP][CinP]           forward B operand
  __nv_bfloat16* wd;        // [taps flipped][CinP][CoutP]   dgrad B operand (nullable)
  float* bpad;              // [CoutP] zero-padded (summed) bias (nullable)
  const float* dwp;         // [taps][CoutP][CinP] packed fp32 weight gradient written by egm_conv2d_wgrad_tc
  float* g[3];              // gradient destinations matching src[]
  int kind;                 // 0 = plain / grouped / zero-padded ("lifted"); 1 = 1x1 conv of cat[x,x]: W[:, :C] + W[:, C:]; 2 = 7x7+5x5+3x3 merged;
                            // 3 = 1x1 conv of (x - avgpool3 x) as one 3x3 conv (src is [Cout][Cin_g], kh = kw = 3; egm_highpass_compose)
  int Cout, Cin_g, groups;  // reference weight [Cout][Cin_g][kh][kw] (kind 1: src is [Cout][2*Cin_g])
  int kh, kw, CoutP, CinP;
  long long prep_begin;     // prefix sums of taps*CoutP*CinP   (prep element space)
  long long unpack_begin;   // prefix sums of Cout*Cin_g*taps   (unpack element space)
  long long reserved;
};
static_assert(sizeof(EgmWJob) == 160, "EgmWJob must stay 20 x 8 bytes (host mirror)");

__device__ __forceinline__ int find_job(const EgmWJob* __restrict__ jobs, int n, long long i, bool unpack) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    long long b = unpack ? jobs[mid].unpack_begin : jobs[mid].prep_begin;
    if (b <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// value of the dense (lifted) weight at (co, ci, tap); zero outside the reference weight / off the group's diagonal block
__device__ __forceinline__ float job_weight(const EgmWJob& j, int co, int ci, int t) {
  const int taps = j.kh * j.kw;
  if (co >= j.Cout) return 0.f;
  if (j.kind == 1) {
    if (ci >= j.Cin_g) return 0.f;
    const float* w = j.src[0] + (long long)co * 2 * j.Cin_g;
    return w[ci] + w[j.Cin_g + ci];
  }
  const int g = co / (j.Cout / j.groups), cil = ci - g * j.Cin_g;
  if (cil < 0 || cil >= j.Cin_g) return 0.f;
  const long long cc = (long long)co * j.Cin_g + cil;
  if (j.kind == 3) {                                  // off-centre taps bf16(-w/9), centre -8x that: exact zero response to constants
    const float wn = __bfloat162float(__float2bfloat16_rn(-j.src[0][cc] * (1.f / 9.f)));
    return t == 4 ? -8.f * wn : wn;
  }
  float v = j.src[0][cc * taps + t];
  if (j.kind == 2) {                                  // 7x7 (+)= centre-embedded 5x5, then 3x3 (same order as the eager path)
    const int r = t / 7, s = t - r * 7;
    if (r >= 1 && r <= 5 && s >= 1 && s <= 5) v += j.src[1][cc * 25 + (r - 1) * 5 + (s - 1)];
    if (r >= 2 && r <= 4 && s >= 2 && s <= 4) v += j.src[2][cc * 9 + (r - 2) * 3 + (s - 2)];
  }
  return v;
}

__global__ void k_weight_prep_batch(const EgmWJob* __restrict__ jobs, int n, long long total) { egm_pdl_enter();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const EgmWJob& j = jobs[find_job(jobs, n, i, false)];
    const long long l = i - j.prep_begin;
    const int ci = (int)(l % j.CinP); const long long q = l / j.CinP; const int co = (int)(q % j.CoutP); const int t = (int)(q / j.CoutP);
    const int taps = j.kh * j.kw;
    const __nv_bfloat16 v = __float2bfloat16_rn(job_weight(j, co, ci, t));
    j.wf[l] = v;
    if (j.wd) j.wd[((long long)(taps - 1 - t) * j.CinP + ci) * j.CoutP + co] = v;
    if (j.bpad && t == 0 && ci == 0) {
      float b = 0.f;
      if (co < j.Cout) {
#pragma unroll
        for (int k = 0; k < 3; ++k) if (j.bsrc[k]) b += j.bsrc[k][co];
      }
      j.bpad[co] = b;
    }
  }
}

__global__ void k_wgrad_unpack_batch(const EgmWJob* __restrict__ jobs, int n, long long total) { egm_pdl_enter();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const EgmWJob& j = jobs[find_job(jobs, n, i, true)];
    const long long l = i - j.unpack_begin;
    if (j.kind == 3) {                                // one element per (co, ci): dw1 = sum_t k[t] * dw3[t]
      const int cil3 = (int)(l % j.Cin_g), co3 = (int)(l / j.Cin_g);
      float s3 = 0.f;
#pragma unroll
      for (int t3 = 0; t3 < 9; ++t3) s3 += j.dwp[((long long)t3 * j.CoutP + co3) * j.CinP + cil3] * (t3 == 4 ? 8.f / 9.f : -1.f / 9.f);
      j.g[0][l] = s3;
      continue;
    }
    const int taps = j.kh * j.kw;
    const int t = (int)(l % taps); const long long cc = l / taps; const int cil = (int)(cc % j.Cin_g); const int co = (int)(cc / j.Cin_g);
    if (j.kind == 1) {                                // both halves of the [Cout][2C] weight receive the folded gradient
      const float v = j.dwp[(long long)co * j.CinP + cil];
      float* g = j.g[0] + (long long)co * 2 * j.Cin_g;
      g[cil] = v; g[j.Cin_g + cil] = v;
      continue;
    }
    const int ci = (co / (j.Cout / j.groups)) * j.Cin_g + cil;
    const float v = j.dwp[((long long)t * j.CoutP + co) * j.CinP + ci];
    j.g[0][l] = v;
    if (j.kind == 2) {
      const int r = t / 7, s = t - r * 7;
      if (r >= 1 && r <= 5 && s >= 1 && s <= 5) j.g[1][cc * 25 + (r - 1) * 5 + (s - 1)] = v;
      if (r >= 2 && r <= 4 && s >= 2 && s <= 4) j.g[2][cc * 9 + (r - 2) * 3 + (s - 2)] = v;
    }
  }
}

extern "C" int egm_wjob_bytes(void) { return (int)sizeof(EgmWJob); }

extern "C" int egm_weight_prep_batch(const void* jobs, int n_jobs, long long total_elems, void* stream) {
  if (n_jobs <= 0 || total_elems <= 0) return EGM_OK;
  EGM_REQUIRE(jobs && ((uintptr_t)jobs & 7) == 0, EGM_E_BADARG, "weight_prep_batch: job table must be an 8-byte aligned device pointer");
  egm_launch(k_weight_prep_batch, egm_grid_for(total_elems, 256), 256, 0, (cudaStream_t)stream, (const EgmWJob*)jobs, n_jobs, total_elems);
  EGM_LAUNCH_CHECK("weight_prep_batch"); return EGM_OK;
}

extern "C" int egm_wgrad_unpack_batch(const void* jobs, int n_jobs, long long total_elems, void* stream) {
  if (n_jobs <= 0 || total_elems <= 0) return EGM_OK;
  EGM_REQUIRE(jobs && ((uintptr_t)jobs & 7) == 0, EGM_E_BADARG, "wgrad_unpack_batch: job table must be an 8-byte aligned device pointer");
  egm_launch(k_wgrad_unpack_batch, egm_grid_for(total_elems, 256), 256, 0, (cudaStream_t)stream, (const EgmWJob*)jobs, n_jobs, total_elems);
  EGM_LAUNCH_CHECK("wgrad_unpack_batch"); return EGM_OK;
}
