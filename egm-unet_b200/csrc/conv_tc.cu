// tcgen05 tensor-core implicit-GEMM convolution for sm_100a: bf16 NHWC activations, fp32 accumulation in TMEM,
// operands staged by TMA (4-D NHWC tensor maps; padding = TMA out-of-bounds zero fill, so there is no im2col buffer).
// Replaces the cuDNN kernels behind nn.Conv2d in DoubleConv (src/EGM-UNet.py:44-55, src/unet.py:7-18) and every other
// dense conv with Cin % 16 == 0 and Cout % 16 == 0.
//
//  forward / dgrad  (k_conv_tc):   D[128 pixels x Cout] += A[128 pixels x 16*k ch] . B[Cout x 16*k ch]^T  per (tap, ch-chunk)
//      M = 128 output pixels (an 8x16 spatial patch), N = Cout (16..256), K = taps*Cin.  Both operands K-major.
//      dgrad is the same kernel on dy with the flipped / transposed weights (egm_pack_conv_weight_tc writes both).
//  wgrad            (k_wgrad_tc):  D[Cout x Cin] += dY[128 pixels x Cout]^T . X_shifted[128 pixels x Cin]   per tap
//      M = Cout chunk (64/128), N = Cin chunk (<= 64), K = pixels.  Both operands MN-major (channels contiguous).
//      Split over pixel ranges across CTAs; partial sums are reduced with 16-byte fp32 vector reductions into dw[tap][co][ci].
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (one TMEM lane quarter each).  smem ring of mbarrier-guarded stages; double-buffered accumulators.
#include "common.cuh"
#include <cuda.h>

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// first use of a tensor map fetches its 128-byte descriptor from global memory (~1 us): start that at kernel entry, under the barrier /
// TMEM set-up, instead of in front of the first TMA load
__device__ __forceinline__ void tmap_prefetch(const CUtensorMap* tm) { asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tm) : "memory"); }
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
               "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
               "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                 "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                 "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
// Output side of a conv: row pointer = y + pixel*cs + coff; only channels < valid are written; acc: y += result.
struct OutView { long long cs, coff; int valid, acc, vec, dense; };  // vec: every row start is 16-byte aligned; dense: all padded channels exist, plain overwrite
// 16 fp32 values -> bf16, written to the channels [0, nvalid) of the row `yp`; fast path = two 16-byte stores
template <bool RELU>
__device__ __forceinline__ void store16f(__nv_bfloat16* yp, float (&f)[16], int nvalid, const OutView& o) {
  if constexpr (RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
  }
  if (nvalid >= 16 && o.vec && !o.acc) {
    uint4 q[2]; __nv_bfloat162* qb = reinterpret_cast<__nv_bfloat162*>(q);
#pragma unroll
    for (int j = 0; j < 8; ++j) qb[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    *reinterpret_cast<uint4*>(yp) = q[0];
    *reinterpret_cast<uint4*>(yp + 8) = q[1];
    return;
  }
  if (nvalid <= 0) return;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    if (nvalid >= hf * 8 + 8 && o.vec) {
      uint4 q; __nv_bfloat162* qb = reinterpret_cast<__nv_bfloat162*>(&q);
      if (o.acc) {
        q = *reinterpret_cast<const uint4*>(yp + hf * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j) { float2 old = __bfloat1622float2(qb[j]); f[hf * 8 + 2 * j] += old.x; f[hf * 8 + 2 * j + 1] += old.y; }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) qb[j] = __floats2bfloat162_rn(f[hf * 8 + 2 * j], f[hf * 8 + 2 * j + 1]);
      *reinterpret_cast<uint4*>(yp + hf * 8) = q;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (hf * 8 + j < nvalid) {
          float t = f[hf * 8 + j];
          if (o.acc) t += __bfloat162float(yp[hf * 8 + j]);
          yp[hf * 8 + j] = __float2bfloat16_rn(t);
        }
    }
  }
}
// convert 16 fp32 accumulator columns (+bias, optional ReLU) to bf16 and store them as two 16-byte vectors (dense rows)
template <bool RELU>
__device__ __forceinline__ void store16(__nv_bfloat16* yp, const uint32_t* v, const float* bp) {
  uint4 o[2]; __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float f0 = __uint_as_float(v[2 * j]), f1 = __uint_as_float(v[2 * j + 1]);
    if (bp) { f0 += bp[2 * j]; f1 += bp[2 * j + 1]; }
    if constexpr (RELU) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
    ob[j] = __floats2bfloat162_rn(f0, f1);
  }
  *reinterpret_cast<uint4*>(yp) = o[0];
  *reinterpret_cast<uint4*>(yp + 8) = o[1];
}
// 16 accumulator columns starting at channel c of the row `yp` (already offset by c), any output view
template <bool RELU>
__device__ __forceinline__ void store16v(__nv_bfloat16* yp, const uint32_t* v, const float* bp, int nvalid, const OutView& o) {
  if (nvalid >= 16 && o.vec && !o.acc) { store16<RELU>(yp, v, bp); return; }
  if (nvalid <= 0) return;
  float f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + (bp ? bp[j] : 0.f);
  store16f<RELU>(yp, f, nvalid, o);
}
// ---- epilogue modes (template parameter EPI of the forward kernels)
//   EPI_PLAIN  y = acc + bias
//   EPI_RELU   y = relu(acc + bias)          inference: BatchNorm folded into the weights (scale) and the bias (shift), ReLU here
//   16/32/64   y = acc + bias AND per-channel sum / sum of squares of the fp32 accumulators for train-mode BatchNorm
//              (Cout == EPI).  Each epilogue thread owns one pixel row of every tile its CTA processes, so it keeps running
//              sums of its row's EPI channels in registers across ALL tiles (2 FP ops per value, no cross-lane traffic per tile);
//              one shuffle + shared-memory + fp64-atomic reduction per CTA at the very end.  (Round 1 reduced per tile with a
//              32-lane butterfly and lost more in the narrow layers' epilogue than the separate statistics pass cost.)
constexpr int EPI_PLAIN = 0, EPI_RELU = 1;
template <int NC>
__device__ __forceinline__ void stats_accum16(float (&s1)[NC], float (&s2)[NC], int c0, const float (&f)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) { s1[c0 + j] += f[j]; s2[c0 + j] = fmaf(f[j], f[j], s2[c0 + j]); }
}
// red: shared [4 warps][2][NC] floats; out: global [2][cvalid] doubles (atomically accumulated); ew = epilogue warp 0..3
template <int NC>
__device__ __forceinline__ void stats_flush(const float (&s1)[NC], const float (&s2)[NC], float* red, double* __restrict__ out, int cvalid, int ew, int lane,
                                            int group = 0) {
  red += group * 8 * NC;                                 // each epilogue group of 4 warps reduces on its own (named barrier 1 + group)
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float a = s1[c], b = s2[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) { red[ew * 2 * NC + c] = a; red[ew * 2 * NC + NC + c] = b; }
  }
  asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
  for (int i = ew * 32 + lane; i < 2 * NC; i += 128) {
    const int k = i / NC, c = i - k * NC;
    if (c < cvalid) atomicAdd(out + (long long)k * cvalid + c, (double)red[i] + (double)red[2 * NC + i] + (double)red[4 * NC + i] + (double)red[6 * NC + i]);
  }
}
// one elected lane of a converged warp (keeps the surrounding control flow warp-uniform, so loop state lives in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 16 consecutive fp32 accumulators -> four 16-byte vector reductions (REDG.ADD.F32x4).  The split-K wgrad epilogues were bound by
// the LSU's ~1.3 cycles per lane and reduction INSTRUCTION (scalar red: 256->256 @30^2 spent 33 of its 45 us per CTA there); the packed
// gradient is laid out [tap][Cout][Cin] so that the 16 columns a thread reads from TMEM are contiguous in memory.
__device__ __forceinline__ void red_add16(float* dst, const uint32_t (&v)[16]) {
#pragma unroll
  for (int m = 0; m < 4; ++m)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * m), "f"(__uint_as_float(v[4 * m])), "f"(__uint_as_float(v[4 * m + 1])),
                 "f"(__uint_as_float(v[4 * m + 2])), "f"(__uint_as_float(v[4 * m + 3]))
                 : "memory");
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version=1 <<46 | layout <<61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32, a=b=BF16, majors, N>>3 at bit 17, M>>4 at bit 24
__device__ __host__ inline uint32_t umma_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
static inline int swz_layout(int row_bytes) { return row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6); }   // SWIZZLE_128B / 64B / 32B

// ------------------------------------------------------------------ tensor-map creation (driver entry point, no libcuda link)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)p;
  }
  return fn;
}
static CUtensorMapSwizzle swz_enum(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}
// Channel-strided view of an NHWC bf16 tensor: channel c (< valid) of pixel p lives at ptr[p*cs + coff + c].  A conv reads such
// a view directly: the TMA box may be wider than `valid` (out-of-bounds channels are zero-filled), which is how thin and
// channel-sliced inputs are padded to the 16-channel MMA granularity without a staging copy.
struct NhwcView { const void* ptr; long long cs, coff; int valid; };
static inline NhwcView dense_view(const void* ptr, int C) { return NhwcView{ptr, C, 0, C}; }
static inline bool view_tma_ok(const NhwcView& v) { return v.cs % 8 == 0 && v.coff % 8 == 0 && (((uintptr_t)v.ptr) & 15) == 0; }
// view [N,H,W,valid] -> 4-D map (C, W, H, N), box (bc, bw, bh, 1)
static int make_map_nhwc(CUtensorMap* tm, const NhwcView& v, int N, int H, int W, int bc, int bw, int bh) {
  PFN_encodeTiled enc = get_encode();
  EGM_REQUIRE(enc, EGM_E_ARCH, "cuTensorMapEncodeTiled unavailable");
  EGM_REQUIRE(view_tma_ok(v), EGM_E_ALIGN, "tcgen05 conv: a TMA operand needs a 16-byte aligned base and channel stride/offset that are multiples of 8");
  const void* ptr = (const __nv_bfloat16*)v.ptr + v.coff;
  const int C = v.valid;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)v.cs * 2, (cuuint64_t)v.cs * 2 * W, (cuuint64_t)v.cs * 2 * W * H};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz_enum(bc * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGM_REQUIRE(r == CUDA_SUCCESS, EGM_E_BADARG, "cuTensorMapEncodeTiled(nhwc %dx%dx%dx%d box %d,%d,%d) failed: %d", N, H, W, C, bc, bw, bh, (int)r);
  return EGM_OK;
}
// packed weights [taps][Cout][Cin] bf16 -> 3-D map (Cin, Cout, taps), box (bc, Cout, 1)
static int make_map_w(CUtensorMap* tm, const void* ptr, int taps, int Cout, int Cin, int bc, int brows) {
  PFN_encodeTiled enc = get_encode();
  EGM_REQUIRE(enc, EGM_E_ARCH, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cin * 2 * Cout};
  cuuint32_t box[3] = {(cuuint32_t)bc, (cuuint32_t)brows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz_enum(bc * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGM_REQUIRE(r == CUDA_SUCCESS, EGM_E_BADARG, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  return EGM_OK;
}

// Persistent CTAs visit tile = blockIdx.x + k * gridDim.x.  Decomposing every tile index into (image, tile row, tile column, Cout slice)
// costs three integer divisions (~100 dependent instructions) per tile in EVERY role warp -- on the thin layers, whose tile needs one
// MMA, that was a third of the epilogue's per-tile critical path (ncu source view: the IABS/MUFU.RCP division sequences right behind
// the accumulator wait).  The walk below divides once and then advances by the constant stride with carries.
struct TileWalk {
  int n, th, tw, ck;            // image, tile row, tile column, Cout slice
  int dn, dth, dtw, dck;        // the grid stride in the same mixed radix
  __device__ __forceinline__ TileWalk(int tile0, int step, int tilesH, int tilesW, int chunks) {
    split(tile0, tilesH, tilesW, chunks, n, th, tw, ck);
    split(step, tilesH, tilesW, chunks, dn, dth, dtw, dck);
  }
  static __device__ __forceinline__ void split(int t, int tilesH, int tilesW, int chunks, int& a, int& b, int& c, int& d) {
    d = t % chunks; t /= chunks;
    c = t % tilesW; t /= tilesW;
    b = t % tilesH; a = t / tilesH;
  }
  __device__ __forceinline__ void next(int tilesH, int tilesW, int chunks) {
    ck += dck; tw += dtw; th += dth; n += dn;
    if (ck >= chunks) { ck -= chunks; ++tw; }
    if (tw >= tilesW) { tw -= tilesW; ++th; }
    if (th >= tilesH) { th -= tilesH; ++n; }
  }
};

constexpr int TILE_H = 8, TILE_W = 16, TILE_PIX = 128;
constexpr int TC_THREADS = 192;

// =================================================================== forward / dgrad (per-tap tiles; any dilation / channel count)
struct ConvTcParams {
  int N, H, W, Cin, Cout, kh, kw, dil, pad;
  int tilesH, tilesW, numTiles, kChunks, bkc;       // bkc = channels per K chunk (64/32/16)
  int stages, aBytes, bStride, tmemCols, accCols;
  int nChunk, coChunks;                               // Cout is processed in coChunks slices of nChunk (<= 256) channels
  int wres;                                           // 1: the whole packed weight tensor stays resident in smem (single Cout slice)
  int kps;                                            // 9: thin 3x3 layers -- all 9 tap tiles of an output tile share ONE pipeline stage
                                                      // (one barrier round trip, straight-line TMA issue and MMA issue per tile); else 1
  OutView out;
  double* stats;                                      // EPI >= 16: [2][out.valid] per-channel sum / sum of squares (atomically accumulated)
};

// KPS consecutive K iterations of one stage, KS MMAs each, as straight-line code
template <int KPS, int KS>
__device__ __forceinline__ void issue_group(uint32_t d, uint64_t ad, uint64_t bd, uint32_t aStep, uint32_t bStep, uint32_t idesc) {
#pragma unroll
  for (int j = 0; j < KPS; ++j) {
#pragma unroll
    for (int k = 0; k < KS; ++k) umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, (j | k) ? 1u : 0u);
    ad += aStep; bd += bStep;
  }
}

template <int EPI, int NG>      // NG epilogue groups, see k_conv_tc_halo
__global__ void __launch_bounds__(64 + 128 * NG, 1) k_conv_tc(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                                                          __nv_bfloat16* __restrict__ y, const float* __restrict__ bias, ConvTcParams p) { egm_pdl_enter();
  if (threadIdx.x == 0) { tmap_prefetch(&tmX); tmap_prefetch(&tmW); }
  constexpr bool RELU = EPI == EPI_RELU;
  constexpr int NSTAT = EPI >= 16 ? EPI : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int taps = p.kh * p.kw;
  const int kIters = taps * p.kChunks;
  uint8_t* sA = smem;
  const size_t stageA = (size_t)p.kps * p.aBytes;                      // one stage holds kps A tiles
  uint8_t* sB = sA + (size_t)p.stages * stageA;                         // ring of weight tiles, or the resident [taps][kChunks] tiles
  const int nB = p.wres ? kIters : p.stages;
  uint64_t* full = (uint64_t*)(sB + (size_t)nB * p.bStride);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = (uint32_t*)(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }       // tempty: one arrival per epilogue WARP
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tilesPerImg = p.tilesH * p.tilesW;

  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t bBytes = (uint32_t)(p.nChunk * p.bkc * 2);
    if (p.wres && leader) {
      mbar_expect_tx(wfull, bBytes * (uint32_t)kIters);
      for (int t = 0; t < taps; ++t)
        for (int kc = 0; kc < p.kChunks; ++kc) tma_load_3d(sB + (size_t)(t * p.kChunks + kc) * p.bStride, &tmW, wfull, kc * p.bkc, 0, t);
    }
    const uint32_t txBytes = (uint32_t)p.aBytes + (p.wres ? 0u : bBytes);
    int s = 0; uint32_t ph = 0;
    TileWalk tw_(blockIdx.x, gridDim.x, p.tilesH, p.tilesW, p.coChunks);
    for (int tile = blockIdx.x; tile < p.numTiles; tile += gridDim.x, tw_.next(p.tilesH, p.tilesW, p.coChunks)) {
      const int n = tw_.n;
      const int h0 = tw_.th * TILE_H, w0 = tw_.tw * TILE_W, co0 = tw_.ck * p.nChunk;
      if (p.kps == 9) {
        // thin 3x3 layer: the per-tap loop below costs ~150 dependent instructions per tap in this single warp (ncu: the producer
        // was the busiest warp) -- here one wait, one expect_tx and nine back-to-back TMA issues with compile-time tap offsets
        mbar_wait(&empty[s], ph ^ 1);
        if (leader) {
          mbar_expect_tx(&full[s], (uint32_t)(9 * p.aBytes));
          uint8_t* dst = sA + (size_t)s * stageA;
#pragma unroll
          for (int tt = 0; tt < 9; ++tt)
            tma_load_4d(dst + (size_t)tt * p.aBytes, &tmX, &full[s], 0, w0 + (tt % 3 - 1) * p.dil, h0 + (tt / 3 - 1) * p.dil, n);
        }
        if (++s == p.stages) { s = 0; ph ^= 1; }
        continue;
      }
      int dh = -p.pad, t = 0;
      for (int kr = 0; kr < p.kh; ++kr, dh += p.dil) {
        int dw = -p.pad;
        for (int kc_ = 0; kc_ < p.kw; ++kc_, dw += p.dil, ++t) {
          for (int kc = 0; kc < p.kChunks; ++kc) {
            mbar_wait(&empty[s], ph ^ 1);
            if (leader) {
              mbar_expect_tx(&full[s], txBytes);
              tma_load_4d(sA + (size_t)s * p.aBytes, &tmX, &full[s], kc * p.bkc, w0 + dw, h0 + dh, n);
              if (!p.wres) tma_load_3d(sB + (size_t)s * p.bStride, &tmW, &full[s], kc * p.bkc, co0, t);
            }
            if (++s == p.stages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc(128, p.nChunk, 0, 0);
    const int rowB = p.bkc * 2;                       // bytes per smem row == swizzle span
    const uint32_t layout = rowB == 128 ? 2u : (rowB == 64 ? 4u : 6u);
    const uint32_t sbo = 8u * rowB;                   // 8-row core-matrix group stride
    const int ksteps = p.bkc / 16;
    const uint64_t adBase = umma_desc(smem_u32(sA), 16, sbo, layout), bdBase = umma_desc(smem_u32(sB), 16, sbo, layout);
    const uint32_t aStep = (uint32_t)p.aBytes >> 4, bStep = (uint32_t)p.bStride >> 4;
    if (p.wres) mbar_wait(wfull, 0);
    int s = 0; uint32_t ph = 0; int acc = 0; uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < p.numTiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], aph ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(acc * p.accCols);
      uint32_t accf = 0;
      // one copy of the K loop per K-steps-per-chunk value: no data-dependent branch between two MMAs of the issuing thread
#define EGM_V1_KLOOP(KS)                                                                                              \
      for (int it = 0; it < kIters; ++it) {                                                                            \
        mbar_wait(&full[s], ph);                                                                                       \
        tc_fence_after();                                                                                              \
        if (leader) {                                                                                                  \
          const uint64_t ad = adBase + (uint64_t)(s * aStep), bd = bdBase + (uint64_t)((p.wres ? it : s) * bStep);    \
          _Pragma("unroll") for (int k = 0; k < KS; ++k) { umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, accf); accf = 1; } \
          umma_commit(&empty[s]);                                                                                      \
        }                                                                                                              \
        accf = 1;                                                                                                      \
        __syncwarp();                                                                                                  \
        if (++s == p.stages) { s = 0; ph ^= 1; }                                                                       \
      }
      if (p.kps == 9) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (leader) {
          const uint64_t ad = adBase + (uint64_t)(s * (uint32_t)(stageA >> 4));
          if (ksteps == 1) issue_group<9, 1>(d, ad, bdBase, aStep, bStep, idesc); else issue_group<9, 2>(d, ad, bdBase, aStep, bStep, idesc);
          umma_commit(&empty[s]);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      } else
      if (ksteps == 4) { EGM_V1_KLOOP(4) } else if (ksteps == 2) { EGM_V1_KLOOP(2) } else { EGM_V1_KLOOP(1) }
#undef EGM_V1_KLOOP
      if (leader) umma_commit(&tfull[acc]);
      __syncwarp();
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  } else {
    const int q = warp & 3;                              // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;                     // epilogue group: tiles grp, grp + NG, ... of this CTA
    const int row = q * 32 + lane;                       // pixel index inside the 8x16 patch
    int acc = grp; uint32_t aph = 0;                     // two accumulator buffers: with NG == 2 each group owns one
    float s1[NSTAT], s2[NSTAT];
#pragma unroll
    for (int c = 0; c < NSTAT; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
    TileWalk tw_(blockIdx.x + grp * gridDim.x, NG * gridDim.x, p.tilesH, p.tilesW, p.coChunks);
    for (int tile = blockIdx.x + grp * gridDim.x; tile < p.numTiles; tile += NG * gridDim.x, tw_.next(p.tilesH, p.tilesW, p.coChunks)) {
      const int chunk = tw_.ck, n = tw_.n;
      const int h = tw_.th * TILE_H + row / TILE_W, w = tw_.tw * TILE_W + row % TILE_W;
      const bool valid = h < p.H && w < p.W;
      __nv_bfloat16* yp = y + (((long long)n * p.H + h) * p.W + w) * p.out.cs + p.out.coff + chunk * p.nChunk;
      const float* bp = bias ? bias + chunk * p.nChunk : nullptr;
      const int nv = p.out.valid - chunk * p.nChunk;                 // channels of this slice that exist in y
      const int cend = nv < p.nChunk ? nv : p.nChunk;
      mbar_wait(&tfull[acc], aph);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.accCols);
      if constexpr (EPI >= 16) {                                      // train-mode BN statistics from the fp32 accumulators (Cout == EPI, one slice)
#pragma unroll
        for (int c = 0; c < EPI; c += 16) {
          uint32_t v[16];
          tmem_ld16(t0 + c, v);
          tmem_ld_wait();
          if (valid) {
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + (bp ? bp[c + j] : 0.f);
            stats_accum16<NSTAT>(s1, s2, c, f);
            store16f<false>(yp + c, f, nv - c, p.out);
          }
        }
      } else if (p.out.dense && (p.nChunk & 31) == 0) {               // hot path of the DoubleConv layers: no per-store decisions
        for (int c = 0; c < p.nChunk; c += 32) {
          uint32_t v[32];
          tmem_ld32(t0 + c, v);
          tmem_ld_wait();
          if (valid) { store16<RELU>(yp + c, v, bp ? bp + c : nullptr); store16<RELU>(yp + c + 16, v + 16, bp ? bp + c + 16 : nullptr); }
        }
      } else if (p.out.dense) {
        for (int c = 0; c < p.nChunk; c += 16) {
          uint32_t v[16];
          tmem_ld16(t0 + c, v);
          tmem_ld_wait();
          if (valid) store16<RELU>(yp + c, v, bp ? bp + c : nullptr);
        }
      } else {
        for (int c = 0; c < cend; c += 16) {
          uint32_t v[16];
          tmem_ld16(t0 + c, v);
          tmem_ld_wait();
          if (valid) store16v<RELU>(yp + c, v, bp ? bp + c : nullptr, nv - c, p.out);
        }
      }
      tc_fence_before();
      __syncwarp();                                      // every lane's tcgen05.ld has retired (wait::ld above) before the warp releases the accumulator
      if (lane == 0) mbar_arrive(&tempty[acc]);          // 4 arrivals per tile instead of 128 serialised shared-memory atomics
      acc += NG; if (acc >= 2) { acc -= 2; aph ^= 1; }
    }
    if constexpr (EPI >= 16) stats_flush<NSTAT>(s1, s2, (float*)(tmem_slot + 4), p.stats, p.out.valid, q, lane, grp);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)p.tmemCols); }
}

#include <stdlib.h>
static int pow2_cols(int c) { int v = 32; while (v < c) v <<= 1; return v; }
// channels per chunk (smem row = 128/64/32 bytes).  A tail chunk that runs past C is zero-filled by TMA (out-of-bounds box).
static int pick_bkc(int C) { return C > 32 ? 64 : (C > 16 ? 32 : 16); }


// =================================================================== forward / dgrad, halo-reuse variant (small-channel layers)
// One (16+2p)x(8+2p) input HALO tile per 16x8 output tile is fetched by a single TMA load; the k*k taps are then consumed by
// UMMA descriptors whose start address is shifted by (r*haloW + s) rows and whose 8-row-group stride (SBO) is the halo row
// pitch -- the 128/64/32B swizzle is a function of the absolute smem address, so shifted starts read exactly what TMA
// wrote.  The whole packed weight tensor [taps][Cout][Cin] stays resident in smem (loaded once per CTA).  Per output tile
// this issues 1 TMA load of ~180 rows instead of k*k loads of 128 rows + k*k weight tiles.
struct ConvHaloParams {
  int N, H, W, Cin, Cout, kh, kw, dil, pad;
  int tilesH, tilesW, numTiles;
  int rowB, haloW, haloH, haloBytes, haloStride, wTapStride, stages, tmemCols, accCols;
  int nacc;                                          // TMEM accumulator buffers in flight (2..8)
  int ksteps;                                        // MMAs (K=16) per tap = ceil(Cin/16); rowB may be wider than Cin*2 (zero-filled)
  int exp;                                           // EGM_DIAG builds only: EGM_EXP bit mask (timing ablations; results are wrong):
                                                     // 1 = no global stores, 4 = no MMAs, 8 = no TMA loads.  Compiled out of the shipped .so.
  OutView out;
  double* stats;                                     // EPI >= 16: [2][out.valid] per-channel sum / sum of squares (atomically accumulated)
};
constexpr int HT_H = 16, HT_W = 8;
#ifdef EGM_DIAG
#define EGM_EXPBIT(b) (p.exp & (b))
#else
#define EGM_EXPBIT(b) (0)
#endif

// NG = epilogue warp GROUPS (4 warps each = the four TMEM lane quarters).  The epilogue of a tile is one dependent chain per warp
// (accumulator wait -> tcgen05.ld -> convert -> store -> release: ~115 instructions at ~8 cycles each, ncu source view of the
// 16->16 1x1 layer: the four epilogue warps were busy 80 % of the time while the TMA and MMA warps idled), so thin layers -- one to
// nine MMAs per tile -- looked bound by it.  With NG groups, group g takes the tiles g, g + NG, ... of its CTA (accumulator buffers
// are handed out round-robin, so consecutive tiles are in different buffers anyway) and NG epilogues overlap.  Measured neutral
// (see launch_conv_halo), kept as a tuning knob.
template <int EPI, int NG>
__global__ void __launch_bounds__(64 + 128 * NG, 1) k_conv_tc_halo(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                                                               __nv_bfloat16* __restrict__ y, const float* __restrict__ bias, ConvHaloParams p) { egm_pdl_enter();
  if (threadIdx.x == 0) { tmap_prefetch(&tmX); tmap_prefetch(&tmW); }
  constexpr bool RELU = EPI == EPI_RELU;
  constexpr int NSTAT = EPI >= 16 ? EPI : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int taps = p.kh * p.kw;
  uint8_t* sW = smem;
  uint8_t* sA = sW + (size_t)taps * p.wTapStride;
  uint64_t* full = (uint64_t*)(sA + (size_t)p.stages * p.haloStride);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint64_t* tempty = tfull + 8;
  uint64_t* wfull = tempty + 8;
  uint32_t* tmem_slot = (uint32_t*)(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < p.nacc; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }  // tempty: one arrival per epilogue WARP
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- TMA producer: all 32 lanes run the (uniform) loop, one elected lane issues
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(wfull, (uint32_t)(taps * p.Cout * p.rowB));
      for (int t = 0; t < taps; ++t) tma_load_3d(sW + (size_t)t * p.wTapStride, &tmW, wfull, 0, 0, t);
    }
    int s = 0; uint32_t ph = 0;
    TileWalk tw_(blockIdx.x, gridDim.x, p.tilesH, p.tilesW, 1);
    for (int tile = blockIdx.x; tile < p.numTiles; tile += gridDim.x, tw_.next(p.tilesH, p.tilesW, 1)) {
      const int n = tw_.n;
      const int h0 = tw_.th * HT_H, w0 = tw_.tw * HT_W;
      mbar_wait(&empty[s], ph ^ 1);
      if (leader) {
        if (EGM_EXPBIT(8)) mbar_arrive(&full[s]);
        else {
          mbar_expect_tx(&full[s], (uint32_t)p.haloBytes);
          tma_load_4d(sA + (size_t)s * p.haloStride, &tmX, &full[s], 0, w0 - p.pad, h0 - p.pad, n);
        }
      }
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: warp-uniform loops; descriptors are pure functions of uniform loop counters
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc(128, p.Cout, 0, 0);
    const uint32_t layout = p.rowB == 128 ? 2u : (p.rowB == 64 ? 4u : 6u);
    const uint32_t sboA = (uint32_t)(p.haloW * p.rowB);
    const int ksteps = p.ksteps;
    const uint64_t bd0 = umma_desc(smem_u32(sW), 16, 8u * p.rowB, layout);
    const uint32_t bTap = (uint32_t)p.wTapStride >> 4;                  // B descriptor advance per tap
    const uint32_t aCol = (uint32_t)(p.dil * p.rowB) >> 4;              // A advance per kernel column
    const uint32_t aRow = (uint32_t)(p.dil * p.haloW * p.rowB) >> 4;    // A advance per kernel row
    mbar_wait(wfull, 0);
    int s = 0; uint32_t ph = 0; int acc = 0; uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < p.numTiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], aph ^ 1);
      mbar_wait(&full[s], ph);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(acc * p.accCols);
      const uint64_t ad0 = umma_desc(smem_u32(sA + (size_t)s * p.haloStride), 16, sboA, layout);
      if (leader) {
        uint32_t accf = 0;
        uint64_t bd = bd0, adr = ad0;
        // 3x3 kernels (the DoubleConv layers): mode decided once per tile, the 9 taps are straight-line code -- every branch in this
        // single-thread issue path costs about as much as an MMA
        if (p.kh == 1 && !EGM_EXPBIT(4)) {
          if (ksteps == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d, ad0 + 2 * k, bd0 + 2 * k, idesc, k ? 1u : 0u);
          } else if (ksteps == 2) { umma_bf16(d, ad0, bd0, idesc, 0u); umma_bf16(d, ad0 + 2, bd0 + 2, idesc, 1u); }
          else umma_bf16(d, ad0, bd0, idesc, 0u);
        } else if (p.kh == 7 && p.kw == 7 && ksteps <= 2 && !EGM_EXPBIT(4)) {          // merged FusionConv 7x7 (16 / 32 channels)
          if (ksteps == 1) {
#pragma unroll
            for (int r = 0; r < 7; ++r) {
              uint64_t ad = adr;
#pragma unroll
              for (int c = 0; c < 7; ++c) { umma_bf16(d, ad, bd, idesc, (r | c) ? 1u : 0u); ad += aCol; bd += bTap; }
              adr += aRow;
            }
          } else {
#pragma unroll
            for (int r = 0; r < 7; ++r) {
              uint64_t ad = adr;
#pragma unroll
              for (int c = 0; c < 7; ++c) {
                umma_bf16(d, ad, bd, idesc, (r | c) ? 1u : 0u); umma_bf16(d, ad + 2, bd + 2, idesc, 1u);
                ad += aCol; bd += bTap;
              }
              adr += aRow;
            }
          }
        } else if (p.kh == 3 && p.kw == 3 && !EGM_EXPBIT(4)) {
          if (ksteps == 2) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              uint64_t ad = adr;
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                umma_bf16(d, ad, bd, idesc, (r | c) ? 1u : 0u); umma_bf16(d, ad + 2, bd + 2, idesc, 1u);
                ad += aCol; bd += bTap;
              }
              adr += aRow;
            }
          } else if (ksteps == 4) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              uint64_t ad = adr;
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                umma_bf16(d, ad, bd, idesc, (r | c) ? 1u : 0u);
#pragma unroll
                for (int k = 1; k < 4; ++k) umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, 1u);
                ad += aCol; bd += bTap;
              }
              adr += aRow;
            }
          } else {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              uint64_t ad = adr;
#pragma unroll
              for (int c = 0; c < 3; ++c) { umma_bf16(d, ad, bd, idesc, (r | c) ? 1u : 0u); ad += aCol; bd += bTap; }
              adr += aRow;
            }
          }
        } else
        for (int r = 0; r < (EGM_EXPBIT(4) ? 0 : p.kh); ++r) {
          uint64_t ad = adr;
          for (int c = 0; c < p.kw; ++c) {
            if (ksteps == 4) {
#pragma unroll
              for (int k = 0; k < 4; ++k) { umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, accf); accf = 1; }
            } else if (ksteps == 2) {
#pragma unroll
              for (int k = 0; k < 2; ++k) { umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, accf); accf = 1; }
            } else { umma_bf16(d, ad, bd, idesc, accf); accf = 1; }
            ad += aCol; bd += bTap;
          }
          adr += aRow;
        }
        umma_commit(&empty[s]);
        umma_commit(&tfull[acc]);
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1; }
      if (++acc == p.nacc) { acc = 0; aph ^= 1; }
    }
  } else {
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;                     // epilogue group: tiles grp, grp + NG, ... of this CTA
    const int row = q * 32 + lane;                       // pixel index inside the 16x8 patch (row-major, 8 wide)
    int acc = grp; uint32_t aph = 0;                     // p.nacc >= NG (host), so the first NG tiles sit in buffers 0 .. NG-1
    float s1[NSTAT], s2[NSTAT];
#pragma unroll
    for (int c = 0; c < NSTAT; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
    TileWalk tw_(blockIdx.x + grp * gridDim.x, NG * gridDim.x, p.tilesH, p.tilesW, 1);
    for (int tile = blockIdx.x + grp * gridDim.x; tile < p.numTiles; tile += NG * gridDim.x, tw_.next(p.tilesH, p.tilesW, 1)) {
      const int n = tw_.n;
      const int h = tw_.th * HT_H + row / HT_W, w = tw_.tw * HT_W + row % HT_W;
      const bool valid = h < p.H && w < p.W;
      __nv_bfloat16* yp = y + (((long long)n * p.H + h) * p.W + w) * p.out.cs + p.out.coff;
      const int nv = p.out.valid, cend = nv < p.Cout ? nv : p.Cout;
      mbar_wait(&tfull[acc], aph);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.accCols);
      if constexpr (EPI >= 16) {                                      // train-mode BN statistics from the fp32 accumulators (Cout == EPI)
#pragma unroll
        for (int c = 0; c < EPI; c += 16) {
          uint32_t v[16];
          tmem_ld16(t0 + c, v);
          tmem_ld_wait();
          if (valid) {
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + (bias ? bias[c + j] : 0.f);
            stats_accum16<NSTAT>(s1, s2, c, f);
            store16f<false>(yp + c, f, nv - c, p.out);
          }
        }
      } else if (p.out.dense && (p.Cout & 31) == 0) {
        for (int c = 0; c < p.Cout; c += 32) {
          uint32_t v[32];
          tmem_ld32(t0 + c, v);
          tmem_ld_wait();
          if (valid && !EGM_EXPBIT(1)) { store16<RELU>(yp + c, v, bias ? bias + c : nullptr); store16<RELU>(yp + c + 16, v + 16, bias ? bias + c + 16 : nullptr); }
        }
      } else if (p.out.dense) {
        for (int c = 0; c < p.Cout; c += 16) {
          uint32_t v[16];
          tmem_ld16(t0 + c, v);
          tmem_ld_wait();
          if (valid) store16<RELU>(yp + c, v, bias ? bias + c : nullptr);
        }
      } else {
        for (int c = 0; c < cend; c += 16) {
          uint32_t v[16];
          tmem_ld16(t0 + c, v);
          tmem_ld_wait();
          if (valid) store16v<RELU>(yp + c, v, bias ? bias + c : nullptr, nv - c, p.out);
        }
      }
      tc_fence_before();
      __syncwarp();                                      // every lane's tcgen05.ld has retired (wait::ld above) before the warp releases the accumulator
      if (lane == 0) mbar_arrive(&tempty[acc]);          // 4 arrivals per tile instead of 128 serialised shared-memory atomics
      acc += NG; if (acc >= p.nacc) { acc -= p.nacc; aph ^= 1; }
    }
    if constexpr (EPI >= 16) stats_flush<NSTAT>(s1, s2, (float*)(tmem_slot + 4), p.stats, p.out.valid, q, lane, grp);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)p.tmemCols); }
}

static bool halo_disabled() { static int v = -1; if (v < 0) { const char* e = getenv("EGM_NO_HALO"); v = (e && e[0] == '1') ? 1 : 0; } return v == 1; }
// eligible: single K chunk (Cin in {16,32,64}), Cout <= 256, halo of <= 3 pixels, resident weights <= 100 KB
static bool halo_eligible(int Cin, int Cout, int kh, int dil) {
  if (halo_disabled()) return false;
  if (!(Cin == 16 || Cin == 32 || Cin == 64) || Cout > 256) return false;
  int pad = dil * (kh - 1) / 2;
  if (pad > 3) return false;
  long long wbytes = (long long)kh * kh * ((Cout * Cin * 2 + 1023) / 1024 * 1024);
  return wbytes <= 100 * 1024;
}
static int launch_conv_halo(const NhwcView& xv, const void* wpk, const float* bias, void* y, const OutView& ov, int N, int H, int W, int Cin, int Cout, int kh,
                            int kw, int dil, int epi, double* stats, cudaStream_t st) {
  ConvHaloParams p{};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kh = kh; p.kw = kw; p.dil = dil; p.pad = dil * (kh - 1) / 2;
  p.tilesH = cdiv(H, HT_H); p.tilesW = cdiv(W, HT_W); p.numTiles = N * p.tilesH * p.tilesW;
  p.rowB = Cin * 2; p.ksteps = Cin / 16;
  p.haloW = HT_W + 2 * p.pad; p.haloH = HT_H + 2 * p.pad;
  p.haloBytes = p.haloW * p.haloH * p.rowB; p.haloStride = (p.haloBytes + 1023) / 1024 * 1024;
  p.wTapStride = (Cout * p.rowB + 1023) / 1024 * 1024;
  size_t wres = (size_t)kh * kw * p.wTapStride;
  static int stage_cap = 0;                          // tuning knob (EGM_HALO_STAGES, read once): ring depth cap; the barrier block holds up to 32 stages
  if (!stage_cap) { const char* e = getenv("EGM_HALO_STAGES"); stage_cap = e ? atoi(e) : 8; if (stage_cap < 2 || stage_cap > 32) stage_cap = 8; }
  p.stages = (int)((198 * 1024 - wres) / p.haloStride); if (p.stages > stage_cap) p.stages = stage_cap; if (p.stages < 2) p.stages = 2;
  p.accCols = (Cout + 31) / 32 * 32;
  p.nacc = 512 / p.accCols; if (p.nacc > 8) p.nacc = 8; if (p.nacc < 2) p.nacc = 2;
  // few tiles per CTA (the 60^2 / 30^2 maps): deep accumulator rings buy nothing, but a CTA that owns all 512 TMEM columns keeps the
  // CTA of a concurrently running branch kernel (engine.Parallel) off its SM -- stay within half of TMEM there
  if (p.numTiles <= 4 * egm_num_sms() && p.nacc * p.accCols > 256 && p.accCols <= 128) p.nacc = 256 / p.accCols;
  p.exp = 0;
#ifdef EGM_DIAG
  { const char* ev = getenv("EGM_EXP"); p.exp = ev ? atoi(ev) : 0; }
#endif
  p.tmemCols = pow2_cols(p.nacc * p.accCols);
  CUtensorMap tmX, tmW;
  p.out = ov;
  int e = make_map_nhwc(&tmX, xv, N, H, W, p.rowB / 2, p.haloW, p.haloH); if (e) return e;
  e = make_map_w(&tmW, wpk, kh * kw, Cout, Cin, p.rowB / 2, Cout); if (e) return e;
  size_t smem = wres + (size_t)p.stages * p.haloStride + 1024 + 1408 + (epi >= 16 ? 2 * 2048 : 0);
  int grid = p.numTiles < egm_num_sms() ? p.numTiles : egm_num_sms();
  p.stats = stats;
  // EGM_EPI_GROUPS=2: two epilogue groups wherever the register file allows (320 threads x <= 204 registers; the 64-channel statistics
  // epilogue keeps 128 running sums per thread and stays at one).  MEASURED NEUTRAL (profiles/step_variants_r2.txt: thin layers +-3 %,
  // step 23.75 vs 23.75 ms) -- the epilogue warps look busy in the stall samples but are not the limiter -- so the default is one group.
  static int ng_max = 0;
  if (!ng_max) { const char* e = getenv("EGM_EPI_GROUPS"); ng_max = e ? atoi(e) : 1; if (ng_max < 1 || ng_max > 2) ng_max = 1; }
  const int ng = (epi == 64 || p.nacc < 2) ? 1 : ng_max;
#define EGM_LAUNCH_HALO_(E, G)                                                                             \
  {                                                                                                        \
    static bool attr_set[64] = {};                                                                         \
    egm_ensure_smem(k_conv_tc_halo<E, G>, 227 * 1024, attr_set);                                           \
    egm_launch(k_conv_tc_halo<E, G>, grid, 64 + 128 * G, smem, st, tmX, tmW, (__nv_bfloat16*)y, bias, p);  \
  }
#define EGM_LAUNCH_HALO(E) { if (ng == 2) EGM_LAUNCH_HALO_(E, 2) else EGM_LAUNCH_HALO_(E, 1) }
  if (epi == EPI_PLAIN) EGM_LAUNCH_HALO(EPI_PLAIN) else if (epi == EPI_RELU) EGM_LAUNCH_HALO(EPI_RELU)
  else if (epi == 16) EGM_LAUNCH_HALO(16) else if (epi == 32) EGM_LAUNCH_HALO(32) else EGM_LAUNCH_HALO_(64, 1)
#undef EGM_LAUNCH_HALO
#undef EGM_LAUNCH_HALO_
  return egm_check_launch("conv2d_tc_halo");
}

extern "C" int egm_conv2d_tc_supported(int Cin, int Cout, int kh, int kw, int dil, int groups) {
  if (groups != 1 || kh != kw || !(kh & 1) || dil < 1) return 0;
  if (Cin % 16 || Cout % 16 || Cout < 16 || Cin < 16 || Cout > 2048 || Cin > 4096) return 0;
  int chunks = (Cout + 255) / 256;
  if (Cout % chunks || (Cout / chunks) % 16) return 0;
  return 1;
}
extern "C" long long egm_conv2d_tc_workspace_bytes(int, int, int, int, int, int, int) { return 0; }

// General form: x and y are channel-strided views (egm_copy_slice semantics).  Cin / Cout are the PADDED channel counts of the
// packed weight [taps][Cout][Cin] (multiples of 16); channels >= cin_valid read as zero (TMA out-of-bounds fill) and channels
// >= cout_valid are not written.  accumulate != 0: y += conv (dgrad into a gradient that already holds other contributions).
// Train-mode BN statistics in the epilogue are available for single-slice layers with 16 / 32 / 64 (padded) output channels --
// the layers whose pre-BN tensors are large; wider layers have small maps and keep the streaming statistics kernel.
extern "C" int egm_conv2d_tc_stats_supported(int Cin, int Cout, int kh, int kw, int dil) {
  return egm_conv2d_tc_supported(Cin, Cout, kh, kw, dil, 1) && (Cout == 16 || Cout == 32 || Cout == 64);
}
// ... and they pay off only where the MMA stream of a tile is long enough to hide the extra epilogue work (2 FP ops per accumulator):
// measured on cfg2 (profiles/conv_epilogue_stats_r2.txt) 3x3 convs with >= 18 MMAs per tile win 13-40 us per layer over a separate
// streaming statistics pass, while 1x1 convs (1-4 MMAs per tile) and the 16-channel layers LOSE 20-40 us -- their epilogue is the
// critical path already.  MMAs per tile = taps * Cin/16; the 64-wide epilogue needs twice the cover of the 32-wide one.
extern "C" int egm_conv2d_tc_stats_profitable(int Cin, int Cout, int kh, int kw, int dil) {
  if (!egm_conv2d_tc_stats_supported(Cin, Cout, kh, kw, dil)) return 0;
  const int mmas = kh * kw * ((Cin + 15) / 16);
  return mmas >= (Cout > 32 ? 36 : 18) ? 1 : 0;
}
// Extended form.  relu != 0: y = relu(conv + bias) (inference with BatchNorm folded into weights / bias).  stats != NULL: in addition to
// writing y, atomically add the per-channel sum and sum of squares of the fp32 results (conv + bias, before rounding) over all N*H*W
// pixels into stats[0 .. cout_valid) and stats[cout_valid .. 2*cout_valid) (doubles, zeroed by the caller) -- the batch statistics of
// the nn.BatchNorm2d that follows the conv (src/EGM-UNet.py:44-55, :958-975), taken in the conv epilogue.
extern "C" int egm_conv2d_tc_ex(const void* x, long long x_cstride, long long x_coff, int cin_valid, const void* w_packed_bf16, const float* bias,
                                void* y, long long y_cstride, long long y_coff, int cout_valid, int accumulate, int N, int H, int W, int Cin,
                                int Cout, int kh, int kw, int dil, int relu, double* stats, void* stream) {
  EGM_REQUIRE(egm_conv2d_tc_supported(Cin, Cout, kh, kw, dil, 1), EGM_E_SHAPE, "conv2d_tc: unsupported shape %d->%d k%d", Cin, Cout, kh);
  EGM_REQUIRE(cin_valid >= 1 && cin_valid <= Cin && cout_valid >= 1 && cout_valid <= Cout, EGM_E_SHAPE, "conv2d_tc: valid channels out of range");
  EGM_REQUIRE(((uintptr_t)w_packed_bf16 & 15) == 0, EGM_E_ALIGN, "conv2d_tc: weights must be 16-byte aligned");
  if ((long long)N * H * W == 0) return EGM_OK;
  const NhwcView xv{x, x_cstride, x_coff, cin_valid};
  OutView ov{y_cstride, y_coff, cout_valid, accumulate ? 1 : 0, (y_cstride % 8 == 0 && y_coff % 8 == 0 && ((uintptr_t)y & 15) == 0) ? 1 : 0, 0};
  ov.dense = (ov.vec && !ov.acc && cout_valid == Cout) ? 1 : 0;
  EGM_REQUIRE(!(relu && stats), EGM_E_BADARG, "conv2d_tc: relu and stats are exclusive (statistics are taken of the pre-BN tensor)");
  EGM_REQUIRE(!stats || egm_conv2d_tc_stats_supported(Cin, Cout, kh, kw, dil), EGM_E_SHAPE, "conv2d_tc: epilogue statistics need Cout in {16,32,64}");
  EGM_REQUIRE(!stats || !accumulate, EGM_E_BADARG, "conv2d_tc: statistics of an accumulating conv are undefined");
  const int epi = stats ? Cout : (relu ? EPI_RELU : EPI_PLAIN);
  if (halo_eligible(Cin, Cout, kh, dil)) return launch_conv_halo(xv, w_packed_bf16, bias, y, ov, N, H, W, Cin, Cout, kh, kw, dil, epi, stats, (cudaStream_t)stream);
  ConvTcParams p{};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kh = kh; p.kw = kw; p.dil = dil; p.pad = dil * (kh - 1) / 2;
  p.coChunks = (Cout + 255) / 256; p.nChunk = Cout / p.coChunks;
  p.tilesH = cdiv(H, TILE_H); p.tilesW = cdiv(W, TILE_W); p.numTiles = N * p.tilesH * p.tilesW * p.coChunks;
  p.bkc = pick_bkc(Cin); p.kChunks = cdiv(Cin, p.bkc);
  p.aBytes = TILE_PIX * p.bkc * 2;                               // 16 KB / 8 KB / 4 KB: multiples of 1024
  p.bStride = (p.nChunk * p.bkc * 2 + 1023) / 1024 * 1024;
  const long long wresBytes = (long long)kh * kw * p.kChunks * p.bStride;
  p.wres = (p.coChunks == 1 && wresBytes <= 72 * 1024) ? 1 : 0;
  p.kps = (p.wres && p.kChunks == 1 && p.bkc <= 32 && kh == 3 && kw == 3) ? 9 : 1;      // thin 3x3 (dilated GRFB branches)
  int per = p.kps * p.aBytes + (p.wres ? 0 : p.bStride);
  p.stages = (int)((198 * 1024 - (p.wres ? wresBytes : 0)) / per); if (p.stages > 8) p.stages = 8; if (p.stages < 2) p.stages = 2;
  p.accCols = (p.nChunk + 31) / 32 * 32; p.tmemCols = pow2_cols(2 * p.accCols);
  CUtensorMap tmX, tmW;
  p.out = ov;
  int e = make_map_nhwc(&tmX, xv, N, H, W, p.bkc, TILE_W, TILE_H); if (e) return e;
  e = make_map_w(&tmW, w_packed_bf16, kh * kw, Cout, Cin, p.bkc, p.nChunk); if (e) return e;
  size_t smem = (size_t)p.stages * per + (p.wres ? (size_t)wresBytes : 0) + 1024 + 1024 + (epi >= 16 ? 2 * 2048 : 0);
  int grid = p.numTiles < egm_num_sms() ? p.numTiles : egm_num_sms();
  p.stats = stats;
  static int ng_max = 0;                              // epilogue groups (see launch_conv_halo): EGM_EPI_GROUPS=2, measured neutral, default 1
  if (!ng_max) { const char* e = getenv("EGM_EPI_GROUPS"); ng_max = e ? atoi(e) : 1; if (ng_max < 1 || ng_max > 2) ng_max = 1; }
  const int ng = epi == 64 ? 1 : ng_max;
#define EGM_LAUNCH_TC_(E, G)                                                                                      \
  {                                                                                                               \
    static bool attr_set[64] = {};                                                                                \
    egm_ensure_smem(k_conv_tc<E, G>, 227 * 1024, attr_set);                                                       \
    egm_launch(k_conv_tc<E, G>, grid, 64 + 128 * G, smem, (cudaStream_t)stream, tmX, tmW, (__nv_bfloat16*)y, bias, p);  \
  }
#define EGM_LAUNCH_TC(E) { if (ng == 2) EGM_LAUNCH_TC_(E, 2) else EGM_LAUNCH_TC_(E, 1) }
  if (epi == EPI_PLAIN) EGM_LAUNCH_TC(EPI_PLAIN) else if (epi == EPI_RELU) EGM_LAUNCH_TC(EPI_RELU)
  else if (epi == 16) EGM_LAUNCH_TC(16) else if (epi == 32) EGM_LAUNCH_TC(32) else EGM_LAUNCH_TC_(64, 1)
#undef EGM_LAUNCH_TC
#undef EGM_LAUNCH_TC_
  EGM_LAUNCH_CHECK("conv2d_tc"); return EGM_OK;
}
extern "C" int egm_conv2d_tc_view(const void* x, long long x_cstride, long long x_coff, int cin_valid, const void* w_packed_bf16, const float* bias,
                                  void* y, long long y_cstride, long long y_coff, int cout_valid, int accumulate, int N, int H, int W, int Cin,
                                  int Cout, int kh, int kw, int dil, void* stream) {
  return egm_conv2d_tc_ex(x, x_cstride, x_coff, cin_valid, w_packed_bf16, bias, y, y_cstride, y_coff, cout_valid, accumulate, N, H, W, Cin, Cout, kh, kw, dil,
                          0, nullptr, stream);
}
extern "C" int egm_conv2d_tc(const void* x, const void* w_packed_bf16, const float* bias, void* y, int N, int H, int W, int Cin, int Cout,
                             int kh, int kw, int dil, void* stream) {
  return egm_conv2d_tc_view(x, Cin, 0, Cin, w_packed_bf16, bias, y, Cout, 0, Cout, 0, N, H, W, Cin, Cout, kh, kw, dil, stream);
}

// weights fp32 [Cout][Cin][kh][kw] -> bf16 wf [taps][Cout][Cin] (forward) and wd [taps_flipped][Cin][Cout] (dgrad: roles swapped)
__global__ void k_pack_w_tc(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd, int Cout, int Cin, int taps) { egm_pdl_enter();
  long long total = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i % taps); long long q = i / taps; int ci = (int)(q % Cin); int co = (int)(q / Cin);
    __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
    if (wf) wf[((long long)t * Cout + co) * Cin + ci] = v;
    if (wd) wd[((long long)(taps - 1 - t) * Cin + ci) * Cout + co] = v;
  }
}
extern "C" int egm_pack_conv_weight_tc(const float* w, void* wf_bf16, void* wd_bf16, int Cout, int Cin, int kh, int kw, void* stream) {
  long long total = (long long)Cout * Cin * kh * kw;
  if (total == 0) return EGM_OK;
  egm_launch(k_pack_w_tc, egm_grid_for(total, 256), 256, 0, (cudaStream_t)stream, w, (__nv_bfloat16*)wf_bf16, (__nv_bfloat16*)wd_bf16, Cout, Cin, kh * kw);
  EGM_LAUNCH_CHECK("pack_conv_weight_tc"); return EGM_OK;
}

// =================================================================== wgrad
// work unit = (tap group of <= 3 taps, Cout chunk mch, Cin chunk nch) x pixel split; stage = dY tile + the group's shifted X tiles.
struct WgradParams {
  int N, H, W, Cin, Cout, kh, kw, dil, pad;
  int tilesH, tilesW, numTiles;
  int tapGroups, coChunks, ciChunks, splits, tilesPerSplit;
  int mch, nch, mAtoms, aAtomBytes, bTileBytes, aBytes, stageBytes, stages, tmemCols, ummaM;
};
constexpr int WG_TAPS = 3;

__global__ void __launch_bounds__(TC_THREADS, 1) k_wgrad_tc(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                                                           float* __restrict__ dwp, WgradParams p) { egm_pdl_enter();
  if (threadIdx.x == 0) { tmap_prefetch(&tmDY); tmap_prefetch(&tmX); }
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + (size_t)p.stages * p.stageBytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint32_t* tmem_slot = (uint32_t*)(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // decode the work unit
  int u = blockIdx.x;
  const int sp = u % p.splits; u /= p.splits;
  const int cic = u % p.ciChunks; u /= p.ciChunks;
  const int coc = u % p.coChunks; u /= p.coChunks;
  const int tg = u;
  const int taps = p.kh * p.kw;
  const int t0 = tg * WG_TAPS;
  const int nt = (taps - t0) < WG_TAPS ? (taps - t0) : WG_TAPS;
  const int tileBeg = sp * p.tilesPerSplit;
  int tileEnd = tileBeg + p.tilesPerSplit; if (tileEnd > p.numTiles) tileEnd = p.numTiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    const bool leader = elect_one();
    const int tilesPerImg = p.tilesH * p.tilesW;
    int s = 0; uint32_t ph = 0;
    for (int tile = tileBeg; tile < tileEnd; ++tile) {
      const int n = tile / tilesPerImg, r = tile - n * tilesPerImg;
      const int h0 = (r / p.tilesW) * TILE_H, w0 = (r % p.tilesW) * TILE_W;
      uint8_t* st = smem + (size_t)s * p.stageBytes;
      mbar_wait(&empty[s], ph ^ 1);
      if (leader) {
        mbar_expect_tx(&full[s], (uint32_t)(p.mAtoms * p.aAtomBytes + nt * p.bTileBytes));
        for (int a = 0; a < p.mAtoms; ++a)
          tma_load_4d(st + (size_t)a * p.aAtomBytes, &tmDY, &full[s], coc * p.mch + a * 64, w0, h0, n);
        for (int j = 0; j < nt; ++j) {
          const int t = t0 + j, dh = (t / p.kw) * p.dil - p.pad, dw = (t % p.kw) * p.dil - p.pad;
          tma_load_4d(st + p.aBytes + (size_t)j * p.bTileBytes, &tmX, &full[s], cic * p.nch, w0 + dw, h0 + dh, n);
        }
      }
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc(p.ummaM, p.nch, 1, 1);
    const int rowA = (p.mch >= 64 ? 64 : p.mch) * 2, rowB = p.nch * 2;
    const uint32_t layA = rowA == 128 ? 2u : (rowA == 64 ? 4u : 6u), layB = rowB == 128 ? 2u : (rowB == 64 ? 4u : 6u);
    const uint32_t kA = (uint32_t)(16 * rowA) >> 4, kB = (uint32_t)(16 * rowB) >> 4, jB = (uint32_t)p.bTileBytes >> 4;
    int s = 0; uint32_t ph = 0; uint32_t accf = 0;
    for (int tile = tileBeg; tile < tileEnd; ++tile) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (leader) {
        const uint32_t a0 = smem_u32(smem + (size_t)s * p.stageBytes);
        const uint64_t ad0 = umma_desc(a0, (uint32_t)p.aAtomBytes, 8u * rowA, layA);
        uint64_t bdj = umma_desc(a0 + p.aBytes, (uint32_t)p.bTileBytes, 8u * rowB, layB);
        uint32_t dcol = tmem_base;
        if (nt == WG_TAPS) {                               // the common case as straight-line code (no branch between MMAs)
#pragma unroll
          for (int j = 0; j < WG_TAPS; ++j) {
#pragma unroll
            for (int k = 0; k < TILE_PIX / 16; ++k)
              umma_bf16(tmem_base + (uint32_t)(j * p.nch), ad0 + k * kA, bdj + (uint64_t)j * jB + k * kB, idesc, k ? 1u : accf);   // 16 pixels per MMA
          }
        } else
        for (int j = 0; j < nt; ++j) {
          uint32_t af = accf;
#pragma unroll
          for (int k = 0; k < TILE_PIX / 16; ++k) { umma_bf16(dcol, ad0 + k * kA, bdj + k * kB, idesc, af); af = 1; }   // 16 pixels per MMA
          bdj += jB; dcol += (uint32_t)p.nch;
        }
        umma_commit(&empty[s]);
      }
      accf = 1;
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
    if (leader) umma_commit(tfull);
    __syncwarp();
  } else if (tileBeg < tileEnd) {
    const int q = warp & 3;
    const int co_local = q * 32 + lane;                  // TMEM lane = output channel inside the chunk
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int co = coc * p.mch + co_local;
    const bool valid = co_local < p.mch && co < p.Cout;
    if (q * 32 < p.ummaM) {
      for (int j = 0; j < nt; ++j) {
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * p.nch);
        for (int c = 0; c < p.nch; c += 16) {
          uint32_t v[16];
          tmem_ld16(ta + c, v);
          tmem_ld_wait();
          if (valid && cic * p.nch + c < p.Cin) red_add16(dwp + ((long long)(t0 + j) * p.Cout + co) * p.Cin + cic * p.nch + c, v);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)p.tmemCols); }
}


// =================================================================== wgrad, halo-reuse variant
// Work unit = (Cout chunk, Cin chunk nch, group of R kernel rows) x pixel split.  Per 16x8 pixel tile ONE dY tile and ONE X
// halo tile are fetched; one MMA per (16-pixel K-step, kernel row) covers the kw taps of that row at once: the B operand is
// MN-major with N = kw*nch, whose N-atoms (one per tap) are the halo tile shifted by `dil` rows (LBO = dil*rowB) -- legal
// because the swizzle is a function of the absolute smem address.  8*R MMAs and 2-3 TMA loads per tile instead of 72 / 12.
// Row stacking (Cout chunk * kh <= 128, i.e. the 16/32-channel layers): an M=128 MMA costs the same whether 32 or 128 of its rows
// are useful (its time is the (M+N) operand fetch), so the kh kernel rows share ONE MMA per K-step: with q = p + r*dil,
//   dW[r][s] = sum_p dY[p] x[p + (r*dil - pad, s*dil - pad)] = sum_q dY[q - r*dil rows] x[q + (-pad, s*dil - pad)],
// i.e. the B operand is the same X window for every r and the A operand gets one atom per r holding the dY tile fetched r*dil rows
// higher (TMA zero-fills rows outside the image; the tile range is extended by (kh-1)*dil rows so every dY row meets every r).
struct WgradHaloParams {
  int N, H, W, Cin, Cout, kh, kw, dil, pad;
  int tilesH, tilesW, numTiles;
  int coChunks, ciChunks, rowGroups, rowsPerGroup, splits, tilesPerSplit;
  int mch, nch, mAtoms, aAtomCh, aAtomBytes, aBytes, rowA, rowB, haloW, haloH, haloBytes, haloStride, stageBytes, stages, tmemCols;
  int stack;   // 1: kernel rows are stacked along M (Cout chunk <= 64): atom j of the A operand = dY shifted up by j*dil rows
  int rpm, nGroups;   // stacked mode: kernel rows per MMA (128 / mch) and MMAs per K-step (ceil(kh / rpm)), one accumulator each
};

__global__ void __launch_bounds__(TC_THREADS, 1) k_wgrad_tc_halo(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                                                                float* __restrict__ dwp, WgradHaloParams p) { egm_pdl_enter();
  if (threadIdx.x == 0) { tmap_prefetch(&tmDY); tmap_prefetch(&tmX); }
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + (size_t)p.stages * p.stageBytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint32_t* tmem_slot = (uint32_t*)(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int u = blockIdx.x;
  const int sp = u % p.splits; u /= p.splits;
  const int cic = u % p.ciChunks; u /= p.ciChunks;
  const int coc = u % p.coChunks; u /= p.coChunks;
  const int r0 = u * p.rowsPerGroup;                                   // first kernel row of this unit
  const int R = (p.kh - r0) < p.rowsPerGroup ? (p.kh - r0) : p.rowsPerGroup;
  const int nN = p.kw * p.nch;                                         // UMMA N: all taps of one kernel row
  const int tileBeg = sp * p.tilesPerSplit;
  int tileEnd = tileBeg + p.tilesPerSplit; if (tileEnd > p.numTiles) tileEnd = p.numTiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    const bool leader = elect_one();
    const int tilesPerImg = p.tilesH * p.tilesW;
    int s = 0; uint32_t ph = 0;
    for (int tile = tileBeg; tile < tileEnd; ++tile) {
      const int n = tile / tilesPerImg, r = tile - n * tilesPerImg;
      const int h0 = (r / p.tilesW) * HT_H, w0 = (r % p.tilesW) * HT_W;
      uint8_t* st = smem + (size_t)s * p.stageBytes;
      mbar_wait(&empty[s], ph ^ 1);
      if (leader) {
        if (p.stack) {
          mbar_expect_tx(&full[s], (uint32_t)(p.kh * p.aAtomBytes + p.haloBytes));
          for (int j = 0; j < p.kh; ++j) tma_load_4d(st + (size_t)j * p.aAtomBytes, &tmDY, &full[s], coc * p.mch, w0, h0 - j * p.dil, n);
        } else {
          mbar_expect_tx(&full[s], (uint32_t)(p.mAtoms * p.aAtomBytes + p.haloBytes));
          for (int a = 0; a < p.mAtoms; ++a)
            tma_load_4d(st + (size_t)a * p.aAtomBytes, &tmDY, &full[s], coc * p.mch + a * p.aAtomCh, w0, h0, n);
        }
        tma_load_4d(st + p.aBytes, &tmX, &full[s], cic * p.nch, w0 - p.pad, h0 - p.pad, n);
      }
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc(128, nN, 1, 1);
    const uint32_t layA = p.rowA == 128 ? 2u : (p.rowA == 64 ? 4u : 6u), layB = p.rowB == 128 ? 2u : (p.rowB == 64 ? 4u : 6u);
    const uint32_t kA = (uint32_t)(16 * p.rowA) >> 4;                 // A advance per 16-pixel K step (16-byte units)
    const uint32_t kB = (uint32_t)(2 * p.haloW * p.rowB) >> 4;        // B advance per K step: two halo rows
    const uint32_t rB = (uint32_t)(p.dil * p.haloW * p.rowB) >> 4;    // B advance per kernel row
    const uint32_t gA = (uint32_t)(p.rpm * p.aAtomBytes) >> 4;         // stacked mode: A advance per group of rpm kernel rows
    int s = 0; uint32_t ph = 0; uint32_t accf = 0;
    for (int tile = tileBeg; tile < tileEnd; ++tile) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (leader) {
        const uint32_t a0 = smem_u32(smem + (size_t)s * p.stageBytes);
        const uint64_t ad0 = umma_desc(a0, (uint32_t)p.aAtomBytes, 8u * p.rowA, layA);
        const uint64_t bd0 = umma_desc(a0 + p.aBytes, (uint32_t)(p.dil * p.rowB), (uint32_t)(p.haloW * p.rowB), layB) + (uint64_t)(r0 * rB);
        // Every branch inside this single-thread issue loop costs about as much as an MMA (measured: +30..45 % on 32->32@480^2),
        // so the mode is decided once per tile and the stacked cases are straight-line code.
        if (p.stack && p.nGroups == 1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) { umma_bf16(tmem_base, ad0 + k * kA, bd0 + k * kB, idesc, accf); accf = 1; }
        } else if (p.stack && p.nGroups == 2) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            umma_bf16(tmem_base, ad0 + k * kA, bd0 + k * kB, idesc, accf);
            umma_bf16(tmem_base + (uint32_t)nN, ad0 + k * kA + gA, bd0 + k * kB, idesc, accf);
            accf = 1;
          }
        } else if (p.stack) {
#pragma unroll 1
          for (int k = 0; k < 8; ++k) {
            for (int g = 0; g < p.nGroups; ++g) umma_bf16(tmem_base + (uint32_t)(g * nN), ad0 + k * kA + (uint64_t)g * gA, bd0 + k * kB, idesc, accf);
            accf = 1;
          }
        } else if (R == 1) {                                             // 1x1 kernels (or one kernel row per unit)
#pragma unroll
          for (int k = 0; k < 8; ++k) { umma_bf16(tmem_base, ad0 + k * kA, bd0 + k * kB, idesc, accf); accf = 1; }
        } else if (R == 3) {                                             // 3x3 kernels, wide Cout: 24 MMAs, straight-line
#pragma unroll
          for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int r = 0; r < 3; ++r) umma_bf16(tmem_base + (uint32_t)(r * nN), ad0 + k * kA, bd0 + k * kB + (uint64_t)r * rB, idesc, accf);
            accf = 1;
          }
        } else {
#pragma unroll 1
          for (int k = 0; k < 8; ++k) {
            const uint64_t ad = ad0 + k * kA;
            uint64_t bd = bd0 + k * kB;
            uint32_t dcol = tmem_base;
            for (int r = 0; r < R; ++r) { umma_bf16(dcol, ad, bd, idesc, accf); bd += rB; dcol += (uint32_t)nN; }
            accf = 1;
          }
        }
        umma_commit(&empty[s]);
      }
      accf = 1;
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
    if (leader) umma_commit(tfull);
    __syncwarp();
  } else if (tileBeg < tileEnd) {
    const int q = warp & 3;
    const int m = q * 32 + lane;                          // accumulator row (TMEM lane)
    mbar_wait(tfull, 0);
    tc_fence_after();
    if (p.stack) {
      // accumulator g, row m = jl*mch + co: kernel row j = g*rpm + jl, output channel co; columns = (s, ci)
      const int jl = m / p.mch, co = coc * p.mch + (m - jl * p.mch);
      if (q * 32 < p.rpm * p.mch) {
        for (int g = 0; g < p.nGroups; ++g) {
          const int j = g * p.rpm + jl;
          const bool valid = jl < p.rpm && j < p.kh && co < p.Cout;
          for (int sidx = 0; sidx < p.kw; ++sidx) {
            const int tap = j * p.kw + sidx;
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * nN + sidx * p.nch);
            for (int c = 0; c < p.nch; c += 16) {
              uint32_t v[16];
              tmem_ld16(ta + c, v);
              tmem_ld_wait();
              if (valid && cic * p.nch + c < p.Cin) red_add16(dwp + ((long long)tap * p.Cout + co) * p.Cin + cic * p.nch + c, v);
            }
          }
        }
      }
    } else {
      const int co_local = m;
      const int co = coc * p.mch + co_local;
      const bool valid = co_local < p.mch && co < p.Cout;
      if (q * 32 < p.mch) {
        for (int r = 0; r < R; ++r)
          for (int sidx = 0; sidx < p.kw; ++sidx) {
            const int tap = (r0 + r) * p.kw + sidx;
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(r * nN + sidx * p.nch);
            for (int c = 0; c < p.nch; c += 16) {
              uint32_t v[16];
              tmem_ld16(ta + c, v);
              tmem_ld_wait();
              if (valid && cic * p.nch + c < p.Cin) red_add16(dwp + ((long long)tap * p.Cout + co) * p.Cin + cic * p.nch + c, v);
            }
          }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)p.tmemCols); }
}

static bool wgrad_stack_disabled() { static int v = -1; if (v < 0) { const char* e = getenv("EGM_NO_WGRAD_STACK"); v = (e && e[0] == '1') ? 1 : 0; } return v == 1; }
// row stacking applies: Cout chunk <= 64 (one A atom per kernel row), >= 2 kernel rows per MMA, accumulators fit in TMEM
static bool wgrad_stackable(int Cin, int Cout, int kh, int kw) {
  if (wgrad_stack_disabled() || kh < 2) return false;
  const int mch = Cout > 64 ? 128 : pick_bkc(Cout), nch = pick_bkc(Cin);
  if (mch > 64) return false;
  int rpm = 128 / mch; if (rpm > kh) rpm = kh;
  return rpm >= 2 && cdiv(kh, rpm) * kw * nch <= 512;
}
static bool wgrad_halo_eligible(int Cin, int Cout, int kh, int kw, int dil) {
  if (halo_disabled()) return false;
  const int pad = dil * (kh - 1) / 2, nch = pick_bkc(Cin);
  if (kw * nch > 256) return false;
  if (pad <= 3) return true;
  // dilated: in stacked mode the X tile needs no vertical halo (16 rows x (8 + 2*pad) columns), so wide dilations fit
  if (!wgrad_stackable(Cin, Cout, kh, kw)) return false;
  const int haloW = HT_W + 2 * pad;
  if (haloW > 256) return false;
  // two pipeline stages (A atoms of every MMA group + the X tile with its tap slack) must fit in shared memory
  const int mch = Cout > 64 ? 128 : pick_bkc(Cout), aAtomCh = mch >= 64 ? 64 : mch;
  int rpm = 128 / mch; if (rpm > kh) rpm = kh;
  const long long aBytes = (long long)cdiv(kh, rpm) * (128 / aAtomCh) * TILE_PIX * aAtomCh * 2;
  const long long halo = ((long long)haloW * HT_H * nch * 2 + (long long)(kw - 1) * dil * nch * 2 + 1023) / 1024 * 1024;
  return 2 * (aBytes + halo) + 2048 <= 220 * 1024;
}
static int launch_wgrad_halo(const NhwcView& xv, const NhwcView& dyv, float* dwp, int N, int H, int W, int Cin, int Cout, int kh, int kw, int dil, cudaStream_t st) {
  WgradHaloParams p{};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kh = kh; p.kw = kw; p.dil = dil; p.pad = dil * (kh - 1) / 2;
  p.tilesH = cdiv(H, HT_H); p.tilesW = cdiv(W, HT_W); p.numTiles = N * p.tilesH * p.tilesW;
  p.mch = Cout > 64 ? 128 : pick_bkc(Cout);
  p.aAtomCh = p.mch >= 64 ? 64 : p.mch; p.mAtoms = p.mch / p.aAtomCh;
  p.rowA = p.aAtomCh * 2; p.aAtomBytes = TILE_PIX * p.rowA;
  p.aBytes = (128 / p.aAtomCh) * p.aAtomBytes;                         // the M=128 MMA addresses 128/aAtomCh atoms (only mAtoms are loaded)
  p.nch = pick_bkc(Cin); p.rowB = p.nch * 2;
  p.haloW = HT_W + 2 * p.pad; p.haloH = HT_H + 2 * p.pad;
  p.haloBytes = p.haloW * p.haloH * p.rowB; p.haloStride = (p.haloBytes + (kw - 1) * dil * p.rowB + 1023) / 1024 * 1024;   // + slack read by the last taps
  p.stageBytes = p.aBytes + p.haloStride;
  p.stages = (200 * 1024) / p.stageBytes; if (p.stages > 6) p.stages = 6; if (p.stages < 2) p.stages = 2;
  int perRow = kw * p.nch;
  p.rowsPerGroup = 512 / perRow; if (p.rowsPerGroup > kh) p.rowsPerGroup = kh;
  p.rowGroups = cdiv(kh, p.rowsPerGroup);
  p.tmemCols = pow2_cols(p.rowsPerGroup * perRow);
  p.rpm = 128 / p.mch; if (p.rpm > kh) p.rpm = kh;
  p.nGroups = cdiv(kh, p.rpm);
  p.stack = wgrad_stackable(Cin, Cout, kh, kw) ? 1 : 0;
  EGM_REQUIRE(p.stack || p.pad <= 3, EGM_E_SHAPE, "wgrad_tc_halo: dilated layers need the stacked mode");
  if (p.stack) {                                         // nGroups accumulators; q rows cover [0, H + (kh-1)*dil)
    p.rowsPerGroup = kh; p.rowGroups = 1; p.tmemCols = pow2_cols(p.nGroups * perRow);
    p.haloH = HT_H;                                      // the X window is not row-shifted in this mode: no vertical halo
    p.haloBytes = p.haloW * p.haloH * p.rowB; p.haloStride = (p.haloBytes + (kw - 1) * dil * p.rowB + 1023) / 1024 * 1024;
    p.tilesH = cdiv(H + (kh - 1) * dil, HT_H); p.numTiles = N * p.tilesH * p.tilesW;
    const int atoms = p.nGroups * (128 / p.aAtomCh);    // every MMA addresses 128/aAtomCh atoms from its first one
    p.aBytes = atoms * p.aAtomBytes;
    p.stageBytes = p.aBytes + p.haloStride;
    p.stages = (200 * 1024) / p.stageBytes; if (p.stages > 6) p.stages = 6; if (p.stages < 2) p.stages = 2;
  }
  p.coChunks = cdiv(Cout, p.mch); p.ciChunks = cdiv(Cin, p.nch);
  long long units = (long long)p.rowGroups * p.coChunks * p.ciChunks;
  long long want = ((long long)egm_num_sms() * 2 + units - 1) / units;
  if (want > p.numTiles) want = p.numTiles; if (want < 1) want = 1;
  p.tilesPerSplit = cdiv(p.numTiles, want); p.splits = cdiv(p.numTiles, p.tilesPerSplit);
  CUtensorMap tmDY, tmX;
  int e = make_map_nhwc(&tmDY, dyv, N, H, W, p.aAtomCh, HT_W, HT_H); if (e) return e;
  e = make_map_nhwc(&tmX, xv, N, H, W, p.nch, p.haloW, p.haloH); if (e) return e;
  size_t smem = (size_t)p.stages * p.stageBytes + 1024 + 256;
  static bool attr_set[64] = {};
  egm_ensure_smem(k_wgrad_tc_halo, 227 * 1024, attr_set);
  long long grid = units * p.splits;
  EGM_REQUIRE(grid < (1ll << 31), EGM_E_SHAPE, "wgrad_tc: grid too large");
  egm_launch(k_wgrad_tc_halo, (unsigned)grid, TC_THREADS, smem, st, tmDY, tmX, dwp, p);
  return egm_check_launch("conv2d_wgrad_tc_halo");
}

// dw_packed fp32 [taps][Cout][Cin] (NOT the [taps][Cin][Cout] of the CUDA-core path: see red_add16; unpack with egm_unpack_conv_wgrad_tc
// or egm_wgrad_unpack_batch; zeroed here)
// General form: x and dy are channel-strided views; Cin / Cout are the padded channel counts of dw_packed.
extern "C" int egm_conv2d_wgrad_tc_view(const void* x, long long x_cstride, long long x_coff, int cin_valid, const void* dy, long long dy_cstride,
                                        long long dy_coff, int cout_valid, float* dw_packed, int N, int H, int W, int Cin, int Cout, int kh, int kw,
                                        int dil, void* stream) {
  EGM_REQUIRE(egm_conv2d_tc_supported(Cin, Cout, kh, kw, dil, 1), EGM_E_SHAPE, "wgrad_tc: unsupported shape %d->%d", Cin, Cout);
  EGM_REQUIRE(cin_valid >= 1 && cin_valid <= Cin && cout_valid >= 1 && cout_valid <= Cout, EGM_E_SHAPE, "wgrad_tc: valid channels out of range");
  const NhwcView xv{x, x_cstride, x_coff, cin_valid}, dyv{dy, dy_cstride, dy_coff, cout_valid};
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(dw_packed, 0, sizeof(float) * (size_t)kh * kw * Cin * Cout, st);
  if ((long long)N * H * W == 0) return EGM_OK;
  if (wgrad_halo_eligible(Cin, Cout, kh, kw, dil)) return launch_wgrad_halo(xv, dyv, dw_packed, N, H, W, Cin, Cout, kh, kw, dil, st);
  WgradParams p{};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kh = kh; p.kw = kw; p.dil = dil; p.pad = dil * (kh - 1) / 2;
  p.tilesH = cdiv(H, TILE_H); p.tilesW = cdiv(W, TILE_W); p.numTiles = N * p.tilesH * p.tilesW;
  p.mch = Cout > 64 ? 128 : pick_bkc(Cout);         // 128 / 64 / 32 / 16 output channels per work unit (tail zero-filled)
  p.ummaM = 128;                                     // always M=128 (lane i == row i); rows >= mch read unused smem and are ignored
  p.mAtoms = p.mch >= 64 ? p.mch / 64 : 1;
  int aAtomCh = p.mch >= 64 ? 64 : p.mch;
  p.aAtomBytes = TILE_PIX * aAtomCh * 2;
  p.aBytes = (p.ummaM / aAtomCh) * p.aAtomBytes;     // room for the atoms the MMA addresses (>= loaded atoms)
  if (p.aBytes < p.mAtoms * p.aAtomBytes) p.aBytes = p.mAtoms * p.aAtomBytes;
  p.nch = pick_bkc(Cin);                            // 64 / 32 / 16 input channels per work unit
  p.bTileBytes = TILE_PIX * p.nch * 2;
  p.stageBytes = p.aBytes + WG_TAPS * p.bTileBytes;
  p.stages = (200 * 1024) / p.stageBytes; if (p.stages > 6) p.stages = 6; if (p.stages < 2) p.stages = 2;
  p.tmemCols = pow2_cols(WG_TAPS * p.nch);
  p.tapGroups = cdiv(kh * kw, WG_TAPS); p.coChunks = cdiv(Cout, p.mch); p.ciChunks = cdiv(Cin, p.nch);
  long long units = (long long)p.tapGroups * p.coChunks * p.ciChunks;
  long long want = ((long long)egm_num_sms() * 2 + units - 1) / units;
  if (want > p.numTiles) want = p.numTiles; if (want < 1) want = 1;
  p.tilesPerSplit = cdiv(p.numTiles, want); p.splits = cdiv(p.numTiles, p.tilesPerSplit);
  CUtensorMap tmDY, tmX;
  int e = make_map_nhwc(&tmDY, dyv, N, H, W, aAtomCh, TILE_W, TILE_H); if (e) return e;
  e = make_map_nhwc(&tmX, xv, N, H, W, p.nch, TILE_W, TILE_H); if (e) return e;
  size_t smem = (size_t)p.stages * p.stageBytes + 1024 + 256;
  static bool attr_set[64] = {};
  egm_ensure_smem(k_wgrad_tc, 227 * 1024, attr_set);
  long long grid = units * p.splits;
  EGM_REQUIRE(grid < (1ll << 31), EGM_E_SHAPE, "wgrad_tc: grid too large");
  egm_launch(k_wgrad_tc, (unsigned)grid, TC_THREADS, smem, st, tmDY, tmX, dw_packed, p);
  EGM_LAUNCH_CHECK("conv2d_wgrad_tc"); return EGM_OK;
}
// dw_packed [taps][Cout][Cin] (tcgen05 wgrad layout) -> dw [Cout][Cin][kh][kw]   (dw = beta*dw + unpacked)
__global__ void k_unpack_dw_tc(const float* __restrict__ dwp, float* __restrict__ dw, int Cout, int Cin, int taps, float beta) { egm_pdl_enter();
  long long total = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i % taps); long long q = i / taps;        // q = co*Cin + ci
    float v = dwp[(long long)t * Cout * Cin + q];
    dw[i] = beta != 0.f ? beta * dw[i] + v : v;
  }
}
extern "C" int egm_unpack_conv_wgrad_tc(const float* dw_packed, float* dw, int Cout, int Cin, int kh, int kw, float beta, void* stream) {
  long long total = (long long)Cout * Cin * kh * kw;
  if (total == 0) return EGM_OK;
  egm_launch(k_unpack_dw_tc, egm_grid_for(total, 256), 256, 0, (cudaStream_t)stream, dw_packed, dw, Cout, Cin, kh * kw, beta);
  EGM_LAUNCH_CHECK("unpack_conv_wgrad_tc"); return EGM_OK;
}
extern "C" int egm_conv2d_wgrad_tc(const void* x, const void* dy, float* dw_packed, int N, int H, int W, int Cin, int Cout, int kh, int kw, int dil,
                                   void* stream) {
  return egm_conv2d_wgrad_tc_view(x, Cin, 0, Cin, dy, Cout, 0, Cout, dw_packed, N, H, W, Cin, Cout, kh, kw, dil, stream);
}
