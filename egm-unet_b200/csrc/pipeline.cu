// Input pipeline between the decoded uint8 image and the model input, on the device.
//
// Reference (PIL + torchvision on DataLoader workers): transforms.py:30-110 (RandomResize -> flips -> RandomCrop with
// pad_if_smaller -> ToTensor -> Normalize), train.py:17-54 (presets), my_dataset.py:103-133 (mask / 255, collate_fn padding).
// ONE kernel per image produces the normalised float32 NCHW crop and the int64 target straight into the batch tensors:
//   out(y, x) <- crop offset -> pad_if_smaller region (raw 0 / target 0) -> flips -> Pillow's two-pass 8-bit bilinear resample
//   (horizontal pass then vertical pass, 22-bit fixed-point coefficients, uint8 rounding between the passes) for the image,
//   Pillow's nearest-neighbour index table for the mask -> /255, (x - mean) / std;  collate padding (0.0 / 255) outside.
// The coefficient / index tables are computed on the host in double precision exactly as Pillow does (egm-unet_b200/data.py);
// the device does integer arithmetic only, so results are bit-identical to the reference pipeline.
#include "common.cuh"

constexpr int PIPE_PRECISION_BITS = 32 - 8 - 2;

struct PipeArgs {
  const unsigned char* img; const unsigned char* mask;      // [H][W][3], [H][W]
  int H, W, rh, rw;                                          // source and resized sizes
  const int* hmin; const int* hcnt; const int* hk; int hks;  // horizontal pass: first tap, tap count, coefficients [rw][hks]   (null: rw == W)
  const int* vmin; const int* vcnt; const int* vk; int vks;  // vertical pass                                              (null: rh == H)
  const int* nnx; const int* nny;                            // nearest-neighbour source index per resized column / row      (null: identity)
  int hflip, vflip, top, left;                               // flips of the resized image; crop origin in the padded resized image
  int vh, vw;                                                // valid output region (crop size, or the resized size in eval mode)
  int OH, OW;                                                // full output plane (>= valid: collate padding)
  float mean[3], stdv[3];
  float* out_img; long long* out_tgt;                        // [3][OH][OW], [OH][OW] of this batch entry
};

__device__ __forceinline__ int clip8(int v) { v >>= PIPE_PRECISION_BITS; return v < 0 ? 0 : (v > 255 ? 255 : v); }

__device__ __forceinline__ int hpass(const PipeArgs& a, int row, int fx, int c) {     // horizontally resampled value at (source row, resized col)
  if (!a.hmin) return a.img[((size_t)row * a.W + fx) * 3 + c];
  int acc = 1 << (PIPE_PRECISION_BITS - 1);
  const int x0 = a.hmin[fx], n = a.hcnt[fx];
  const int* k = a.hk + (size_t)fx * a.hks;
  const unsigned char* p = a.img + ((size_t)row * a.W + x0) * 3 + c;
  for (int i = 0; i < n; ++i) acc += (int)p[i * 3] * k[i];
  return clip8(acc);
}

__global__ void __launch_bounds__(256) k_input_transform(PipeArgs a) { egm_pdl_enter();
  const long long total = (long long)a.OH * a.OW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int oy = (int)(i / a.OW), ox = (int)(i - (long long)oy * a.OW);
    float v[3]; long long t;
    if (oy >= a.vh || ox >= a.vw) {                           // collate_fn padding: image 0.0 (after normalisation), target 255
      v[0] = v[1] = v[2] = 0.f; t = 255;
    } else {
      const int py = a.top + oy, px = a.left + ox;
      int raw[3] = {0, 0, 0}; t = 0;                           // pad_if_smaller: raw 0 before ToTensor / Normalize, target 0
      if (py < a.rh && px < a.rw) {
        const int fy = a.vflip ? a.rh - 1 - py : py, fx = a.hflip ? a.rw - 1 - px : px;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (!a.vmin) raw[c] = hpass(a, fy, fx, c);
          else {
            int acc = 1 << (PIPE_PRECISION_BITS - 1);
            const int y0 = a.vmin[fy], n = a.vcnt[fy];
            const int* k = a.vk + (size_t)fy * a.vks;
            for (int j = 0; j < n; ++j) acc += hpass(a, y0 + j, fx, c) * k[j];
            raw[c] = clip8(acc);
          }
        }
        const int sy = a.nny ? a.nny[fy] : fy, sx = a.nnx ? a.nnx[fx] : fx;
        t = a.mask[(size_t)sy * a.W + sx] == 255 ? 1 : 0;    // int64(float32(v / 255)): only 255 maps to 1
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)raw[c], 255.f), a.mean[c]), a.stdv[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) a.out_img[(size_t)c * total + i] = v[c];
    a.out_tgt[i] = t;
  }
}

extern "C" int egm_input_transform(const unsigned char* img, const unsigned char* mask, int H, int W, int rh, int rw, const int* hmin, const int* hcnt,
                                   const int* hk, int hks, const int* vmin, const int* vcnt, const int* vk, int vks, const int* nnx, const int* nny,
                                   int hflip, int vflip, int top, int left, int valid_h, int valid_w, int out_h, int out_w, const float* mean_std_host,
                                   float* out_img, long long* out_tgt, void* stream) {
  EGM_REQUIRE(H > 0 && W > 0 && rh > 0 && rw > 0 && out_h >= valid_h && out_w >= valid_w && valid_h >= 0 && valid_w >= 0, EGM_E_SHAPE, "input_transform: bad sizes");
  EGM_REQUIRE((rw == W) == (hmin == nullptr) && (rh == H) == (vmin == nullptr), EGM_E_BADARG, "input_transform: a resample table is needed exactly when the size changes");
  if ((long long)out_h * out_w == 0) return EGM_OK;
  PipeArgs a{img, mask, H, W, rh, rw, hmin, hcnt, hk, hks, vmin, vcnt, vk, vks, nnx, nny, hflip, vflip, top, left, valid_h, valid_w, out_h, out_w,
             {mean_std_host[0], mean_std_host[1], mean_std_host[2]}, {mean_std_host[3], mean_std_host[4], mean_std_host[5]}, out_img, out_tgt};
  egm_launch(k_input_transform, egm_grid_for((long long)out_h * out_w, 256, 8), 256, 0, (cudaStream_t)stream, a);
  EGM_LAUNCH_CHECK("input_transform"); return EGM_OK;
}
