// Frame preprocessing on the GPU: uint8 video frames -> the bf16 [N, 3, S, S] tensor the vision
// tower consumes, bit-identical to `ImageProcessor.process_images`
// (long_vita/data/processor/image_processor.py:183-223: expand2square with the mean colour :192-205, PIL BICUBIC
// resize :207-209, 1/255 scaling and mean / std normalisation in float32 :211-216, channel-first :218-221) followed by
// the `.to(bfloat16)` of the model's input.  4096 frames are 4.9 GB of bf16 pixels: produced here from ~0.8-25 GB of
// decoded uint8 frames without a host-side float tensor.
//
// The resize is Pillow's 8-bit ImagingResample (Resample.c), which is integer arithmetic: per output coordinate a
// window of input pixels with fixed-point weights (22 fractional bits, computed ON THE HOST in float64 exactly as
// Pillow does - long_vita_b200/preprocess.py), horizontal pass to a uint8 intermediate, vertical pass, each pass
// accumulating in int32 from 1 << 21 and clipping (acc >> 22) to [0, 255].  Both passes resample the SQUARE canvas
// (side n = max(H, W)) to S, so they share one coefficient table.  HBM-bound byte work, coalesced along x.
#include <cuda_bf16.h>

#include "common.cuh"
#include "preprocess_core.h"

namespace lv {

constexpr int PRE_BITS = 22;

__device__ __forceinline__ int clip8(int acc) {
  const int v = acc >> PRE_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// tmp[f, y, ox, c] = horizontal pass of canvas row y.  canvas(y, x) = frame(y - top, x - left) inside the pasted
// rectangle, the background colour outside.  One thread per (f, y, ox).
__global__ void __launch_bounds__(256) pre_hpass_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ tmp,
                                                        const int* __restrict__ xmin, const int* __restrict__ cnt,
                                                        const int* __restrict__ kk, int ksize, int n_frames, int H, int W,
                                                        int n, int top, int left, int S, int bg0, int bg1, int bg2) {
  const long long total = (long long)n_frames * n * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % S);
    const long long r = i / S;
    const int y = (int)(r % n);
    const long long f = r / n;
    const int x0 = xmin[ox], c = cnt[ox];
    const int* k = kk + (long long)ox * ksize;
    int a0 = 1 << (PRE_BITS - 1), a1 = a0, a2 = a0;
    const int fy = y - top;
    const bool row_in = fy >= 0 && fy < H;
    const uint8_t* src = frames + ((f * H + (row_in ? fy : 0)) * W) * 3;
    for (int t = 0; t < c; ++t) {
      const int fx = x0 + t - left;
      int p0 = bg0, p1 = bg1, p2 = bg2;
      if (row_in && fx >= 0 && fx < W) {
        p0 = src[fx * 3 + 0];
        p1 = src[fx * 3 + 1];
        p2 = src[fx * 3 + 2];
      }
      const int w = k[t];
      a0 += p0 * w;
      a1 += p1 * w;
      a2 += p2 * w;
    }
    uint8_t* d = tmp + i * 3;
    d[0] = (uint8_t)clip8(a0);
    d[1] = (uint8_t)clip8(a1);
    d[2] = (uint8_t)clip8(a2);
  }
}

// out[f, c, oy, ox] = bf16(((vertical pass)[oy, ox, c] * 1.0f / 255.0f - mean[c]) / std[c]) with IEEE float32
// operations in the reference's order (no FMA contraction, true division).
__global__ void __launch_bounds__(256) pre_vpass_norm_kernel(const uint8_t* __restrict__ tmp, __nv_bfloat16* __restrict__ out,
                                                             const int* __restrict__ ymin, const int* __restrict__ cnt,
                                                             const int* __restrict__ kk, int ksize, int n_frames, int n, int S,
                                                             float m0, float m1, float m2, float s0, float s1, float s2) {
  const long long total = (long long)n_frames * S * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % S);
    const long long r = i / S;
    const int oy = (int)(r % S);
    const long long f = r / S;
    const int y0 = ymin[oy], c = cnt[oy];
    const int* k = kk + (long long)oy * ksize;
    int a0 = 1 << (PRE_BITS - 1), a1 = a0, a2 = a0;
    const uint8_t* src = tmp + ((f * n + y0) * S + ox) * 3;
    for (int t = 0; t < c; ++t) {
      const int w = k[t];
      a0 += src[0] * w;
      a1 += src[1] * w;
      a2 += src[2] * w;
      src += (long long)S * 3;
    }
    const float v0 = __fdiv_rn(__fsub_rn(__fdiv_rn(__fmul_rn((float)clip8(a0), 1.0f), 255.0f), m0), s0);
    const float v1 = __fdiv_rn(__fsub_rn(__fdiv_rn(__fmul_rn((float)clip8(a1), 1.0f), 255.0f), m1), s1);
    const float v2 = __fdiv_rn(__fsub_rn(__fdiv_rn(__fmul_rn((float)clip8(a2), 1.0f), 255.0f), m2), s2);
    const long long plane = (long long)S * S;
    __nv_bfloat16* d = out + f * 3 * plane + (long long)oy * S + ox;
    d[0] = __float2bfloat16_rn(v0);
    d[plane] = __float2bfloat16_rn(v1);
    d[2 * plane] = __float2bfloat16_rn(v2);
  }
}

// Dynamic-patch tiling of ONE image (image_processor.py:263-285 process_dynamic, :404-448 dynamic_preprocess): the image
// is resized to a grid of S x S tiles (out_w x out_h, aspect ratio NOT kept - the host picks the grid) and cut into the
// tiles; no padding.  Two passes like the frame path, but with separate tables for the two axes; the per-element bodies
// live in preprocess_core.h (also compiled and tested on the host).
__global__ void __launch_bounds__(256) pre_resize_h_kernel(const uint8_t* __restrict__ image, uint8_t* __restrict__ tmp,
                                                           const int* __restrict__ xmin, const int* __restrict__ cnt,
                                                           const int* __restrict__ kk, int ksize, int H, int W, int OW) {
  const long long total = (long long)H * OW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    lv_pre_resize_h_item(i, image, tmp, xmin, cnt, kk, ksize, W, OW);
}

__global__ void __launch_bounds__(256) pre_resize_v_tiles_kernel(const uint8_t* __restrict__ tmp, uint16_t* __restrict__ out,
                                                                 const int* __restrict__ ymin, const int* __restrict__ cnt,
                                                                 const int* __restrict__ kk, int ksize, int OH, int OW, int S,
                                                                 int tile_base, float m0, float m1, float m2, float s0, float s1,
                                                                 float s2) {
  const long long total = (long long)OH * OW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    lv_pre_resize_v_tile_item(i, tmp, out, ymin, cnt, kk, ksize, OW, S, tile_base, m0, m1, m2, s0, s1, s2);
}

}  // namespace lv

using namespace lv;

extern "C" int64_t lv_frame_preprocess_ws_bytes(int64_t n_frames, int64_t H, int64_t W, int64_t S) {
  const int64_t n = H > W ? H : W;
  return n_frames * n * S * 3;
}

extern "C" int lv_frame_preprocess(const void* frames, void* out, void* ws, const int32_t* win_min, const int32_t* win_cnt,
                                   const int32_t* coeff, int64_t ksize, int64_t n_frames, int64_t H, int64_t W, int64_t S,
                                   const int32_t* background, const float* mean, const float* std, lv_stream_t stream) {
  LV_CHECK_ARG(n_frames >= 0 && H > 0 && W > 0 && S > 0 && ksize > 0, "lv_frame_preprocess: empty shape");
  LV_CHECK_ARG(H < (1 << 15) && W < (1 << 15) && S < (1 << 15), "lv_frame_preprocess: image side too large");
  if (n_frames == 0) return LV_OK;      // an empty batch has no buffers to check
  LV_CHECK_ARG(frames && out && ws && win_min && win_cnt && coeff && background && mean && std, "lv_frame_preprocess: null pointer");
  LV_BIND_DEVICE(frames);
  const int n = (int)(H > W ? H : W);
  const int top = W > H ? (int)((W - H) / 2) : 0;       // expand2square: result.paste(img, (0, (width - height) // 2))
  const int left = H > W ? (int)((H - W) / 2) : 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t cap = 32 * (int64_t)sm_count();
  {
    const int64_t total = n_frames * n * S;
    const int64_t blocks = (total + 255) / 256;
    pre_hpass_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(
        reinterpret_cast<const uint8_t*>(frames), reinterpret_cast<uint8_t*>(ws), win_min, win_cnt, coeff, (int)ksize, (int)n_frames,
        (int)H, (int)W, n, top, left, (int)S, background[0], background[1], background[2]);
    LV_CHECK_LAUNCH("pre_hpass_kernel");
  }
  {
    const int64_t total = n_frames * S * S;
    const int64_t blocks = (total + 255) / 256;
    pre_vpass_norm_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(
        reinterpret_cast<const uint8_t*>(ws), reinterpret_cast<__nv_bfloat16*>(out), win_min, win_cnt, coeff, (int)ksize, (int)n_frames, n,
        (int)S, mean[0], mean[1], mean[2], std[0], std[1], std[2]);
    LV_CHECK_LAUNCH("pre_vpass_norm_kernel");
  }
  return LV_OK;
}

extern "C" int64_t lv_image_tiles_ws_bytes(int64_t H, int64_t out_w) { return H * out_w * 3; }

extern "C" int lv_image_tiles_preprocess(const void* image, void* out, void* ws, const int32_t* x_min, const int32_t* x_cnt,
                                         const int32_t* x_coeff, int64_t x_ksize, const int32_t* y_min, const int32_t* y_cnt,
                                         const int32_t* y_coeff, int64_t y_ksize, int64_t H, int64_t W, int64_t out_h,
                                         int64_t out_w, int64_t S, int64_t tile_base, const float* mean, const float* std,
                                         lv_stream_t stream) {
  LV_CHECK_ARG(H > 0 && W > 0 && out_h > 0 && out_w > 0 && S > 0 && x_ksize > 0 && y_ksize > 0 && tile_base >= 0,
               "lv_image_tiles_preprocess: empty shape");
  LV_CHECK_ARG(H < (1 << 15) && W < (1 << 15) && out_h < (1 << 15) && out_w < (1 << 15), "lv_image_tiles_preprocess: image side too large");
  LV_CHECK_ARG(out_h % S == 0 && out_w % S == 0, "lv_image_tiles_preprocess: the resized image (%lld x %lld) is not a grid of %lld-pixel tiles",
               (long long)out_w, (long long)out_h, (long long)S);
  LV_CHECK_ARG(image && out && ws && x_min && x_cnt && x_coeff && y_min && y_cnt && y_coeff && mean && std,
               "lv_image_tiles_preprocess: null pointer");
  LV_BIND_DEVICE(image);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t cap = 32 * (int64_t)sm_count();
  {
    const int64_t blocks = (H * out_w + 255) / 256;
    pre_resize_h_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(
        reinterpret_cast<const uint8_t*>(image), reinterpret_cast<uint8_t*>(ws), x_min, x_cnt, x_coeff, (int)x_ksize, (int)H, (int)W,
        (int)out_w);
    LV_CHECK_LAUNCH("pre_resize_h_kernel");
  }
  {
    const int64_t blocks = (out_h * out_w + 255) / 256;
    pre_resize_v_tiles_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(
        reinterpret_cast<const uint8_t*>(ws), reinterpret_cast<uint16_t*>(out), y_min, y_cnt, y_coeff, (int)y_ksize, (int)out_h,
        (int)out_w, (int)S, (int)tile_base, mean[0], mean[1], mean[2], std[0], std[1], std[2]);
    LV_CHECK_LAUNCH("pre_resize_v_tiles_kernel");
  }
  return LV_OK;
}
