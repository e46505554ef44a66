// Per-output-element bodies of the image-tiling kernels (csrc/preprocess.cu: pre_resize_h_kernel,
// pre_resize_v_tiles_kernel).  Plain C-style code without CUDA types so that the SAME text also compiles with the host
// compiler: tests/test_preprocess_tiles_host.py builds it with gcc (LV_HD empty) and runs the loops on the CPU against
// the oracle, bit for bit - index arithmetic, fixed-point accumulation, clipping, float32 normalisation and the bf16
// rounding are then the ones the kernels execute.
#ifndef LV_PREPROCESS_CORE_H_
#define LV_PREPROCESS_CORE_H_

#include <stdint.h>

#ifdef __CUDACC__
#define LV_HD __host__ __device__ __forceinline__
#else
#define LV_HD static inline
#endif

#define LV_PRE_BITS 22 /* Pillow: PRECISION_BITS = 32 - 8 - 2 (Resample.c) */

LV_HD int lv_pre_clip8(int acc) {
  const int v = acc >> LV_PRE_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// float32 -> bf16 bits, round to nearest even (inputs here are finite).
LV_HD uint16_t lv_pre_bf16_bits(float f) {
  union {
    float f;
    uint32_t u;
  } c;
  c.f = f;
  return (uint16_t)((c.u + 0x7FFFu + ((c.u >> 16) & 1u)) >> 16);
}

// ((v * 1.0f) / 255.0f - mean) / std with IEEE float32 operations in the reference's order
// (image_processor.py:211-216); the device build pins the roundings with intrinsics, the host build relies on
// -ffp-contract=off (there is no multiply-add pattern to contract anyway).
LV_HD float lv_pre_normalise(int v, float mean, float std) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(__fsub_rn(__fdiv_rn(__fmul_rn((float)v, 1.0f), 255.0f), mean), std);
#else
  volatile float a = (float)v * 1.0f;
  volatile float b = a / 255.0f;
  volatile float c = b - mean;
  return c / std;
#endif
}

// Horizontal pass, item i = (y, ox) of a [H, OW] intermediate: tmp[y, ox, c] = clip8(sum_t image[y, xmin[ox] + t, c] * k[ox, t]).
LV_HD void lv_pre_resize_h_item(long long i, const uint8_t* image, uint8_t* tmp, const int* xmin, const int* cnt, const int* kk,
                                int ksize, int W, int OW) {
  const int ox = (int)(i % OW);
  const long long y = i / OW;
  const int x0 = xmin[ox], c = cnt[ox];
  const int* k = kk + (long long)ox * ksize;
  const uint8_t* src = image + (y * W + x0) * 3;
  int a0 = 1 << (LV_PRE_BITS - 1), a1 = a0, a2 = a0;
  for (int t = 0; t < c; ++t) {
    const int w = k[t];
    a0 += src[0] * w;
    a1 += src[1] * w;
    a2 += src[2] * w;
    src += 3;
  }
  uint8_t* d = tmp + i * 3;
  d[0] = (uint8_t)lv_pre_clip8(a0);
  d[1] = (uint8_t)lv_pre_clip8(a1);
  d[2] = (uint8_t)lv_pre_clip8(a2);
}

// Vertical pass + normalisation + tiling, item i = (oy, ox) of the [OH, OW] resized image: the pixel lands in tile
// tile_base + (oy / S) * (OW / S) + ox / S (dynamic_preprocess's crop boxes, image_processor.py:431-440, row-major
// over the grid) at (oy % S, ox % S), channel-first.  out holds bf16 bit patterns.
LV_HD void lv_pre_resize_v_tile_item(long long i, const uint8_t* tmp, uint16_t* out, const int* ymin, const int* cnt, const int* kk,
                                     int ksize, int OW, int S, int tile_base, float m0, float m1, float m2, float s0, float s1,
                                     float s2) {
  const int ox = (int)(i % OW);
  const int oy = (int)(i / OW);
  const int y0 = ymin[oy], c = cnt[oy];
  const int* k = kk + (long long)oy * ksize;
  const uint8_t* src = tmp + ((long long)y0 * OW + ox) * 3;
  int a0 = 1 << (LV_PRE_BITS - 1), a1 = a0, a2 = a0;
  for (int t = 0; t < c; ++t) {
    const int w = k[t];
    a0 += src[0] * w;
    a1 += src[1] * w;
    a2 += src[2] * w;
    src += (long long)OW * 3;
  }
  const long long plane = (long long)S * S;
  const long long tile = tile_base + (long long)(oy / S) * (OW / S) + ox / S;
  uint16_t* d = out + tile * 3 * plane + (long long)(oy % S) * S + (ox % S);
  d[0] = lv_pre_bf16_bits(lv_pre_normalise(lv_pre_clip8(a0), m0, s0));
  d[plane] = lv_pre_bf16_bits(lv_pre_normalise(lv_pre_clip8(a1), m1, s1));
  d[2 * plane] = lv_pre_bf16_bits(lv_pre_normalise(lv_pre_clip8(a2), m2, s2));
}

#endif  // LV_PREPROCESS_CORE_H_
