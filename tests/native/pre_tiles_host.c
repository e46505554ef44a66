/* Host build of the image-tiling kernels' per-element bodies (long-vita_b200/csrc/preprocess_core.h) - TEST
 * INFRASTRUCTURE: tests/test_preprocess_tiles_host.py compiles this with gcc and runs the same loops the CUDA kernels
 * run (one item per thread there, a plain for loop here) against the oracle on a box without a GPU. */
#include "preprocess_core.h"

void lv_host_resize_h(const uint8_t* image, uint8_t* tmp, const int* xmin, const int* cnt, const int* kk, int ksize, int H, int W,
                      int OW) {
  for (long long i = 0; i < (long long)H * OW; ++i) lv_pre_resize_h_item(i, image, tmp, xmin, cnt, kk, ksize, W, OW);
}

void lv_host_resize_v_tiles(const uint8_t* tmp, uint16_t* out, const int* ymin, const int* cnt, const int* kk, int ksize, int OH,
                            int OW, int S, int tile_base, const float* mean, const float* std) {
  for (long long i = 0; i < (long long)OH * OW; ++i)
    lv_pre_resize_v_tile_item(i, tmp, out, ymin, cnt, kk, ksize, OW, S, tile_base, mean[0], mean[1], mean[2], std[0], std[1], std[2]);
}
