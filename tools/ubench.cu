// Micro-benchmarks behind the attention kernel's design decisions (B200, sm_100a): throughput of the
// units one (query tile, key tile) step keeps busy, measured with clock64 inside ONE CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I long-vita_b200/csrc tools/ubench.cu -o tools/ubench
//   gpurun -- ./tools/ubench
// Prints one line per experiment: cycles per operation and the implied bytes (or ops) per cycle per SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace lv;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int ITERS = 256;

// ---- TMEM load / store: `nw` warps (warp w touches lane quadrant w % 4), x32 columns per instruction ----
template <int INFLIGHT, bool STORE>
__global__ void k_tmem(long long* out, int nw) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t r[INFLIGHT][32];
#pragma unroll
  for (int i = 0; i < INFLIGHT; ++i)
#pragma unroll
    for (int j = 0; j < 32; ++j) r[i][j] = threadIdx.x + j;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nw) {
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < INFLIGHT; ++i) {
        if (STORE)
          tmem_st32(base + (i % 4) * 32, r[i]);
        else
          tmem_ld32(base + (i % 4) * 32, r[i]);
      }
      if (STORE)
        tmem_wait_st();
      else
        tmem_wait_ld();
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < INFLIGHT; ++i)
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += r[i][j];
  if (acc == 0x12345678u) out[63] = acc;
  if (threadIdx.x == 0) out[0] = t1 - t0;
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

// ---- MUFU ex2 (+ the FFMA / FADD / pack around it, like one softmax element) ----
template <int MODE>   // 0: ex2 only, 1: ffma + ex2 + fadd + pack (softmax element), 2: polynomial exp2 only, 3: 3 of 4 MUFU + 1 poly
__global__ void k_exp(long long* out, float* sink, int nw) {
  const int warp = threadIdx.x >> 5;
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = -0.001f * (threadIdx.x + j);
  float l = 0.f;
  uint32_t pk = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nw) {
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float a, b;
        if (MODE == 0) {
          a = ex2(x[j]);
          b = ex2(x[j + 1]);
          x[j] = a - 1.5f;
          x[j + 1] = b - 1.5f;
        } else {
          const float xa = fmaf(x[j], 0.99f, -0.25f), xb = fmaf(x[j + 1], 0.99f, -0.25f);
          auto poly = [](float v) {
            v = fmaxf(v, -125.f);
            const float t = v + 12582912.f;
            const float f = v - (t - 12582912.f);
            float p = fmaf(0.055171321f, f, 0.24261054f);
            p = fmaf(p, f, 0.69326099f);
            p = fmaf(p, f, 0.99992811f);
            return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
          };
          if (MODE == 1) {
            a = ex2(xa);
            b = ex2(xb);
          } else if (MODE == 2) {
            a = poly(xa);
            b = poly(xb);
          } else {
            a = ex2(xa);
            b = ((j & 2) ? poly(xb) : ex2(xb));
          }
          l += a;
          l += b;
          pk ^= pack_bf16(a, b);
          x[j] = a - 1.5f;
          x[j + 1] = b - 1.5f;
        }
      }
    }
  }
  long long t1 = clock64();
  float acc = l + __uint_as_float(pk);
#pragma unroll
  for (int j = 0; j < 32; ++j) acc += x[j];
  if (acc == 123.456f) sink[0] = acc;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

// ---- UMMA: NMMA back-to-back 128 x N x 16 MMAs, SS (both operands in smem) or TS (A from TMEM) ----
template <bool TS, int N, bool B_MN>
__global__ void k_mma(long long* out, int nmma) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, B_MN ? 1 : 0);
    const uint64_t adesc = make_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t bdesc = B_MN ? make_smem_desc(smem_u32(smem + 32768), 16384, 1024) : make_smem_desc(smem_u32(smem + 32768), 16, 1024);
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < nmma; ++i) {
        const int kk = i & 7;
        if (TS)
          umma_ts(tm + 256, tm + kk * 8, bdesc + (uint64_t)(B_MN ? kk * 128 : (((kk / 4) * 16384 + (kk % 4) * 32) >> 4)), idesc, 1u);
        else
          umma_ss(tm + 256, adesc + (((kk / 4) * 16384 + (kk % 4) * 32) >> 4), bdesc + (((kk / 4) * 16384 + (kk % 4) * 32) >> 4), idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static long long run_and_get(long long* d_out) {
  CK(cudaDeviceSynchronize());
  long long h;
  CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
  return h;
}

int main() {
  long long* d_out;
  float* d_sink;
  CK(cudaMalloc(&d_out, 64 * 8));
  CK(cudaMalloc(&d_sink, 64));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);

  for (int rep = 0; rep < 2; ++rep) {   // rep 0 warms up
    for (int nw : {1, 4, 8}) {
      k_tmem<1, false><<<1, 256>>>(d_out, nw);
      long long c1 = run_and_get(d_out);
      k_tmem<4, false><<<1, 256>>>(d_out, nw);
      long long c4 = run_and_get(d_out);
      k_tmem<1, true><<<1, 256>>>(d_out, nw);
      long long s1 = run_and_get(d_out);
      k_tmem<4, true><<<1, 256>>>(d_out, nw);
      long long s4 = run_and_get(d_out);
      if (rep)
        printf("tmem x32 (4 KB / warp-instr) warps=%d: ld 1-in-flight %.1f cyc/instr, ld 4-in-flight %.1f cyc/instr (%.0f B/cyc/SM); "
               "st 1-in-flight %.1f, st 4-in-flight %.1f cyc/instr (%.0f B/cyc/SM)\n",
               nw, (double)c1 / ITERS, (double)c4 / ITERS / 4, 4096.0 * nw / ((double)c4 / ITERS / 4), (double)s1 / ITERS,
               (double)s4 / ITERS / 4, 4096.0 * nw / ((double)s4 / ITERS / 4));
    }
    for (int nw : {1, 4, 8}) {
      long long c[4];
      k_exp<0><<<1, 256>>>(d_out, d_sink, nw);
      c[0] = run_and_get(d_out);
      k_exp<1><<<1, 256>>>(d_out, d_sink, nw);
      c[1] = run_and_get(d_out);
      k_exp<2><<<1, 256>>>(d_out, d_sink, nw);
      c[2] = run_and_get(d_out);
      k_exp<3><<<1, 256>>>(d_out, d_sink, nw);
      c[3] = run_and_get(d_out);
      if (rep)
        printf("exp warps=%d (cycles per 32-lane element): ex2 only %.2f | softmax element (ffma+ex2+fadd+pack) %.2f | all-poly %.2f | 3 MUFU : 1 poly %.2f  "
               "-> 128x128 tile on 4 warps: %.0f / %.0f / %.0f / %.0f cyc\n",
               nw, (double)c[0] / ITERS / 32, (double)c[1] / ITERS / 32, (double)c[2] / ITERS / 32, (double)c[3] / ITERS / 32,
               (double)c[0] / ITERS / 32 * 128, (double)c[1] / ITERS / 32 * 128, (double)c[2] / ITERS / 32 * 128, (double)c[3] / ITERS / 32 * 128);
    }
    const int smem = 97 * 1024 + 1024;
    CK(cudaFuncSetAttribute(k_mma<false, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_mma<true, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_mma<false, 256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_mma<false, 64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_mma<true, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int n : {8, 64, 512}) {
      k_mma<false, 128, false><<<1, 64, smem>>>(d_out, n);
      long long a = run_and_get(d_out);
      k_mma<true, 128, true><<<1, 64, smem>>>(d_out, n);
      long long b = run_and_get(d_out);
      k_mma<false, 256, false><<<1, 64, smem>>>(d_out, n);
      long long c = run_and_get(d_out);
      k_mma<false, 64, false><<<1, 64, smem>>>(d_out, n);
      long long d = run_and_get(d_out);
      k_mma<true, 64, true><<<1, 64, smem>>>(d_out, n);
      long long e = run_and_get(d_out);
      if (rep)
        printf("umma x%d (issue -> commit -> mbarrier): SS 128x128x16 %.1f cyc/mma | TS 128x128x16 (B MN-major) %.1f | SS 128x256x16 %.1f | SS 128x64x16 %.1f | TS 128x64x16 %.1f\n",
               n, (double)a / n, (double)b / n, (double)c / n, (double)d / n, (double)e / n);
    }
  }
  return 0;
}
