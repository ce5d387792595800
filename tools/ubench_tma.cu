// Microbenchmark: per-SM ingest bandwidth of cp.async.bulk (TMA bulk engine) from L2 into shared memory.
// Each CTA streams `iters` stage-fills through a ring of S stages of B bytes; a consumer thread frees each stage at once.
// mode 0: all CTAs read the SAME region (weights-like), mode 1: each CTA reads a private region (activations-like).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../sdrm_b200/csrc/ptx_sm100.cuh"
using namespace sdrm;

__global__ void __launch_bounds__(128, 1) stream_kernel(const uint8_t* src, size_t region, int private_region, int S, int B,
                                                        int copies, int iters, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + S * B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (S + s), 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const uint8_t* my = src + (private_region ? (size_t)blockIdx.x * region : 0);
  long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    uint32_t st = 0, ph = 0; size_t off = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(bars + 8 * (S + st), ph ^ 1, nullptr, 1);
      mbar_arrive_expect_tx(bars + 8 * st, B);
      const int piece = B / copies;
      for (int c = 0; c < copies; ++c) bulk_g2s(base + st * B + c * piece, my + off + c * piece, piece, bars + 8 * st);
      off += B; if (off + B > region) off = 0;
      if (++st == S) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    uint32_t st = 0, ph = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(bars + 8 * st, ph, nullptr, 2);
      mbar_arrive(bars + 8 * (S + st));
      if (++st == S) { st = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint8_t* src; cudaMalloc(&src, (size_t)(1 << 20) * sms); cudaMemset(src, 1, (size_t)(1 << 20) * sms);
  unsigned long long* cyc; cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 2048);
  const int cfgs[][3] = {{4, 49152, 2}, {4, 49152, 1}, {8, 24576, 1}, {2, 98304, 2}, {3, 65536, 4}};
  for (int priv = 0; priv < 3; ++priv)
    for (auto& c : cfgs) {
      const size_t region = priv == 1 ? (256 << 10) : (1 << 20);  // priv 1: 37 MB total (L2 resident); priv 2: 148 MB (HBM)
      const int S = c[0], B = c[1], copies = c[2], iters = 4000;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      stream_kernel<<<sms, 128, S * B + 2048>>>(src, region, priv > 0, S, B, copies, 200, cyc);
      cudaEventRecord(e0);
      stream_kernel<<<sms, 128, S * B + 2048>>>(src, region, priv > 0, S, B, copies, iters, cyc);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      cudaError_t err = cudaGetLastError();
      const double bytes = (double)B * iters;
      printf("private=%d stages=%2d stage=%6d B copies=%d : %7.3f ms  %6.1f GB/s/SM  %6.2f TB/s chip  (%s)\n", priv, S, B, copies, ms,
             bytes / ms / 1e6, bytes * sms / ms / 1e9, cudaGetErrorString(err));
    }
  return 0;
}
