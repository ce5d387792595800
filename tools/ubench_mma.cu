// Microbenchmark: tcgen05.mma issue throughput (cta_group::1, M=128, kind::f16, SS) on resident smem operands,
// optionally with a concurrent TMA bulk stream into other smem (interference test).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../sdrm_b200/csrc/ptx_sm100.cuh"
using namespace sdrm;

__global__ void __launch_bounds__(128, 1) mma_kernel(int N, int iters, int stages, const uint8_t* src, int tma_bytes, int tma_iters, int commit_each, int alt = 0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + 4 * 49152;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bars, 1); mbar_init(bars + 8, 1); mbar_init(bars + 16, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  // zero operands
  for (uint32_t i = threadIdx.x; i < 4 * 49152 / 16; i += 128) reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    for (int i = 0; i < iters; ++i) {
      const uint32_t st = (stages > 1) ? (i % stages) : 0;
      const uint64_t a = umma_desc_sw128(base + st * 49152), b = umma_desc_sw128(base + st * 49152 + 16384);
      const uint32_t d = alt ? tmem + (i & 1) * 256 : tmem + ((i / 15) & 1) * 256;   // alt: consecutive k-blocks accumulate into different TMEM columns
      for (int k = 0; k < 4; ++k) umma_bf16_ss(d, a + 2u * k, b + 2u * k, idesc, (i % 15) | k);
      if (commit_each) umma_commit(bars + 16);
      if (commit_each == 2) { tc_fence_after(); }
    }
    umma_commit(bars);
    mbar_wait(bars, 0, nullptr, 1);
  } else if (warp == 0 && lane == 0 && tma_iters > 0) {
    // free-running TMA stream into stage 3's W region (never read by the MMAs when stages <= 3)
    uint32_t ph = 0;
    for (int i = 0; i < tma_iters; ++i) {
      mbar_arrive_expect_tx(bars + 8, tma_bytes);
      bulk_g2s(base + 3 * 49152, src + (size_t)(i & 15) * 49152, tma_bytes, bars + 8);
      mbar_wait(bars + 8, ph, nullptr, 2); ph ^= 1;
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint8_t* src; cudaMalloc(&src, 1 << 20); cudaMemset(src, 0, 1 << 20);
  cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 49152 + 2048);
  const int iters = 20000;
  for (int alt = 0; alt < 2; ++alt)
  for (int tma = 1; tma < 2; ++tma)
    for (int stages = 3; stages <= 3; stages += 2)
      for (int N : {64, 112, 128, 208, 256}) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        mma_kernel<<<sms, 128, 4 * 49152 + 2048>>>(N, 100, stages, src, 49152, 0, tma, alt);
        cudaEventRecord(e0);
        mma_kernel<<<sms, 128, 4 * 49152 + 2048>>>(N, iters, stages, src, 49152, 0, tma, alt);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 128 * N * 64 * (double)iters * sms;
        printf("alt=%d commit_each=%d stages=%d N=%3d: %7.3f ms  %7.1f ns per 64-K block  %7.1f TFLOP/s  (%s)\n", alt, tma, stages, N, ms, ms * 1e6 / iters,
               flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
