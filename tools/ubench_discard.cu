// Does discard.global.L2 keep dead dirty lines from being written back to HBM on B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_discard tools/ubench_discard.cu
//   ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum tools/ubench_discard MODE
// Every CTA owns a private slab of `slab` bytes inside a footprint far above the L2 (148 x 1 MB) and loops: write the
// slab (dirty lines), read it back, then (mode 1) discard it.  mode 0: no discard (baseline: every written byte is evicted
// dirty -> DRAM writes ~ bytes written); mode 1: discard.global.L2 after the last read; mode 2: like 1 but through a second
// "rotating" slab so that a discarded line is only rewritten after the whole footprint cycled.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void k(float4* base, size_t slab_f4, int iters, int mode, int nslab, float* sink) {
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    float4* p = base + (static_cast<size_t>(blockIdx.x) * nslab + (it % nslab)) * slab_f4;
    for (size_t i = threadIdx.x; i < slab_f4; i += blockDim.x) p[i] = make_float4(it, i, 1.f, 2.f);
    __syncthreads();
    for (size_t i = threadIdx.x; i < slab_f4; i += blockDim.x) { float4 v = p[i]; acc += v.x + v.y; }
    __syncthreads();
    if (mode >= 1) {
      for (size_t i = threadIdx.x * 8; i < slab_f4; i += blockDim.x * 8)   // 8 float4 = 128 B
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(p + i) : "memory");
    }
    __syncthreads();
  }
  if (acc == 123.456f) *sink = acc;
}

int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 0;
  int nslab = argc > 2 ? atoi(argv[2]) : 4;           // slabs per CTA (footprint = 148 * nslab * 256 KB)
  int iters = argc > 3 ? atoi(argv[3]) : 400;
  const size_t slab = 256 << 10;
  float4* buf; float* sink;
  cudaMalloc(&buf, 148 * nslab * slab);
  cudaMalloc(&sink, 4);
  cudaMemset(buf, 0, 148 * nslab * slab);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<148, 512>>>(buf, slab / 16, 8, mode, nslab, sink);
  cudaEventRecord(e0);
  k<<<148, 512>>>(buf, slab / 16, iters, mode, nslab, sink);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("DISCARD mode %d nslab %d footprint %.0f MB: wrote %.1f GB in %.2f ms (%s)\n", mode, nslab, 148.0 * nslab * slab / 1e6,
         148.0 * iters * slab / 1e9, ms, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
