// Counter-based Philox4x32-10 (Salmon et al., SC'11) and the SDRM noise streams built on it.
// The same arithmetic is restated in oracle/philox_ref.py; tests pin one against the other.
//
// Stream layout (key = 64-bit seed, counter = {c0, c1, c2, c3}):
//   normals : c0 = column / 4, c1 = step, c2 = global row (low 32), c3 = STREAM_NORMAL | (row >> 32) << 8
//             -> 4 x u32 -> two Box-Muller pairs -> N(0,1) for columns 4*c0 .. 4*c0+3
//             step 0 is x_T; step i >= 2 is the z of reverse step i (train_SDRM.py:51,56)
//   dropout : (sampler, STREAM_MASK) c0 = column / 128, c1 = step, c2 = row, c3 = STREAM_MASK | ...
//             -> keep bit of column (128*c0 + 32*w + b) is bit b of output word w   (F.dropout p = 0.5, train_SDRM.py:100)
//             (training step, STREAM_TRAIN_MASK) c0 = column / 16, c1 = 0 -> keep bit of column (16*c0 + b) for the k-th of
//             the three denoiser inputs is bit b of output word k
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace sdrm {

enum : uint32_t { STREAM_NORMAL = 0u, STREAM_MASK = 1u, STREAM_TRAIN_NOISE = 2u, STREAM_TRAIN_MASK = 3u };

struct u32x4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
    uint64_t p0 = static_cast<uint64_t>(M0) * c0, p1 = static_cast<uint64_t>(M1) * c2;
    uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
    uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return u32x4{c0, c1, c2, c3};
}

#ifdef __CUDACC__
// Same function with the ten round keys precomputed (they only depend on the seed): 2 IMAD.WIDE + 2 LOP3 per round.
struct PhiloxKeys {
  uint32_t k0[10], k1[10];
};
__device__ __forceinline__ PhiloxKeys philox_make_keys(uint64_t seed) {
  PhiloxKeys K;
  uint32_t a = static_cast<uint32_t>(seed), b = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) { K.k0[r] = a; K.k1[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
  return K;
}
__device__ __forceinline__ u32x4 philox4x32_10_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& K) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0, p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ K.k0[r];
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ K.k1[r];
    c1 = static_cast<uint32_t>(p1); c3 = static_cast<uint32_t>(p0); c0 = n0; c2 = n2;
  }
  return u32x4{c0, c1, c2, c3};
}
// Box-Muller on the same 24-bit uniforms as box_muller below, with MUFU-only math (lg2, sqrt.approx, sin, cos) and the
// constants folded; differs from it by ~1 ulp.  u1 in (0,1] so lg2(u1) <= 0 and the radicand is never negative.
__device__ __forceinline__ void box_muller_fast(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float u1 = static_cast<float>((a >> 8) + 1u) * 5.9604644775390625e-8f;     // 2^-24
  const float r2 = __log2f(u1) * -1.3862943611198906f;                             // -2 ln(u1)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
  const float ang = static_cast<float>(b >> 8) * 3.7450702829239286e-7f;           // 2 pi 2^-24
  z0 = r * __cosf(ang);
  z1 = r * __sinf(ang);
}
__device__ __forceinline__ void philox_normal4_keys(const PhiloxKeys& K, uint32_t stream, uint64_t row, uint32_t step,
                                                    uint32_t col_quad, float (&z)[4]) {
  const u32x4 r = philox4x32_10_keys(col_quad, step, static_cast<uint32_t>(row), stream | (static_cast<uint32_t>(row >> 32) << 8), K);
  box_muller_fast(r.x, r.y, z[0], z[1]);
  box_muller_fast(r.z, r.w, z[2], z[3]);
}
// sampler dropout stream: 128 keep bits (columns 128*block .. 128*block+127) of (row, step)
__device__ __forceinline__ u32x4 philox_mask128(const PhiloxKeys& K, uint32_t stream, uint64_t row, uint32_t step, uint32_t block128) {
  return philox4x32_10_keys(block128, step, static_cast<uint32_t>(row), stream | (static_cast<uint32_t>(row >> 32) << 8), K);
}
#endif

}  // namespace sdrm
