// K5 — tcgen05 GEMM of the diffusion TRAINING step (reference: the three denoiser forwards and their autograd backward,
// train_SDRM.py:191-199 + 331-337, i.e. 3 forward and 6 backward dense products per layer and minibatch).
//
//   C[M, N] = A[M, K] . B[N, K]^T      A, B: bf16 row-major (K contiguous), given as hi and lo halves of an fp32 matrix
//                                      (x = hi + lo, hi = bf16(x), lo = bf16(x - hi)); fp32 accumulate in TMEM.
//   passes = 3: hi.hi + hi.lo + lo.hi  ("bf16x3": ~2^-16 relative per product -- the training step's parity bar against the
//               reference's fp32 autograd is 2e-4 of max|grad|, which single-pass bf16 operands do not meet)
//   passes = 1: hi.hi only             (plain bf16 operands, 3x less tensor work, for callers that accept mixed precision)
//
// The same product serves the forward layers (A = activations, B = W), the data gradients (A = dY, B = W^T) and the weight
// gradients (A = dY^T, B = X^T, split over K = rows into slabs); the transposed operand images are written by
// operand_prep_kernel (train_gemm.cu), so every operand is K-major and one kernel covers all nine products.
//
// Structure: persistent clusters of 2 CTAs = one tcgen05 cta_group::2 pair per 256 x BN output tile (UMMA M = 256, N = BN <= 256,
// K = 16).  Each CTA stages its own 128 rows of A (hi + lo) and HALF of the B tile (hi + lo): a 64-wide k-block is 64 KB per
// CTA and serves 3 x 4 UMMAs (1536 tensor cycles) = 42 B/clk of L2 -> SM traffic per SM, under the SM's ~64 B/clk port.
// Three such stages; two 256-column accumulators in TMEM so the epilogue of tile i overlaps the UMMAs of tile i+1.
// warp 0 = TMA producer, warp 1 = UMMA issuer (leader CTA) + TMEM allocation, warps 2-9 = epilogue (two per TMEM lane quarter,
// alternating 16-column groups; the next group's accumulator columns and PReLU' inputs are requested one group ahead).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "ptx_sm100.cuh"

namespace sdrm {

enum GemmEpi : int {
  GEMM_EPI_STORE = 0,        // C = acc (+ bias)                                          (weight-gradient slabs, plain products)
  GEMM_EPI_PRELU_SPLIT = 1,  // C = pre = acc + bias;  o_hi/o_lo = bf16 split of PReLU(pre)  (forward hidden layers)
  GEMM_EPI_TANH = 2,         // C = tanh(acc + bias)                                      (forward output layer)
  GEMM_EPI_DPRELU = 3        // C = g = acc * PReLU'(aux);  o_hi/o_lo = split of g;  slope_grad += sum acc * min(aux, 0)
};

struct GemmParams {
  CUtensorMap tmA[2];        // hi, lo: dims {K, M}, box {64, 128}, SWIZZLE_128B
  CUtensorMap tmB[2];        // hi, lo: dims {K, N}, box {64, BN / 2}
  int M, N, K;
  int BN;                    // tile width, multiple of 16, <= 256
  int kb_total, kb_per_split, splits;
  int passes;                // 1 or 3
  int m_tiles, n_tiles;      // tiles of 256 rows x BN columns
  float* C;
  long long ldc, slab_stride;   // split s writes C + s * slab_stride
  const float* bias;            // [N] or nullptr
  const float* bias_table;      // [*, bias_ld]: row m adds bias_table[bias_rows[m]] (hoisted time embedding, train_SDRM.py:98-101) or nullptr
  const long long* bias_rows;
  long long bias_ld;
  int epi;
  const float* slope;           // PReLU slope (device scalar)
  __nv_bfloat16* o_hi;          // bf16 row-major outputs [M, ldo] (PRELU_SPLIT, DPRELU) or nullptr
  __nv_bfloat16* o_lo;
  long long ldo;
  const float* aux;             // DPRELU: pre-activation the gradient flows through [M, ld_aux]
  long long ld_aux;
  double* slope_grad;           // DPRELU: accumulated d loss / d slope (or nullptr)
  int* err_word;
};

constexpr int GEMM_EPI_WARPS = 8;                 // two per TMEM lane quarter: the epilogue is latency-bound (aux / bias loads, stores)
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_NSTG = 3;
constexpr uint32_t GEMM_TILE_A = 128 * 128;                 // one 128-row x 64-column bf16 operand image
constexpr uint32_t GEMM_STG_BYTES = 4 * GEMM_TILE_A;        // A hi | A lo | B hi (<= 128 rows) | B lo
constexpr int GEMM_SMEM_BYTES = GEMM_NSTG * GEMM_STG_BYTES + 1024 + 256;
enum : int { WD_G_PRODUCER = 501, WD_G_MMA_FULL = 502, WD_G_MMA_ACC = 503, WD_G_EPI = 504 };

__global__ void __launch_bounds__(GEMM_THREADS, 1) sdrm_gemm_pair_kernel(const __grid_constant__ GemmParams P) {
  extern __shared__ uint8_t gsm_raw[];
  const uint32_t raw_addr = smem_u32(gsm_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  const uint32_t bar_base = base + GEMM_NSTG * GEMM_STG_BYTES;
  auto stage_addr = [&](uint32_t s, uint32_t which) { return base + s * GEMM_STG_BYTES + which * GEMM_TILE_A; };   // which: 0 A hi, 1 A lo, 2 B hi, 3 B lo
  auto bar_full = [&](uint32_t s) { return bar_base + 8u * s; };
  auto bar_empty = [&](uint32_t s) { return bar_base + 8u * (GEMM_NSTG + s); };
  auto bar_acc_full = [&](uint32_t b) { return bar_base + 8u * (2 * GEMM_NSTG + b); };
  auto bar_acc_empty = [&](uint32_t b) { return bar_base + 8u * (2 * GEMM_NSTG + 2 + b); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gsm_raw + (bar_base - raw_addr) + 8 * (2 * GEMM_NSTG + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();   // 0 = leader: issues the UMMAs, owns the full / accumulator barriers
  int* err = P.err_word;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < GEMM_NSTG; ++s) {
      mbar_init(bar_full(s), 1);     // the leader's producer arms the bytes of both CTAs
      mbar_init(bar_empty(s), 1);    // one multicast commit per consumed stage
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), 2 * GEMM_EPI_WARPS);   // the epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_clusters = gridDim.x >> 1;
  const int my_cluster = blockIdx.x >> 1;
  const long long tiles_mn = static_cast<long long>(P.m_tiles) * P.n_tiles;
  const long long n_work = tiles_mn * P.splits;
  const bool x3 = P.passes == 3;
  const uint32_t b_half_bytes = static_cast<uint32_t>(P.BN / 2) * 128u;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    const uint64_t pol = l2_policy_evict_last();
    uint32_t stage = 0, sphase = 0;
    const uint32_t tx_bytes = 2u * (x3 ? 2u : 1u) * (GEMM_TILE_A + b_half_bytes);   // both CTAs
    for (long long wk = my_cluster; wk < n_work; wk += n_clusters) {
      const int split = static_cast<int>(wk / tiles_mn);
      const long long tmn = wk - split * tiles_mn;
      const int mt = static_cast<int>(tmn / P.n_tiles), nt = static_cast<int>(tmn % P.n_tiles);   // n fastest: the clusters running together share A row blocks (L2 hits)
      const int a_row = mt * 256 + static_cast<int>(rank) * 128;
      const int b_row = nt * P.BN + static_cast<int>(rank) * (P.BN / 2);
      const int kb0 = split * P.kb_per_split, kb1 = min(P.kb_total, kb0 + P.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_empty(stage), sphase ^ 1u, err, WD_G_PRODUCER);
        if (elect_one()) {
          const uint32_t fb0 = mapa_cluster(bar_full(stage), 0);
          if (rank == 0) mbar_arrive_expect_tx(bar_full(stage), tx_bytes);
          tma_load_2d_pair_hint(mapa_cluster(stage_addr(stage, 0), rank), &P.tmA[0], kb * 64, a_row, fb0, pol);
          tma_load_2d_pair_hint(mapa_cluster(stage_addr(stage, 2), rank), &P.tmB[0], kb * 64, b_row, fb0, pol);
          if (x3) {
            tma_load_2d_pair_hint(mapa_cluster(stage_addr(stage, 1), rank), &P.tmA[1], kb * 64, a_row, fb0, pol);
            tma_load_2d_pair_hint(mapa_cluster(stage_addr(stage, 3), rank), &P.tmB[1], kb * 64, b_row, fb0, pol);
          }
        }
        __syncwarp();
        if (++stage == GEMM_NSTG) { stage = 0; sphase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ================================ UMMA issuer (leader CTA) ================================
      uint32_t stage = 0, sphase = 0, cc = 0;
      const uint32_t idesc = umma_idesc_bf16(256, P.BN);
      for (long long wk = my_cluster; wk < n_work; wk += n_clusters) {
        const int split = static_cast<int>(wk / tiles_mn);
        const int kb0 = split * P.kb_per_split, kb1 = min(P.kb_total, kb0 + P.kb_per_split);
        const uint32_t buf = cc & 1u;
        mbar_wait(bar_acc_empty(buf), ((cc >> 1) & 1u) ^ 1u, err, WD_G_MMA_ACC);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256u;
        uint32_t acc = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full(stage), sphase, err, WD_G_MMA_FULL);
          tc_fence_after();
          const uint64_t a_hi = umma_desc_sw128(stage_addr(stage, 0)), a_lo = umma_desc_sw128(stage_addr(stage, 1));
          const uint64_t b_hi = umma_desc_sw128(stage_addr(stage, 2)), b_lo = umma_desc_sw128(stage_addr(stage, 3));
          if (elect_one()) {
            // +32 B (16 bf16) along K inside the swizzle row = +2 in the 16-byte address field
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
              umma_bf16_ss_pair(d_tmem, a_hi + 2u * k, b_hi + 2u * k, idesc, acc | k);
            }
            if (x3) {
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k) umma_bf16_ss_pair(d_tmem, a_hi + 2u * k, b_lo + 2u * k, idesc, 1u);
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k) umma_bf16_ss_pair(d_tmem, a_lo + 2u * k, b_hi + 2u * k, idesc, 1u);
            }
            umma_commit_pair(bar_empty(stage), static_cast<uint16_t>(0x3u));
          }
          __syncwarp();
          acc = 1;
          if (++stage == GEMM_NSTG) { stage = 0; sphase ^= 1u; }
        }
        if (elect_one()) umma_commit_pair(bar_acc_full(buf), static_cast<uint16_t>(0x3u));
        __syncwarp();
        ++cc;
      }
    }
  } else {
    // ================================ epilogue warps (both CTAs) ================================
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int sub = (warp - 2) >> 2;             // which of the quarter's warps: owns the column groups g = sub (mod 2)
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint32_t cc = 0;
    const float slope = P.slope ? __ldg(P.slope) : 0.0f;
    double dslope = 0.0;
    const int ngroups = P.BN >> 4;
    for (long long wk = my_cluster; wk < n_work; wk += n_clusters) {
      const int split = static_cast<int>(wk / tiles_mn);
      const long long tmn = wk - split * tiles_mn;
      const int mt = static_cast<int>(tmn / P.n_tiles), nt = static_cast<int>(tmn % P.n_tiles);   // n fastest: the clusters running together share A row blocks (L2 hits)
      const long long row = static_cast<long long>(mt) * 256 + rank * 128 + r;
      const bool row_ok = row < P.M;
      const uint32_t buf = cc & 1u;
      float* crow = P.C + split * P.slab_stride + row * P.ldc;
      const bool vec_c = ((P.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0) && ((P.slab_stride & 3) == 0);
      const float* brow = nullptr;
      if (P.bias_table && row_ok) brow = P.bias_table + P.bias_rows[row] * P.bias_ld;
      const float* arow = (P.aux && row_ok) ? P.aux + row * P.ld_aux : nullptr;
      float tile_ds = 0.0f;
      // PReLU' inputs of this warp's first group: they do not depend on the accumulator, so they are in flight during the wait
      const bool aux_vec = arow && ((P.ld_aux & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.aux) & 15) == 0);
      float an[16];
      auto load_aux = [&](int g) {
        const int c0 = nt * P.BN + g * 16;
        if (!arow || c0 >= P.N) return;
        if (aux_vec && c0 + 16 <= P.N) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 q4 = __ldcs(reinterpret_cast<const float4*>(arow + c0) + j);
            an[4 * j] = q4.x; an[4 * j + 1] = q4.y; an[4 * j + 2] = q4.z; an[4 * j + 3] = q4.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) an[e] = (c0 + e < P.N) ? arow[c0 + e] : 1.0f;
        }
      };
      load_aux(sub);
      mbar_wait(bar_acc_full(buf), (cc >> 1) & 1u, err, WD_G_EPI);
      tc_fence_after();
      const uint32_t t_tile = tmem_base + lane_addr + buf * 256u;
      uint32_t v[16];
      if (sub < ngroups) tmem_ld16(t_tile + sub * 16u, v);
      for (int g = sub; g < ngroups; g += GEMM_EPI_WARPS / 4) {
        const int col0 = nt * P.BN + g * 16;
        tmem_ld_wait();
        float h[16], a16[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) { h[e] = __uint_as_float(v[e]); a16[e] = an[e]; }
        if (g + GEMM_EPI_WARPS / 4 < ngroups) {
          tmem_ld16(t_tile + (g + GEMM_EPI_WARPS / 4) * 16u, v);
          load_aux(g + GEMM_EPI_WARPS / 4);
        }
        if (!row_ok || col0 >= P.N) continue;
        const bool full = col0 + 16 <= P.N;
        if (P.bias || brow) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            if (full || col0 + e < P.N) h[e] += brow ? __ldg(brow + col0 + e) : __ldg(P.bias + col0 + e);
          }
        }
        float o[16];   // what goes to the bf16 operand image
        if (P.epi == GEMM_EPI_PRELU_SPLIT) {
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = h[e] > 0.0f ? h[e] : slope * h[e];
        } else if (P.epi == GEMM_EPI_TANH) {
#pragma unroll
          for (int e = 0; e < 16; ++e) h[e] = tanhf(h[e]);
        } else if (P.epi == GEMM_EPI_DPRELU) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float a = (full || col0 + e < P.N) ? a16[e] : 1.0f;
            tile_ds += h[e] * fminf(a, 0.0f);
            h[e] = a > 0.0f ? h[e] : slope * h[e];
            o[e] = h[e];
          }
        }
        if (full && vec_c) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(crow + col0 + 4 * j) = make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (col0 + e < P.N) crow[col0 + e] = h[e];
        }
        if (P.o_hi && (P.epi == GEMM_EPI_PRELU_SPLIT || P.epi == GEMM_EPI_DPRELU)) {
          uint32_t ph[8], pl[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float x0 = (full || col0 + 2 * e < P.N) ? o[2 * e] : 0.0f, x1 = (full || col0 + 2 * e + 1 < P.N) ? o[2 * e + 1] : 0.0f;
            const float h0 = bf16_round(x0), h1 = bf16_round(x1);
            ph[e] = pack_bf16x2(h0, h1);
            pl[e] = pack_bf16x2(x0 - h0, x1 - h1);
          }
          // ldo is a multiple of 8 and col0 of 16: 16-byte stores; columns N..ldo of the last group are written as zeros
          if (col0 + 16 <= P.ldo) {
            uint4* dh = reinterpret_cast<uint4*>(P.o_hi + row * P.ldo + col0);
            uint4* dl = reinterpret_cast<uint4*>(P.o_lo + row * P.ldo + col0);
            dh[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]); dh[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
            dl[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]); dl[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
          } else {
            for (int e = 0; e < 16 && col0 + e < P.ldo; ++e) {
              reinterpret_cast<uint16_t*>(P.o_hi)[row * P.ldo + col0 + e] = static_cast<uint16_t>(ph[e >> 1] >> (16 * (e & 1)));
              reinterpret_cast<uint16_t*>(P.o_lo)[row * P.ldo + col0 + e] = static_cast<uint16_t>(pl[e >> 1] >> (16 * (e & 1)));
            }
          }
        }
      }
      dslope += static_cast<double>(tile_ds);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_cluster(bar_acc_empty(buf), 0));
      ++cc;
    }
    if (P.epi == GEMM_EPI_DPRELU && P.slope_grad) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dslope += __shfl_xor_sync(0xffffffffu, dslope, o);
      if (lane == 0 && dslope != 0.0) atomicAdd(P.slope_grad, dslope);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // nobody exits while the peer may still signal its barriers / read its operands
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace sdrm
