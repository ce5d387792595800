// K3 — warp-level top-k and Recall/NDCG counters (reference: utilities.recall_at_k_batch,
// utilities.py:149-171; NDCG_binary_at_k_batch 123-146; mask_training_examples 116-120).
//
// One warp owns one score row and reads it exactly once (HBM-bound: 4*I bytes per row).  The running
// top-k (k <= 64) is a sorted list distributed over the warp (entry e lives in lane e%32, slot e/32);
// an element is inserted only if it beats the current k-th entry, so after a short warm-up almost
// every 32-element load costs one compare + one ballot.
// Order: descending score, ties -> lower item index first (== stable argsort of -score), NaN never wins.
// Three kernels: topk_rows_kernel (list insertion; fp64, k <= 16), topk_pool_kernel (candidate pool + bitonic merge on
// (value, index) pairs; fp64, k > 16) and topk_pool_f32_kernel (the same on 64-bit keys: every fp32 call).
#include <cuda_runtime.h>
#include <float.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/sdrm_b200.h"
#include "host_util.h"

namespace sdrm {

template <typename T>
__device__ __forceinline__ bool before(T av, int ai, T bv, int bi) {
  // true when (av, ai) ranks strictly ahead of (bv, bi)
  return (av > bv) || (av == bv && ai < bi);
}

// VEC: the row start is 16-byte aligned and ld * sizeof(T) is a multiple of 16: a lane reads two 16-byte vectors per
// iteration (1 KB in flight per warp) instead of four scalars -- the scalar form reached 0.56 of the measured HBM peak at
// k = 10 (65 536 x 20 000 fp32) because of the load count, not the insertions.  Element order inside an iteration does not
// matter: every comparison carries the item index, so ties still resolve to the lower index.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) topk_rows_kernel(const T* __restrict__ scores, long long rows, int n_items,
                                                        long long ld, int k, int* __restrict__ idx_out,
                                                        T* __restrict__ val_out) {
  const T NEG_INF = static_cast<T>(-CUDART_INF_F);
  constexpr int VE = VEC ? static_cast<int>(16 / sizeof(T)) : 1;   // elements per load
  constexpr int NV = VEC ? 2 : 4;                                  // loads in flight per lane
  constexpr int SPAN = 32 * VE * NV;                               // elements per warp iteration
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* x = scores + row * ld;

  T v0 = NEG_INF, v1 = NEG_INF;  // entries lane and lane+32
  int i0 = INT_MAX, i1 = INT_MAX;
  const int kth_lane = (k - 1) & 31;
  const bool kth_hi = (k - 1) >= 32;
  T thr_v = NEG_INF;
  int thr_i = INT_MAX;

  for (int base = 0; base < n_items; base += SPAN) {
    T c[NV * VE];
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int j0 = base + (u * 32 + lane) * VE;
      if (VEC && j0 + VE <= n_items) {
        if (sizeof(T) == 4) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(x + j0));
          c[u * VE + 0] = static_cast<T>(q.x); c[u * VE + (VE > 1 ? 1 : 0)] = static_cast<T>(q.y);
          c[u * VE + (VE > 2 ? 2 : 0)] = static_cast<T>(q.z); c[u * VE + (VE > 3 ? 3 : 0)] = static_cast<T>(q.w);
        } else {
          const double2 q = __ldg(reinterpret_cast<const double2*>(x + j0));
          c[u * VE + 0] = static_cast<T>(q.x); c[u * VE + (VE > 1 ? 1 : 0)] = static_cast<T>(q.y);
        }
      } else {
#pragma unroll
        for (int e = 0; e < VE; ++e) c[u * VE + e] = (j0 + e < n_items) ? __ldg(x + j0 + e) : NEG_INF;
      }
    }
    // Quick reject: after the first few hundred items almost no element reaches the k-th score, so the whole iteration costs
    // one comparison per element and ONE ballot (was a ballot per 32 elements: the scan, not the insertions, bounded the kernel).
    // `!(c < thr)` keeps equal scores and NaN (= -inf below) for the exact index-aware test.
    bool any = false;
#pragma unroll
    for (int q = 0; q < NV * VE; ++q) any |= !(c[q] < thr_v);
    if (!__any_sync(0xffffffffu, any)) continue;
#pragma unroll
    for (int q = 0; q < NV * VE; ++q) c[q] = (c[q] != c[q]) ? NEG_INF : c[q];  // NaN -> -inf
#pragma unroll
    for (int u = 0; u < NV; ++u) {
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        const T cu = c[u * VE + e];
        const int j = base + (u * 32 + lane) * VE + e;
        bool cand = (j < n_items) && before(cu, j, thr_v, thr_i);
        unsigned m = __ballot_sync(0xffffffffu, cand);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const T cv = __shfl_sync(0xffffffffu, cu, src);
          const int ci = base + (u * 32 + src) * VE + e;
          if (!before(cv, ci, thr_v, thr_i)) continue;  // threshold moved since the ballot
          const int pos = __popc(__ballot_sync(0xffffffffu, before(v0, i0, cv, ci))) +
                          __popc(__ballot_sync(0xffffffffu, before(v1, i1, cv, ci)));
          const T up_v0 = __shfl_up_sync(0xffffffffu, v0, 1);
          const int up_i0 = __shfl_up_sync(0xffffffffu, i0, 1);
          T up_v1 = __shfl_up_sync(0xffffffffu, v1, 1);
          int up_i1 = __shfl_up_sync(0xffffffffu, i1, 1);
          const T wrap_v = __shfl_sync(0xffffffffu, v0, 31);
          const int wrap_i = __shfl_sync(0xffffffffu, i0, 31);
          if (lane == 0) { up_v1 = wrap_v; up_i1 = wrap_i; }
          if (lane == pos) { v0 = cv; i0 = ci; }
          else if (lane > pos) { v0 = up_v0; i0 = up_i0; }
          if (lane + 32 == pos) { v1 = cv; i1 = ci; }
          else if (lane + 32 > pos) { v1 = up_v1; i1 = up_i1; }
          thr_v = __shfl_sync(0xffffffffu, kth_hi ? v1 : v0, kth_lane);
          thr_i = __shfl_sync(0xffffffffu, kth_hi ? i1 : i0, kth_lane);
        }
      }
    }
  }
  if (lane < k) {
    idx_out[row * k + lane] = i0;
    if (val_out) val_out[row * k + lane] = v0;
  }
  if (lane + 32 < k) {
    idx_out[row * k + lane + 32] = i1;
    if (val_out) val_out[row * k + lane + 32] = v1;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k > 16: buffered variant.  A streaming top-k sees ~k (1 + ln(I / k)) elements that beat the running k-th score (350 at
// k = 50, I = 20 000) and the kernel above pays a serial ~45-instruction list insertion for each: issue-bound at 0.35 of the
// HBM peak.  Here a candidate is only APPENDED to a 64-entry pool in shared memory (one ballot + one store); a full pool is
// sorted by a bitonic network in registers (2 entries per lane) and merged with the sorted list in one step
// (max(L_desc[p], C_asc[p]) is a bitonic sequence holding the 64 best of both), which moves the threshold.  ~8 flushes of
// ~450 instructions per row instead of 350 insertions.  Same total order as above (descending score, lower index first).
template <typename T>
__device__ __forceinline__ void cx_lane(T& v, int& i, int d, bool keep_better) {
  const T ov = __shfl_xor_sync(0xffffffffu, v, d);
  const int oi = __shfl_xor_sync(0xffffffffu, i, d);
  if (before(ov, oi, v, i) == keep_better) { v = ov; i = oi; }
}
template <typename T>
__device__ __forceinline__ void cx_slots(T& v0, int& i0, T& v1, int& i1, bool slot0_keeps_better) {
  if (before(v1, i1, v0, i0) == slot0_keeps_better) {
    const T tv = v0; v0 = v1; v1 = tv;
    const int ti = i0; i0 = i1; i1 = ti;
  }
}

// (not inlined: the scan loop calls it from every unrolled element position)
template <typename T>
__device__ __noinline__ void pool_flush(const T* pv, const int* pi, int npool, int lane, int kth_lane, bool kth_hi,
                                        T& v0, int& i0, T& v1, int& i1, T& thr_v, int& thr_i) {
  const T NEG_INF = static_cast<T>(-CUDART_INF_F);
    __syncwarp();
    T c0 = (lane < npool) ? pv[lane] : NEG_INF;
    int j0 = (lane < npool) ? pi[lane] : INT_MAX;
    T c1 = (lane + 32 < npool) ? pv[lane + 32] : NEG_INF;
    int j1 = (lane + 32 < npool) ? pi[lane + 32] : INT_MAX;
    __syncwarp();
    // bitonic sort of the 64 pool entries, ASCENDING (worst first); position p = lane + 32 * slot
#pragma unroll
    for (int s = 2; s <= 32; s <<= 1) {
#pragma unroll
      for (int d = s >> 1; d >= 1; d >>= 1) {
        const bool lower = (lane & d) == 0;
        const bool asc0 = (s == 32) ? true : ((lane & s) == 0);    // slot 0: positions 0..31
        const bool asc1 = (s == 32) ? false : ((lane & s) == 0);   // slot 1: positions 32..63 (bit 5 set)
        cx_lane(c0, j0, d, asc0 != lower);
        cx_lane(c1, j1, d, asc1 != lower);
      }
    }
    cx_slots(c0, j0, c1, j1, false);   // s = 64, d = 32: ascending, the lower position keeps the worse entry
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const bool lower = (lane & d) == 0;
      cx_lane(c0, j0, d, !lower);
      cx_lane(c1, j1, d, !lower);
    }
    // the 64 best of list + pool, as a bitonic sequence; then a descending bitonic merge
    if (before(c0, j0, v0, i0)) { v0 = c0; i0 = j0; }
    if (before(c1, j1, v1, i1)) { v1 = c1; i1 = j1; }
    cx_slots(v0, i0, v1, i1, true);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const bool lower = (lane & d) == 0;
      cx_lane(v0, i0, d, lower);
      cx_lane(v1, i1, d, lower);
    }
    thr_v = __shfl_sync(0xffffffffu, kth_hi ? v1 : v0, kth_lane);
    thr_i = __shfl_sync(0xffffffffu, kth_hi ? i1 : i0, kth_lane);
}


template <typename T, bool VEC>
__global__ void __launch_bounds__(256) topk_pool_kernel(const T* __restrict__ scores, long long rows, int n_items,
                                                        long long ld, int k, int* __restrict__ idx_out,
                                                        T* __restrict__ val_out) {
  const T NEG_INF = static_cast<T>(-CUDART_INF_F);
  constexpr int VE = VEC ? static_cast<int>(16 / sizeof(T)) : 1;
  constexpr int NV = VEC ? 2 : 4;
  constexpr int NE = NV * VE;                                      // elements per lane and iteration
  constexpr int SPAN = 32 * NE;
  constexpr int POOL = 64;
  __shared__ T pool_v[8][POOL];
  __shared__ int pool_i[8][POOL];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + wib;
  if (row >= rows) return;
  const T* x = scores + row * ld;
  T* pv = pool_v[wib];
  int* pi = pool_i[wib];
  const unsigned lt_mask = (1u << lane) - 1u;

  T v0 = NEG_INF, v1 = NEG_INF;  // sorted list, entries lane and lane + 32
  int i0 = INT_MAX, i1 = INT_MAX;
  const int kth_lane = (k - 1) & 31;
  const bool kth_hi = (k - 1) >= 32;
  T thr_v = NEG_INF;
  int thr_i = INT_MAX;
  int npool = 0;

  for (int base = 0; base < n_items; base += SPAN) {
    T c[NE];
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int j0 = base + (u * 32 + lane) * VE;
      if (VEC && j0 + VE <= n_items) {
        if (sizeof(T) == 4) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(x + j0));
          c[u * VE + 0] = static_cast<T>(q.x); c[u * VE + (VE > 1 ? 1 : 0)] = static_cast<T>(q.y);
          c[u * VE + (VE > 2 ? 2 : 0)] = static_cast<T>(q.z); c[u * VE + (VE > 3 ? 3 : 0)] = static_cast<T>(q.w);
        } else {
          const double2 q = __ldg(reinterpret_cast<const double2*>(x + j0));
          c[u * VE + 0] = static_cast<T>(q.x); c[u * VE + (VE > 1 ? 1 : 0)] = static_cast<T>(q.y);
        }
      } else {
#pragma unroll
        for (int e = 0; e < VE; ++e) c[u * VE + e] = (j0 + e < n_items) ? __ldg(x + j0 + e) : NEG_INF;
      }
    }
    bool any = false;
#pragma unroll
    for (int q = 0; q < NE; ++q) any |= !(c[q] < thr_v);   // keeps equal scores and NaN for the exact test
    if (!__any_sync(0xffffffffu, any)) continue;
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const T cq = (c[q] != c[q]) ? NEG_INF : c[q];   // NaN -> -inf
      const int j = base + ((q / VE) * 32 + lane) * VE + (q % VE);
      bool cand = (j < n_items) && before(cq, j, thr_v, thr_i);
      unsigned m = __ballot_sync(0xffffffffu, cand);
      if (m == 0) continue;
      if (npool + __popc(m) > POOL) {
        pool_flush<T>(pv, pi, npool, lane, kth_lane, kth_hi, v0, i0, v1, i1, thr_v, thr_i);
        npool = 0;
        cand = cand && before(cq, j, thr_v, thr_i);
        m = __ballot_sync(0xffffffffu, cand);
      }
      if (cand) {
        const int slot = npool + __popc(m & lt_mask);
        pv[slot] = cq;
        pi[slot] = j;
      }
      npool += __popc(m);
    }
  }
  if (npool > 0) pool_flush<T>(pv, pi, npool, lane, kth_lane, kth_hi, v0, i0, v1, i1, thr_v, thr_i);
  if (lane < k) {
    idx_out[row * k + lane] = i0;
    if (val_out) val_out[row * k + lane] = v0;
  }
  if (lane + 32 < k) {
    idx_out[row * k + lane + 32] = i1;
    if (val_out) val_out[row * k + lane + 32] = v1;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// fp32 scores, k > 16 (the evaluators' k = 20 / 50 and the multi-k pass): the same pool idea on 64-bit keys.
//   key = ord(score) << 32 | (0xFFFFFFFF - item): one unsigned 64-bit comparison IS the total order (descending score, lower index
//   first; -0.0 is folded into +0.0 and NaN into -inf first), so a compare-exchange of the bitonic networks is 2 shuffles + a
//   64-bit max/min.  Candidates (score not below the running k-th score, then the exact key test) take a slot of the warp's
//   320-entry pool with one shared-memory atomic; after an iteration that leaves >= 64 entries the pool is merged into the sorted
//   64-entry list (2 keys per lane) in chunks of 64 and the threshold moves.  The next iteration's loads are issued before the
//   current one is examined.  Measured (65 536 x 20 000, B200): see profiles/k3_topk_r02.txt.
__device__ __forceinline__ uint32_t ord_f32(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord_f32(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ void cxk(unsigned long long& k, int d, bool keep_larger) {
  const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, d);
  if ((o > k) == keep_larger) k = o;
}

constexpr int POOL_CAP = 320;   // < 64 entries left by a flush + one iteration of 256 candidates

// Merges the 64-entry chunk pool[off .. off + 64) (entries at or past n read as key 0) into the sorted list.  Position p of a
// 64-entry sequence lives in lane p / 2, register p % 2, so the distance-1 exchanges of the networks (6 of the 21 sorting steps)
// stay inside a lane: 15 + 5 shuffle steps per merge instead of 20 + 5 (the shuffle pipe is the scarce resource here).
__device__ __forceinline__ void pool_merge_f32(const unsigned long long* pool, int off, int n, int lane, unsigned long long& k0, unsigned long long& k1) {
  unsigned long long c0 = (off + 2 * lane < n) ? pool[off + 2 * lane] : 0ull;
  unsigned long long c1 = (off + 2 * lane + 1 < n) ? pool[off + 2 * lane + 1] : 0ull;
  // bitonic sort of the chunk, ASCENDING
#pragma unroll
  for (int s = 2; s <= 64; s <<= 1) {
    const bool asc = (s == 64) ? true : ((lane & (s >> 1)) == 0);
#pragma unroll
    for (int d = s >> 1; d >= 2; d >>= 1) {
      const bool lower = (lane & (d >> 1)) == 0;
      cxk(c0, d >> 1, asc != lower);
      cxk(c1, d >> 1, asc != lower);
    }
    if ((c0 > c1) == asc) { const unsigned long long t = c0; c0 = c1; c1 = t; }   // d = 1
  }
  // max(L_desc[p], C_asc[p]) holds the 64 best of both as a bitonic sequence; descending bitonic merge
  k0 = (c0 > k0) ? c0 : k0;
  k1 = (c1 > k1) ? c1 : k1;
#pragma unroll
  for (int d = 32; d >= 2; d >>= 1) {
    const bool lower = (lane & (d >> 1)) == 0;
    cxk(k0, d >> 1, lower);
    cxk(k1, d >> 1, lower);
  }
  if (k1 > k0) { const unsigned long long t = k0; k0 = k1; k1 = t; }
}

template <bool VEC>
__global__ void __launch_bounds__(256) topk_pool_f32_kernel(const float* __restrict__ scores, long long rows, int n_items,
                                                            long long ld, int k, int* __restrict__ idx_out,
                                                            float* __restrict__ val_out) {
  constexpr int VE = VEC ? 4 : 1;
  constexpr int NV = VEC ? 2 : 8;
  constexpr int NE = NV * VE;                                      // 8 elements per lane and iteration
  constexpr int SPAN = 32 * NE;
  __shared__ __align__(16) unsigned long long pool_s[8][POOL_CAP];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long row_raw = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + wib;
  const bool row_ok = row_raw < rows;
  const long long row = row_ok ? row_raw : rows - 1;   // (no early exit: a surplus warp repeats the last row and stores nothing)
  const float* x = scores + row * ld;
  unsigned long long* pool = pool_s[wib];
  const unsigned lt_mask = (1u << lane) - 1u;

  unsigned long long k0 = 0ull, k1 = 0ull, thr_key = 0ull;   // list entries 2 * lane and 2 * lane + 1, descending
  float thr_v = -CUDART_INF_F;
  const int kth_lane = (k - 1) >> 1;
  const bool kth_odd = ((k - 1) & 1) != 0;
  int npool = 0;   // warp-uniform (built from ballots only: the flush branch and the merge loop stay provably uniform, so ptxas
                   // does not wrap the shuffles of the networks in WARPSYNC.COLLECTIVE / ENDCOLLECTIVE)

  auto load = [&](int base, float (&c)[NE]) {
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int j0 = base + (u * 32 + lane) * VE;
      if (VEC && j0 + VE <= n_items) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(x + j0));
        c[u * VE + 0] = q.x; c[u * VE + (VE > 1 ? 1 : 0)] = q.y; c[u * VE + (VE > 2 ? 2 : 0)] = q.z; c[u * VE + (VE > 3 ? 3 : 0)] = q.w;
      } else {
#pragma unroll
        for (int e = 0; e < VE; ++e) c[u * VE + e] = (j0 + e < n_items) ? __ldg(x + j0 + e) : -CUDART_INF_F;   // (j >= n_items never enters the pool)
      }
    }
  };
  auto new_threshold = [&]() {
    thr_key = __shfl_sync(0xffffffffu, kth_odd ? k1 : k0, kth_lane);
    const uint32_t hi = static_cast<uint32_t>(thr_key >> 32);
    thr_v = hi ? unord_f32(hi) : -CUDART_INF_F;
  };
  // merge the full 64-entry chunks; the (< 64) entries behind them that still beat the new threshold move to the pool's front
  auto flush = [&]() {
    __syncwarp();
    const int full = npool & ~63;
    for (int off = 0; off < full; off += 64) pool_merge_f32(pool, off, npool, lane, k0, k1);
    new_threshold();
    const int rem = npool - full;
    const unsigned long long r0 = (lane < rem) ? pool[full + lane] : 0ull;
    const unsigned long long r1 = (lane + 32 < rem) ? pool[full + 32 + lane] : 0ull;
    const unsigned m0 = __ballot_sync(0xffffffffu, r0 > thr_key), m1 = __ballot_sync(0xffffffffu, r1 > thr_key);   // (key 0 never passes)
    __syncwarp();
    if (r0 > thr_key) pool[__popc(m0 & lt_mask)] = r0;
    if (r1 > thr_key) pool[__popc(m0) + __popc(m1 & lt_mask)] = r1;
    npool = __popc(m0) + __popc(m1);
  };

  float c[NE], nx[NE];
  load(0, c);
  for (int base = 0; base < n_items; base += SPAN) {
    if (base + SPAN < n_items) load(base + SPAN, nx);
    bool p[NE];
    bool any = false;
#pragma unroll
    for (int q = 0; q < NE; ++q) { p[q] = !(c[q] < thr_v); any |= p[q]; }   // keeps equal scores and NaN for the exact key test
    if (__any_sync(0xffffffffu, any)) {
#pragma unroll
      for (int q = 0; q < NE; ++q) {
        if (!__any_sync(0xffffffffu, p[q])) continue;
        const int j = base + ((q / VE) * 32 + lane) * VE + (q % VE);
        float v = c[q] + 0.0f;                       // -0.0 -> +0.0: equal scores must have equal keys
        v = (v != v) ? -CUDART_INF_F : v;            // NaN -> -inf
        const unsigned long long key = (static_cast<unsigned long long>(ord_f32(v)) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(j));
        const bool cand = p[q] && j < n_items && key > thr_key;
        const unsigned m = __ballot_sync(0xffffffffu, cand);
        if (cand) pool[npool + __popc(m & lt_mask)] = key;
        npool += __popc(m);
      }
      if (npool >= 64) flush();
    }
#pragma unroll
    for (int q = 0; q < NE; ++q) c[q] = nx[q];
  }
  __syncwarp();
  for (int off = 0; off < npool; off += 64) pool_merge_f32(pool, off, npool, lane, k0, k1);
  if (row_ok && 2 * lane < k) {
    idx_out[row * k + 2 * lane] = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(k0));
    if (val_out) val_out[row * k + 2 * lane] = unord_f32(static_cast<uint32_t>(k0 >> 32));
  }
  if (row_ok && 2 * lane + 1 < k) {
    idx_out[row * k + 2 * lane + 1] = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(k1));
    if (val_out) val_out[row * k + 2 * lane + 1] = unord_f32(static_cast<uint32_t>(k1 >> 32));
  }
}

// per row: hits = #{e < k : heldout[row, idx[e]] > 0}, nrel = #{heldout[row, :] > 0},
// dcg = sum_e heldout[row, idx[e]] / log2(e + 2)   (fp64, sequential over e)
__global__ void __launch_bounds__(256) recall_ndcg_kernel(const int* __restrict__ topk, int k_stored, int k,
                                                          const float* __restrict__ heldout, long long rows,
                                                          int n_items, long long ld_h, int* __restrict__ hits_out,
                                                          int* __restrict__ nrel_out, double* __restrict__ dcg_out) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* hrow = heldout + row * ld_h;
  int nrel = 0;
  for (int j = lane; j < n_items; j += 32) nrel += (__ldg(hrow + j) > 0.0f) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nrel += __shfl_xor_sync(0xffffffffu, nrel, o);
  int hits = 0;
  double dcg = 0.0;
  if (lane == 0) {
    for (int e = 0; e < k; ++e) {
      const int it = topk[row * k_stored + e];
      const float hv = (it >= 0 && it < n_items) ? hrow[it] : 0.0f;
      hits += hv > 0.0f ? 1 : 0;
      dcg += static_cast<double>(hv) * (1.0 / log2(static_cast<double>(e + 2)));
    }
    if (hits_out) hits_out[row] = hits;
    if (nrel_out) nrel_out[row] = nrel;
    if (dcg_out) dcg_out[row] = dcg;
  }
}

}  // namespace sdrm

using namespace sdrm;

// smallest k that takes a pooled kernel is g_topk_pool_min_k + 1 (SDRM_TOPK_POOL_MIN_K in the environment: A/B measurements; -1 = default)
static int g_topk_pool_min_k = [] {
  const char* e = getenv("SDRM_TOPK_POOL_MIN_K");
  return (e && *e) ? atoi(e) : -1;
}();

static const bool g_topk_pool_generic = getenv("SDRM_TOPK_POOL_GENERIC") != nullptr;   // A/B: (value, index) pairs instead of 64-bit keys

template <typename T>
static int topk_impl(const T* d_scores, int64_t rows, int n_items, int64_t ld, int k, int32_t* d_idx_out, T* d_val_out,
                     void* stream) {
  if (!d_scores || !d_idx_out) return sdrm_fail(SDRM_ERR_BAD_ARG, "sdrm_topk: null pointer");
  if (k < 1 || k > 64) return sdrm_fail(SDRM_ERR_UNSUPPORTED, "sdrm_topk: k must be in [1, 64]");
  if (n_items < 1 || ld < n_items) return sdrm_fail(SDRM_ERR_BAD_ARG, "sdrm_topk: bad shape");
  if (rows <= 0) return SDRM_OK;
  const int warps = 8;
  const long long blocks = (rows + warps - 1) / warps;
  const bool vec = (reinterpret_cast<uintptr_t>(d_scores) % 16 == 0) && ((static_cast<size_t>(ld) * sizeof(T)) % 16 == 0);
  // fp32: the pooled 64-bit-key kernel for every k (measured faster than list insertion from k = 10 on: 0.97 vs 0.91 of the HBM
  // peak at k = 10, 0.78 vs 0.35 at k = 50); fp64: list insertion up to k = 16, (value, index) pool above
  const bool pooled = k > (g_topk_pool_min_k >= 0 ? g_topk_pool_min_k : (sizeof(T) == 4 ? 0 : 16));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned nb = static_cast<unsigned>(blocks);
  if (pooled && sizeof(T) == 4 && !g_topk_pool_generic) {
    const float* sc = reinterpret_cast<const float*>(d_scores);
    float* vo = reinterpret_cast<float*>(d_val_out);
    if (vec) topk_pool_f32_kernel<true><<<nb, warps * 32, 0, st>>>(sc, rows, n_items, ld, k, d_idx_out, vo);
    else topk_pool_f32_kernel<false><<<nb, warps * 32, 0, st>>>(sc, rows, n_items, ld, k, d_idx_out, vo);
  } else if (pooled && vec) topk_pool_kernel<T, true><<<nb, warps * 32, 0, st>>>(d_scores, rows, n_items, ld, k, d_idx_out, d_val_out);
  else if (pooled) topk_pool_kernel<T, false><<<nb, warps * 32, 0, st>>>(d_scores, rows, n_items, ld, k, d_idx_out, d_val_out);
  else if (vec) topk_rows_kernel<T, true><<<nb, warps * 32, 0, st>>>(d_scores, rows, n_items, ld, k, d_idx_out, d_val_out);
  else topk_rows_kernel<T, false><<<nb, warps * 32, 0, st>>>(d_scores, rows, n_items, ld, k, d_idx_out, d_val_out);
  SDRM_CUDA(cudaGetLastError());
  return SDRM_OK;
}

extern "C" {

int sdrm_topk(const float* d_scores, int64_t rows, int n_items, int64_t ld, int k, int32_t* d_idx_out,
              float* d_val_out, void* stream) {
  return topk_impl<float>(d_scores, rows, n_items, ld, k, d_idx_out, d_val_out, stream);
}

int sdrm_topk_f64(const double* d_scores, int64_t rows, int n_items, int64_t ld, int k, int32_t* d_idx_out,
                  double* d_val_out, void* stream) {
  return topk_impl<double>(d_scores, rows, n_items, ld, k, d_idx_out, d_val_out, stream);
}

int sdrm_recall_ndcg_at_k(const int32_t* d_topk_idx, int k_stored, int k, const float* d_heldout, int64_t rows,
                          int n_items, int64_t ld_h, int32_t* d_hits_out, int32_t* d_nrel_out, double* d_dcg_out,
                          void* stream) {
  if (!d_topk_idx || !d_heldout) return sdrm_fail(SDRM_ERR_BAD_ARG, "sdrm_recall_ndcg_at_k: null pointer");
  if (k < 1 || k > k_stored) return sdrm_fail(SDRM_ERR_BAD_ARG, "sdrm_recall_ndcg_at_k: k must be in [1, k_stored]");
  if (rows <= 0) return SDRM_OK;
  const int warps = 8;
  const long long blocks = (rows + warps - 1) / warps;
  recall_ndcg_kernel<<<static_cast<unsigned>(blocks), warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      d_topk_idx, k_stored, k, d_heldout, rows, n_items, ld_h, d_hits_out, d_nrel_out, d_dcg_out);
  SDRM_CUDA(cudaGetLastError());
  return SDRM_OK;
}

}  // extern "C"
