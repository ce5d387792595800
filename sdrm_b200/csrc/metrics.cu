// K3 — warp-level top-k and Recall/NDCG counters (reference: utilities.recall_at_k_batch,
// utilities.py:149-171; NDCG_binary_at_k_batch 123-146; mask_training_examples 116-120).
//
// One warp owns one score row and reads it exactly once (HBM-bound: 4*I bytes per row).  The running
// top-k (k <= 64) is a sorted list distributed over the warp (entry e lives in lane e%32, slot e/32);
// an element is inserted only if it beats the current k-th entry, so after a short warm-up almost
// every 32-element load costs one compare + one ballot.
// Order: descending score, ties -> lower item index first (== stable argsort of -score), NaN never wins.
#include <cuda_runtime.h>
#include <float.h>
#include <math_constants.h>
#include <stdint.h>

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
  if (vec)
    topk_rows_kernel<T, true><<<static_cast<unsigned>(blocks), warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        d_scores, rows, n_items, ld, k, d_idx_out, d_val_out);
  else
    topk_rows_kernel<T, false><<<static_cast<unsigned>(blocks), warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        d_scores, rows, n_items, ld, k, d_idx_out, d_val_out);
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
