"""Recall@k / NDCG@k with the reference's signatures (utilities.py:116-171), computed by the K3 warp
top-k kernel.  Deterministic top-k (descending score, lower index wins ties) replaces
bottleneck.argpartition, whose tie order is implementation-defined; results are identical whenever the
k-th and (k+1)-th scores differ, and Recall is formed from integer counts so it is bit-exact.
"""
import numpy as np
import torch

from . import _lib

MAX_K = 64


def _dev():
    if not torch.cuda.is_available():
        raise _lib.SdrmError("SDRM metrics need a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _topk_pass(scores, k, want_vals):
    rows, n_items = scores.shape
    lib = _lib.load()
    idx = torch.empty((rows, k), dtype=torch.int32, device=scores.device)
    vals = torch.empty((rows, k), dtype=scores.dtype, device=scores.device) if want_vals else None
    if scores.dtype == torch.float32:
        fn = lib.sdrm_topk
    elif scores.dtype == torch.float64:
        fn = lib.sdrm_topk_f64
    else:
        raise TypeError("scores must be float32 or float64")
    _lib.check(fn(_lib.ptr(scores), rows, n_items, scores.stride(0), k, _lib.ptr(idx), _lib.ptr(vals),
                  _lib.stream_ptr()), "sdrm_topk")
    return idx, vals


def topk_device(scores, k, return_values=False):
    """scores: CUDA float32/float64 [rows, I] (row stride arbitrary) -> int32 [rows, k] sorted indices.

    k <= 64 is one pass of the warp kernel.  Larger k (the reference wrappers default to k=100, utilities.py:123,149) takes
    ceil(k / 64) passes over a working copy in which the entries already emitted are lowered to -inf; the concatenation is
    the same descending / lower-index-first order.  (Rows with fewer than k entries above -inf list -inf entries at the end,
    whose order is arbitrary in the reference's argpartition too.)"""
    if k < 1:
        raise ValueError("k must be >= 1")
    if scores.dim() != 2 or scores.stride(1) != 1:
        scores = scores.contiguous()
    rows, n_items = scores.shape
    if k > n_items:
        raise ValueError("k larger than the number of items")
    if k <= MAX_K:
        idx, vals = _topk_pass(scores, k, return_values)
        return (idx, vals) if return_values else idx
    work = scores.clone()
    idx_parts, val_parts, left = [], [], k
    while left > 0:
        kk = min(MAX_K, left)
        idx, vals = _topk_pass(work, kk, return_values)
        idx_parts.append(idx)
        val_parts.append(vals)
        left -= kk
        if left > 0:
            work.scatter_(1, idx.long(), float("-inf"))
    idx = torch.cat(idx_parts, dim=1)
    return (idx, torch.cat(val_parts, dim=1)) if return_values else idx


def _counters(X_pred, heldout, k):
    idx = topk_device(X_pred, k)
    rows, n_items = heldout.shape
    if heldout.dtype != torch.float32 or heldout.stride(1) != 1:
        heldout = heldout.to(torch.float32).contiguous()
    hits = torch.empty(rows, dtype=torch.int32, device=heldout.device)
    nrel = torch.empty(rows, dtype=torch.int32, device=heldout.device)
    dcg = torch.empty(rows, dtype=torch.float64, device=heldout.device)
    lib = _lib.load()
    _lib.check(lib.sdrm_recall_ndcg_at_k(_lib.ptr(idx), k, k, _lib.ptr(heldout), rows, n_items, heldout.stride(0),
                                         _lib.ptr(hits), _lib.ptr(nrel), _lib.ptr(dcg), _lib.stream_ptr()),
               "sdrm_recall_ndcg_at_k")
    return hits.cpu().numpy(), nrel.cpu().numpy(), dcg.cpu().numpy()


def recall_ndcg_multi_k(X_pred, heldout_batch, ks):
    """Recall@k and NDCG@k for every k in `ks` from ONE pass over the score matrix: the sorted top-max(ks)
    list of K3 contains the top-k list of every smaller k, so the six evaluator cut-offs (svd_benchmark.py:57-68) cost one
    read of the scores instead of twelve.  Returns {k: (recall[rows], ndcg[rows])}, values identical to the per-k calls."""
    held, stored = _to_device_heldout(heldout_batch)
    scores = _to_device_scores(X_pred)
    kmax = max(ks)
    idx = topk_device(scores, kmax)
    rows, n_items = held.shape
    lib = _lib.load()
    out = {}
    for k in ks:
        hits = torch.empty(rows, dtype=torch.int32, device=held.device)
        nrel = torch.empty(rows, dtype=torch.int32, device=held.device)
        dcg = torch.empty(rows, dtype=torch.float64, device=held.device)
        _lib.check(lib.sdrm_recall_ndcg_at_k(_lib.ptr(idx), kmax, k, _lib.ptr(held), rows, n_items, held.stride(0),
                                             _lib.ptr(hits), _lib.ptr(nrel), _lib.ptr(dcg), _lib.stream_ptr()),
                   "sdrm_recall_ndcg_at_k")
        hits, nrel, dcg = hits.cpu().numpy(), nrel.cpu().numpy(), dcg.cpu().numpy()
        tp = 1.0 / np.log2(np.arange(2, k + 2))
        cnt = nrel if stored is None else np.asarray(stored)
        idcg = np.array([tp[: min(int(n), k)].sum() for n in cnt])
        with np.errstate(invalid="ignore", divide="ignore"):
            out[k] = (hits.astype(np.float32) / np.minimum(k, nrel.astype(np.int64)), dcg / idcg)
    return out


def recall_at_k_device(X_pred, heldout, k):
    """Both arguments are CUDA tensors [rows, I]; returns float64 ndarray[rows] (0/0 -> NaN like the reference)."""
    hits, nrel, _ = _counters(X_pred, heldout, k)
    with np.errstate(invalid="ignore", divide="ignore"):
        return hits.astype(np.float32) / np.minimum(k, nrel.astype(np.int64))


def ndcg_at_k_device(X_pred, heldout, k, n_stored=None):
    """n_stored: per-row count of STORED held-out entries (csr getnnz, utilities.py:141); default = non-zeros."""
    _, nrel, dcg = _counters(X_pred, heldout, k)
    tp = 1.0 / np.log2(np.arange(2, k + 2))
    cnt = nrel if n_stored is None else np.asarray(n_stored)
    idcg = np.array([tp[: min(int(n), k)].sum() for n in cnt])
    with np.errstate(invalid="ignore", divide="ignore"):
        return dcg / idcg


def _to_device_scores(X_pred):
    dev = _dev()
    if isinstance(X_pred, torch.Tensor):
        return X_pred.to(dev)
    X_pred = np.asarray(X_pred)
    if X_pred.dtype not in (np.float32, np.float64):
        X_pred = X_pred.astype(np.float64)
    return torch.from_numpy(np.ascontiguousarray(X_pred)).to(dev)


def _to_device_heldout(heldout_batch):
    dev = _dev()
    if isinstance(heldout_batch, torch.Tensor):
        return heldout_batch.to(dev, torch.float32), None
    if isinstance(heldout_batch, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(heldout_batch, dtype=np.float32)).to(dev), None
    stored = heldout_batch.getnnz(axis=1)  # scipy sparse
    return torch.from_numpy(np.asarray(heldout_batch.toarray(), dtype=np.float32)).to(dev), stored


def recall_at_k_batch(X_pred, heldout_batch, k=100):
    """Reference signature (utilities.py:149-171): ndarray/csr in, float64 ndarray[rows] out."""
    held, _ = _to_device_heldout(heldout_batch)
    return recall_at_k_device(_to_device_scores(X_pred), held, k)


def NDCG_binary_at_k_batch(X_pred, heldout_batch, k=100):
    """Reference signature (utilities.py:123-146)."""
    held, stored = _to_device_heldout(heldout_batch)
    return ndcg_at_k_device(_to_device_scores(X_pred), held, k, n_stored=stored)


def mask_training_examples(sparse_training_set, dense_matrix):
    """Set already-seen items to -inf (utilities.py:116-120); in place on ndarray, like the reference."""
    dense_matrix[sparse_training_set.nonzero()] = -np.inf
    return dense_matrix
