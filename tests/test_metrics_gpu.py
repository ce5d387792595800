"""GPU parity of K3 (warp top-k, Recall/NDCG counters) against the oracle: indices and Recall bit-exact."""
import numpy as np
import pytest
import torch

from oracle import sdrm_oracle as orc

pytestmark = pytest.mark.gpu


def _scores(rows, items, seed, dtype=np.float32, ties=False):
    rng = np.random.RandomState(seed)
    x = rng.randn(rows, items).astype(dtype)
    if ties:
        x = np.round(x * 4) / 4          # heavy ties
    return x


@pytest.mark.parametrize("rows,items", [(95, 1008), (136, 729), (64, 3125), (40, 8582), (3, 20000), (1, 1), (7, 33)])
@pytest.mark.parametrize("k", [1, 3, 5, 10, 20, 50, 64])
def test_topk_indices_bit_exact(rows, items, k):
    from sdrm_b200 import metrics
    if k > items:
        pytest.skip("k > items")
    x = _scores(rows, items, seed=rows + k)
    idx, vals = metrics.topk_device(torch.from_numpy(x).cuda(), k, return_values=True)
    ref = orc.topk_oracle(x, k)
    assert np.array_equal(idx.cpu().numpy(), ref)
    assert np.array_equal(vals.cpu().numpy(), np.take_along_axis(x, ref, axis=1))


@pytest.mark.parametrize("k", [65, 100, 150])
def test_topk_above_64_multi_pass(k):
    """the reference wrappers default to k=100 (utilities.py:123,149): ceil(k/64) passes of the warp kernel"""
    from sdrm_b200 import metrics
    x = _scores(37, 1008, seed=k, ties=(k == 100))
    idx = metrics.topk_device(torch.from_numpy(x).cuda(), k).cpu().numpy()
    assert np.array_equal(idx, orc.topk_oracle(x, k))
    held = (np.random.RandomState(k).rand(37, 1008) < 0.05).astype(np.float32)
    r = metrics.recall_at_k_batch(torch.from_numpy(x).cuda(), held)          # default k = 100 like the reference
    ref_idx = orc.topk_oracle(x, 100)
    hits = np.take_along_axis(held, ref_idx, axis=1).sum(axis=1).astype(np.float32)
    assert np.array_equal(r, hits / np.minimum(100, held.sum(axis=1).astype(np.int64)))


def test_topk_ties_neg_inf_nan_and_strided_rows():
    from sdrm_b200 import metrics
    x = _scores(50, 400, seed=1, ties=True)
    x[:, 5:60] = -np.inf            # masked training items (utilities.mask_training_examples)
    x[3, :] = -np.inf               # a fully masked row: lowest indices win
    x[4, 7] = np.nan
    x[5, :] = 1.0                   # all equal
    big = torch.full((50, 512), 9.0)
    big[:, :400] = torch.from_numpy(x)
    view = big.cuda()[:, :400]      # row stride 512 != 400
    for k in (1, 10, 50):
        assert np.array_equal(metrics.topk_device(view, k).cpu().numpy(), orc.topk_oracle(x, k))


def test_topk_float64_scores():
    from sdrm_b200 import metrics
    x = _scores(20, 1008, seed=3, dtype=np.float64)
    x[:, 100] = x[:, 200] + 1e-12   # distinguishable only in float64
    got = metrics.topk_device(torch.from_numpy(x).cuda(), 10).cpu().numpy()
    order = np.argsort(-x, axis=1, kind="stable")[:, :10]
    assert np.array_equal(got, order)


@pytest.mark.parametrize("k", [17, 40, 64])
def test_topk_pooled_kernel_worst_cases(k):
    """k > 16 runs the pooled kernel (candidate pool + bitonic merge): ascending rows make EVERY element a candidate (a flush
    per 64 elements), descending rows none after the first flush; float64 scores; against the stable argsort."""
    from sdrm_b200 import metrics
    rng = np.random.RandomState(k)
    base = np.sort(rng.randn(6, 2500).astype(np.float32), axis=1)
    x = np.concatenate([base, base[:, ::-1], np.repeat(base[:, :1], 2500, axis=1)], axis=0).copy()
    x[1, 100:164] = x[1, 2400]                     # a run of ties inside an ascending row
    got = metrics.topk_device(torch.from_numpy(x).cuda(), k).cpu().numpy()
    assert np.array_equal(got, np.argsort(-x, axis=1, kind="stable")[:, :k])
    z = rng.randn(5, 900).astype(np.float32)       # signed zeros are EQUAL scores (ties -> lower index), denormals are not zero
    z[:, ::3] = 0.0
    z[:, 1::6] = -0.0
    z[:, 2::9] = np.float32(1e-42)
    z[:, 5::11] = np.float32(-1e-42)
    z[2] = np.where(np.arange(900) % 2 == 0, np.float32(-0.0), np.float32(0.0))
    got = metrics.topk_device(torch.from_numpy(z).cuda(), k).cpu().numpy()
    assert np.array_equal(got, np.argsort(-z, axis=1, kind="stable")[:, :k])
    xd = rng.randn(9, 1777)                        # float64, odd width (scalar tail loads)
    xd[:, 300] = xd[:, 900] + 1e-13
    got = metrics.topk_device(torch.from_numpy(xd).cuda(), k).cpu().numpy()
    assert np.array_equal(got, np.argsort(-xd, axis=1, kind="stable")[:, :k])


@pytest.mark.parametrize("k", [1, 3, 5, 10, 20, 50])
def test_recall_ndcg_match_reference_formulas(k):
    """recall_at_k_batch / NDCG_binary_at_k_batch with the reference's numpy formulas (utilities.py:123-171)."""
    from scipy.sparse import csr_matrix
    from sdrm_b200 import metrics
    rng = np.random.RandomState(k)
    x = rng.randn(120, 700).astype(np.float32)
    held = (rng.rand(120, 700) < 0.02).astype(np.float64)
    held[0] = 0                                            # 0 relevant items -> NaN like the reference
    seen = rng.rand(120, 700) < 0.05
    x[seen] = -np.inf
    rec = metrics.recall_at_k_batch(x, csr_matrix(held), k=k)
    ndcg = metrics.NDCG_binary_at_k_batch(x, csr_matrix(held), k=k)
    # reference formulas with np.argpartition (ties are measure-zero for continuous scores)
    part = np.argpartition(-x, k, axis=1)[:, :k]
    pb = np.zeros_like(x, dtype=bool)
    pb[np.arange(120)[:, None], part] = True
    tb = held > 0
    with np.errstate(invalid="ignore", divide="ignore"):
        ref_rec = np.logical_and(tb, pb).sum(axis=1).astype(np.float32) / np.minimum(k, tb.sum(axis=1))
    assert rec.dtype == ref_rec.dtype
    assert np.array_equal(np.isnan(rec), np.isnan(ref_rec))
    assert np.array_equal(rec[~np.isnan(rec)], ref_rec[~np.isnan(ref_rec)])        # bit-exact
    assert np.allclose(ndcg[~np.isnan(ndcg)], orc.ndcg_at_k_oracle(x, held, k)[~np.isnan(ndcg)], rtol=1e-12, atol=0)
    assert np.array_equal(rec[~np.isnan(rec)], orc.recall_at_k_oracle(x, held, k)[~np.isnan(rec)])


def test_topk_full_scale_properties():
    """BASELINE-sized rows (20 000 items): size-independent properties instead of an O(n log n) CPU check."""
    from sdrm_b200 import metrics
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(4096, 20000, device="cuda", generator=g)
    idx, vals = metrics.topk_device(x, 50, return_values=True)
    assert bool((vals[:, :-1] >= vals[:, 1:]).all())                       # sorted
    assert bool((torch.gather(x, 1, idx.long()) == vals).all())            # indices point at the values
    kth = vals[:, -1:]
    assert bool(((x > kth).sum(dim=1) == 49).all())                        # exactly k-1 elements beat the k-th
    torch_idx = torch.topk(x, 50, dim=1).indices
    assert bool((torch.sort(torch_idx, dim=1).values == torch.sort(idx.long(), dim=1).values).all())


def test_multi_k_single_pass_equals_per_k_calls():
    """One top-50 pass serves all six evaluator cut-offs: values identical to the per-k entry points."""
    from scipy.sparse import csr_matrix
    from sdrm_b200 import metrics
    rng = np.random.RandomState(9)
    x = np.round(rng.randn(95, 1008) * 8) / 8          # float64 SVD-like scores with ties
    x[rng.rand(95, 1008) < 0.05] = -np.inf
    held = csr_matrix((rng.rand(95, 1008) < 0.03).astype(np.float64))
    ks = [1, 3, 5, 10, 20, 50]
    both = metrics.recall_ndcg_multi_k(x, held, ks)
    for k in ks:
        r, n = metrics.recall_at_k_batch(x, held, k=k), metrics.NDCG_binary_at_k_batch(x, held, k=k)
        assert np.array_equal(both[k][0], r, equal_nan=True)
        assert np.array_equal(both[k][1], n, equal_nan=True)
