"""CPU: host logic of the equal-sparsity thresholding against NumPy itself — the quantile rule
restatement, the radix-select walk over digit histograms and the key map; plus the world-size-2 histogram merge (gloo)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sparsify_oracle as so
from sdrm_b200 import sparsify as sp


def _cases():
    rng = np.random.RandomState(0)
    yield rng.randn(1).astype(np.float32)
    yield rng.randn(2).astype(np.float32)
    yield rng.randn(5).astype(np.float32)
    yield (rng.randn(843 * 1008) * 3 - 4).astype(np.float32)            # cfg 1 sized score matrix
    yield np.round(rng.randn(20000) * 2).astype(np.float32)             # heavy ties, +-0
    a = rng.randn(4096).astype(np.float32); a[::7] = 0.0; a[1::7] = -0.0; a[5] = np.inf; a[6] = -np.inf
    yield a


@pytest.mark.parametrize("q", [0.0, 1.0, 0.5, 0.9424673, 1 - 0.9424673, 0.0575, 0.999999, 1e-7])
def test_quantile_rule_matches_numpy(q):
    for a in _cases():
        s = np.sort(a)
        p, n, g = sp.quantile_plan(a.size, q, np.float32)
        got = sp.lerp(s[p], s[n], g)
        ref = np.quantile(a, q)
        assert got.dtype == ref.dtype and (got == ref or (np.isnan(got) and np.isnan(ref))), (a.size, q, got, ref)


def test_quantile_rule_large_n_float32_index():
    """Above 2^24 values NumPy's float32 virtual index is no longer exact; the restatement must follow it anyway."""
    rng = np.random.RandomState(1)
    a = rng.randn(20_000_003).astype(np.float32)
    s = np.sort(a)
    for q in (0.9424673, 0.987, 0.3333333):
        p, n, g = sp.quantile_plan(a.size, q, np.float32)
        assert sp.lerp(s[p], s[n], g) == np.quantile(a, q)


def test_key_map_roundtrip_and_order():
    a = np.array([-np.inf, -3.5, -1e-30, -0.0, 0.0, 1e-30, 2.0, np.inf], dtype=np.float32)
    k = so.score_keys(a)
    assert (np.diff(k.astype(np.int64)) > 0).all()
    for x, kk in zip(a, k):
        assert sp.float_to_key(x) == int(kk)
        back = sp.key_to_float(int(kk))
        assert back == x and np.signbit(back) == np.signbit(x)


def test_radix_select_matches_sort():
    for a in _cases():
        s = np.sort(a)
        ranks = sorted(set([0, a.size - 1, a.size // 2, min(a.size - 1, a.size // 2 + 1), (a.size * 9) // 10]))
        got = sp.select_ranks(so.numpy_histogram_fn(a), ranks)
        for r, g in zip(ranks, got):
            assert g == s[r], (a.size, r, g, s[r])
    with pytest.raises(ValueError):
        sp.select_ranks(so.numpy_histogram_fn(np.zeros(4, np.float32)), [4])


def test_pack_bits_layout():
    rng = np.random.RandomState(2)
    d = (rng.rand(5, 70) > 0.5).astype(np.uint8)
    w = so.pack_bits(d)
    assert w.shape == (5, 3)
    for r in range(5):
        for c in range(70):
            assert (int(w[r, c // 32]) >> (c % 32)) & 1 == d[r, c]
    pm = sp.PackedMatrix(torch.from_numpy(w.view(np.int32).copy()), 70, 0.0, None)
    assert np.array_equal(pm.numpy(np.uint8), d)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(3)
    full = (rng.randn(1001, 37) * 2 - 1).astype(np.float32)
    lo, hi = (0, 389) if rank == 0 else (389, 1001)          # ragged shards
    local = so.numpy_histogram_fn(full[lo:hi])

    def hist_fn(*a):   # what device_histogram_fn does after the kernel: sum the digit counts over the ranks
        h = torch.from_numpy(local(*a))
        dist.all_reduce(h)
        return h.numpy()

    n = torch.tensor([(hi - lo) * 37]); dist.all_reduce(n)
    res = []
    for qq in (0.9424673, 0.5, 0.0, 1.0):
        p, nx, g = sp.quantile_plan(int(n), qq, np.float32)
        v = sp.select_ranks(hist_fn, [p] if p == nx else [p, nx])
        res.append(float(sp.lerp(v[0], v[-1], g)) == float(np.quantile(full.flatten(), qq)))
    q.put((rank, res))
    dist.destroy_process_group()


def test_sharded_threshold_is_the_global_quantile():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    out = [q.get(timeout=120) for _ in ps]
    [p.join(60) for p in ps]
    assert all(all(r) for _, r in out), out
