"""GPU parity of K4 (sdrm_key_histogram / sdrm_threshold_pack through the C ABI) with NumPy: the threshold is
np.quantile's bit for bit and the packed matrix equals the reference's `(S >= threshold).astype(int)`."""
import numpy as np
import pytest
import torch

from oracle import sparsify_oracle as so

pytestmark = pytest.mark.gpu


def _mats():
    rng = np.random.RandomState(0)
    yield "cfg1", (rng.randn(843, 1008) * 3 - 4).astype(np.float32)
    yield "ragged-cols", rng.randn(300, 729).astype(np.float32)          # 729 % 4 != 0 -> scalar path
    yield "one-row", rng.randn(1, 17).astype(np.float32)
    yield "single", rng.randn(1, 1).astype(np.float32)
    yield "ties", np.round(rng.randn(257, 96) * 2).astype(np.float32)
    a = rng.randn(64, 128).astype(np.float32); a[::3, ::5] = 0.0; a[1::3, ::5] = -0.0; a[0, 0] = np.inf; a[1, 1] = -np.inf
    yield "specials", a


@pytest.mark.parametrize("name,a", list(_mats()), ids=[n for n, _ in _mats()])
@pytest.mark.parametrize("sparsity", [0.9424673, 0.5, 0.0133, 0.99885, 0.0, 1.0])
def test_equal_sparsity_matches_reference_lines(name, a, sparsity):
    from sdrm_b200.sparsify import equal_sparsity_device, quantile_device
    x = torch.from_numpy(a).cuda()
    for lower in (False, True):
        ref_bin, ref_thr = so.equal_sparsity_reference(a, sparsity, lower=lower)
        pm = equal_sparsity_device(x, sparsity, lower=lower)
        # a quantile that lands on +-inf interpolates inf - inf = NaN in NumPy (and every comparison is then False)
        assert (pm.threshold == ref_thr or (np.isnan(pm.threshold) and np.isnan(ref_thr))) and pm.threshold.dtype == ref_thr.dtype
        assert np.array_equal(pm.numpy(int), ref_bin)
        assert int(pm.ones.item()) == int(ref_bin.sum())
        assert np.array_equal(pm.bits.cpu().numpy().view(np.uint32), so.pack_bits(ref_bin))
    got, ref = quantile_device(x, sparsity), np.quantile(a.flatten(), sparsity)
    assert got == ref or (np.isnan(got) and np.isnan(ref))


def test_device_walk_fallback_nan_and_host_walk_agree():
    """(a) the two order statistics in different top-digit bins (the device walk flags it and the host walk takes over);
    (b) NaN scores make the threshold NaN and every bit 0 like np.quantile; (c) host_walk=True gives the same bits."""
    from sdrm_b200.sparsify import equal_sparsity_device, quantile_device
    a = np.concatenate([np.full(50, 1.0, np.float32), np.full(50, 1.0e20, np.float32)]).reshape(10, 10)   # ranks 49 | 50 straddle
    x = torch.from_numpy(a).cuda()
    q = 49.5 / 99
    ref = np.quantile(a.flatten(), np.float32(q))
    assert quantile_device(x, q) == np.quantile(a.flatten(), q)
    ref_bin, ref_thr = so.equal_sparsity_reference(a, q)
    pm = equal_sparsity_device(x, q)
    assert pm.threshold == ref_thr and np.array_equal(pm.numpy(int), ref_bin)
    rng = np.random.RandomState(9)
    b = rng.randn(300, 500).astype(np.float32)
    xb = torch.from_numpy(b).cuda()
    for sp_ in (0.3, 0.977):
        p1, p2 = equal_sparsity_device(xb, sp_), equal_sparsity_device(xb, sp_, host_walk=True)
        assert p1.threshold == p2.threshold and torch.equal(p1.bits, p2.bits)
    b[17, 33] = np.nan
    b[200, 1] = -np.nan
    xn = torch.from_numpy(b).cuda()
    with np.errstate(invalid="ignore"):
        refq = np.quantile(b.flatten(), 0.9)
    assert np.isnan(refq) and np.isnan(quantile_device(xn, 0.9))
    pn = equal_sparsity_device(xn, 0.9)
    assert np.isnan(pn.threshold) and int(pn.ones.item()) == 0 and not pn.numpy(int).any()


def test_strided_view_and_histogram_digits():
    """ld > n_cols (a column slice of a wider tensor) and the raw digit histograms against the NumPy stand-in."""
    from sdrm_b200.sparsify import device_histogram_fn, quantile_device
    rng = np.random.RandomState(5)
    wide = (rng.randn(500, 1100) * 5).astype(np.float32)
    xw = torch.from_numpy(wide).cuda()
    for cols in (1008, 1001):
        view, ref = xw[:, :cols], np.ascontiguousarray(wide[:, :cols])
        assert quantile_device(view, 0.9) == np.quantile(ref.flatten(), 0.9)
        dev, host = device_histogram_fn(view), so.numpy_histogram_fn(ref)
        h0 = host(0, 0, 21, 11)
        assert np.array_equal(dev(0, 0, 21, 11), h0)
        b = int(np.argmax(h0))
        assert np.array_equal(dev(b, 11, 10, 11), host(b, 11, 10, 11))
        h1 = host(b, 11, 10, 11)
        b1 = (b << 11) | int(np.argmax(h1))
        assert np.array_equal(dev(b1, 22, 0, 10), host(b1, 22, 0, 10))


def test_full_size_order_statistic_property():
    """At a size the host cannot sort in seconds: the selected value v at rank r satisfies
    #(S < v) <= r < #(S <= v), and the packed matrix has exactly #(S >= thr) ones."""
    from sdrm_b200.sparsify import device_histogram_fn, equal_sparsity_device, select_ranks
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(20000, 20000, device="cuda", generator=g) * 2.5 - 3.0      # 1.6 GB, cfg 5 item count
    n = x.numel()
    for r in (0, n // 3, int(0.99 * n), n - 1):
        v = float(select_ranks(device_histogram_fn(x), [r])[0])
        assert int((x < v).sum()) <= r < int((x <= v).sum())
    pm = equal_sparsity_device(x, 0.99)
    assert int(pm.ones.item()) == int((x >= float(pm.threshold)).sum())
    assert abs(int(pm.ones.item()) / n - 0.01) < 1e-6


def test_sampler_output_end_to_end():
    """Scores straight out of sdrm_sample -> equal sparsity on the device == the reference's host lines."""
    from helpers import random_modules
    from sdrm_b200.engine import SamplerEngine
    from sdrm_b200.models import make_schedule
    from sdrm_b200.sparsify import equal_sparsity_device
    diff, vae = random_modules(1008, 120, 72, 9, 1, seed=5, device="cuda")
    eng = SamplerEngine()
    eng.pack_denoiser(diff, make_schedule(9, device="cuda"), 1.0)
    eng.pack_decoder(vae)
    S = eng.sample(843, seed=3, check=True)
    ref_bin, ref_thr = so.equal_sparsity_reference(S.cpu().numpy(), 0.9424673)
    pm = equal_sparsity_device(S, 0.9424673)
    assert pm.threshold == ref_thr and np.array_equal(pm.numpy(int), ref_bin)


def test_no_cpu_fallback():
    from sdrm_b200 import _lib
    from sdrm_b200.sparsify import equal_sparsity_device
    with pytest.raises(_lib.SdrmError):
        equal_sparsity_device(torch.zeros(4, 4), 0.5)
