"""GPU parity of the MultiVAE++ training step on the tcgen05 GEMM (reference train_SDRM.py:136-150, 206-256) and of the
device-resident CSR staging (8f-4; reference dataloaders.py:46-79, train_SDRM.py:323-324)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

pytestmark = pytest.mark.gpu


def test_tc_linear_matches_torch_linear():
    from sdrm_b200.training import TcLinear
    torch.manual_seed(0)
    lin = torch.nn.Linear(333, 217).cuda()
    x = torch.randn(150, 333, device="cuda", requires_grad=True)
    g = torch.randn(150, 217, device="cuda")
    y = TcLinear.apply(x, lin.weight, lin.bias, 3)
    y.backward(g)
    got = (y.detach(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    x64 = x.detach().double().requires_grad_(True)
    W64 = lin.weight.detach().double().requires_grad_(True)
    b64 = lin.bias.detach().double().requires_grad_(True)
    y64 = torch.nn.functional.linear(x64, W64, b64)
    y64.backward(g.double())
    for a, r in zip(got, (y64.detach(), x64.grad, W64.grad, b64.grad)):
        assert (a.double() - r).abs().max().item() <= 3e-5 * max(1.0, r.abs().max().item())


@pytest.mark.parametrize("B,I,H,L", [(40, 300, 64, 24), (550, 1008, 930, 830)])
def test_vae_training_step_matches_reference_autograd(B, I, H, L):
    """One MultiVAE++ step (train_SDRM.py:136-146): neg_ll + anneal * KL, backward.  Same seed -> same dropout mask and
    reparameterisation noise in both paths; the reference path is the module's own torch forward (nn.Linear / cuBLAS fp32)."""
    from sdrm_b200.models import VAE
    from sdrm_b200.training import multinomial_nll, vae_forward_tc
    torch.manual_seed(5)
    vae = VAE(I, H, L).cuda()
    vae.train()
    vae.is_training = 1
    X = (torch.rand(B, I, device="cuda") < 0.06).float()
    anneal = 0.137

    def step(fwd):
        vae.zero_grad()
        torch.manual_seed(99)
        out, kl = fwd(X)
        logp = torch.log_softmax(out.double(), dim=1)     # the reference's expression (train_SDRM.py:143), in float64
        ref_nll = -torch.mean(torch.sum(logp * X.double(), dim=1))
        nll = multinomial_nll(out, X)
        loss = nll + anneal * kl
        loss.backward()
        return loss.item(), ref_nll.item() + anneal * kl.item(), {k: p.grad.detach().clone() for k, p in vae.named_parameters()}

    loss_tc, loss_tc_ref, g_tc = step(lambda x: vae_forward_tc(vae, x, passes=3))
    loss_t, _, g_t = step(lambda x: vae(x))
    assert abs(loss_tc - loss_tc_ref) <= 2e-5 * abs(loss_tc_ref)
    assert abs(loss_tc - loss_t) <= 2e-5 * abs(loss_t), (loss_tc, loss_t)
    for k in g_t:
        scale = g_t[k].abs().max().item() + 1e-12
        assert (g_tc[k] - g_t[k]).abs().max().item() <= 2e-4 * scale, k


def _random_csr(rows, cols, density, seed, explicit_zeros=True):
    rng = np.random.RandomState(seed)
    m = sp.random(rows, cols, density=density, format="csr", random_state=rng, data_rvs=lambda n: rng.randint(1, 6, n).astype(np.int64))
    if explicit_zeros:     # the reference pickles hold explicit stored zeros
        m.data[rng.rand(m.nnz) < 0.3] = 0
    return m.astype(np.int64)


def test_device_csr_row_slices_match_scipy():
    from sdrm_b200.training import DeviceCSR
    m = _random_csr(500, 321, 0.05, 1)
    m[7] = 0          # an empty row
    m = m.tocsr()
    d = DeviceCSR(m, "cuda")
    rng = np.random.RandomState(2)
    for n in (1, 17, 300):
        idx = rng.permutation(500)[:n]
        assert np.array_equal(d.dense_rows(idx).cpu().numpy(), m[idx].toarray().astype(np.float32))
        assert np.array_equal(d.csr_rows(idx).to_dense().cpu().numpy(), m[idx].toarray().astype(np.float32))
    assert np.array_equal(d.dense_rows([7, 7]).cpu().numpy(), np.zeros((2, 321), np.float32))


def test_staged_loader_mu_once_equals_per_batch_encode():
    """train_SDRM computes mu for all rows once from the device CSR and slices it with the loader's own batch sampler; that must
    equal encoding every collated batch (train_SDRM.py:323-324) in the order the loader yields them."""
    from torch.utils.data import DataLoader
    from sdrm_b200.data import SparseDataset, sparse_batch_collate
    from sdrm_b200.models import VAE
    from sdrm_b200.training import FrozenEncoder, stage_loader
    m = _random_csr(230, 150, 0.08, 3).astype(np.float64)
    ds = SparseDataset(m, m)
    sampler = torch.utils.data.sampler.BatchSampler(
        torch.utils.data.sampler.RandomSampler(ds, generator=torch.Generator(device="cpu")), batch_size=64, drop_last=False)
    dl = DataLoader(ds, batch_size=1, collate_fn=sparse_batch_collate, generator=torch.Generator(device="cpu"), sampler=sampler, shuffle=False)
    torch.manual_seed(1)
    vae = VAE(150, 48, 20).cuda().eval()
    enc = FrozenEncoder(vae)
    staged = stage_loader(dl, torch.device("cuda"))
    assert staged is not None
    rows_all, bs = staged
    mu_all = enc(rows_all.as_torch_csr())
    # the two iterations draw the permutation from the sampler's own generator: replay it with a fixed state
    state = sampler.sampler.generator.get_state()
    staged_mus = [mu_all[torch.as_tensor(idx, device="cuda")] for idx in bs]
    sampler.sampler.generator.set_state(state)
    loader_mus = [enc(x) for x, _ in iter(dl)]
    assert len(staged_mus) == len(loader_mus) == 4
    for a, b in zip(staged_mus, loader_mus):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
