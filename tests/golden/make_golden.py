"""Generate the golden fixtures in this directory by running the REAL reference (/root/reference,
imported through oracle/refstub.py) in the build container.  The reference has no tests or golden
vectors of its own, so these files are what pins the oracle:

  for every case the reference's own functions are run under a fixed torch/numpy seed; the same seed is
  then replayed with the reference's RNG draw order to capture the exact noise tensors it consumed, and
  the pair (inputs incl. noise, reference outputs) is stored.

Run:  python tests/golden/make_golden.py        (needs /root/reference; writes tests/golden/*.pt)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refstub  # noqa: E402
from oracle import sdrm_oracle as orc  # noqa: E402

# name, n, I, H, L, T, nh, nd
SAMPLER_CASES = [
    ("s_nh0_T6_L20", 24, 37, 28, 20, 6, 0, 0.2),
    ("s_nh1_T7_L24", 33, 50, 32, 24, 7, 1, 1.0),
    ("s_nh2_T5_L150", 20, 130, 96, 150, 5, 2, 1.0),
    ("s_nh5_T9_L40", 40, 64, 40, 40, 9, 5, 1.0),
]
TRAIN_CASES = [
    ("t_nh2_T7_L24", 30, 50, 32, 24, 7, 2, 1.0),
    ("t_nh0_T6_L20", 17, 37, 28, 20, 6, 0, 0.2),
]


def build_models(ref, I, H, L, T, nh, seed):
    torch.manual_seed(seed)
    vae = ref.VAE(input_dim=I, hidden_dim=H, latent_dim=L)
    diff = ref.SDRM(N_ITEMS=L, EMB_DIM=T, LATENT_DIM=L, n_hidden_layers=nh)
    # perturb the PReLU slopes / biases a little so shared-vs-unshared mistakes are visible
    with torch.no_grad():
        diff.dnn[1].weight.fill_(0.21)
        if nh > 0:
            diff.dnn[3].weight.fill_(0.31)
    vae.eval()
    vae.is_training = 0
    vae.model_is_trained = True
    diff.eval()
    return diff, vae


def set_schedule(ref, T):
    b_t, a_t, ab_t = orc.make_schedule(T)
    ref.b_t, ref.a_t, ref.ab_t = b_t, a_t, ab_t


def replay_full(n, L, T, seed):
    torch.manual_seed(seed)
    xT = torch.randn(n, L)
    z = torch.zeros(T + 1, n, L)
    keep = torch.zeros(T + 1, n, L, dtype=torch.uint8)
    for i in range(T, 0, -1):
        if i > 1:
            z[i] = torch.randn_like(xT)
        keep[i] = torch.empty_like(xT).bernoulli_(0.5).to(torch.uint8)
    return xT, z, keep


def replay_random(n, L, T, seed, np_seed):
    torch.manual_seed(seed)
    np.random.seed(np_seed)
    xT = torch.randn(n, L)
    z = torch.zeros(T + 1, n, L)
    keep = torch.zeros(T + 1, n, L, dtype=torch.uint8)
    t_start = np.zeros(n, dtype=np.int32)
    for j in range(n):
        tj = np.random.randint(1, T)
        t_start[j] = tj
        for i in range(tj, 0, -1):
            if i > 1:
                z[i, j] = torch.randn(L)
            keep[i, j] = torch.empty(1, L).bernoulli_(0.5).to(torch.uint8)[0]
    return xT, z, keep, torch.from_numpy(t_start)


def make_sampler_case(ref, name, n, I, H, L, T, nh, nd):
    diff, vae = build_models(ref, I, H, L, T, nh, seed=100 + T)
    set_schedule(ref, T)
    seed = 7
    torch.manual_seed(seed)
    full_ref = ref.sample_ddpm(n, diff, vae, L, nd, n_timesteps=T).clone()
    xT, z, keep = replay_full(n, L, T, seed)
    torch.manual_seed(seed)
    np.random.seed(11)
    rand_ref = ref.sample_ddpm(n, diff, vae, L, nd, timesteps="random", n_timesteps=T).clone()
    xT_r, z_r, keep_r, t_start = replay_random(n, L, T, seed, 11)
    dsd = {k: v.detach().clone() for k, v in diff.state_dict().items()}
    vsd = {k: v.detach().clone() for k, v in vae.state_dict().items()}
    # sanity: the oracle must reproduce the reference before the fixture is written
    full_or = orc.sample_full(dsd, vsd, T, nd, xT, z, keep)
    rand_or = orc.sample_random(dsd, vsd, T, nd, xT_r, z_r, keep_r, t_start)
    e1 = (full_or - full_ref).abs().max().item()
    e2 = (rand_or - rand_ref).abs().max().item()
    print(f"{name}: oracle vs reference  full {e1:.2e}  random {e2:.2e}")
    assert e1 < 1e-5 and e2 < 1e-5, "oracle does not reproduce the reference"
    torch.save({
        "kind": "sampler", "n": n, "I": I, "H": H, "L": L, "T": T, "nh": nh, "nd": nd,
        "denoiser": dsd, "vae": vsd,
        "full": {"xT": xT, "z": z, "keep": keep, "logits_ref": full_ref},
        "random": {"xT": xT_r, "z": z_r, "keep": keep_r, "t_start": t_start, "logits_ref": rand_ref},
    }, os.path.join(HERE, name + ".pt"))


def make_train_case(ref, name, B, I, H, L, T, nh, nd):
    diff, vae = build_models(ref, I, H, L, T, nh, seed=200 + T)
    set_schedule(ref, T)
    diff.train()
    g = torch.Generator().manual_seed(5)
    X = (torch.rand(B, I, generator=g) < 0.15).float()
    X[:, 0] = 1.0  # no empty rows
    seed = 13
    # --- the reference's training-step body (train_SDRM.py:322-336), run with its own functions
    torch.manual_seed(seed)
    diff.zero_grad()
    encode_x, _ = vae.encode(X)
    noise = torch.randn_like(encode_x, dtype=torch.float) * nd
    t = torch.randint(1, T + 1, (encode_x.shape[0],))
    x_pert = ref.perturb_input(encode_x, t, noise)
    pred_noise = diff.forward(x_pert, t)
    loss = ref.score_matching_loss(diff, XT=encode_x, t=t, epsilon_theta=pred_noise, epsilon=noise, mu=.1)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in diff.named_parameters()}
    # --- replay the draw order: reparam eps, noise, t, three dropout masks
    torch.manual_seed(seed)
    _ = torch.randn(B, L)
    noise_r = torch.randn(B, L) * nd
    t_r = torch.randint(1, T + 1, (B,))
    keeps = torch.stack([torch.empty(B, L).bernoulli_(0.5).to(torch.uint8) for _ in range(3)])
    assert torch.equal(t, t_r) and torch.equal(noise, noise_r)
    dsd = {k: v.detach().clone() for k, v in diff.state_dict().items()}
    vsd = {k: v.detach().clone() for k, v in vae.state_dict().items()}
    mu = orc.vae_encode_mu(vsd, X)
    assert torch.allclose(mu, encode_x.detach(), atol=1e-6)
    params = {k: v.clone().requires_grad_(True) for k, v in dsd.items()}
    nh_ = orc.n_hidden_of(dsd)
    for j in range(1, nh_):  # aliases of the shared layer share storage, like the reference
        params[f"dnn.{2 + 2 * j}.weight"] = params["dnn.2.weight"]
        params[f"dnn.{2 + 2 * j}.bias"] = params["dnn.2.bias"]
        params[f"dnn.{3 + 2 * j}.weight"] = params["dnn.3.weight"]
    loss_o, (pred_o, sx_o, psx_o) = orc.training_loss(params, mu, t_r, noise_r, keeps, T)
    loss_o.backward()
    print(f"{name}: loss ref {loss.item():.8f} oracle {loss_o.item():.8f}")
    assert abs(loss.item() - loss_o.item()) <= 1e-6 * max(1.0, abs(loss.item()))
    for k, gref in grads.items():
        go = params[k].grad
        assert torch.allclose(go, gref, rtol=1e-4, atol=1e-7), k
    torch.save({
        "kind": "train", "B": B, "I": I, "H": H, "L": L, "T": T, "nh": nh, "nd": nd,
        "denoiser": dsd, "vae": vsd, "X": X, "mu": encode_x.detach().clone(), "noise": noise_r, "t": t_r,
        "keeps": keeps, "loss_ref": loss.detach().clone(), "grads_ref": grads,
        "pred_ref": pred_noise.detach().clone(),
    }, os.path.join(HERE, name + ".pt"))


def main():
    ref = refstub.import_reference()
    for case in SAMPLER_CASES:
        make_sampler_case(ref, *case)
    for case in TRAIN_CASES:
        make_train_case(ref, *case)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
