"""Shared helpers for the parity tests (build reference-shaped modules from golden state_dicts)."""
import torch

from sdrm_b200.models import SDRM, VAE


def modules_from_golden(g, device="cpu"):
    diff = SDRM(N_ITEMS=g["L"], EMB_DIM=g["T"], LATENT_DIM=g["L"], n_hidden_layers=g["nh"])
    diff.load_state_dict(g["denoiser"])
    vae = VAE(input_dim=g["I"], hidden_dim=g["H"], latent_dim=g["L"])
    vae.load_state_dict(g["vae"])
    return diff.to(device).eval(), vae.to(device).eval()


def random_modules(I, H, L, T, nh, seed=0, device="cpu", scale_out=1.0):
    torch.manual_seed(seed)
    vae = VAE(input_dim=I, hidden_dim=H, latent_dim=L)
    diff = SDRM(N_ITEMS=L, EMB_DIM=T, LATENT_DIM=L, n_hidden_layers=nh)
    return diff.to(device).eval(), vae.to(device).eval()


def state_dicts(diff, vae):
    return ({k: v.detach().cpu().float() for k, v in diff.state_dict().items()},
            {k: v.detach().cpu().float() for k, v in vae.state_dict().items()})


def rel_fro(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def max_scaled_err(a, b):
    """max |a-b| / max(1, |b|)  — the 1e-3 bar of BASELINE.json north_star."""
    return float(((a.double() - b.double()).abs() / b.double().abs().clamp_min(1.0)).max())
