"""Per-parameter relative gradient errors of the tcgen05 training path vs float64 autograd for several shapes (debug)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdrm_b200.models import SDRM
from sdrm_b200.training import denoiser_gemms

def run(L, T, nh, B, passes=3):
    torch.manual_seed(L + T)
    net = SDRM(N_ITEMS=L, EMB_DIM=T, LATENT_DIM=L, n_hidden_layers=nh).cuda()
    rows = 3 * B
    x = torch.randn(rows, L, device="cuda") * (torch.rand(rows, L, device="cuda") < 0.5) * 2.0
    t = torch.randint(1, T + 1, (B,), device="cuda").repeat(3)
    g_out = torch.randn(rows, L, device="cuda") / rows
    net.zero_grad()
    out = denoiser_gemms(net, x, t, passes=passes)
    (out * g_out).sum().backward()
    got = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    ref_net = SDRM(N_ITEMS=L, EMB_DIM=T, LATENT_DIM=L, n_hidden_layers=nh).cuda().double()
    ref_net.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
    emb = ref_net.emb_layer(ref_net.timestep_embedding(t, T).double())
    ref_out = ref_net.dnn(torch.cat([x.double(), emb], dim=-1))
    (ref_out * g_out.double()).sum().backward()
    rep = {k: float((got[k].double() - p.grad).abs().max() / (p.grad.abs().max() + 1e-30)) for k, p in ref_net.named_parameters()}
    print(f"GRADS L={L} T={T} nh={nh} B={B} rows={rows}: out err {float((out.double() - ref_out).abs().max()):.1e}  " +
          "  ".join(f"{k}:{v:.1e}" for k, v in rep.items()), flush=True)

import ast
for cfg in ast.literal_eval(sys.argv[1]):
    run(*cfg)
