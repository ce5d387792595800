import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdrm_b200.models import SDRM
from sdrm_b200.training import denoiser_gemms, DenoiserGemms
L, T, nh, B = int(sys.argv[1]), int(sys.argv[3]), 0, int(sys.argv[2])
sparse = len(sys.argv) > 4
torch.manual_seed(1)
net = SDRM(N_ITEMS=L, EMB_DIM=T, LATENT_DIM=L, n_hidden_layers=nh).cuda()
rows = 3 * B
x = torch.randn(rows, L, device="cuda")
if sparse:
    x = x * (torch.rand(rows, L, device="cuda") < 0.5) * 2.0
t = torch.randint(1, T + 1, (B,), device="cuda").repeat(3)
g_out = torch.randn(rows, L, device="cuda") / rows
for rep in range(2):
    net.zero_grad()
    out = denoiser_gemms(net, x, t, passes=3)
    lt = net.layer_tensors()
    steps = torch.arange(T + 1, device="cuda")
    (out * g_out).sum().backward()
    # reference G0 rows in float64
    W0, Wo = lt["W0"].double(), lt["Wo"].double()
    emb = net.emb_layer(net.timestep_embedding(t, T)).double()
    pre0 = torch.cat([x.double(), emb], -1) @ W0.T + lt["b0"].double()
    a0 = lt["a0"].double()
    h0 = torch.where(pre0 > 0, pre0, a0 * pre0)
    o = torch.tanh(h0 @ Wo.T + lt["bo"].double())
    G1 = g_out.double() * (1 - o * o)
    G0 = (G1 @ Wo) * torch.where(pre0 > 0, torch.ones_like(pre0), a0.expand_as(pre0))
    gW0_ref = G0.T @ x.double()
    got = net.dnn[0].weight.grad[:, :L].double()
    err = (got - gW0_ref).abs()
    print(f"G0DBG rep {rep}: gW0 max err {err.max():.2e} of {gW0_ref.abs().max():.2e}; bad rows of dW0 (out features) {(err.max(1).values > 1e-4 * gW0_ref.abs().max()).nonzero().flatten().tolist()[:40]}; "
          f"bad cols {(err.max(0).values > 1e-4 * gW0_ref.abs().max()).nonzero().flatten().tolist()[:40]}")
    gb = net.dnn[0].bias.grad.double()
    eb = (gb - G0.sum(0)).abs()
    print(f"G0DBG rep {rep}: gb0 max err {eb.max():.2e} of {G0.sum(0).abs().max():.2e}; bad features {(eb > 1e-4 * G0.sum(0).abs().max()).nonzero().flatten().tolist()[:40]}")
