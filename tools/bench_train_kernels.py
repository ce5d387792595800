"""K2 family timing (HBM-bound kernels of the training side): fused NLL forward / backward, noising, loss statistics,
CSR encoder.  Prints algorithmic GB/s against the measured HBM peak.  Shapes: a scale-up VAE batch (8192 x 20 000)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdrm_b200 import _lib
from sdrm_b200.models import VAE
from sdrm_b200.training import CudaLossBackend, FrozenEncoder

peak = 6535.7
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p)).get("hbm_gbs", peak)
lib = _lib.load()


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(n))


def line(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(f"K2 {name:44s} {ms:8.3f} ms  {gbs:8.1f} GB/s  {gbs / peak:5.2f} of measured HBM peak ({peak} GB/s)")


B, I = 8192, 20000
logits = torch.randn(B, I, device="cuda") * 3
X = (torch.rand(B, I, device="cuda") < 0.01).float()
lse, sx, dot = (torch.empty(B, device="cuda") for _ in range(3))
grad = torch.empty_like(logits)
ms = timed(lambda: lib.sdrm_multinomial_nll_fwd(_lib.ptr(logits), _lib.ptr(X), B, I, I, I, _lib.ptr(lse), _lib.ptr(sx), _lib.ptr(dot), _lib.stream_ptr()))
line(f"multinomial_nll_fwd {B}x{I}", ms, 8.0 * B * I)
ms = timed(lambda: lib.sdrm_multinomial_nll_bwd(_lib.ptr(logits), _lib.ptr(X), B, I, I, I, _lib.ptr(lse), _lib.ptr(sx), None, 1.0 / B, _lib.ptr(grad), I, _lib.stream_ptr()))
line(f"multinomial_nll_bwd {B}x{I}", ms, 12.0 * B * I)
ms = timed(lambda: torch.nn.functional.log_softmax(logits, dim=1).mul(X).sum(1).mean())
line("  (torch log_softmax*X.sum.mean forward, for scale)", ms, 8.0 * B * I)

Bl, L, T = 1 << 20, 950, 178
be = CudaLossBackend()
mu = torch.randn(Bl, L, device="cuda")
t = torch.randint(1, T + 1, (Bl,), device="cuda")
ab = torch.rand(T + 1, device="cuda")
ms = timed(lambda: be.noise_inputs(mu, t, ab, 1.0, 0.1, 7, 0), n=5)
line(f"noise_inputs {Bl}x{L} (1 read, 4 writes)", ms, 20.0 * Bl * L)
pred, sxx, psx = (torch.randn(Bl, L, device="cuda") for _ in range(3))
ms = timed(lambda: be.stats(pred, sxx, psx, mu, 0.1), n=5)
line(f"loss_stats {Bl}x{L} (4 reads)", ms, 16.0 * Bl * L)
st = be.stats(pred, sxx, psx, mu, 0.1)
ms = timed(lambda: be.seeds(pred, sxx, psx, mu, 0.1, st), n=5)
line(f"loss_grad_seeds {Bl}x{L} (4 reads, 3 writes)", ms, 28.0 * Bl * L)

vae = VAE(20000, 1000, 950).cuda().eval()
enc = FrozenEncoder(vae)
xs = (torch.rand(8192, 20000, device="cuda") < 0.01).float().to_sparse_csr()
nnz = xs.values().numel()
ms = timed(lambda: enc.hidden(xs))
line(f"encode_csr 8192 rows, nnz {nnz}, H 1000 (gather bytes)", ms, 4.0 * nnz * 1000)

# K3: warp top-k over a score matrix (4 bytes read per score)
from sdrm_b200 import metrics
del logits, X, grad, mu, pred, sxx, psx
torch.cuda.empty_cache()
for rows, items in ((65536, 20000), (9558, 8582)):
    sc = torch.randn(rows, items, device="cuda")
    for k in (10, 50):
        ms = timed(lambda: metrics.topk_device(sc, k))
        print(f"K3 topk k={k:<2d} {rows}x{items} fp32".ljust(47), f"{ms:8.3f} ms  {4.0 * rows * items / ms / 1e6:8.1f} GB/s  {4.0 * rows * items / ms / 1e6 / peak:5.2f} of measured HBM peak ({peak} GB/s)")
    del sc
