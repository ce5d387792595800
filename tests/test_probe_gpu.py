"""GPU: the tcgen05 layer engine in isolation (one dense layer through sdrm_probe_linear) vs torch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K,N,split3", [(128, 64, 64, 0), (100, 40, 40, 0), (300, 200, 300, 0), (1000, 950, 950, 0),
                                          (1000, 950, 950, 1), (19000, 340, 490, 0), (257, 1000, 5000, 1), (1, 5, 17, 1)])
def test_probe_linear(M, K, N, split3):
    from sdrm_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda")
    wsb = lib.sdrm_probe_linear_workspace_bytes(M, K, N)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    _lib.check(lib.sdrm_probe_linear(_lib.ptr(A), _lib.ptr(W), _lib.ptr(b), _lib.ptr(out), M, K, N, split3, _lib.ptr(ws),
                                     wsb, _lib.stream_ptr()), "probe")
    torch.cuda.synchronize()
    if split3:
        ref = (A.double() @ W.double().T + b.double()).float()
        tol = 1e-4
    else:
        ref = (A.bfloat16().double() @ W.bfloat16().double().T + b.double()).float()
        tol = 2e-5
    assert torch.isfinite(out).all()
    assert float((out - ref).norm() / ref.norm()) < tol
