"""CPU, gloo, world_size 2: the multi-GPU host logic (row sharding, global loss statistics, gradient all-reduce).
The CUDA kernels are replaced by a test-local torch stub with the same three methods, so what is under test is the
orchestration in sdrm_b200.distributed / sdrm_b200.training, compared with the single-process oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden


class TorchStubBackend:
    """Same contract as training.CudaLossBackend, plain torch (test-only)."""

    def noise_inputs(self, mu, t, ab_t, nd, mu_coef, seed, row_offset, inj_noise=None, inj_masks=None, want_masks=False):
        assert inj_noise is not None and inj_masks is not None, "tests inject the noise"
        k = inj_masks.to(mu.dtype)
        x_t = ab_t.sqrt()[t, None] * mu + (1 - ab_t[t, None]) * inj_noise
        x_p = mu + mu_coef * inj_noise
        return inj_noise, x_t * k[0] * 2, mu * k[1] * 2, x_p * k[2] * 2, None

    def stats(self, pred, sx, psx, mu, mu_coef):
        r = (pred - mu).double()
        sd = ((psx - sx) / (mu_coef ** 2)).double()
        return torch.stack([r.sum(), (r * r).sum(), ((sd - r) ** 2).sum(), ((r - sx.double()) ** 2).sum(),
                            torch.tensor(float(r.numel()), dtype=torch.float64)])

    def seeds(self, pred, sx, psx, mu, mu_coef, st):
        N = st[4]
        mean_r = st[0] / N
        V = (st[1] - N * mean_r ** 2) / (N - 1)
        A, Bm = st[2] / N, st[3] / N
        den = 1e-8 + V
        c = 0.5 / den
        r = (pred - mu).double()
        sd = ((psx - sx) / (mu_coef ** 2)).double()
        g_sd = 2 * c * (sd - r) / N
        g_psx = g_sd / (mu_coef ** 2)
        g_sx = -2 * c * (r - sx.double()) / N - g_sd / (mu_coef ** 2)
        g_pred = c * (-2 * (sd - r) + 2 * (r - sx.double())) / N - 0.5 * (A + Bm) / den ** 2 * 2 * (r - mean_r) / (N - 1)
        return g_pred.float(), g_sx.float(), g_psx.float(), (0.5 * (A + Bm) / den).float().reshape(1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import modules_from_golden
        from oracle import sdrm_oracle as orc
        from sdrm_b200 import distributed as sdd
        from sdrm_b200.training import DiffusionTrainStep
        g = load_golden("t_nh2_T7_L24")
        diff, _ = modules_from_golden(g)
        diff.train()
        _, _, ab_t = orc.make_schedule(g["T"])
        stepper = DiffusionTrainStep(diff, ab_t, g["T"], g["nd"], backend=TorchStubBackend(), seed=1)
        opt = torch.optim.SGD(diff.parameters(), lr=0.0)
        loss = sdd.dp_train_step(stepper, opt, g["mu"], g["t"], inj_noise=g["noise"], inj_masks=g["keeps"])
        grads = {k: p.grad.clone() for k, p in diff.named_parameters()}
        # sharded "sampling" with a fake sampler: every rank returns its global row ids -> gather is the identity
        fake = lambda n, *a, row_offset=0, **k: torch.arange(row_offset, row_offset + n, dtype=torch.float32)[:, None].repeat(1, 3)
        rows, span = sdd.sample_ddpm_sharded(11, None, None, 0, seed=0, gather=True, sampler=fake)
        torch.save((rank, float(loss), grads, rows, span), os.path.join(out_dir, f"rank{rank}.pt"))
    except Exception:
        import traceback
        with open(os.path.join(out_dir, f"rank{rank}.err"), "w") as fh:
            fh.write(traceback.format_exc())
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_dp_step_equals_single_process_reference():
    world = 2
    import tempfile
    ctx = mp.get_context("spawn")
    port = _free_port()
    out_dir = tempfile.mkdtemp(prefix="sdrm_dp_")
    procs = [ctx.Process(target=_worker, args=(r, world, port, out_dir)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
    for r, p in enumerate(procs):
        err = os.path.join(out_dir, f"rank{r}.err")
        assert p.exitcode == 0, open(err).read() if os.path.exists(err) else f"rank {r} exit code {p.exitcode}"
    res = [torch.load(os.path.join(out_dir, f"rank{r}.pt"), weights_only=False) for r in range(world)]
    g = load_golden("t_nh2_T7_L24")
    for rank, loss, grads, rows, span in res:
        # the loss and EVERY gradient equal the reference's single-process step on the whole minibatch
        assert abs(loss - g["loss_ref"].item()) <= 1e-5 * abs(g["loss_ref"].item())
        for k, gref in g["grads_ref"].items():
            assert torch.allclose(grads[k], gref, rtol=2e-3, atol=1e-7), (rank, k)
        assert span == (0, 11) and torch.equal(rows[:, 0], torch.arange(11, dtype=torch.float32))
