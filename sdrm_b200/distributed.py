"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the box,
gloo in the CPU tests).  The reference has no distributed code at all; this is new.

Sampling shards naturally: rows are independent and the Philox streams are keyed by the GLOBAL row id, so rank r
simply generates rows [lo_r, hi_r) and the union is bit-identical to a single-GPU run.  No collective happens
during compute; gathering the rows afterwards is optional.

Training is data-parallel with ONE real exchange step: the score-matching loss divides by the variance of the
residual over the GLOBAL minibatch, so the five fp64 partial sums are all-reduced before the gradient seeds are
formed (sdrm_b200.training.ScoreMatchingLoss), and the flat parameter gradient is all-reduced (SUM — the seeds
already carry the global 1/N) before the identical Adam step on every rank.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous block partition of n rows: the first n % world ranks get one extra row."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, extra = divmod(int(n), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sample_ddpm_sharded(n_sample, diff_net, vae_net, diff_latent_dim, noise_divider=1.0, timesteps=None,
                        n_timesteps=None, *, seed, group=None, gather=False, sampler=None):
    """Each rank samples its block of the n_sample rows (same `seed` on every rank).  Returns (rows, (lo, hi)),
    or the full [n_sample, I] matrix on every rank when gather=True (all_gather of equal-size padded blocks)."""
    if sampler is None:
        from .train_SDRM import sample_ddpm as sampler
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_bounds(n_sample, rank, world)
    rows = sampler(hi - lo, diff_net, vae_net, diff_latent_dim, noise_divider, timesteps=timesteps,
                   n_timesteps=n_timesteps, seed=seed, row_offset=lo)
    if not gather or world == 1:
        return rows, (lo, hi)
    width = rows.shape[1]
    cap = shard_bounds(n_sample, 0, world)[1]  # largest block
    padded = rows.new_zeros((cap, width))
    padded[: hi - lo] = rows
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    out = torch.cat([p[: shard_bounds(n_sample, r, world)[1] - shard_bounds(n_sample, r, world)[0]]
                     for r, p in enumerate(parts)], dim=0)
    return out, (0, n_sample)


def unique_parameters(module):
    """Parameters de-duplicated by identity (the shared hidden layer appears once, like in Adam's param list)."""
    seen, out = set(), []
    for p in module.parameters():
        if id(p) not in seen:
            seen.add(id(p))
            out.append(p)
    return out


def allreduce_gradients(module, group=None):
    """SUM-all-reduce all gradients as one flat bucket (8-12 MB for the SDRM denoiser: latency-, not bandwidth-bound)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    params = [p for p in unique_parameters(module) if p.grad is not None]
    if not params:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off: off + n].view_as(p.grad))
        off += n


def dp_train_step(stepper, optimizer, mu_global, t_global=None, group=None, inj_noise=None, inj_masks=None):
    """One data-parallel diffusion training step on a GLOBAL minibatch of latents [B, L] known to every rank:
    rank r works on its block of rows; returns the global loss (identical on all ranks)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_bounds(mu_global.shape[0], rank, world)
    stepper.group = (group if group is not None else dist.group.WORLD) if world > 1 else None
    stepper.row_offset = lo
    optimizer.zero_grad()
    loss = stepper.loss(mu_global[lo:hi], None if t_global is None else t_global[lo:hi],
                        None if inj_noise is None else inj_noise[lo:hi].contiguous(),
                        None if inj_masks is None else inj_masks[:, lo:hi].contiguous())
    loss.backward()
    allreduce_gradients(stepper.net, group)
    optimizer.step()
    return loss.detach()
