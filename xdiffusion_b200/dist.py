"""Batch-sharded sampling over the GPUs of one box: one process per GPU, samples are independent
(every op on the path is per-sample), so rank r owns rows [r*B/P, (r+1)*B/P) of the initial
latents, per-step noise and conditioning; there is no communication inside the loop and ONE
all-gather of the finished samples at the end (NCCL over NVLink; ``gloo`` in the CPU tests)."""
from typing import Dict, Optional

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous, balanced row ranges (first ``total % world`` ranks get one extra row)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_context(context: Optional[Dict], total: int, lo: int, hi: int):
    """Row-slice every batch-major tensor / list of the conditioning context."""
    if context is None:
        return None
    out = {}
    for k, v in context.items():
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == total:
            out[k] = v[lo:hi].contiguous()
        elif isinstance(v, (list, tuple)) and len(v) == total:
            out[k] = list(v[lo:hi])
        else:
            out[k] = v
    return out


def gather_rows(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather row shards (possibly ragged by one row) into the full [total, ...] tensor."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    full = torch.empty((world * maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, pad, group=group)
    return torch.cat([full[r * maxn: r * maxn + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], 0)


def sample_sharded(model, context: Optional[Dict] = None, num_samples: int = 16, initial_noise=None, noise=None,
                   group=None, gather: bool = True, **kwargs):
    """``model.sample`` with the batch split across the process group.  Every rank passes the FULL
    context / initial_noise / noise (or None); each takes its rows.  Returns the full batch on every
    rank (``gather=True``) or only the local rows."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(num_samples, world, rank)
    ctx = shard_context(context, num_samples, lo, hi)
    x0 = None if initial_noise is None else initial_noise[lo:hi].contiguous()
    z = None if noise is None else noise[:, lo:hi].contiguous()
    # row_offset: the in-kernel Philox counter (and a seeded x_T) is indexed by the GLOBAL row, so every rank draws
    # the noise its rows get in the one-GPU run with the same seed: shards differ from each other and
    # world = 1 / world = N produce the same batch.
    local, inter = model.sample(context=ctx, num_samples=hi - lo, initial_noise=x0, noise=z, row_offset=lo, **kwargs)
    if gather and world > 1:
        return gather_rows(local, num_samples, group), inter
    return local, inter
