"""Multi-GPU plumbing: columns are independent, so they are sharded contiguously over
the ranks (one process per GPU) and the forward/reverse kernels need NO collective.  The only
exchange is the calibration step with parameters SHARED by all columns: one all-reduce of
[loss_sum, n_columns, dL/dalpha[L], dL/dn[L], dL/dksat[L]] (1 + 1 + 3L doubles) per optimiser
step, over NCCL (NVLink 5 / NVSwitch) on the GPU box or gloo in the CPU tests.  The reference has
no distributed code at all (agents/DifferentiableLGAR.py trains one column on one CPU)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(num_columns: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of columns owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(num_columns, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_loss_and_grads(loss_sum: torch.Tensor, count, grads: list[torch.Tensor], group=None):
    """Sum a local loss sum, a column count and the local gradient sums of the shared parameters over
    all ranks with ONE collective.  Returns (mean loss, [mean gradients])."""
    flat = torch.cat([loss_sum.reshape(1).to(torch.float64),
                      torch.as_tensor([float(count)], dtype=torch.float64, device=loss_sum.device)]
                     + [g.reshape(-1).to(torch.float64) for g in grads])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    total = flat[1]
    out, o = [], 2
    for g in grads:
        out.append((flat[o:o + g.numel()] / total).reshape(g.shape))
        o += g.numel()
    return flat[0] / total, out


def calibration_step(shared_params, ensemble_fn, loss_fn, optimizer, group=None):
    """One data-parallel calibration step with parameters shared by every column of every rank.
    `ensemble_fn(params) -> per-column outputs` runs this rank's shard (lgar_columns);
    `loss_fn(outputs) -> (loss_sum, n_columns)` on the local shard."""
    optimizer.zero_grad(set_to_none=True)
    outputs = ensemble_fn(shared_params)
    loss_sum, count = loss_fn(outputs)
    loss_sum.backward()
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in shared_params]
    loss, mean_grads = allreduce_loss_and_grads(loss_sum.detach(), count, grads, group)
    for p, g in zip(shared_params, mean_grads):
        p.grad = g.to(p.dtype)
    optimizer.step()
    return loss
