"""Stage-2 (PCD) training step with the reference's structure (run.py:574-721) on the B200 kernels, plus the
one parallel strategy this path has: rays sharded across the GPUs of one box, one NCCL all-reduce (AVG) of a
flat fp32 gradient bucket per step.  torch.distributed is plumbing only.

  create_optimizer         lib/utils.py:480-513  (groups from the `lrate_*` keys)
  STAGE2_LRATES            configs/nerf/default.py:80-92
  train_step               run.py:575,615-631,713-721 (zero_grad, forward, 200*MSE, backward, step, lr decay)
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch
import torch.nn.functional as F

from .masked_adam import MaskedAdam

# configs/nerf/default.py:80-92 (keys without a matching attribute on the model are skipped, lib/utils.py:490)
STAGE2_LRATES = dict(rgbnet=1e-4, densitynet=1e-4, featurenet=1e-4, canonical_feat=1e-4, gammas=1e-3, weights=1e-4,
                     theta_weight=1e-4, forward_warp=1e-4, joints=1e-5, theta=1e-5, feat_net=1e-3)
WEIGHT_RENDER = 2e2          # configs/nerf/default.py:95
LRATE_DECAY = 160            # N_iters // 1000 (configs/nerf/default.py:76)


def create_optimizer(model, lrates: Optional[Dict[str, float]] = None, global_step: int = 0,
                     lrate_decay: int = LRATE_DECAY, skip_zero_grad_fields: Iterable[str] = ()) -> MaskedAdam:
    """lib/utils.py:480-513."""
    lrates = dict(STAGE2_LRATES if lrates is None else lrates)
    decay_factor = 0.1 ** (global_step / (lrate_decay * 1000))
    groups = []
    for k, lr0 in lrates.items():
        if not hasattr(model, k):
            continue
        param = getattr(model, k)
        if param is None:
            continue
        lr = lr0 * decay_factor
        if lr > 0:
            if isinstance(param, torch.nn.Module):
                param = list(param.parameters())
            else:
                param = [param]
            groups.append({'params': param, 'name': k, 'lr': lr, 'skip_zero_grad': k in skip_zero_grad_fields})
        else:
            param.requires_grad = False
    return MaskedAdam(groups)


class GradBucket:
    """All gradients of the optimiser's parameters as views of ONE flat fp32 buffer: zeroing is one memset and the
    data-parallel exchange is one all-reduce ([canonical_feat | weights | joints | theta_weight |
    transform_net | feat_net | rgbnet | densitynet])."""

    def __init__(self, optimizer: torch.optim.Optimizer):
        params = [p for g in optimizer.param_groups for p in g['params'] if p.requires_grad]
        assert params, "no trainable parameters"
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 63) // 64 * 64          # 256-byte aligned slices
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.params = params
        for p, o in zip(params, offs):
            p.grad = self.flat[o:o + p.numel()].view_as(p)
        self.numel = sum(p.numel() for p in params)
        # gradients of the decoder parameters are accumulated by the backward kernels directly into these slices
        from . import ops
        ops.DIRECT_GRAD_ACCUM = True

    def zero(self):
        self.flat.zero_()

    def all_reduce_avg(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG if self.flat.is_cuda else dist.ReduceOp.SUM, group=group)
            if not self.flat.is_cuda:                      # gloo has no AVG
                self.flat.div_(dist.get_world_size(group))


def shard_rays(n_rays: int, rank: int, world: int):
    """Contiguous, near-equal ray ranges per rank: [start, stop)."""
    base, rem = divmod(n_rays, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def train_step(model, optimizer: MaskedAdam, bucket: GradBucket, t, render_kwargs, target, *, decay_factor: float = 1.0):
    """One stage-2 iteration on one rank's ray batch.  Returns the (device) loss tensor."""
    bucket.zero()
    res = model(t, False, render_kwargs, render_pcd_direct=False)
    loss = WEIGHT_RENDER * F.mse_loss(res['rgb_marched'], target)
    loss.backward()
    bucket.all_reduce_avg()
    optimizer.step()
    if decay_factor != 1.0:
        for g in optimizer.param_groups:                  # run.py:718-721
            g['lr'] = g['lr'] * decay_factor
    return loss
