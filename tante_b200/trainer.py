"""Host-side mirrors of the reference training drivers for the hot path.

`rollout_train` restates `Trainer.rollout_model` (trainer/trainer.py:144-159: whole batch, fixed step) and
`R_Trainer.rollout_model` (trainer/r_trainer.py:112-133: per-sample B=1 while-loops with out_T=1.5, window NOT
detached => BPTT through the chained model calls) around the drop-in module; `train_step` is the body of
`train_one_epoch` (trainer.py:178-198 / r_trainer.py:145-159): rollout -> MSE(+rt penalty) -> backward -> clip ->
optimizer step.  All model arithmetic (forward AND backward) runs in libtante_b200.so; loss, clipping and the
optimizer are the reference's own torch calls (optim is out of scope), and data-parallel training adds
exactly one collective: an all-reduce of ONE flat gradient bucket before clipping.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def mse_loss(y_pred, y_ref, rts=None, eps: float = 0.5, n: int = 2):
    """trainer/metrics.py:19-80: MSE.eval (mean over H, W) .mean() [+ MSE.eval_rt penalty on mean R_t]."""
    loss = torch.mean((y_pred - y_ref) ** 2, dim=(-3, -2)).mean()
    if rts is None:
        return loss
    avg = torch.mean(rts)
    up, down = min(1 + eps, 4), max(1 + eps, 4)
    a = float(avg)                                   # the reference's Python `if` on a tensor: a host sync
    if a < up:
        loss = loss + 5e-3 * (up - avg) ** n
    if a > down:
        loss = loss + 1e-1 * (avg - down) ** n
    return loss


def _roll(model, window, n_steps: int, out_T):
    moving, ys, rts, cum = window, [], [], 0
    while cum < n_steps:
        if model.deg:
            y, rt = model(moving), None
        else:
            y, rt = model(moving, out_T)
        cum += y.shape[1]
        if cum < n_steps:
            moving = torch.cat([moving[:, y.shape[1]:], y], dim=1)      # window NOT detached (r_trainer.py:126)
        ys.append(y.permute(0, 1, 3, 4, 2))                             # formatter.process_output
        if rt is not None:
            rts.append(rt)
    return torch.cat(ys, dim=1)[:, :n_steps], (torch.cat(rts, dim=0) if rts else None)


def rollout_train(model, x, n_steps: int, out_T: float = 1.5):
    """x (B,T,D,H,W) channels-first -> (y_pred (B,n_steps,H,W,D), Rts or None)."""
    if model.deg:
        return _roll(model, x, n_steps, out_T)
    outs, rts = [], []
    for b in range(x.shape[0]):                                         # r_trainer.py:118 ("TODO: batch size > 1")
        y, r = _roll(model, x[b:b + 1], n_steps, out_T)
        outs.append(y)
        rts.append(r)
    return torch.cat(outs, dim=0), torch.cat(rts, dim=0)


class GradBucket:
    """One flat fp32 gradient bucket: every `p.grad` is a view into it, so autograd accumulates in place and
    data-parallel training needs a single all-reduce (NCCL over NVLink) per optimizer step."""

    def __init__(self, model: torch.nn.Module):
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        named = dict(model.named_parameters())
        if hasattr(model, "grad_layout") and dev.type == "cuda" and all(p.requires_grad for p in named.values()):
            # the library's own flat layout: the backward then accumulates with ONE add per model call
            names, offsets, n = model.grad_layout(dev)
            self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
            for name, off in zip(names, offsets):
                p = named[name]
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            return
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / dist.get_world_size())


def train_step(model, optimizer, x, y_ref, n_steps: int = 4, bucket: Optional[GradBucket] = None,
               clip: str = "norm", rt_eps: float = 0.5, rt_n: int = 2):
    """One optimizer step.  clip = "norm": clip_grad_norm_(1.0) (trainer.py:192-193); "value": clip_grad_value_(1.0)
    (r_trainer.py:155).  Returns the (device) loss tensor."""
    y_pred, rts = rollout_train(model, x, n_steps)
    loss = mse_loss(y_pred, y_ref, rts, rt_eps, rt_n)
    if bucket is not None:
        bucket.zero()
    else:
        optimizer.zero_grad(set_to_none=True)
    loss.backward()
    if bucket is not None:
        bucket.all_reduce_mean()          # before clipping: clipping must see the averaged gradient
    params = bucket.params if bucket is not None else list(model.parameters())
    if clip == "norm":
        torch.nn.utils.clip_grad_norm_(params, 1.0)
    elif clip == "value":
        torch.nn.utils.clip_grad_value_(params, 1.0)
    optimizer.step()
    return loss.detach()
