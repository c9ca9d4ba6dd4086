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


class _FusedMSE(torch.autograd.Function):
    """mean((y_pred - y_ref)^2) over every element -- what `MSE.eval(...).mean()` (trainer/metrics.py:53-80) computes
    on the concatenated, channels-last predictions -- evaluated directly on the per-call channels-first frame tensors
    by libtante_b200's `tante_mse_cl`: one pass for the loss, one for the gradients; no permute / cat / sub / pow / mean
    kernels and none of their autograd counterparts."""

    @staticmethod
    def forward(ctx, y_ref, n_steps, *frames):
        from . import _abi
        lib = _abi.load()
        dev = frames[0].device
        y_ref = y_ref.to(torch.float32).contiguous()
        B, n_ref, H, W, D = y_ref.shape
        stream = torch.cuda.current_stream(dev).cuda_stream
        acc = torch.zeros(1, device=dev, dtype=torch.float32)
        plan, f0 = [], 0
        for f in frames:
            if f.dtype != torch.float32 or not f.is_contiguous() or f.shape[0] != B or tuple(f.shape[2:]) != (D, H, W):
                raise ValueError("fused MSE expects contiguous fp32 frames (B, n, D, H, W) matching y_ref (B, n, H, W, D)")
            use = max(0, min(int(f.shape[1]), n_steps - f0))
            plan.append((f0, use))
            if use > 0:
                _abi.check(lib.tante_mse_cl(f.data_ptr(), y_ref.data_ptr(), B, int(f.shape[1]), use, D, H * W, n_ref, f0,
                                            0.0, None, acc.data_ptr(), None, stream))
            f0 += use
        ctx.numel = float(B) * n_steps * H * W * D
        ctx.plan, ctx.n_ref = plan, n_ref
        ctx.save_for_backward(y_ref, *frames)
        return (acc / ctx.numel).reshape(())

    @staticmethod
    def backward(ctx, gout):
        from . import _abi
        lib = _abi.load()
        y_ref, *frames = ctx.saved_tensors
        B, n_ref, H, W, D = y_ref.shape
        dev = y_ref.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        g = gout.to(torch.float32).contiguous()
        grads = []
        for f, (f0, use) in zip(frames, ctx.plan):
            gf = torch.empty_like(f)
            _abi.check(lib.tante_mse_cl(f.data_ptr(), y_ref.data_ptr(), B, int(f.shape[1]), use, D, H * W, n_ref, f0,
                                        2.0 / ctx.numel, g.data_ptr(), None, gf.data_ptr(), stream))
            grads.append(gf)
        return (None, None, *grads)


def mse_loss_frames(frames, y_ref, n_steps: int):
    """MSE.eval(...).mean() on the per-call channels-first predictions (list of (B, n_i, D, H, W)) against the
    channels-last targets y_ref (B, >= n_steps, H, W, D); frames beyond n_steps are ignored (r_trainer.py:130)."""
    return _FusedMSE.apply(y_ref, int(n_steps), *frames)


def _roll_frames(model, window, n_steps: int):
    """`_roll` for the fixed-step model without the formatter permute and the cat over calls."""
    moving, ys, cum = window, [], 0
    while cum < n_steps:
        y = model(moving)
        cum += y.shape[1]
        if cum < n_steps:
            moving = torch.cat([moving[:, y.shape[1]:], y], dim=1)      # window NOT detached (BPTT)
        ys.append(y)
    return ys


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
    if getattr(model, "deg", False) and x.is_cuda and hasattr(model, "grad_layout"):
        # fixed-step model on the CUDA path: loss and its gradient straight from the per-call frames (tante_mse_cl)
        loss = mse_loss_frames(_roll_frames(model, x, n_steps), y_ref, n_steps)
    else:
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
