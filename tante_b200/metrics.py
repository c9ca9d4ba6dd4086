"""On-device evaluation metrics -- drop-ins for the reference's `trainer.metrics` classes the Hydra config instantiates
(configs/tante.yaml:50-52,67-74: `trainer.MSE`, `trainer.L2RE`, `trainer.NNMSE`, `trainer.VRMSE`).

Same call contract as reference trainer/metrics.py:18-51: `metric(x, y, rt[, eps, n])` on channels-last tensors
`(B, T, H, W, C)`; with `rt is None` the per-(batch, frame, field) values are returned, otherwise
`eval(x, y).mean() + eval_rt(rt, eps, n)`.

Every metric below is a function of three spatial moments per (b, t, c) -- sum (x-y)^2, sum y^2, sum y over (H, W) --
which `tante_metric_moments` (libtante_b200.so) produces in ONE pass over prediction and target.  The reference
makes a sub / pow / mean (/ std / norm) pass per metric, four metrics per evaluated batch (r_evaler.py:134-137).
The moments of one (x, y) pair are cached, so the evaluators' four metrics cost a single pass.  There is no CPU path.
"""
from __future__ import annotations

import torch

from . import _abi

_cache = {"key": None, "val": None}


def spatial_moments(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """f64 (B, T, C, 3): sum (x-y)^2, sum y^2, sum y over (H, W) of channels-last (B, T, H, W, C) tensors."""
    if x.device.type != "cuda" or y.device != x.device:
        raise RuntimeError("tante_b200.metrics run only on CUDA tensors (libtante_b200.so); there is no CPU path")
    if x.shape != y.shape or x.dim() != 5:
        raise ValueError(f"expected two channels-last (B, T, H, W, C) tensors of equal shape, got {tuple(x.shape)} "
                         f"and {tuple(y.shape)}")
    if torch.is_grad_enabled() and (x.requires_grad or y.requires_grad):
        # used as a TRAINING loss (trainer.py:188, r_trainer.py:150): the same moments as differentiable torch ops
        xd, yd = x.to(torch.float64), y.to(torch.float64)
        return torch.stack([((xd - yd) ** 2).sum(dim=(2, 3)), (yd * yd).sum(dim=(2, 3)), yd.sum(dim=(2, 3))], dim=-1)
    x = x.detach().to(torch.float32).contiguous()
    y = y.detach().to(torch.float32).contiguous()
    def version(t):
        try:
            return t._version
        except RuntimeError:          # inference tensors carry no version counter (immutable outside inference mode)
            return -1
    key = (x.data_ptr(), y.data_ptr(), tuple(x.shape), version(x), version(y))
    if _cache["key"] == key:
        return _cache["val"]
    B, T, H, W, C = x.shape
    out = torch.empty((B, T, C, 3), device=x.device, dtype=torch.float64)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _abi.check(_abi.load().tante_metric_moments(x.data_ptr(), y.data_ptr(), B * T, H * W, C, out.data_ptr(), stream))
    _cache["key"], _cache["val"] = key, out
    _cache["keep"] = (x, y)          # the key holds raw pointers: keep their storages alive
    return out


class Metric(torch.nn.Module):
    """reference trainer/metrics.py:18-51."""

    def forward(self, *args, **kwargs):
        assert len(args) >= 3, "At least three arguments required (x, y, rt)"
        x, y, rt = args[:3]
        eps, n = (args[3], args[4]) if len(args) >= 5 else (0.5, 2)
        loss_spatial = self.eval(x, y, **kwargs)
        if rt is not None:
            return loss_spatial.mean() + self.eval_rt(rt, eps, n)
        return loss_spatial

    @staticmethod
    def eval(x, y, **kwargs):
        raise NotImplementedError

    @staticmethod
    def eval_rt(rt, eps=0.5, n=2.0):
        # MSE.eval_rt (metrics.py:62-80); only MSE defines it in the reference
        raise NotImplementedError


def _n_hw(x):
    return float(x.shape[2] * x.shape[3])


def _var_hw(m, n):
    """torch.std(y, dim=(H, W)) ** 2 (unbiased) from the moments."""
    return (m[..., 1] - m[..., 2] ** 2 / n) / (n - 1.0)


class MSE(Metric):
    @staticmethod
    def eval(x, y):                                  # metrics.py:53-60 -> (B, T, C)
        return (spatial_moments(x, y)[..., 0] / _n_hw(x)).to(torch.float32)

    @staticmethod
    def eval_rt(rt, eps=0.5, n=2.0):                 # metrics.py:62-80
        rt_loss = 0
        rt_avg = torch.mean(rt)
        up, down = min(1 + eps, 4), max(1 + eps, 4)
        if rt_avg < up:
            rt_loss = rt_loss + 5e-3 * (up - rt_avg) ** n
        if rt_avg > down:
            rt_loss = rt_loss + 1e-1 * (rt_avg - down) ** n
        return rt_loss


class NMSE(Metric):
    @staticmethod
    def eval(x, y, eps: float = 1e-7, norm_mode: str = "norm"):      # metrics.py:82-98
        m, n = spatial_moments(x, y), _n_hw(x)
        if norm_mode == "norm":
            norm = m[..., 1] / n
        elif norm_mode == "std":
            norm = _var_hw(m, n)
        else:
            raise ValueError(f"Invalid norm_mode: {norm_mode}")
        return (m[..., 0] / n / (norm + eps)).to(torch.float32)


class L2RE(Metric):
    @staticmethod
    def eval(x, y, eps: float = 1e-7):               # metrics.py:100-111 -> (B, C): norms over (T, H, W)
        m = spatial_moments(x, y)
        return (torch.sqrt(m[..., 0].sum(dim=1)) / (torch.sqrt(m[..., 1].sum(dim=1)) + eps)).to(torch.float32)


class NNMSE(Metric):
    @staticmethod
    def eval(x, y, eps: float = 1e-7, norm_mode: str = "norm"):      # metrics.py:114-130 -> (B, T): norm over (H, W, C)
        m, n = spatial_moments(x, y), _n_hw(x)
        C = x.shape[-1]
        if norm_mode == "norm":
            norm = m[..., 1].sum(dim=-1) / (n * C)
        elif norm_mode == "std":
            N = n * C
            norm = (m[..., 1].sum(dim=-1) - m[..., 2].sum(dim=-1) ** 2 / N) / (N - 1.0)
        else:
            raise ValueError(f"Invalid norm_mode: {norm_mode}")
        return ((m[..., 0] / n).mean(dim=-1) / (norm + eps)).to(torch.float32)


class RMSE(Metric):
    @staticmethod
    def eval(x, y):                                  # metrics.py:132-138
        return torch.sqrt(MSE.eval(x, y))


class NRMSE(Metric):
    @staticmethod
    def eval(x, y, eps: float = 1e-7, norm_mode: str = "norm"):      # metrics.py:140-148
        return torch.sqrt(NMSE.eval(x, y, eps=eps, norm_mode=norm_mode))


class VMSE(Metric):
    @staticmethod
    def eval(x, y):                                  # metrics.py:150-156
        return NMSE.eval(x, y, norm_mode="std")


class VRMSE(Metric):
    @staticmethod
    def eval(x, y):                                  # metrics.py:158-164
        return NRMSE.eval(x, y, norm_mode="std")
