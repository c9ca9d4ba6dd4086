"""Host-side mirrors of the reference rollout drivers for the adaptive path.

`R_Evaler` / `Evaler` keep the reference constructor kwargs and the `rollout_model`
return tuples (trainer/r_evaler.py:46-105, trainer/evaler.py:121-138) but run the whole
while-loop on the device through `TANTE.rollout` (ring-buffer window, device-side step
counters, no per-step host sync, no torch.cat window shifts, no formatter transposes).
Metrics stay host-side torch one-liners (plumbing).
"""
from __future__ import annotations

import time
from typing import Callable, Optional

import torch
from einops import rearrange


class DefaultChannelsFirstFormatter:
    """data/datamodule.py:184-192."""

    def __init__(self, metadata=None):
        self.metadata = metadata

    def process_input(self, data):
        x = rearrange(data["input"], "b t ... c -> b t c ...")
        return (torch.nan_to_num(x),), torch.nan_to_num(data["output"])

    def process_output(self, output):
        return rearrange(output, "b t c ... -> b t ... c")


def rollout_eval(model, window: torch.Tensor, n_steps_rollout: int, out_T=None, per_sample: bool = False):
    """Functional form of `R_Evaler.rollout_model` (r_evaler.py:87-105) on a channels-first window.

    Returns (y_pred (B,n_roll,H,W,D), Rts, ns, steps).  `Rts` is flattened exactly like the
    reference's `torch.cat(Rts, dim=0)` -- one R_t per sample per model call, call-major --
    when per_sample=False; with per_sample=True it is sample-major like R_Trainer's (r_trainer.py:132)."""
    y, rts, ns, steps = model.rollout(window, n_steps_rollout, out_T=out_T, per_sample=per_sample)
    if model.deg:
        return y, None, ns, steps
    steps_h = steps.tolist()
    if per_sample:
        flat = torch.cat([rts[:steps_h[b], b] for b in range(len(steps_h))], dim=0)
    else:
        flat = rts[:steps_h[0]].reshape(-1)
    return y, flat, ns, steps


class R_Evaler:
    """Mirror of trainer.R_Evaler (r_evaler.py:46-177) for the hot path."""

    def __init__(self, checkpoint_folder: str = "", formatter: str = "channels_first_default", model=None,
                 datamodule=None, eval_loss_fn1: Optional[Callable] = None, eval_loss_fn2: Optional[Callable] = None,
                 eval_loss_fn3: Optional[Callable] = None, eval_loss_fn4: Optional[Callable] = None,
                 device=torch.device("cuda"), enable_amp: bool = False, amp_type: str = "float16",
                 checkpoint_path: str = "", n_steps_rollout: int = 8, batch_size: int = 4, rt_eps: float = 0.5,
                 rt_n: int = 2, per_sample: bool = False):
        self.model = model
        self.datamodule = datamodule
        self.device = torch.device(device)
        self.enable_amp = enable_amp
        self.amp_type = torch.bfloat16 if amp_type == "bfloat16" else torch.float16
        self.n_steps_rollout = n_steps_rollout
        self.per_sample = per_sample
        self.eval_loss_fns = [eval_loss_fn1, eval_loss_fn2, eval_loss_fn3, eval_loss_fn4]
        if formatter != "channels_first_default":
            raise NotImplementedError("only channels_first_default is wired to the device rollout")
        self.formatter = DefaultChannelsFirstFormatter(getattr(getattr(datamodule, "train_dataset", None), "metadata", None))
        if checkpoint_path:
            self.load_checkpoint(checkpoint_path)

    def load_checkpoint(self, checkpoint_path: str):
        checkpoint = torch.load(checkpoint_path, weights_only=False)
        self.model.load_state_dict(checkpoint["model_state_dict"])

    def rollout_model(self, model, batch, formatter):
        moving_batch, y_ref = formatter.process_input(batch)
        moving_batch = moving_batch[0].to(self.device)
        start_time = time.time()
        y_pred_out, Rts, _, _ = rollout_eval(model, moving_batch, self.n_steps_rollout, per_sample=self.per_sample)
        forward_time = time.time() - start_time     # rollout() synchronises, so this is device time
        return y_pred_out, y_ref.to(self.device), Rts, forward_time

    @torch.inference_mode()
    def validation_loop(self, dataloader):
        self.model.eval()
        seq = [[] for _ in self.eval_loss_fns]
        rt_list, step_list, time_used = [], [], []
        with torch.autocast(self.device.type, enabled=self.enable_amp, dtype=self.amp_type):
            for batch in dataloader:
                y_pred, y_ref, rts, ftime = self.rollout_model(self.model, batch, self.formatter)
                assert y_ref.shape == y_pred.shape
                for acc, fn in zip(seq, self.eval_loss_fns):
                    if fn is not None:
                        acc.append(fn(y_pred, y_ref, None).mean().item())
                time_used.append(ftime)
                if rts is not None:
                    rt_list.append(torch.mean(rts).item())
                    step_list.append(len(rts))
        n = max(len(time_used), 1)
        return ([sum(s) / n if s else None for s in seq], rt_list, step_list, sum(time_used) / n)


class Evaler(R_Evaler):
    """Mirror of trainer.Evaler.rollout_model (evaler.py:121-138): fixed-step model, no R_t."""

    def rollout_model(self, model, batch, formatter):
        y, y_ref, _, _ = super().rollout_model(model, batch, formatter)
        return y, y_ref
