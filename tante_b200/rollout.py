"""Host-side mirrors of the reference rollout drivers for the adaptive path.

`R_Evaler` / `Evaler` keep the reference constructor kwargs and the `rollout_model`
return tuples (trainer/r_evaler.py:46-105, trainer/evaler.py:121-138) but run the whole
while-loop on the device through `TANTE.rollout` (ring-buffer window, device-side step
counters, no per-step host sync, no torch.cat window shifts, no formatter transposes).
`Eval()` / `validation_loop()` keep the reference's return tuples (7 values for `R_Evaler`, r_evaler.py:108-177;
3 for `Evaler`, evaler.py:186-230) including its loss ordering quirk (the second slot holds eval_loss_fn3), so
`eval.py:48-56` runs unchanged with `evaler._target_: tante_b200.R_Evaler`.  The metrics themselves are
`tante_b200.metrics` (one fused moments pass on the device).
"""
from __future__ import annotations

import logging
import statistics
import time
from typing import Callable, Optional

import numpy as np
import torch
from einops import rearrange

logger = logging.getLogger(__name__)


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


class DefaultChannelsLastFormatter:
    """data/datamodule.py:194-201: channels-last models; identity on both sides."""

    def __init__(self, metadata=None):
        self.metadata = metadata

    def process_input(self, data):
        return (torch.nan_to_num(data["input"]),), torch.nan_to_num(data["output"])

    def process_output(self, output):
        return output


def _five_number_summary(data):
    data = np.array(data)        # r_evaler.py:164-172
    return {"min": np.min(data), "q1": np.percentile(data, 25), "median": np.median(data),
            "q3": np.percentile(data, 75), "max": np.max(data)}


def _variance(seq):
    # statistics.variance raises for fewer than two batches (as the reference would); report nan instead of dying
    return statistics.variance(seq) if len(seq) > 1 else float("nan")


class R_Evaler:
    """Drop-in for trainer.R_Evaler (r_evaler.py:46-177): same constructor kwargs, `rollout_model` 4-tuple, `Eval(mode)`
    and the 7-tuple `validation_loop`.  `per_sample=True` (extension) gives every trajectory its own step sequence."""

    def __init__(self, checkpoint_folder: str = "", formatter: str = "channels_first_default", model=None,
                 datamodule=None, eval_loss_fn1: Optional[Callable] = None, eval_loss_fn2: Optional[Callable] = None,
                 eval_loss_fn3: Optional[Callable] = None, eval_loss_fn4: Optional[Callable] = None,
                 device=torch.device("cuda"), enable_amp: bool = False, amp_type: str = "float16",
                 checkpoint_path: str = "", n_steps_rollout: int = 8, batch_size: int = 4, rt_eps: float = 0.5,
                 rt_n: int = 2, per_sample: bool = False):
        params = dict(locals())
        for k, v in params.items():
            if k != "self" and not k.startswith("_"):
                setattr(self, k, v)
        self.device = torch.device(device)
        self.amp_type = torch.bfloat16 if amp_type == "bfloat16" else torch.float16
        self.dset_metadata = getattr(getattr(datamodule, "train_dataset", None), "metadata", None)
        if formatter != "channels_first_default":
            # DefaultChannelsLastFormatter feeds (B,T,H,W,C) models; TANTE is channels-first (tante.yaml:49,65)
            raise NotImplementedError("only channels_first_default is wired to the device rollout")
        self.formatter = DefaultChannelsFirstFormatter(self.dset_metadata)
        if checkpoint_path:           # the reference loads unconditionally (r_evaler.py:77); "" = keep the given weights
            self.load_checkpoint(checkpoint_path)

    @property
    def eval_loss_fns(self):
        return [self.eval_loss_fn1, self.eval_loss_fn2, self.eval_loss_fn3, self.eval_loss_fn4]

    def load_checkpoint(self, checkpoint_path: str):
        logger.info(f"Loading checkpoint from {checkpoint_path}")
        checkpoint = torch.load(checkpoint_path, weights_only=False)
        if self.model is not None:
            self.model.load_state_dict(checkpoint["model_state_dict"])

    def rollout_model(self, model, batch, formatter):
        moving_batch, y_ref = formatter.process_input(batch)
        moving_batch = moving_batch[0].to(self.device)
        start_time = time.time()
        y_pred_out, Rts, _, _ = rollout_eval(model, moving_batch, self.n_steps_rollout, per_sample=self.per_sample)
        forward_time = time.time() - start_time     # rollout() synchronises, so this is device time
        return y_pred_out, y_ref.to(self.device), Rts, forward_time

    def Eval(self, mode="common"):
        test_dataloader = self.datamodule.test_dataloader()
        if mode == "common":
            test_loss, std, RT, Step, time_used, summary_error, summary_rt = self.validation_loop(test_dataloader)
            logger.info(f"Test Loss: {test_loss}")
            logger.info(f"std:{std}")
            logger.info(f"rt: {RT}, Step: {Step}, Time used: {time_used}")
            logger.info(f"error: {summary_error}, rt: {summary_rt}")
            return test_loss, std, RT, Step, time_used, summary_error, summary_rt

    @torch.inference_mode()
    def validation_loop(self, dataloader):
        """(validation_loss[4], Std_error[4], RT, Step, time_used, summary_error, summary_rt), r_evaler.py:117-177."""
        self.model.eval()
        seq = [[], [], [], []]
        rt_list, step_list, time_used = [], [], []
        with torch.autocast(self.device.type, enabled=self.enable_amp, dtype=self.amp_type):
            for batch in dataloader:
                y_pred, y_ref, rts, ftime = self.rollout_model(self.model, batch, self.formatter)
                assert y_ref.shape == y_pred.shape, \
                    f"Mismatching shapes between reference {y_ref.shape} and prediction {y_pred.shape}"
                loss1 = self.eval_loss_fn1(y_pred, y_ref, None)
                loss2 = self.eval_loss_fn2(y_pred, y_ref, None)
                loss3 = self.eval_loss_fn3(y_pred, y_ref, None)
                loss4 = self.eval_loss_fn4(y_pred, y_ref, None)
                seq[0].append(loss1.mean().item())
                seq[1].append(loss3.mean().item())          # sic: r_evaler.py:139-140 store fn3 second, fn2 third
                seq[2].append(loss2.mean().item())
                seq[3].append(loss4.mean().item())
                time_used.append(ftime)
                if rts is not None:
                    rt_list.append(torch.mean(rts).item())
                    step_list.append(len(rts))
        n = len(dataloader)
        validation_loss = [sum(s) / n for s in seq]
        std_error = [_variance(s) for s in seq]
        RT = sum(rt_list) / len(rt_list) if rt_list else float("nan")
        Step = sum(step_list) / len(step_list) if step_list else float("nan")
        time_avg = sum(time_used) / len(time_used)
        summary_error = _five_number_summary(seq[1])
        summary_rt = _five_number_summary(rt_list) if rt_list else None
        return validation_loss, std_error, RT, Step, time_avg, summary_error, summary_rt


class Evaler(R_Evaler):
    """Drop-in for trainer.Evaler (evaler.py:85-230): fixed-step model, `rollout_model` 3-tuple, `validation_loop`
    3-tuple.  The CViT query-point branch (evaler.py:140-184) belongs to another model and is not supported."""

    def __init__(self, checkpoint_folder: str = "", formatter: str = "channels_first_default", model=None,
                 datamodule=None, eval_loss_fn1: Optional[Callable] = None, eval_loss_fn2: Optional[Callable] = None,
                 eval_loss_fn3: Optional[Callable] = None, eval_loss_fn4: Optional[Callable] = None,
                 device=torch.device("cuda"), enable_amp: bool = False, amp_type: str = "float16",
                 checkpoint_path: str = "", n_steps_rollout: int = 8, batch_size: int = 4, cvit: bool = False,
                 num_query_points: int = 1024):
        if cvit:
            raise NotImplementedError("cvit=True selects the CViT query-point rollout (evaler.py:140-184), not TANTE")
        super().__init__(checkpoint_folder, formatter, model, datamodule, eval_loss_fn1, eval_loss_fn2, eval_loss_fn3,
                         eval_loss_fn4, device, enable_amp, amp_type, checkpoint_path, n_steps_rollout, batch_size)
        self.cvit, self.num_query_points = cvit, num_query_points

    def rollout_model(self, model, batch, formatter):
        y, y_ref, _, ftime = super().rollout_model(model, batch, formatter)
        return y, y_ref, ftime

    def Eval(self, mode="common"):
        test_dataloader = self.datamodule.test_dataloader()
        if mode == "common":
            test_loss, std, time_used = self.validation_loop(test_dataloader)
            logger.info(f"Test Loss: {test_loss}")
            logger.info(f"std:{std}")
            logger.info(f"Time used: {time_used}")
            return test_loss, std, time_used

    @torch.inference_mode()
    def validation_loop(self, dataloader, epoch: int = 0):
        """(validation_loss[4], Std_error[4], time_used), evaler.py:194-230."""
        self.model.eval()
        seq = [[], [], [], []]
        time_used = []
        with torch.autocast(self.device.type, enabled=self.enable_amp, dtype=self.amp_type):
            for batch in dataloader:
                y_pred, y_ref, ftime = self.rollout_model(self.model, batch, self.formatter)
                assert y_ref.shape == y_pred.shape, \
                    f"Mismatching shapes between reference {y_ref.shape} and prediction {y_pred.shape}"
                seq[0].append(self.eval_loss_fn1(y_pred, y_ref, None).mean().item())
                seq[1].append(self.eval_loss_fn3(y_pred, y_ref, None).mean().item())     # sic: evaler.py:204-205
                seq[2].append(self.eval_loss_fn2(y_pred, y_ref, None).mean().item())
                seq[3].append(self.eval_loss_fn4(y_pred, y_ref, None).mean().item())
                time_used.append(ftime)
        n = len(dataloader)
        return [sum(s) / n for s in seq], [_variance(s) for s in seq], sum(time_used) / len(time_used)
