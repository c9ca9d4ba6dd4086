"""Host-side drop-in contracts (no GPU): the reference's own call sequences -- eval.py:48-56 (`instantiate(cfg.evaler,
...)` then `Eval("common")`) and train.py:55-77 (`instantiate(cfg.trainer, ...)` then `train()`) -- run against
`tante_b200.R_Evaler / Evaler / R_Trainer / Trainer` with the model's device entry points mocked.  Checks the return
tuples (r_evaler.py:177, evaler.py:230), the loss-slot ordering quirk, and the checkpoint contract
(`recent.pt` / `best.pt`, `optimizer_state_dit`, resume; r_trainer.py:85-110,208-230)."""
import os

import numpy as np
import pytest
import torch

import tante_b200
from tante_b200 import rollout as R
from tante_b200 import trainer as TR


class _DS:
    metadata = None


class _DM:
    """The three attributes/methods the reference drivers touch (data/datamodule.py:29-169)."""
    train_dataset = _DS()

    def __init__(self, n_batches=3, B=2, T=4, n_out=4, H=8, W=6, D=2):
        g = torch.Generator().manual_seed(0)
        self.batches = [{"input": torch.randn(B, T, H, W, D, generator=g), "output": torch.randn(B, n_out, H, W, D, generator=g)}
                        for _ in range(n_batches)]

    def test_dataloader(self):
        return self.batches

    val_dataloader = test_dataloader

    def train_dataloader(self):       # n_steps_output = 4 target frames (tante.yaml:14), eval loaders give 8 (:15)
        return [{"input": b["input"], "output": b["output"][:, :4]} for b in self.batches]


class _MockAdaptive(torch.nn.Module):
    """Stands in for tante_b200.TANTE on a CPU box: `rollout` returns what tante_rollout would (per-call R_t rows)."""
    deg = False

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.ones(1))

    def rollout(self, window, n_roll, out_T=None, per_sample=False, sync=True):
        B, T, D, H, W = window.shape
        y = window[:, -1:].permute(0, 1, 3, 4, 2).repeat(1, n_roll, 1, 1, 1) * self.w.detach()
        steps = torch.full((B,), 2, dtype=torch.int32)
        rts = torch.zeros(n_roll, B)
        rts[0], rts[1] = 4.25, 4.75
        ns = torch.zeros(n_roll, B, dtype=torch.int32)
        ns[0], ns[1] = 4, n_roll - 4
        return y, rts, ns, steps


class _Fn:
    def __init__(self, v):
        self.v = v

    def __call__(self, x, y, rt):
        return torch.full((x.shape[0], x.shape[1], x.shape[-1]), float(self.v))


def test_r_evaler_runs_eval_py_call_sequence(tmp_path):
    ck = tmp_path / "recent.pt"
    model = _MockAdaptive()
    torch.save({"model_state_dict": model.state_dict()}, ck)
    ev = tante_b200.R_Evaler(checkpoint_folder=str(tmp_path), formatter="channels_first_default", model=model,
                             datamodule=_DM(n_out=8), eval_loss_fn1=_Fn(1), eval_loss_fn2=_Fn(2), eval_loss_fn3=_Fn(3),
                             eval_loss_fn4=_Fn(4), device="cpu", checkpoint_path=str(ck), n_steps_rollout=8, batch_size=2)
    out = ev.Eval(mode="common")                  # eval.py:56
    assert len(out) == 7                          # r_evaler.py:177
    loss, std, RT, Step, t, s_err, s_rt = out
    assert loss == [1.0, 3.0, 2.0, 4.0]           # sic: slot 2 holds eval_loss_fn3 (r_evaler.py:139-140)
    assert std == [0.0, 0.0, 0.0, 0.0]
    assert RT == pytest.approx(4.5) and Step == 4           # 2 calls x 2 samples of R_t per batch (torch.cat(Rts), :103)
    assert set(s_err) == {"min", "q1", "median", "q3", "max"} and s_err["median"] == 3.0
    assert s_rt["max"] == pytest.approx(4.5)
    y, y_ref, rts, ftime = ev.rollout_model(model, _DM(n_out=8).batches[0], ev.formatter)      # r_evaler.py:87-105
    assert y.shape == y_ref.shape == (2, 8, 8, 6, 2) and rts.shape == (4,) and ftime >= 0


def test_evaler_contract():
    model = _MockAdaptive()
    model.deg = True
    ev = tante_b200.Evaler(model=model, datamodule=_DM(n_out=4), eval_loss_fn1=_Fn(1), eval_loss_fn2=_Fn(2),
                           eval_loss_fn3=_Fn(3), eval_loss_fn4=_Fn(4), device="cpu", n_steps_rollout=4)
    loss, std, t = ev.Eval("common")              # evaler.py:186-192,230
    assert loss == [1.0, 3.0, 2.0, 4.0] and len(std) == 4
    y, y_ref, ftime = ev.rollout_model(model, _DM().batches[0], ev.formatter)                  # evaler.py:121-138
    assert y.shape == y_ref.shape
    with pytest.raises(NotImplementedError):
        tante_b200.Evaler(model=model, datamodule=_DM(), cvit=True)


class _TinyFixed(torch.nn.Module):
    """Fixed-step stand-in with the module call contract of models.TANTE(deg=True): (B,T,D,H,W) -> (B,1,D,H,W)."""
    deg = True

    def __init__(self):
        super().__init__()
        self.mix = torch.nn.Parameter(torch.tensor([0.1, 0.2, 0.3, 0.4]))

    def forward(self, x):
        return (x * self.mix.view(1, -1, 1, 1, 1)).sum(dim=1, keepdim=True)


class _TinyAdaptive(_TinyFixed):
    deg = False

    def forward(self, x, out_T):
        y = super().forward(x)
        return y, 1.2 + 0.0 * self.mix.sum().reshape(1).expand(x.shape[0])

    def rollout(self, window, n_roll, out_T=None, per_sample=False, sync=True):
        ys, moving = [], window
        for _ in range(n_roll):
            y, _ = self.forward(moving, out_T)
            moving = torch.cat([moving[:, 1:], y], dim=1)
            ys.append(y.permute(0, 1, 3, 4, 2))
        B = window.shape[0]
        return (torch.cat(ys, 1), torch.full((n_roll, B), 1.2), torch.ones(n_roll, B, dtype=torch.int32),
                torch.full((B,), n_roll, dtype=torch.int32))


def _mse(x, y, rt, eps=0.5, n=2):
    l = torch.mean((x - y) ** 2, dim=(-3, -2))
    return l if rt is None else l.mean() + 5e-3 * (1.5 - rt.mean()) ** 2


@pytest.mark.parametrize("cls,model_cls", [(TR.Trainer, _TinyFixed), (TR.R_Trainer, _TinyAdaptive)])
def test_trainer_train_checkpoints_and_resume(tmp_path, cls, model_cls):
    torch.manual_seed(0)
    model = model_cls()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2)
    kw = dict(checkpoint_folder=str(tmp_path), formatter="channels_first_default", model=model, datamodule=_DM(n_out=8),
              optimizer=opt, train_loss_fn=_mse, eval_loss_fn=_mse, max_epoch=2, device="cpu", n_steps_output=4,
              n_steps_rollout=8)
    tr = cls(**kw)
    w0 = model.mix.detach().clone()
    tr.train()                                    # train.py:77
    assert not torch.equal(w0, model.mix.detach())
    for f in ("recent.pt", "best.pt", "saved_loss.txt"):
        assert os.path.exists(tmp_path / f), f
    ck = torch.load(tmp_path / "recent.pt", weights_only=False)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dit", "validation_loss", "best_validation_loss"}
    assert ck["epoch"] == 2
    # resume (utils.set_ckpt -> checkpoint_path = recent.pt): starts at epoch 3, nothing left to do for max_epoch = 2
    model2 = model_cls()
    tr2 = cls(**{**kw, "model": model2, "optimizer": torch.optim.AdamW(model2.parameters(), lr=1e-2),
                 "checkpoint_path": str(tmp_path / "recent.pt")})
    assert tr2.starting_epoch == 3 and torch.equal(model2.mix.detach(), ck["model_state_dict"]["mix"])
    if cls is TR.R_Trainer:
        y, y_ref, rts = tr.rollout_model(model, _DM(n_out=4).batches[0], tr.formatter, "train")   # r_trainer.py:112-133
        assert y.shape == y_ref.shape and rts.shape == (2 * 4,)
        assert os.path.exists(tmp_path / "saved_rt.txt")


def test_metric_classes_mirror_reference_names_and_have_no_cpu_path():
    for name in ("MSE", "NMSE", "L2RE", "NNMSE", "RMSE", "NRMSE", "VMSE", "VRMSE"):
        assert issubclass(getattr(tante_b200, name), tante_b200.metrics.Metric)
    x = torch.zeros(1, 1, 2, 2, 1)
    with pytest.raises(RuntimeError):
        tante_b200.MSE()(x, x, None)
    # the rt penalty is host-side scalar logic (metrics.py:62-80)
    assert float(tante_b200.MSE.eval_rt(torch.tensor([1.0, 1.2]), 0.5, 2)) == pytest.approx(5e-3 * 0.4 ** 2)
    assert float(tante_b200.MSE.eval_rt(torch.tensor([5.0]), 0.5, 2)) == pytest.approx(1e-1)
