"""`FusedAdamW`: torch.optim.AdamW (reference configs/tante.yaml:38-41) with the clip of the reference drivers folded in
(`clip_grad_norm_(1.0)` trainer/trainer.py:192-193, `clip_grad_value_(1.0)` trainer/r_trainer.py:155) as ONE pass over the
library's flat gradient (tante_optimizer_step, csrc/optimizer.cuh) followed by the repack of the GEMM-ready weights.

Same constructor, `param_groups` (schedulers keep working) and `state_dict()` keys as torch.optim.AdamW: `state[p]` holds
`step`, `exp_avg`, `exp_avg_sq`, the moments being views of two flat buffers laid out like the gradient.  Needs the
parameters' gradients in the library's flat layout -- `tante_b200.trainer.GradBucket(model)` -- and fails loudly otherwise."""
from __future__ import annotations

from typing import Optional

import torch

from . import _abi

_CLIP = {None: 0, "none": 0, "norm": 1, "value": 2}


def _find_model(params):
    from .tante import live_models
    ptrs = {p.data_ptr() for p in params}
    for m in live_models():
        if {p.data_ptr() for p in m.parameters()} == ptrs:
            return m
    raise RuntimeError("FusedAdamW: the parameters are not those of one live tante_b200.TANTE module (pass model=...)")


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 amsgrad: bool = False, *, model=None, **_ignored):
        if amsgrad:
            raise ValueError("FusedAdamW does not implement amsgrad")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdamW supports a single parameter group (the reference uses one)")
        self._model = model
        self._eng = None
        self._m: Optional[torch.Tensor] = None
        self._v: Optional[torch.Tensor] = None
        self._step_t = torch.zeros((), dtype=torch.float32)      # shared by every state entry

    # ---- binding to the module's engine and flat gradient -------------------------------------------------------------
    def _bind(self):
        params = self.param_groups[0]["params"]
        if self._model is None:
            self._model = _find_model(params)
        model = self._model
        dev = params[0].device
        eng = model._engine(dev)
        named = dict(model.named_parameters())
        if {p.data_ptr() for p in params} != {named[n].data_ptr() for n in eng.names}:
            raise RuntimeError("FusedAdamW must own every parameter of the model (one group)")
        self._eng = eng
        self._m = torch.zeros(eng.grad_numel, device=dev, dtype=torch.float32)
        self._v = torch.zeros(eng.grad_numel, device=dev, dtype=torch.float32)
        self._views(named)

    def _views(self, named, load_from=None):
        eng = self._eng
        for n, off in zip(eng.names, eng.grad_offsets):
            p = named[n]
            cnt = p.numel() * (2 if p.is_complex() else 1)
            shape = tuple(p.shape) + ((2,) if p.is_complex() else ())
            m, v = self._m[off:off + cnt].view(shape), self._v[off:off + cnt].view(shape)
            if load_from is not None and p in load_from:
                old = load_from[p]
                m.copy_(torch.view_as_real(old["exp_avg"]) if old["exp_avg"].is_complex() else old["exp_avg"].reshape(shape))
                v.copy_(torch.view_as_real(old["exp_avg_sq"]) if old["exp_avg_sq"].is_complex() else old["exp_avg_sq"].reshape(shape))
            self.state[p] = {"step": self._step_t, "exp_avg": m, "exp_avg_sq": v}

    # ---- the step ----------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, *, clip: Optional[str] = None, clip_value: float = 1.0, grad_scale: float = 1.0):
        """clip = None (the caller already clipped, as the unmodified reference drivers do), "norm" or "value";
        grad_scale multiplies the gradient first (1 / world size after a SUM all-reduce)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._eng is None:
            self._bind()
        model, eng = self._model, self._eng
        g = model._flat_grad_view(eng)
        if g is None:
            raise RuntimeError("FusedAdamW: gradients are not in the library's flat layout: create tante_b200.trainer.GradBucket("
                               "model) before the first backward and clear it with GradBucket.zero(), not zero_grad(set_to_none=True)")
        grp = self.param_groups[0]
        self._step_t += 1
        stream = torch.cuda.current_stream(g.device).cuda_stream
        _abi.check(eng.lib.tante_optimizer_step(
            eng.handle, g.data_ptr(), self._m.data_ptr(), self._v.data_ptr(), float(grp["lr"]), float(grp["betas"][0]),
            float(grp["betas"][1]), float(grp["eps"]), float(grp["weight_decay"]), int(self._step_t.item()), _CLIP[clip],
            float(clip_value), float(grad_scale), None, stream))
        _abi.check(eng.lib.tante_pack_params(eng.handle, stream))
        # the masters changed behind torch's back: bump their version counters and tell the engine it is in sync
        params = grp["params"]
        try:
            torch.autograd.graph.increment_version(params)
        except TypeError:
            for p in params:
                torch.autograd.graph.increment_version(p)
        eng.mark_synced(model)
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        loaded = {p: dict(s) for p, s in self.state.items()}
        if loaded:
            any_state = next(iter(loaded.values()))
            self._step_t = torch.as_tensor(float(any_state["step"]), dtype=torch.float32).reshape(())
        if self._eng is None:
            self._bind()
        self._views(dict(self._model.named_parameters()), load_from=loaded)
