"""Stock-PyTorch `nn.Module` restatement of the reference TANTE -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

Purpose: the "stock kernel" bar of SURVEY.md §2 / §8(d) -- what the reference module costs on the SAME B200 when it runs
the way the reference runs it (`.cuda()`, bf16 autocast, TF32 matmuls as utils.py:29 sets them): cuDNN convolutions,
cuBLAS linears, `nn.MultiheadAttention`'s fused path, `nn.LayerNorm`, `einops.rearrange` copies around every axial layer,
one elementwise launch per Taylor term.  /root/reference does not exist on the GPU box, so `bench.py`'s
`gpu_eager_baseline` leg times THIS module; `tests/test_oracle_golden.py` pins it to the reference's goldens (same
`state_dict` keys, same outputs), so it is the reference's op sequence, not an approximation of it.

Unlike `oracle/tante_oracle.py` (elementary tensor algebra, the parity checker) this file deliberately calls the same
`torch.nn` layers the reference is built from, each citing its reference line.  The product never imports it.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from einops import rearrange

from .tante_oracle import PATCH_MAP, OracleConfig, s_emb_init, t_emb_init, t_series


class _Conv(nn.Module):                       # RealConv2d (enc_dec_cnn.py:49-110)
    def __init__(self, cin, cout, k):
        super().__init__()
        self.k = k
        self.conv = nn.Conv2d(cin, cout, kernel_size=k, stride=k, padding=(k - 1) // 2)

    def forward(self, x):
        y = self.conv(x)
        return F.adaptive_avg_pool2d(y, (x.shape[-2] // self.k, x.shape[-1] // self.k))     # :109 (identity for k <= 2)


class _Deconv(nn.Module):                     # RealTransConv2d (enc_dec_cnn.py:113-184)
    def __init__(self, cin, cout, k):
        super().__init__()
        self.k = k
        self.deconv = nn.ConvTranspose2d(cin, cout, kernel_size=k, stride=k, padding=(k - 1) // 2)

    def forward(self, x):
        y = self.deconv(x)
        th, tw = x.shape[-2] * self.k, x.shape[-1] * self.k
        if y.shape[-2] == th and y.shape[-1] == tw:
            return y
        return F.interpolate(y, size=(th, tw), mode="bilinear", align_corners=False)             # :176-184


class _Enc(nn.Module):                        # enc_CNN (enc_dec_cnn.py:187-229)
    def __init__(self, D, C, ks):
        super().__init__()
        self.enc_conv_1, self.enc_conv_2, self.enc_conv_3 = _Conv(D, C // 4, ks[0]), _Conv(C // 4, C // 2, ks[1]), _Conv(C // 2, C, ks[2])
        self.act = nn.GELU()

    def forward(self, x):
        B, T = x.shape[:2]
        z = rearrange(x, "b t d h w -> (b t) d h w")
        z = self.act(self.enc_conv_1(z))
        z = self.act(self.enc_conv_2(z))
        z = self.enc_conv_3(z)
        return rearrange(z, "(b t) c h w -> b t h w c", b=B, t=T)


class _Dec(nn.Module):                        # dec_CNN (enc_dec_cnn.py:232-277)
    def __init__(self, D, C, ks):
        super().__init__()
        self.dec_conv_1, self.dec_conv_2, self.dec_conv_3 = _Deconv(C, C // 2, ks[2]), _Deconv(C // 2, C // 4, ks[1]), _Deconv(C // 4, D, ks[0])
        self.act = nn.GELU()

    def forward(self, x):
        B, T = x.shape[:2]
        z = rearrange(x, "b t h w c -> (b t) c h w")
        z = self.act(self.dec_conv_1(z))
        z = self.act(self.dec_conv_2(z))
        z = self.dec_conv_3(z)
        return rearrange(z, "(b t) d h w -> b t d h w", b=B, t=T)


class _Block(nn.Module):                      # TransformerBlock (attn_backbone.py:38-83)
    def __init__(self, C, n_head, dropout, mlp_ratio=1.0):
        super().__init__()
        self.ln1 = nn.LayerNorm(C)
        self.attn = nn.MultiheadAttention(C, n_head, batch_first=True, dropout=dropout, bias=True)
        self.ln2 = nn.LayerNorm(C)
        hidden = int(C * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(C, hidden), nn.GELU(approximate="tanh"), nn.Linear(hidden, C))
        self.drop = nn.Dropout(dropout)

    def forward(self, x, causal):
        h = self.ln1(x)
        mask = torch.triu(torch.ones(x.shape[1], x.shape[1], dtype=torch.bool, device=x.device), diagonal=1) if causal else None
        y, _ = self.attn(h, h, h, attn_mask=mask, need_weights=False, is_causal=causal)
        x = x + self.drop(y)
        return x + self.drop(self.mlp(self.ln2(x)))


class _Backbone(nn.Module):                   # Attn_Backbone (attn_backbone.py:88-191), axes T / H / W / L / Y / A
    def __init__(self, T, Hp, Wp, C, axes, n_head, dropout, mlp_ratio=1.0):
        super().__init__()
        self.axes = axes
        self.blocks = nn.ModuleList([_Block(C, n_head, dropout, mlp_ratio) for _ in axes])
        self.vertical_propagator = nn.Sequential(nn.Linear(Hp, Hp), nn.GELU(), nn.Linear(Hp, Hp))
        self.horizontal_propagator = nn.Sequential(nn.Linear(Wp, Wp), nn.GELU(), nn.Linear(Wp, Wp))
        self.temporal_propagator = nn.Sequential(nn.Linear(T, T), nn.GELU(), nn.Linear(T, T))

    PAT = {"T": "(b h w) t c", "H": "(b t w) h c", "W": "(b t h) w c", "L": "(b t) (h w) c", "Y": "(b w) (t h) c",
           "A": "b (t h w) c"}

    def forward(self, x):
        B, T, H, W, C = x.shape
        d = dict(b=B, t=T, h=H, w=W, c=C)
        x = rearrange(x, "b t h w c -> b t w c h")
        x = x + self.vertical_propagator(x)
        x = rearrange(x, "b t w c h -> b t h c w")
        x = x + self.horizontal_propagator(x)
        x = rearrange(x, "b t h c w -> b (h w c) t")
        x = x + self.temporal_propagator(x)
        x = rearrange(x, "b (h w c) t -> b t h w c", **d)
        for blk, axis in zip(self.blocks, self.axes):
            pat = self.PAT[axis]
            x = rearrange(x, f"b t h w c -> {pat}")
            x = blk(x, causal=axis == "T")
            x = rearrange(x, f"{pat} -> b t h w c", **d)
        return x


class _Film(nn.Module):                       # film (tante.py:203-230)
    def __init__(self, C):
        super().__init__()
        self.condition_to_scale = nn.Sequential(nn.Linear(1, C // 2), nn.ReLU(), nn.Linear(C // 2, C))
        self.condition_to_shift = nn.Sequential(nn.Linear(1, C // 2), nn.ReLU(), nn.Linear(C // 2, C))

    def forward(self, x, t):
        scale, shift = self.condition_to_scale(t[..., None]), self.condition_to_shift(t[..., None])
        if x.dim() == 3:
            scale, shift = scale[:, None, :], shift[:, None, :]
        else:
            scale, shift = scale[None, :, None, None, :], shift[None, :, None, None, :]
        return x + (x * scale + shift)


class _Interprator(nn.Module):                # interprator (tante.py:178-201)
    def __init__(self, C, L):
        super().__init__()
        self.L = L
        self.interprete = nn.Sequential(nn.Linear(C, C // 2), nn.ReLU(), nn.Linear(C // 2, C // 4), nn.ReLU(), nn.Linear(C // 4, 1))

    def forward(self, x, out_T):
        t = self.interprete(x).reshape(-1, self.L)
        td = t.detach()
        t = t + torch.relu(-td) - torch.relu(td - (out_T - 1))
        return torch.mean(t, dim=1) + 1.001


class EagerTANTE(nn.Module):
    """reference models/tante.py:37-176 (with the F5 repair of the adaptive branch)."""

    def __init__(self, cfg: OracleConfig, dropout: float = 0.0):
        super().__init__()
        self.cfg = cfg
        C, D, T = cfg.embed_dim, cfg.n_fields, cfg.in_T
        ks = PATCH_MAP[cfg.patch_scale]
        self.decoders = nn.ModuleList()
        self.encoder = _Enc(D, C, ks)
        for _ in range(cfg.taylor_order):
            self.decoders.append(_Dec(D, C, ks))
        self.blocks = nn.ModuleList([_Backbone(T, cfg.Hp, cfg.Wp, C, seg, cfg.n_head, dropout, cfg.mlp_ratio) for seg in cfg.segments])
        self.t_emb = nn.Parameter(t_emb_init(C, T))
        self.s_emb = nn.Parameter(s_emb_init(C, cfg.Hp, cfg.Wp))
        self.t_encode = _Film(C)
        if not cfg.deg:
            self.interprators = nn.ModuleList([_Interprator(C, cfg.Hp * cfg.Wp) for _ in range(cfg.taylor_order)])
            self.modifiers = nn.ModuleList([_Film(C) for _ in range(cfg.taylor_order)])
        self.register_buffer("t_seq", t_series(T, cfg.frame_interval, torch.float32), persistent=False)

    def forward(self, inp, out_T=1):
        cfg = self.cfg
        if inp.shape[1] != cfg.in_T:
            inp = inp[:, -cfg.in_T:]
        B = inp.shape[0]
        x = self.encoder(inp)
        x = self.t_encode(x, self.t_seq)
        x = x + self.s_emb
        x = rearrange(x, "b t h w c -> (b h w) t c")
        x = x + self.t_emb
        x = rearrange(x, "(b h w) t c -> b t h w c", b=B, h=cfg.Hp, w=cfg.Wp)
        derivs, rts = [], []
        for i in range(cfg.taylor_order):
            x = self.blocks[i](x)
            d = x[:, -1:]
            if not cfg.deg:
                dl = rearrange(d, "b 1 h w c -> b (h w) c")
                rt = self.interprators[i](dl, out_T)
                rts.append(rt)
                d = rearrange(self.modifiers[i](dl, rt), "b (h w) c -> b 1 h w c", h=cfg.Hp, w=cfg.Wp)
            derivs.append(self.decoders[i](d))
        R_t = None
        if not cfg.deg:
            R_t = torch.mean(torch.stack(rts, dim=1), dim=1)
        n = cfg.output_length if cfg.deg else math.floor(R_t[0])            # the reference's host sync (tante.py:163)
        outs = []
        for i in range(1, n + 1):
            out = 0
            for order in range(1, cfg.taylor_order + 1):
                out = out + derivs[order - 1] * (i * cfg.frame_interval) ** order / math.factorial(order)
            outs.append(out + inp[:, -1:])
        outs = torch.cat(outs, dim=1)
        return outs if cfg.deg else (outs, R_t)


def eager_rollout(model: EagerTANTE, window, n_steps: int, out_T=None):
    """R_Evaler.rollout_model / Evaler.rollout_model loop (r_evaler.py:87-105, evaler.py:121-138) around the module."""
    out_T = n_steps if out_T is None else out_T
    moving, ys, cum = window, [], 0
    while cum < n_steps:
        y = model(moving) if model.cfg.deg else model(moving, out_T)[0]
        cum += y.shape[1]
        if cum < n_steps:
            moving = torch.cat([moving[:, y.shape[1]:], y], dim=1)
        ys.append(rearrange(y, "b t c h w -> b t h w c"))
    return torch.cat(ys, dim=1)[:, :n_steps]
