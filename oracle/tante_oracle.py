"""CPU oracle for the TANTE hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional, plain-PyTorch (CPU, fp32 or fp64) restatement of the reference's
neural-Taylor forward and adaptive rollout.  Only `tests/`, `bench.py`'s
`cpu_baseline` / `--impl reference` legs and `__graft_entry__.smoke()` may import
this file; the product (`tante_b200/`) never does and has no CPU fallback.

Every function cites the reference lines (under /root/reference) it restates.
The arithmetic of the reference lives in stock PyTorch ATen (`nn.Conv2d`,
`nn.ConvTranspose2d`, `nn.MultiheadAttention`, `nn.LayerNorm`, `nn.Linear`,
`nn.GELU`; torch is pinned only as unversioned `torch` in requirements.txt:1,
2.11.0 here), so the restatement is written in elementary tensor algebra
(matmul / softmax / erf / tanh) over a reference-layout `state_dict`.

Pinning: the reference ships no tests or golden vectors, so this
oracle is pinned against the live reference module imported in the build
container (`oracle/ref_shim.py`; the `deg=False` branch needs the one-line-class
repair documented there).  `oracle/make_golden.py` wrote `tests/golden/*.npz`
from the *reference* outputs; `tests/test_oracle_golden.py` checks this file
against them everywhere, `tests/test_oracle_vs_reference.py` checks it against
the live reference when /root/reference exists.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

# reference models/enc_dec_cnn.py:39-46
PATCH_MAP = {64: (4, 4, 4), 32: (4, 4, 2), 16: (4, 2, 2), 8: (2, 2, 2), 4: (2, 2, 1), 2: (2, 1, 1)}
# reference models/enc_dec_fno.py:39-46 (two patch stages, spectral layers in between)
PATCH_MAP_FNO = {64: (8, 8), 32: (8, 4), 16: (4, 4), 8: (4, 2), 4: (2, 2), 2: (2, 1)}


@dataclass
class OracleConfig:
    """Mirror of the TANTE ctor arguments that matter (reference models/tante.py:38-60)."""
    in_T: int = 4
    n_fields: int = 4
    H: int = 128
    W: int = 384
    taylor_order: int = 1
    frame_interval: float = 1.0
    output_length: int = 1
    attn_axes: str = "THWTHWTHW"
    n_head: int = 8
    embed_dim: int = 256
    patch_scale: int = 8
    deg: bool = True
    enc_dec_type: str = "cnn"      # 'cnn' (enc_dec_cnn.py) | 'fno' (enc_dec_fno.py)
    modes1: int = 32
    modes2: int = 32
    mlp_ratio: float = 1.0         # hidden width of the block MLP = int(embed_dim * mlp_ratio) (attn_backbone.py:52)
    expanded_channel: int = 128    # embedding width of the channel-attention blocks, axis 'C' (attn_backbone.py:124-130)
    overlap_ratio: float = 0.0     # patch overlap of the cnn encoder / decoder: stride = max(1, round(k * (1 - r))) (enc_dec_cnn.py:64-66)

    @property
    def Hp(self):
        return self.H // self.patch_scale

    @property
    def Wp(self):
        return self.W // self.patch_scale

    @property
    def segments(self) -> List[str]:
        axes = self.attn_axes.replace(" ", "")
        segs = [p.strip() for p in axes.split("-")]
        if len(segs) != self.taylor_order:
            raise ValueError("Block allocation doesn't match expansion order")
        return segs


# --------------------------------------------------------------------------
# elementary pieces
# --------------------------------------------------------------------------
def gelu_erf(x):
    # nn.GELU() default, used in enc/dec (enc_dec_cnn.py:215,261) and propagators (attn_backbone.py:111-119)
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def gelu_tanh(x):
    # nn.GELU(approximate="tanh") in the transformer MLP (attn_backbone.py:52-56)
    k = math.sqrt(2.0 / math.pi)
    return 0.5 * x * (1.0 + torch.tanh(k * (x + 0.044715 * x * x * x)))


def linear(x, w, b=None):
    y = x @ w.transpose(-1, -2)
    return y if b is None else y + b


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def t_series(T: int, fi: float, dtype):
    # reference models/tante.py:279-285 -- note the duplicated zero: [-(T-2)..-1,-0,0]*fi
    seq = [0.0] + [-i * fi for i in range(T - 1)]
    seq.reverse()
    return torch.tensor(seq, dtype=dtype)


def sincos_1d(embed_dim: int, pos: torch.Tensor) -> torch.Tensor:
    # reference models/tante.py:232-242
    omega = torch.arange(embed_dim // 2, dtype=torch.float32)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = torch.einsum("m,d->md", pos.reshape(-1), omega)
    return torch.cat([torch.sin(out), torch.cos(out)], dim=1)


def t_emb_init(C: int, T: int) -> torch.Tensor:
    # reference models/tante.py:243-249
    return sincos_1d(C, torch.arange(T, dtype=torch.float32)).unsqueeze(0)


def s_emb_init(C: int, Hp: int, Wp: int) -> torch.Tensor:
    # reference models/tante.py:251-276 (meshgrid "w goes first", then a raw reshape to (2,1,H,W))
    grid_h = torch.arange(Hp, dtype=torch.float32)
    grid_w = torch.arange(Wp, dtype=torch.float32)
    gw, gh = torch.meshgrid(grid_w, grid_h, indexing="ij")
    grid = torch.stack([gh, gw], dim=0).reshape(2, 1, Hp, Wp)
    emb_h = sincos_1d(C // 2, grid[0])
    emb_w = sincos_1d(C // 2, grid[1])
    return torch.cat([emb_h, emb_w], dim=1).view(Hp, Wp, C).unsqueeze(0)


# --------------------------------------------------------------------------
# encoder / decoder (reference models/enc_dec_cnn.py)
# --------------------------------------------------------------------------
def patch_stride(k: int, overlap_ratio: float) -> int:
    """enc_dec_cnn.py:64-66 / 130-132: Python's round() (half to even), at least 1."""
    return max(1, int(round(k * (1.0 - overlap_ratio))))


def _patch_conv(x, w, b, k: int, s: Optional[int] = None):
    """RealConv2d forward (enc_dec_cnn.py:96-110): kernel k, stride s (k without overlap), pad (k-1)//2, then
    adaptive_avg_pool2d to (H//k, W//k) (an identity for s = k in {1,2,4} on divisible sizes)."""
    pad = (k - 1) // 2
    s = k if s is None else s
    if s != k:
        y = torch.nn.functional.conv2d(x, w, b, stride=s, padding=pad)
        return torch.nn.functional.adaptive_avg_pool2d(y, (x.shape[-2] // k, x.shape[-1] // k))
    if pad == 0:
        Bn, Ci, H, W = x.shape
        xp = x.reshape(Bn, Ci, H // k, k, W // k, k)
        y = torch.einsum("bcidje,ocde->boij", xp, w) + b[None, :, None, None]
        return y
    y = torch.nn.functional.conv2d(x, w, b, stride=k, padding=pad)
    return torch.nn.functional.adaptive_avg_pool2d(y, (x.shape[-2] // k, x.shape[-1] // k))


def _patch_deconv(x, w, b, k: int, s: Optional[int] = None):
    """RealTransConv2d forward (enc_dec_cnn.py:162-184): ConvTranspose2d kernel k stride s (k without overlap) pad (k-1)//2,
    bilinear(align_corners=False) resize to (k*H, k*W) when the deconv misses the patch grid (k=4, or any overlap)."""
    pad = (k - 1) // 2
    s = k if s is None else s
    Bn, Ci, H, W = x.shape
    if s != k:
        y = torch.nn.functional.conv_transpose2d(x, w, b, stride=s, padding=pad)
        if y.shape[-2] != H * k or y.shape[-1] != W * k:
            y = torch.nn.functional.interpolate(y, size=(H * k, W * k), mode="bilinear", align_corners=False)
        return y
    if pad == 0:
        y = torch.einsum("bcij,code->boidje", x, w).reshape(Bn, w.shape[1], H * k, W * k)
        return y + b[None, :, None, None]
    y = torch.nn.functional.conv_transpose2d(x, w, b, stride=k, padding=pad)
    if y.shape[-2] != H * k or y.shape[-1] != W * k:
        y = torch.nn.functional.interpolate(y, size=(H * k, W * k), mode="bilinear", align_corners=False)
    return y


def spectral_layer(sd, p: str, x, modes1: int, modes2: int):
    """SpectralLayer.forward (enc_dec_fno.py:184-222): rfft2 (ortho) -> the low modes (top and bottom m1 rows, first m2
    columns) mixed across channels by a complex weight -> irfft2, plus a 1x1 convolution.  x (N, Cin, H, W)."""
    N, Cin, H, W = x.shape
    w = sd[p + "weight"]                                            # complex (Cin, Cout, modes1, modes2)
    x_ft = torch.fft.rfft2(x, dim=(-2, -1), norm="ortho")
    Wf = x_ft.shape[-1]
    m1, m2 = min(modes1, H), min(modes2, Wf)
    out_ft = torch.zeros(N, w.shape[1], H, Wf, dtype=x_ft.dtype)
    wc = w[:, :, :m1, :m2].to(x_ft.dtype)
    out_ft[:, :, :m1, :m2] = torch.einsum("bcij,coij->boij", x_ft[:, :, :m1, :m2], wc)
    out_ft[:, :, -m1:, :m2] = torch.einsum("bcij,coij->boij", x_ft[:, :, -m1:, :m2], wc)
    y = torch.fft.irfft2(out_ft, s=(H, W), dim=(-2, -1), norm="ortho")
    s = torch.einsum("bchw,oc->bohw", x, sd[p + "w0.weight"][:, :, 0, 0]) + sd[p + "w0.bias"][None, :, None, None]
    return s + y


def encoder_fno(sd, cfg: OracleConfig, x):
    """enc_FNO.forward (enc_dec_fno.py:254-271)."""
    B, T, D, H, W = x.shape
    ps = PATCH_MAP_FNO[cfg.patch_scale]
    z = x.reshape(B * T, D, H, W)
    z = gelu_erf(spectral_layer(sd, "encoder.enc_spectral_1.", z, cfg.modes1, cfg.modes2))
    st = [patch_stride(k, cfg.overlap_ratio) for k in ps]
    z = gelu_erf(_patch_conv(z, sd["encoder.enc_conv_1.conv.weight"], sd["encoder.enc_conv_1.conv.bias"], ps[0], st[0]))
    z = gelu_erf(spectral_layer(sd, "encoder.enc_spectral_2.", z, cfg.modes1 // ps[0], cfg.modes2 // ps[0]))
    z = _patch_conv(z, sd["encoder.enc_conv_2.conv.weight"], sd["encoder.enc_conv_2.conv.bias"], ps[1], st[1])
    return z.reshape(B, T, z.shape[1], z.shape[2], z.shape[3]).permute(0, 1, 3, 4, 2).contiguous()


def decoder_fno(sd, cfg: OracleConfig, k: int, d):
    """dec_FNO.forward (enc_dec_fno.py:303-323) on the last-frame latent: (B,Hp,Wp,C) -> (B,D,H,W)."""
    ps = PATCH_MAP_FNO[cfg.patch_scale]
    p = f"decoders.{k}."
    z = d.permute(0, 3, 1, 2)
    st = [patch_stride(kk, cfg.overlap_ratio) for kk in ps]
    z = gelu_erf(_patch_deconv(z, sd[p + "dec_conv_1.deconv.weight"], sd[p + "dec_conv_1.deconv.bias"], ps[1], st[1]))
    z = gelu_erf(spectral_layer(sd, p + "dec_spectral_1.", z, cfg.modes1 // ps[0], cfg.modes2 // ps[0]))
    z = gelu_erf(_patch_deconv(z, sd[p + "dec_conv_2.deconv.weight"], sd[p + "dec_conv_2.deconv.bias"], ps[0], st[0]))
    return spectral_layer(sd, p + "dec_spectral_2.", z, cfg.modes1, cfg.modes2)


def encoder(sd, cfg: OracleConfig, x):
    """enc_CNN.forward (enc_dec_cnn.py:217-229): (B,T,D,H,W) -> (B,T,Hp,Wp,C)."""
    if cfg.enc_dec_type == "fno":
        return encoder_fno(sd, cfg, x)
    B, T, D, H, W = x.shape
    ks = PATCH_MAP[cfg.patch_scale]
    z = x.reshape(B * T, D, H, W)
    st = [patch_stride(k, cfg.overlap_ratio) for k in ks]
    z = gelu_erf(_patch_conv(z, sd["encoder.enc_conv_1.conv.weight"], sd["encoder.enc_conv_1.conv.bias"], ks[0], st[0]))
    z = gelu_erf(_patch_conv(z, sd["encoder.enc_conv_2.conv.weight"], sd["encoder.enc_conv_2.conv.bias"], ks[1], st[1]))
    z = _patch_conv(z, sd["encoder.enc_conv_3.conv.weight"], sd["encoder.enc_conv_3.conv.bias"], ks[2], st[2])
    return z.reshape(B, T, z.shape[1], z.shape[2], z.shape[3]).permute(0, 1, 3, 4, 2).contiguous()


def decoder(sd, cfg: OracleConfig, k: int, d):
    """dec_CNN.forward (enc_dec_cnn.py:263-277) on the last-frame latent: (B,Hp,Wp,C) -> (B,D,H,W)."""
    if cfg.enc_dec_type == "fno":
        return decoder_fno(sd, cfg, k, d)
    ks = PATCH_MAP[cfg.patch_scale]
    p = f"decoders.{k}."
    z = d.permute(0, 3, 1, 2)
    st = [patch_stride(kk, cfg.overlap_ratio) for kk in ks]
    z = gelu_erf(_patch_deconv(z, sd[p + "dec_conv_1.deconv.weight"], sd[p + "dec_conv_1.deconv.bias"], ks[2], st[2]))
    z = gelu_erf(_patch_deconv(z, sd[p + "dec_conv_2.deconv.weight"], sd[p + "dec_conv_2.deconv.bias"], ks[1], st[1]))
    z = _patch_deconv(z, sd[p + "dec_conv_3.deconv.weight"], sd[p + "dec_conv_3.deconv.bias"], ks[0], st[0])
    return z


# --------------------------------------------------------------------------
# FiLM / interprator (reference models/tante.py:178-230)
# --------------------------------------------------------------------------
def film_scale_shift(sd, prefix: str, t):
    """film.condition_to_{scale,shift} (tante.py:206-220): t (N,) -> scale, shift (N,C)."""
    t = t[..., None]
    out = []
    for name in ("condition_to_scale", "condition_to_shift"):
        h = torch.relu(linear(t, sd[f"{prefix}{name}.0.weight"], sd[f"{prefix}{name}.0.bias"]))
        out.append(linear(h, sd[f"{prefix}{name}.2.weight"], sd[f"{prefix}{name}.2.bias"]))
    return out[0], out[1]


def interprator(sd, k: int, d, out_T):
    """interprator.forward (tante.py:191-201): d (B,L,C) -> rt (B,).
    Forward value of the straight-through clamp is clamp(t, 0, out_T-1)."""
    p = f"interprators.{k}.interprete."
    h = torch.relu(linear(d, sd[p + "0.weight"], sd[p + "0.bias"]))
    h = torch.relu(linear(h, sd[p + "2.weight"], sd[p + "2.bias"]))
    t = linear(h, sd[p + "4.weight"], sd[p + "4.bias"]).reshape(d.shape[0], -1)
    td = t.detach()
    t = t + torch.relu(-td) - torch.relu(td - (out_T - 1))
    return t.mean(dim=1) + 1.001


# --------------------------------------------------------------------------
# backbone (reference models/attn_backbone.py)
# --------------------------------------------------------------------------
def transformer_block(sd, p: str, x, n_head: int, causal: bool, drop=None):
    """TransformerBlock.forward (attn_backbone.py:59-83) on (N, S, C) sequences.  `drop` (train mode, dropout > 0): explicit
    multipliers {0, 1/(1-p)} -- "attn" (N, heads, S, S) on the softmax probabilities (nn.MultiheadAttention(dropout=p), :47),
    "res1" / "res2" (N, S, C) on the two residual branches (self.drop, :81-83)."""
    N, S, C = x.shape
    hd = C // n_head
    h = layer_norm(x, sd[p + "ln1.weight"], sd[p + "ln1.bias"])
    qkv = linear(h, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
    q, k, v = qkv.split(C, dim=-1)
    q = q.reshape(N, S, n_head, hd).transpose(1, 2)
    k = k.reshape(N, S, n_head, hd).transpose(1, 2)
    v = v.reshape(N, S, n_head, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd))
    if causal:  # causal_mask (attn_backbone.py:35-36): True above the diagonal = masked
        m = torch.triu(torch.ones(S, S, dtype=torch.bool), diagonal=1)
        s = s.masked_fill(m, float("-inf"))
    pr = torch.softmax(s, dim=-1)
    if drop is not None:
        pr = pr * drop["attn"]
    a = pr @ v
    a = a.transpose(1, 2).reshape(N, S, C)
    y = linear(a, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
    x = x + (y if drop is None else y * drop["res1"])
    h = layer_norm(x, sd[p + "ln2.weight"], sd[p + "ln2.bias"])
    h = gelu_tanh(linear(h, sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"]))
    y = linear(h, sd[p + "mlp.2.weight"], sd[p + "mlp.2.bias"])
    return x + (y if drop is None else y * drop["res2"])


def _axis_mlp(sd, p: str, v):
    return linear(gelu_erf(linear(v, sd[p + "0.weight"], sd[p + "0.bias"])), sd[p + "2.weight"], sd[p + "2.bias"])


def backbone(sd, cfg: OracleConfig, k: int, axes: str, x, drop_fn=None):
    """Attn_Backbone.forward (attn_backbone.py:134-191): x (B,T,Hp,Wp,C).  `drop_fn(k, i, tok)` (tests of the training
    dropout): returns the explicit dropout multipliers of layer i for sequences whose tokens are `tok` (N, S) -- indices
    into the (B,T,Hp,Wp) token order -- see transformer_block."""
    B, T, H, W, C = x.shape
    p = f"blocks.{k}."
    tok_all = torch.arange(B * T * H * W).reshape(B, T, H, W)
    # three residual axis MLPs, h then w then t (:140-146)
    v = x.permute(0, 1, 3, 4, 2)
    x = (v + _axis_mlp(sd, p + "vertical_propagator.", v)).permute(0, 1, 4, 2, 3)
    v = x.permute(0, 1, 2, 4, 3)
    x = (v + _axis_mlp(sd, p + "horizontal_propagator.", v)).permute(0, 1, 2, 4, 3)
    v = x.permute(0, 2, 3, 4, 1)
    x = (v + _axis_mlp(sd, p + "temporal_propagator.", v)).permute(0, 4, 1, 2, 3)
    channel_index = 0
    for i, axis in enumerate(axes):
        bp = f"{p}blocks.{i}."
        if axis == "C":      # (b t h w) c 1 -> lift 1 -> E -> block over the C channel tokens -> last feature (:184-189)
            if drop_fn is not None:
                raise ValueError("explicit dropout masks are not defined for the channel axis")
            s = x.reshape(B * T * H * W, C, 1)
            s = _axis_mlp(sd, f"{p}channel_blocks.{channel_index}.", s)
            channel_index += 1
            s = transformer_block(sd, bp, s, cfg.n_head, False)[..., -1]
            x = s.reshape(B, T, H, W, C)
        elif axis == "T":    # (b h w) t c, causal (:149-152)
            s = x.permute(0, 2, 3, 1, 4).reshape(B * H * W, T, C)
            dr = None if drop_fn is None else drop_fn(k, i, tok_all.permute(0, 2, 3, 1).reshape(B * H * W, T))
            s = transformer_block(sd, bp, s, cfg.n_head, True, dr)
            x = s.reshape(B, H, W, T, C).permute(0, 3, 1, 2, 4)
        elif axis == "H":    # (b t w) h c (:154-157)
            s = x.permute(0, 1, 3, 2, 4).reshape(B * T * W, H, C)
            dr = None if drop_fn is None else drop_fn(k, i, tok_all.permute(0, 1, 3, 2).reshape(B * T * W, H))
            s = transformer_block(sd, bp, s, cfg.n_head, False, dr)
            x = s.reshape(B, T, W, H, C).permute(0, 1, 3, 2, 4)
        elif axis == "W":    # (b t h) w c (:159-162)
            s = x.reshape(B * T * H, W, C)
            dr = None if drop_fn is None else drop_fn(k, i, tok_all.reshape(B * T * H, W))
            s = transformer_block(sd, bp, s, cfg.n_head, False, dr)
            x = s.reshape(B, T, H, W, C)
        elif axis == "L":    # (b t) (h w) c (:164-167)
            s = x.reshape(B * T, H * W, C)
            dr = None if drop_fn is None else drop_fn(k, i, tok_all.reshape(B * T, H * W))
            s = transformer_block(sd, bp, s, cfg.n_head, False, dr)
            x = s.reshape(B, T, H, W, C)
        elif axis == "Y":    # (b w) (t h) c (:169-172)
            s = x.permute(0, 3, 1, 2, 4).reshape(B * W, T * H, C)
            dr = None if drop_fn is None else drop_fn(k, i, tok_all.permute(0, 3, 1, 2).reshape(B * W, T * H))
            s = transformer_block(sd, bp, s, cfg.n_head, False, dr)
            x = s.reshape(B, W, T, H, C).permute(0, 2, 3, 1, 4)
        elif axis == "A":    # b (t h w) c (:179-182)
            s = x.reshape(B, T * H * W, C)
            dr = None if drop_fn is None else drop_fn(k, i, tok_all.reshape(B, T * H * W))
            s = transformer_block(sd, bp, s, cfg.n_head, False, dr)
            x = s.reshape(B, T, H, W, C)
        else:
            raise ValueError(f"axis {axis!r} not covered by the oracle")
        x = x.contiguous()
    return x


# --------------------------------------------------------------------------
# TANTE.forward (reference models/tante.py:125-176, with the F5 repair)
# --------------------------------------------------------------------------
def embed(sd, cfg: OracleConfig, x):
    """encoder + t_encode FiLM + s_emb + t_emb (tante.py:132-141)."""
    dtype = x.dtype
    z = encoder(sd, cfg, x)
    tseq = t_series(cfg.in_T, cfg.frame_interval, dtype)
    scale, shift = film_scale_shift(sd, "t_encode.", tseq)          # (T,C)
    z = z + (z * scale[None, :, None, None, :] + shift[None, :, None, None, :])
    z = z + sd["s_emb"]
    z = z + sd["t_emb"][0][None, :, None, None, :]
    return z


def taylor_coefs(K: int, n: int, fi: float):
    """(i*fi)**k / k! for i=1..n, k=1..K (tante.py:165-169)."""
    return [[(i * fi) ** k / math.factorial(k) for k in range(1, K + 1)] for i in range(1, n + 1)]


def forward(sd: Dict[str, torch.Tensor], cfg: OracleConfig, inp: torch.Tensor, out_T=1,
            per_sample: bool = False, return_parts: bool = False, drop_fn=None):
    """One TANTE step.

    Returns frames (B,n,D,H,W) [and R_t (B,) when deg=False].  With `per_sample=False`
    n = floor(R_t[0]) governs the whole batch exactly as the reference does
    (tante.py:163); with `per_sample=True` a list of per-sample frame tensors is
    returned (each sample behaves like a reference B=1 call).
    """
    if inp.shape[1] != cfg.in_T:
        inp = inp[:, -cfg.in_T:]
    B = inp.shape[0]
    x = embed(sd, cfg, inp)
    derivs, rts = [], []
    for k, axes in enumerate(cfg.segments):
        x = backbone(sd, cfg, k, axes, x, drop_fn)
        d = x[:, -1]                                               # (B,Hp,Wp,C)
        if not cfg.deg:
            dl = d.reshape(B, -1, d.shape[-1])
            rt = interprator(sd, k, dl, out_T)
            rts.append(rt)
            scale, shift = film_scale_shift(sd, f"modifiers.{k}.", rt)
            dl = dl + (dl * scale[:, None, :] + shift[:, None, :])
            d = dl.reshape(d.shape)
        derivs.append(decoder(sd, cfg, k, d))                      # (B,D,H,W)
    R_t = None
    if not cfg.deg:
        R_t = torch.stack(rts, dim=1).mean(dim=1)
    u0 = inp[:, -1]

    def emit(dlist, u, n):
        outs = []
        for i in range(1, n + 1):
            acc = 0
            for order in range(1, cfg.taylor_order + 1):
                acc = acc + dlist[order - 1] * (i * cfg.frame_interval) ** order / math.factorial(order)
            outs.append(acc + u)
        return torch.stack(outs, dim=1)

    if per_sample and not cfg.deg:
        ns = [int(math.floor(float(R_t[b].detach()))) for b in range(B)]
        frames = [emit([d[b:b + 1] for d in derivs], u0[b:b + 1], ns[b]) for b in range(B)]
    else:
        n = cfg.output_length if cfg.deg else int(math.floor(float(R_t[0].detach())))
        frames = emit(derivs, u0, n)
    if return_parts:
        return frames, R_t, dict(latent=x, derivatives=derivs, rts=rts)
    if cfg.deg:
        return frames
    return frames, R_t


# --------------------------------------------------------------------------
# rollouts (reference trainer/r_evaler.py:87-105, trainer/r_trainer.py:112-133,
#           trainer/evaler.py:121-138 / trainer/trainer.py:144-159)
# --------------------------------------------------------------------------
def rollout_eval(sd, cfg: OracleConfig, window: torch.Tensor, n_steps_rollout: int, out_T=None):
    """R_Evaler.rollout_model: whole batch, n governed by sample 0, out_T=n_steps_rollout.
    window (B,T,D,H,W) channels-first.  Returns y (B,n_roll,H,W,D) channels-last,
    Rts (m*B,) concatenated like the reference, and the list of n per model call."""
    out_T = n_steps_rollout if out_T is None else out_T
    moving = window
    ys, Rts, ns = [], [], []
    cum = 0
    while cum < n_steps_rollout:
        if cfg.deg:
            y = forward(sd, cfg, moving)
            rt = None
        else:
            y, rt = forward(sd, cfg, moving, out_T)
        n = y.shape[1]
        cum += n
        if cum < n_steps_rollout:
            moving = torch.cat([moving[:, n:], y], dim=1)
        ys.append(y.permute(0, 1, 3, 4, 2))      # DefaultChannelsFirstFormatter.process_output
        ns.append(n)
        if rt is not None:
            Rts.append(rt)
    y = torch.cat(ys, dim=1)[:, :n_steps_rollout]
    return y, (torch.cat(Rts, dim=0) if Rts else None), ns


def rollout_per_sample(sd, cfg: OracleConfig, window: torch.Tensor, n_steps: int, out_T):
    """R_Trainer.rollout_model: per-sample (B=1) while-loops; the per-trajectory semantics
    the B200 rollout engine implements for batches.  Returns y (B,n,H,W,D), Rts (m,), ns per sample."""
    outs, Rts, all_ns = [], [], []
    for b in range(window.shape[0]):
        y, r, ns = rollout_eval(sd, cfg, window[b:b + 1], n_steps, out_T)
        outs.append(y)
        all_ns.append(ns)
        if r is not None:
            Rts.append(r)
    return torch.cat(outs, dim=0), (torch.cat(Rts, dim=0) if Rts else None), all_ns


# --------------------------------------------------------------------------
# loss / metrics (reference trainer/metrics.py)
# --------------------------------------------------------------------------
def mse_eval(x, y):
    # MSE.eval (metrics.py:53-60): mean over (H,W) of channels-last (B,T,H,W,C) -> (B,T,C)
    return torch.mean((x - y) ** 2, dim=(-3, -2))


def rt_penalty(rt, eps=0.5, n=2):
    # MSE.eval_rt (metrics.py:62-80)
    beta1, beta2 = 5e-3, 1e-1
    loss = 0
    avg = torch.mean(rt)
    up, down = min(1 + eps, 4), max(1 + eps, 4)
    if avg < up:
        loss = loss + beta1 * (up - avg) ** n
    if avg > down:
        loss = loss + beta2 * (avg - down) ** n
    return loss


def train_loss(y_pred, y_ref, rts, eps=0.5, n=2):
    # Metric.forward (metrics.py:19-41) as called at r_trainer.py:150
    l = mse_eval(y_pred, y_ref)
    if rts is None:
        return l.mean()   # trainer.py:186
    return l.mean() + rt_penalty(rts, eps, n)


def l2re_eval(x, y, eps=1e-7):
    # L2RE.eval (metrics.py:100-111)
    B = x.shape[0]
    C = x.shape[-1]
    xf, yf = x.reshape(B, -1, C), y.reshape(B, -1, C)
    return torch.linalg.vector_norm(xf - yf, dim=1) / (torch.linalg.vector_norm(yf, dim=1) + eps)


def nmse_eval(x, y, eps=1e-7, norm_mode="norm"):
    # NMSE.eval (metrics.py:82-98)
    if norm_mode == "norm":
        norm = torch.mean(y ** 2, dim=(-3, -2))
    else:
        norm = torch.std(y, dim=(-3, -2)) ** 2
    return mse_eval(x, y) / (norm + eps)


def nnmse_eval(x, y, eps=1e-7):
    # NNMSE.eval (metrics.py:114-130), norm_mode="norm": normaliser over (H, W, C)
    return torch.mean(mse_eval(x, y), dim=-1) / (torch.mean(y ** 2, dim=(-3, -2, -1)) + eps)


def vrmse_eval(x, y):
    # VRMSE.eval -> NRMSE.eval(norm_mode="std") -> sqrt(NMSE.eval(norm_mode="std")) (metrics.py:140-164)
    return torch.sqrt(nmse_eval(x, y, norm_mode="std"))


# --------------------------------------------------------------------------
# deterministic synthetic weights (shared by goldens and parity tests)
# --------------------------------------------------------------------------
def param_shapes(cfg: OracleConfig) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys and shapes of reference TANTE (measured from the module)."""
    C, D, T, Hp, Wp = cfg.embed_dim, cfg.n_fields, cfg.in_T, cfg.Hp, cfg.Wp
    ks = PATCH_MAP[cfg.patch_scale]
    sh: Dict[str, Tuple[int, ...]] = {}
    sh["t_emb"] = (1, T, C)
    sh["s_emb"] = (1, Hp, Wp, C)
    fno = cfg.enc_dec_type == "fno"
    if fno:
        ps = PATCH_MAP_FNO[cfg.patch_scale]
        m1, m2 = cfg.modes1, cfg.modes2

        def spectral(pfx, ci, co, a, b):
            sh[pfx + "weight"] = ("complex", ci, co, a, b)
            sh[pfx + "w0.weight"] = (co, ci, 1, 1)
            sh[pfx + "w0.bias"] = (co,)
        spectral("encoder.enc_spectral_1.", D, C // 8, m1, m2)
        sh["encoder.enc_conv_1.conv.weight"] = (C // 4, C // 8, ps[0], ps[0]); sh["encoder.enc_conv_1.conv.bias"] = (C // 4,)
        spectral("encoder.enc_spectral_2.", C // 4, C // 2, m1 // ps[0], m2 // ps[0])
        sh["encoder.enc_conv_2.conv.weight"] = (C, C // 2, ps[1], ps[1]); sh["encoder.enc_conv_2.conv.bias"] = (C,)
    else:
        chans = [D, C // 4, C // 2, C]
        for i in range(3):
            sh[f"encoder.enc_conv_{i+1}.conv.weight"] = (chans[i + 1], chans[i], ks[i], ks[i])
            sh[f"encoder.enc_conv_{i+1}.conv.bias"] = (chans[i + 1],)
    for k, axes in enumerate(cfg.segments):
        if fno:
            p = f"decoders.{k}."
            sh[p + "dec_conv_1.deconv.weight"] = (C, C // 2, ps[1], ps[1]); sh[p + "dec_conv_1.deconv.bias"] = (C // 2,)
            spectral(p + "dec_spectral_1.", C // 2, C // 4, m1 // ps[0], m2 // ps[0])
            sh[p + "dec_conv_2.deconv.weight"] = (C // 4, C // 8, ps[0], ps[0]); sh[p + "dec_conv_2.deconv.bias"] = (C // 8,)
            spectral(p + "dec_spectral_2.", C // 8, D, m1, m2)
        else:
            dch = [C, C // 2, C // 4, D]
            for i in range(3):
                kk = ks[2 - i]
                sh[f"decoders.{k}.dec_conv_{i+1}.deconv.weight"] = (dch[i], dch[i + 1], kk, kk)
                sh[f"decoders.{k}.dec_conv_{i+1}.deconv.bias"] = (dch[i + 1],)
        n_chan = 0
        for i, ax in enumerate(axes):
            p = f"blocks.{k}.blocks.{i}."
            Cb = cfg.embed_dim
            if ax == "C":      # channel attention: the block works on expanded_channel features (attn_backbone.py:124-132)
                Cb = cfg.expanded_channel
                q = f"blocks.{k}.channel_blocks.{n_chan}."
                n_chan += 1
                sh[q + "0.weight"] = (Cb // 4, 1); sh[q + "0.bias"] = (Cb // 4,)
                sh[q + "2.weight"] = (Cb, Cb // 4); sh[q + "2.bias"] = (Cb,)
            sh[p + "ln1.weight"] = (Cb,); sh[p + "ln1.bias"] = (Cb,)
            sh[p + "attn.in_proj_weight"] = (3 * Cb, Cb); sh[p + "attn.in_proj_bias"] = (3 * Cb,)
            sh[p + "attn.out_proj.weight"] = (Cb, Cb); sh[p + "attn.out_proj.bias"] = (Cb,)
            sh[p + "ln2.weight"] = (Cb,); sh[p + "ln2.bias"] = (Cb,)
            Hm = int(Cb * cfg.mlp_ratio)
            sh[p + "mlp.0.weight"] = (Hm, Cb); sh[p + "mlp.0.bias"] = (Hm,)
            sh[p + "mlp.2.weight"] = (Cb, Hm); sh[p + "mlp.2.bias"] = (Cb,)
        for name, n in (("vertical", Hp), ("horizontal", Wp), ("temporal", T)):
            for j in (0, 2):
                sh[f"blocks.{k}.{name}_propagator.{j}.weight"] = (n, n)
                sh[f"blocks.{k}.{name}_propagator.{j}.bias"] = (n,)
    films = ["t_encode."] + ([f"modifiers.{k}." for k in range(cfg.taylor_order)] if not cfg.deg else [])
    for p in films:
        for nm in ("condition_to_scale", "condition_to_shift"):
            sh[f"{p}{nm}.0.weight"] = (C // 2, 1); sh[f"{p}{nm}.0.bias"] = (C // 2,)
            sh[f"{p}{nm}.2.weight"] = (C, C // 2); sh[f"{p}{nm}.2.bias"] = (C,)
    if not cfg.deg:
        for k in range(cfg.taylor_order):
            p = f"interprators.{k}.interprete."
            sh[p + "0.weight"] = (C // 2, C); sh[p + "0.bias"] = (C // 2,)
            sh[p + "2.weight"] = (C // 4, C // 2); sh[p + "2.bias"] = (C // 4,)
            sh[p + "4.weight"] = (1, C // 4); sh[p + "4.bias"] = (1,)
    return sh


def _name_seed(name: str, seed: int) -> int:
    h = 1469598103934665603
    for ch in f"{seed}:{name}".encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h & 0x7FFFFFFF


def make_state_dict(cfg: OracleConfig, seed: int = 211, rt_bias: float = 0.0,
                    dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights: every tensor is drawn from its own generator seeded by
    (seed, key name), with PyTorch-default-like scales (uniform +-1/sqrt(fan_in)), *non-trivial*
    LayerNorm affine / biases so every parameter is exercised, and sincos + noise embeddings.
    `rt_bias` is added to each interprator's last bias to make the adaptive-dt check
    non-vacuous."""
    sd = {}
    shapes = param_shapes(cfg)
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed(_name_seed(name, seed))
        if shape and shape[0] == "complex":       # SpectralLayer.weight (enc_dec_fno.py:190-193): cfloat, scale 1/sqrt(Cin*Cout)
            cs = shape[1:]
            v = torch.complex(torch.randn(cs, generator=g), torch.randn(cs, generator=g)) * (1.0 / math.sqrt(cs[0] * cs[1]))
            sd[name] = v.to(torch.complex128 if dtype == torch.float64 else torch.complex64)
            continue
        if name == "t_emb":
            v = t_emb_init(cfg.embed_dim, cfg.in_T) + 0.02 * torch.randn(shape, generator=g)
        elif name == "s_emb":
            v = s_emb_init(cfg.embed_dim, cfg.Hp, cfg.Wp) + 0.02 * torch.randn(shape, generator=g)
        elif ".ln" in name and name.endswith("weight"):
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif ".ln" in name and name.endswith("bias"):
            v = 0.1 * torch.randn(shape, generator=g)
        else:
            wshape = _weight_shape_for(shapes, name) if name.endswith("bias") else shape
            if "deconv" in name:
                fan_in = wshape[1] * wshape[2] * wshape[3]   # ConvTranspose2d: fan_in from dim 1
            elif len(wshape) == 4:
                fan_in = wshape[1] * wshape[2] * wshape[3]
            else:
                fan_in = wshape[-1]
            bound = 1.0 / math.sqrt(max(fan_in, 1))
            v = (torch.rand(shape, generator=g) * 2 - 1) * bound
        if rt_bias and name.startswith("interprators.") and name.endswith("interprete.4.bias"):
            v = v + rt_bias
        sd[name] = v.to(dtype)
    return sd


def _weight_shape_for(shapes, bias_name: str):
    if bias_name.endswith("in_proj_bias"):
        return shapes[bias_name.replace("in_proj_bias", "in_proj_weight")]
    return shapes[bias_name[:-4] + "weight"]


def make_input(cfg: OracleConfig, B: int, seed: int = 212, dtype=torch.float32, T: Optional[int] = None):
    g = torch.Generator().manual_seed(seed)
    T = cfg.in_T if T is None else T
    return torch.randn(B, T, cfg.n_fields, cfg.H, cfg.W, generator=g).to(dtype)
