import torch

from oracle import tante_oracle as O
from tante_b200 import TANTE, TanteMetadata


def make_model(cfg: O.OracleConfig, sd, precision="fp32", device="cuda:0"):
    m = TANTE(cfg.in_T, TanteMetadata(spatial_resolution=(cfg.H, cfg.W), n_fields=cfg.n_fields),
              taylor_order=cfg.taylor_order, frame_interval=cfg.frame_interval, output_length=cfg.output_length,
              attn_axes=cfg.attn_axes, n_head=cfg.n_head, embed_dim=cfg.embed_dim, patch_scale=cfg.patch_scale,
              deg=cfg.deg, precision=precision, enc_dec_type=cfg.enc_dec_type, modes1=cfg.modes1, modes2=cfg.modes2,
              mlp_ratio=getattr(cfg, "mlp_ratio", 1.0), expanded_channel=getattr(cfg, "expanded_channel", 128),
              overlap_ratio=getattr(cfg, "overlap_ratio", 0.0))
    m.load_state_dict(sd)
    return m.to(device).eval()


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
