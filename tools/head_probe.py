"""One configuration of the fused Taylor head, for ncu: python tools/head_probe.py trl 8 1 1"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tante_b200 import TANTE, TanteMetadata

shape, P, K, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
D, H, W, B = {"trl": (4, 128, 384, 64), "active_matter": (11, 256, 256, 32)}[shape]
m = TANTE(4, TanteMetadata(spatial_resolution=(H, W), n_fields=D), taylor_order=K, attn_axes="-".join(["T"] * K),
          patch_scale=P, deg=False, precision="bf16").cuda().eval()
x = torch.randn(B, 4, D, H, W, device="cuda")
ms = m.bench_head(x, n, iters=int(sys.argv[5]) if len(sys.argv) > 5 else 5)
rows = B * H * W // 4
by = rows * K * 64 * 2 + B * D * H * W * 4 * (1 + n)
print(f"{shape} P={P} K={K} n={n}: {ms*1e3:.1f} us, {by/ms/1e6:.0f} GB/s")
