"""Forward time of a model whose backbone uses the long axes (L / Y / A): the tiled attention kernel vs the general one
(TANTE_ATT_FLASH=0)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tante_b200 import TANTE, TanteMetadata
torch.manual_seed(0)
m = TANTE(4, TanteMetadata(spatial_resolution=(128, 384), n_fields=4), taylor_order=1, attn_axes="LYA", patch_scale=8,
          deg=True, precision="bf16").cuda().eval()
x = torch.randn(4, 4, 4, 128, 384, device="cuda")
with torch.inference_mode():
    for _ in range(2):
        y = m(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        y = m(x)
    torch.cuda.synchronize()
print(f"axes LYA, TRL shape, B=4: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms per forward, out norm {y.float().norm().item():.4f}")
