"""Timing probe of one training step (Trainer.train_one_epoch body, trainer/trainer.py:174-207) on the CUDA path.
usage: python tools/train_probe.py [shape] [batch] [steps] [precision] [K]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tante_b200 import TANTE, TanteMetadata  # noqa: E402

SHAPES = {"trl": (4, 128, 384), "active_matter": (11, 256, 256), "rayleigh_benard": (4, 512, 128),
          "viscoelastic": (8, 512, 512)}


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "active_matter"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    prec = sys.argv[4] if len(sys.argv) > 4 else "bf16"
    D, H, W = SHAPES[shape]
    dev = torch.device("cuda:0")
    torch.manual_seed(211)
    model = TANTE(4, TanteMetadata(spatial_resolution=(H, W), n_fields=D), taylor_order=1, attn_axes="THWTHWTHW",
                  patch_scale=8, deg=True, dropout=0.0, precision=prec).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=5e-5, weight_decay=1e-5)
    x = torch.randn(B, 4, D, H, W, device=dev)
    y_ref = torch.randn(B, 4, H, W, D, device=dev)

    def step():
        moving, ys = x, []
        for _ in range(4):
            y = model(moving)
            moving = torch.cat([moving[:, 1:], y], dim=1)
            ys.append(y.permute(0, 1, 3, 4, 2))
        yp = torch.cat(ys, dim=1)
        loss = torch.mean((yp - y_ref) ** 2, dim=(-3, -2)).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{shape} B={B} {prec}: {ms:.2f} ms/step  {B / ms * 1e3:.1f} samples/s  loss {float(loss):.5f} "
          f"wall {(time.perf_counter() - t0) / steps * 1e3:.1f} ms  mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB torch, "
          f"launches {model.launch_count()}")


if __name__ == "__main__":
    main()
