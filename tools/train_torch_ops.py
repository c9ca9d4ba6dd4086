"""Which torch ops still launch kernels in one training step (torch.profiler, one step after warm-up)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from tante_b200 import TANTE, TanteMetadata
from tante_b200.trainer import GradBucket, train_step

D, H, W, B = 11, 256, 256, 16
torch.manual_seed(211)
model = TANTE(4, TanteMetadata(spatial_resolution=(H, W), n_fields=D), taylor_order=1, attn_axes="THWTHWTHW",
              patch_scale=8, deg=True, dropout=0.0, precision="bf16").cuda().train()
opt = torch.optim.AdamW(model.parameters(), lr=5e-5, weight_decay=1e-5, fused=True)
bucket = GradBucket(model)
x = torch.randn(B, 4, D, H, W, device="cuda")
y = torch.randn(B, 4, H, W, D, device="cuda")
for _ in range(3):
    train_step(model, opt, x, y, 4, bucket)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    train_step(model, opt, x, y, 4, bucket)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_stack_n=4).table(sort_by="self_cuda_time_total", row_limit=25, max_name_column_width=50,
                                                   max_src_column_width=90))
