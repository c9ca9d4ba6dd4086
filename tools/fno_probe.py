"""One fno forward at the TRL shape (B = 4, bf16) for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import tante_oracle as O
from gpu_util import make_model
cfg = O.OracleConfig(n_fields=4, H=128, W=384, taylor_order=1, deg=True, attn_axes="THW", enc_dec_type="fno", patch_scale=8, modes1=32, modes2=32)
model = make_model(cfg, O.make_state_dict(cfg, 211, 0.0), "bf16")
x = O.make_input(cfg, 4, 212).cuda()
with torch.inference_mode():
    for _ in range(3):
        model(x)
torch.cuda.synchronize()
