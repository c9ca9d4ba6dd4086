import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import tante_oracle as O
from gpu_util import make_model
cfg = O.OracleConfig(n_fields=2, H=32, W=48, taylor_order=2, attn_axes="THW-HWT", deg=False)
sd = O.make_state_dict(cfg, 5, rt_bias=3.0585)
B = 12
x = O.make_input(cfg, B, 6)
scale = torch.tensor([8.0, 0.05, 3.0, 1.0, 20.0, 8.0, 0.05, 1.0, 30.0, 8.0, 1.0, 50.0])
x = (x * scale.view(B, 1, 1, 1, 1)).cuda()
model = make_model(cfg, sd, precision=sys.argv[1] if len(sys.argv) > 1 else "fp32")
with torch.inference_mode():
    y, R, ns, steps = model.rollout(x, 8, per_sample=True)
    print("mode", os.environ.get("TANTE_ROLLOUT_COMPACT"), "steps", steps.tolist())
    print("ns", [ns[: int(steps[b]), b].tolist() for b in range(B)])
    for b in range(B):
        y1, R1, ns1, steps1 = model.rollout(x[b:b + 1], 8, per_sample=True)
        d = (y[b] - y1[0]).abs().amax(dim=(1, 2, 3)).tolist()
        print(b, "ns1", ns1[: int(steps1[0]), 0].tolist(), "max diff per frame", [f"{v:.2e}" for v in d])
