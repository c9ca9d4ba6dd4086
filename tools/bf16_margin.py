import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import rel_l2
import test_gpu_parity as T
for name in ["fwd_deg_k1_p8", "fwd_stages_k2_p8", "fwd_adp_k2_p8_b27", "fwd_adp_k3_p4", "fwd_adp_k1_p2", "trl_k1_b13", "trl_k2_b52"]:
    z, meta, cfg, sd, x, model = T._setup(name, precision="bf16")
    with torch.inference_mode():
        out = model(x.cuda(), meta["out_T"])
    y = out if cfg.deg else out[0]
    if y.shape[1] != meta["n"]:
        print(name, "n differs", y.shape[1], meta["n"]); continue
    s = meta["stride"]; yc = y.cpu()
    u0 = x[:, -1:].expand_as(yc)
    print(name, "field rel", rel_l2(yc.reshape(-1)[::s].numpy(), z["frames"]), "deriv rel",
          rel_l2((yc - u0).reshape(-1)[::s].numpy(), z["frames"] - u0.reshape(-1)[::s].numpy()),
          "" if cfg.deg else ("rt err %.4f" % float(np.abs(out[1].cpu().numpy() - z["R_t"]).max())))
