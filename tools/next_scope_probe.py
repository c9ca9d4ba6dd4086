"""Timing of the SURVEY.md 8(f) rows ("next" scope: patch_scale 16 / 32, fno, composite axes, axis C, embed_dim 512) on the CUDA path,
beside the stock-torch restatement of the reference module (oracle/eager_module.py, bf16 autocast + TF32) where that covers the
configuration.  One inference call and one training call (forward + backward) per configuration at the TRL shape (4 fields, 128 x 384).

usage: python tools/next_scope_probe.py [out.json]
Test / evidence tooling: the eager module is the checker's sibling, never part of the product path."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import tante_oracle as O  # noqa: E402
from oracle.eager_module import EagerTANTE  # noqa: E402
from gpu_util import make_model  # noqa: E402

CASES = [
    # name, config kwargs, batch, eager available
    ("p8_thw (shipped configuration)", dict(attn_axes="THWTHWTHW"), 8, True),
    ("patch_scale 16", dict(attn_axes="THWTHWTHW", patch_scale=16), 8, True),
    ("patch_scale 32", dict(attn_axes="THWTHWTHW", patch_scale=32), 8, True),
    ("axes LYA (S = 768 / 64 / 3072)", dict(attn_axes="LYATHW"), 4, True),
    ("embed_dim 512", dict(attn_axes="THWTHWTHW", embed_dim=512, n_head=16), 8, True),
    ("fno, patch_scale 8", dict(attn_axes="THWTHWTHW", enc_dec_type="fno", patch_scale=8, modes1=32, modes2=32), 4, False),
    ("axis C (THCW)", dict(attn_axes="THCW"), 1, False),
]


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    recs = []
    for name, kw, B, has_eager in CASES:
        cfg = O.OracleConfig(n_fields=4, H=128, W=384, taylor_order=1, deg=True, **kw)
        sd = O.make_state_dict(cfg, 211, 0.0)
        x = O.make_input(cfg, B, 212).cuda()
        rec = {"case": name, "batch": B, "shape": "trl (4 fields, 128x384)", "precision": "bf16"}
        model = make_model(cfg, sd, "bf16")

        def infer():
            with torch.inference_mode():
                model(x)
        rec["forward_ms"] = timed(infer)
        model.train()
        xg = x.clone().requires_grad_(True)

        def train():
            for p in model.parameters():
                p.grad = None
            model(xg).square().mean().backward()
        rec["train_call_ms"] = timed(train, iters=3, warm=1)
        rec["forward_samples_per_s"] = B / rec["forward_ms"] * 1e3
        rec["train_samples_per_s"] = B / rec["train_call_ms"] * 1e3
        del model
        if has_eager:
            try:
                em = EagerTANTE(cfg).cuda()
                em.load_state_dict({k: v.cuda() for k, v in sd.items()})

                def einfer():
                    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
                        em(x)

                def etrain():
                    em.zero_grad(set_to_none=True)
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        y = em(xg)
                    y.float().square().mean().backward()
                rec["eager_forward_ms"] = timed(einfer)
                em.train()
                rec["eager_train_call_ms"] = timed(etrain, iters=3, warm=1)
                rec["forward_vs_eager"] = rec["eager_forward_ms"] / rec["forward_ms"]
                rec["train_vs_eager"] = rec["eager_train_call_ms"] / rec["train_call_ms"]
                del em
            except Exception as e:      # stock kernels refuse some shapes (see bench.py: batches_refused_by_stock_torch)
                rec["eager_error"] = str(e)[:200]
        torch.cuda.empty_cache()
        recs.append(rec)
        print(json.dumps(rec), flush=True)
    if out_path:
        with open(out_path, "w") as f:
            json.dump({"what": __doc__.split("\n\n")[0], "records": recs}, f, indent=1)


if __name__ == "__main__":
    main()
