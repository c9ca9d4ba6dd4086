"""K x patch-size sweep of the fused Taylor head vs the HBM roofline (BASELINE.json configs[4]).

bytes (boundary B of SURVEY.md 8(d)) per launch = rows*K*64*s_act + B*D*H*W*4*(1 + n), rows = B*H*W/k0^2.
Writes one JSON object per case to stdout (and profiles/head_sweep_r1.json when --out is given)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tante_b200 import TANTE, TanteMetadata

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--precision", default="bf16")
args = ap.parse_args()
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
peak = peaks.get("hbm_gbs", 6650.0)
res = []
# batch sizes chosen so that every case moves > 2x the 126 MB L2 per launch (no explicit flush needed)
shapes = {"trl": (4, 128, 384, 64), "active_matter": (11, 256, 256, 32)}
for sname, (D, H, W, B) in shapes.items():
    for P in (2, 4, 8):
        for K in (1, 2, 3, 4):
            axes = "-".join(["T"] * K)
            try:
                m = TANTE(4, TanteMetadata(spatial_resolution=(H, W), n_fields=D), taylor_order=K, attn_axes=axes,
                          patch_scale=P, deg=False, precision=args.precision).cuda().eval()
            except Exception as e:   # e.g. axis length > 64 at small patch sizes
                res.append({"shape": sname, "P": P, "K": K, "skipped": str(e)[:80]}); continue
            x = torch.randn(B, 4, D, H, W, device="cuda")
            for n in (1, 4, 8):
                try:
                    ms = m.bench_head(x, n, iters=20)
                except Exception as e:
                    res.append({"shape": sname, "P": P, "K": K, "n": n, "skipped": str(e)[:120]}); continue
                s_act = 2 if args.precision == "bf16" else 4
                rows = B * H * W // 4
                by = rows * K * 64 * s_act + B * D * H * W * 4 * (1 + n)
                gbs = by / (ms * 1e-3) / 1e9
                res.append({"shape": sname, "D": D, "HxW": [H, W], "B": B, "P": P, "K": K, "n": n, "ms": ms,
                            "bytes": by, "GBps": gbs, "frac_of_measured_hbm": gbs / peak})
                print(json.dumps(res[-1]), flush=True)
            del m
            torch.cuda.empty_cache()
if args.out:
    json.dump({"peak_hbm_gbs": peak, "precision": args.precision, "cases": res}, open(args.out, "w"), indent=1)
