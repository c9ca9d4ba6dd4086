"""Training dropout (configs/tante.yaml:29 p = 0.1; nn.MultiheadAttention(dropout=p) + self.drop on both residual branches,
models/attn_backbone.py:47-57,81-83).  The masks are counter-based (dropout.cuh), so the test regenerates them on the host
(tests/dropout_ref.py) and demands PARITY with the oracle run under exactly those masks -- frames and every gradient, i.e.
the forward and the hand-written backward use the same bits at all three sites -- plus the usual statistics."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from dropout_ref import DropMasks, multipliers, next_call_seed
from oracle import tante_oracle as O

pytestmark = pytest.mark.gpu


def _model(cfg, sd, prec, p):
    from tante_b200 import TANTE, TanteMetadata
    m = TANTE(cfg.in_T, TanteMetadata(spatial_resolution=(cfg.H, cfg.W), n_fields=cfg.n_fields), taylor_order=cfg.taylor_order,
              attn_axes=cfg.attn_axes, patch_scale=cfg.patch_scale, deg=cfg.deg, dropout=p, precision=prec,
              mlp_ratio=cfg.mlp_ratio, embed_dim=cfg.embed_dim, n_head=cfg.n_head)
    m.load_state_dict(sd)
    return m.cuda().train()


@pytest.mark.parametrize("prec,tol_y,tol_g", [("fp32", 1e-5, 3e-4), ("bf16", 2e-2, 8e-2)])
@pytest.mark.parametrize("case", ["deg_thw", "adp_k2", "adp_lya", "deg_mlp2", "deg_c512"])
def test_dropout_training_matches_oracle_with_same_masks(case, prec, tol_y, tol_g):
    p = 0.25
    if case == "deg_thw":
        cfg = O.OracleConfig(n_fields=3, H=32, W=48, taylor_order=1, attn_axes="THW", deg=True)
        rt_bias, out_T = 0.0, 1
    elif case == "deg_mlp2":      # mlp_ratio 2: the block outside the fused tail (tensor mode: GEMM + masked residual pass)
        cfg = O.OracleConfig(n_fields=2, H=32, W=32, taylor_order=1, attn_axes="HWT", deg=True, mlp_ratio=2.0)
        rt_bias, out_T = 0.0, 1
    elif case == "deg_c512":      # embed_dim 512, head_dim 32
        cfg = O.OracleConfig(n_fields=2, H=32, W=32, taylor_order=1, attn_axes="TH", deg=True, embed_dim=512, n_head=16)
        rt_bias, out_T = 0.0, 1
    elif case == "adp_lya":      # composite axes: L = 96 and A = 384 tokens (general forward kernel + tiled recompute backward)
        cfg = O.OracleConfig(n_fields=2, H=64, W=96, taylor_order=2, attn_axes="LT-AY", deg=False)
        rt_bias, out_T = 1.3, 4
    else:
        cfg = O.OracleConfig(n_fields=2, H=64, W=32, taylor_order=2, attn_axes="WT-H", deg=False)
        rt_bias, out_T = 1.3, 4
    sd = O.make_state_dict(cfg, 511, rt_bias)
    B = 2
    x = O.make_input(cfg, B, 512)
    seed = next_call_seed(77)
    masks = DropMasks(seed, p, cfg.n_head, cfg.embed_dim)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xg = x.clone().requires_grad_(True)
    out = O.forward(sdg, cfg, xg, out_T, drop_fn=masks)
    y_ref, rt_ref = (out, None) if cfg.deg else out
    g = torch.Generator().manual_seed(513)
    gy = torch.randn(y_ref.shape, generator=g)
    grt = torch.randn(B, generator=g)
    loss = (y_ref * gy).sum() + (0 if rt_ref is None else (rt_ref * grt).sum())
    loss.backward()

    model = _model(cfg, sd, prec, p)
    xc = x.cuda().requires_grad_(True)
    torch.manual_seed(77)                       # the module draws the same key
    out = model(xc) if cfg.deg else model(xc, out_T)
    y, rt = (out, None) if cfg.deg else out
    assert y.shape == y_ref.shape
    ((y * gy.cuda()).sum() + (0 if rt is None else (rt * grt.cuda()).sum())).backward()
    assert rel_l2(y.detach().cpu().numpy(), y_ref.detach().numpy()) < tol_y
    # dropout really happened: the eval-mode output differs
    with torch.no_grad():
        y_eval = O.forward(sd, cfg, x, out_T)
        y_eval = y_eval if cfg.deg else y_eval[0]
    if y_eval.shape == y_ref.shape:
        u0 = x[:, -1:]
        assert rel_l2((y.detach().cpu() - u0).numpy(), (y_eval - u0).numpy()) > 5e-2
    bad = []
    for n, prm in model.named_parameters():
        ref = sdg[n].grad
        if ref is None or float(ref.norm()) == 0.0:
            continue
        e = rel_l2(prm.grad.cpu().numpy(), ref.numpy())
        if not e <= tol_g:
            bad.append(f"{n}: rel {e:.3e}")
    if prec == "fp32":
        assert not bad, "gradients differ from oracle autograd under the same masks:\n" + "\n".join(bad)
    else:   # bf16: cancellation-dominated tensors (biases, LayerNorm affines) are noisy; the bulk must agree
        assert len(bad) <= max(2, len(list(model.parameters())) // 10), "\n".join(bad)
    assert rel_l2(xc.grad.cpu().numpy(), xg.grad.numpy()) < tol_g


def test_dropout_masks_are_fresh_per_call_and_reproducible():
    cfg = O.OracleConfig(n_fields=2, H=32, W=32, taylor_order=1, attn_axes="TH", deg=True)
    sd = O.make_state_dict(cfg, 11)
    model = _model(cfg, sd, "fp32", 0.1)
    x = O.make_input(cfg, 2, 12).cuda()
    with torch.no_grad():                       # train() mode drops under no_grad too, like the reference module
        torch.manual_seed(5)
        a, b = model(x), model(x)               # two calls draw two keys
        torch.manual_seed(5)
        a2 = model(x)
        model.eval()
        e1, e2 = model(x), model(x)
    assert torch.equal(a, a2) and not torch.equal(a, b)
    assert torch.equal(e1, e2) and not torch.equal(e1, a)
    eng = next(iter(model._engines.values()))
    assert len(eng.free_slots) == eng.n_slots, "tape slots of the no_grad train-mode calls were not returned"


def test_dropout_keep_statistics():
    p = 0.1
    n = 1 << 20
    e = np.arange(n, dtype=np.uint64)
    for site in (1, 6, 130):
        m = multipliers(123456789012345, site, e >> np.uint64(3), (e & np.uint64(7)).astype(np.int64), p)
        keep = (m > 0).mean()
        assert abs(keep - (1 - p)) < 4 * np.sqrt(p * (1 - p) / n) + 2e-5
        assert abs(m.mean() - 1.0) < 2e-3
        # neighbours are uncorrelated
        k = (m > 0).astype(np.float64)
        assert abs(np.corrcoef(k[:-1], k[1:])[0, 1]) < 5e-3


def test_dropout_output_distribution_matches_stock_module():
    """Different generators, same law: mean and spread of the train-mode output over many mask draws vs the stock torch
    module (oracle/eager_module.py, nn.Dropout / nn.MultiheadAttention dropout) on the CPU."""
    from oracle.eager_module import EagerTANTE
    p, n = 0.2, 24
    cfg = O.OracleConfig(n_fields=2, H=32, W=32, taylor_order=1, attn_axes="HWT", deg=True)
    sd = O.make_state_dict(cfg, 21)
    x = O.make_input(cfg, 1, 22)
    ref = EagerTANTE(cfg, dropout=p).train()
    ref.load_state_dict(sd)
    model = _model(cfg, sd, "fp32", p)
    u0 = x[:, -1:]
    with torch.no_grad():
        torch.manual_seed(1)
        R = torch.stack([ref(x) - u0 for _ in range(n)])
        torch.manual_seed(2)
        Y = torch.stack([model(x.cuda()).cpu() - u0 for _ in range(n)])
    # per-pixel mean of the derivative field: the two sample means differ by ~ sqrt(2/n) sigma
    sig = R.std(0).mean()
    assert float((R.mean(0) - Y.mean(0)).abs().mean()) < 1.5 * float(sig) * np.sqrt(2.0 / n)
    assert abs(float(Y.std(0).mean()) / float(sig) - 1.0) < 0.15
