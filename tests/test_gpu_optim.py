"""tante_optimizer_step / tante_b200.FusedAdamW against the reference's optimizer tail -- torch.nn.utils.clip_grad_norm_ /
clip_grad_value_ (trainer/trainer.py:192-193, trainer/r_trainer.py:155) followed by torch.optim.AdamW.step()
(configs/tante.yaml:38-41) -- on identical gradients."""
import copy
import io

import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(seed=0):
    from tante_b200 import TANTE, TanteMetadata
    torch.manual_seed(seed)
    m = TANTE(4, TanteMetadata(spatial_resolution=(32, 64), n_fields=3), taylor_order=1, attn_axes="THW", patch_scale=8,
              deg=True, precision="bf16")
    return m.cuda().train()


@pytest.mark.parametrize("clip", ["norm", "value", None])
def test_fused_adamw_matches_torch(clip):
    from tante_b200 import FusedAdamW
    from tante_b200.trainer import GradBucket
    ma, mb = _model(), _model()
    ba, bb = GradBucket(ma), GradBucket(mb)
    assert [n for n, _ in ma.named_parameters()] == [n for n, _ in mb.named_parameters()]
    oa = torch.optim.AdamW(ma.parameters(), lr=3e-3, weight_decay=1e-2)
    ob = FusedAdamW(mb.parameters(), lr=3e-3, weight_decay=1e-2)          # finds its module by itself
    g = torch.Generator(device="cuda").manual_seed(5)
    scale = 0.25 if clip is None else 1.0
    for step in range(4):
        flat = torch.randn(ba.flat.numel(), device="cuda", generator=g) * (3.0 if step % 2 else 0.01)
        ba.flat.copy_(flat * scale)
        bb.flat.copy_(flat)
        if clip == "norm":
            torch.nn.utils.clip_grad_norm_(ba.params, 1.0)
        elif clip == "value":
            torch.nn.utils.clip_grad_value_(ba.params, 1.0)
        oa.step()
        ob.step(clip=clip, clip_value=1.0, grad_scale=scale)
        if clip is not None:        # the clipped gradient is left in place, as clip_grad_* does
            assert torch.allclose(ba.flat, bb.flat, rtol=1e-5, atol=1e-7)
    pa, pb = dict(ma.named_parameters()), dict(mb.named_parameters())
    for n in pa:
        assert torch.allclose(pa[n], pb[n], rtol=2e-5, atol=2e-6), n
    sa, sb = oa.state_dict()["state"], ob.state_dict()["state"]
    assert sa.keys() == sb.keys()
    for k in sa:
        assert float(sa[k]["step"]) == float(sb[k]["step"]) == 4.0
        assert torch.allclose(sa[k]["exp_avg"], sb[k]["exp_avg"], rtol=1e-5, atol=1e-7)
        assert torch.allclose(sa[k]["exp_avg_sq"], sb[k]["exp_avg_sq"], rtol=1e-4, atol=1e-9)      # fma contraction


def test_fused_adamw_repacks_and_resumes():
    """After a step the module computes with the UPDATED weights (the library repacked them), and an optimizer
    state_dict written with torch.save resumes to the same trajectory."""
    from tante_b200 import FusedAdamW
    from tante_b200.trainer import GradBucket, train_step
    m = _model(1)
    bucket = GradBucket(m)
    opt = FusedAdamW(m.parameters(), lr=1e-3, weight_decay=1e-5, model=m)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 4, 3, 32, 64, generator=g).cuda()
    y = torch.randn(2, 2, 32, 64, 3, generator=g).cuda()
    for _ in range(2):
        train_step(m, opt, x, y, 2, bucket)
    with torch.no_grad():
        out = m.eval()(x)
        fresh = _model(1)
        fresh.load_state_dict(m.state_dict())
        ref = fresh.eval()(x)
    assert torch.equal(out, ref)
    # resume: two further steps from a checkpoint of model + optimizer, on IDENTICAL synthetic gradients (the model's own
    # weight-gradient reductions are not bit-reproducible, and Adam's sign-like update amplifies that)
    buf = io.BytesIO()
    torch.save({"model": m.state_dict(), "opt": opt.state_dict()}, buf)
    buf.seek(0)
    ck = torch.load(buf, weights_only=False)
    m2 = _model(1)
    m2.load_state_dict(ck["model"])
    b2 = GradBucket(m2)
    o2 = FusedAdamW(m2.parameters(), lr=1e-3, weight_decay=1e-5)
    o2.load_state_dict(ck["opt"])
    assert float(next(iter(o2.state_dict()["state"].values()))["step"]) == 2.0
    gg = torch.Generator(device="cuda").manual_seed(11)
    for _ in range(2):
        flat = torch.randn(bucket.flat.numel(), device="cuda", generator=gg)
        bucket.flat.copy_(flat)
        b2.flat.copy_(flat)
        opt.step(clip="norm")
        o2.step(clip="norm")
    p1, p2 = dict(m.named_parameters()), dict(m2.named_parameters())
    for n in p1:
        assert torch.allclose(p1[n], p2[n], rtol=1e-6, atol=1e-7), n
