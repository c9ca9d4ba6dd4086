"""The fused block-tail kernel (block_tail_tc.cuh: attention out-projection + residual + LN2 + MLP + residual + next LN1 as
three chained tcgen05 GEMMs, reference models/attn_backbone.py:81-83,68) stand-alone against an fp64 torch restatement that
rounds where the kernel rounds (bf16 operands of every GEMM)."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16)


def _ln(x, g, b):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-5) * g + b


def _gelu_tanh(x):
    return 0.5 * x * (1.0 + torch.tanh(0.7978845608028654 * (x + 0.044715 * x ** 3)))


def _reference(att, Wo, W1, W2, v, x_in, train):
    d = torch.float64
    bo, g2, be2, b1, b2, gn, ben = [v[i].to(d) for i in range(7)]
    x_mid = x_in.to(d) + att.to(d) @ Wo.to(d).t() + bo
    ln2 = _bf(_ln(x_mid, g2, be2).float())
    hpre = ln2.to(d) @ W1.to(d).t() + b1
    hpre_r = _bf(hpre.float())
    hact = _bf(_gelu_tanh(hpre_r.to(d) if train else hpre).float())
    x_out = x_mid + hact.to(d) @ W2.to(d).t() + b2
    ln_out = _ln(x_out, gn, ben)
    return x_mid, ln2, hpre_r, hact, x_out, ln_out


@pytest.mark.parametrize("M", [128, 1000, 148 * 128 * 2 + 77])
@pytest.mark.parametrize("train", [False, True])
@pytest.mark.parametrize("with_ln", [True, False])
def test_block_tail_matches_torch(M, train, with_ln):
    from tante_b200 import _abi
    lib = _abi.load()
    g = torch.Generator(device="cuda").manual_seed(7 + M)
    C = 256
    att = _bf(torch.randn(M, C, device="cuda", generator=g))
    Wo, W1, W2 = [_bf(torch.randn(C, C, device="cuda", generator=g) / 16) for _ in range(3)]
    v = torch.randn(7, C, device="cuda", generator=g) * 0.2
    v[1] += 1.0
    v[5] += 1.0
    x_in = torch.randn(M, C, device="cuda", generator=g) * 2 + 0.5
    x_out = torch.full((M, C), float("nan"), device="cuda")
    ln_out = torch.zeros(M, C, device="cuda", dtype=torch.bfloat16) if with_ln else None
    x_mid = torch.full((M, C), float("nan"), device="cuda") if train else None
    ln2, hpre, hact = [torch.zeros(M, C, device="cuda", dtype=torch.bfloat16) if train else None for _ in range(3)]
    ptr = lambda t: None if t is None else t.data_ptr()
    _abi.check(lib.tante_test_block_tail(att.data_ptr(), Wo.data_ptr(), W1.data_ptr(), W2.data_ptr(), v.data_ptr(), x_in.data_ptr(),
                                         x_out.data_ptr(), ptr(ln_out), ptr(x_mid), ptr(ln2), ptr(hpre), ptr(hact), M, 1,
                                         torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    r_mid, r_ln2, r_hpre, r_hact, r_out, r_ln = _reference(att, Wo, W1, W2, v, x_in, train)
    assert rel_l2(x_out.cpu().numpy(), r_out.cpu().numpy()) < 2e-3          # hidden rounded to bf16 inside, tanh.approx
    if with_ln:
        assert rel_l2(ln_out.float().cpu().numpy(), r_ln.cpu().numpy()) < 6e-3
    if train:
        assert rel_l2(x_mid.cpu().numpy(), r_mid.cpu().numpy()) < 1e-5
        assert rel_l2(ln2.float().cpu().numpy(), r_ln2.float().cpu().numpy()) < 3e-3
        assert rel_l2(hpre.float().cpu().numpy(), r_hpre.float().cpu().numpy()) < 5e-3
        assert rel_l2(hact.float().cpu().numpy(), r_hact.float().cpu().numpy()) < 8e-3


def test_block_tail_in_place_equals_out_of_place():
    """Inference updates the residual stream in place (x_out aliases x_in)."""
    from tante_b200 import _abi
    lib = _abi.load()
    g = torch.Generator(device="cuda").manual_seed(3)
    M, C = 148 * 128 + 5, 256
    att = _bf(torch.randn(M, C, device="cuda", generator=g))
    Wo, W1, W2 = [_bf(torch.randn(C, C, device="cuda", generator=g) / 16) for _ in range(3)]
    v = torch.randn(7, C, device="cuda", generator=g) * 0.2
    x = torch.randn(M, C, device="cuda", generator=g)
    outs = []
    for inplace in (False, True):
        xi = x.clone()
        xo = xi if inplace else torch.empty_like(xi)
        ln = torch.zeros(M, C, device="cuda", dtype=torch.bfloat16)
        _abi.check(lib.tante_test_block_tail(att.data_ptr(), Wo.data_ptr(), W1.data_ptr(), W2.data_ptr(), v.data_ptr(), xi.data_ptr(),
                                             xo.data_ptr(), ln.data_ptr(), None, None, None, None, M, 1,
                                             torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        outs.append((xo.clone(), ln.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
