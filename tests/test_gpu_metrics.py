"""On-device evaluation metrics (tante_metric_moments) vs the oracle's restatement of
reference trainer/metrics.py, and the evaluator drop-ins end to end on the CUDA model."""
import pytest
import torch

from oracle import tante_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(2, 3, 16, 24, 4), (1, 8, 128, 384, 4), (3, 2, 64, 64, 11), (2, 1, 40, 56, 1)])
def test_metrics_match_oracle(shape):
    import tante_b200 as tb
    g = torch.Generator().manual_seed(5)
    y = torch.randn(shape, generator=g) * 1.7 + 0.3
    x = y + 0.1 * torch.randn(shape, generator=g)
    xd, yd = x.cuda(), y.cuda()
    want = {"MSE": O.mse_eval(x.double(), y.double()), "L2RE": O.l2re_eval(x.double(), y.double()),
            "NNMSE": O.nnmse_eval(x.double(), y.double()), "VRMSE": O.vrmse_eval(x.double(), y.double()),
            "NMSE": O.nmse_eval(x.double(), y.double()), "RMSE": torch.sqrt(O.mse_eval(x.double(), y.double())),
            "VMSE": O.nmse_eval(x.double(), y.double(), norm_mode="std")}
    for name, ref in want.items():
        got = getattr(tb, name)()(xd, yd, None).cpu().double()
        assert got.shape == ref.shape, name
        assert float((got - ref).abs().max() / ref.abs().max()) < 1e-5, name
    # Metric.forward with rt: eval().mean() + eval_rt (metrics.py:37-41)
    rt = torch.tensor([1.2, 1.3], device="cuda")
    got = float(tb.MSE()(xd, yd, rt, 0.5, 2))
    assert abs(got - float(O.train_loss(x, y, rt.cpu()))) < 1e-6 * max(1.0, abs(got))


def test_metrics_as_differentiable_training_loss():
    import tante_b200 as tb
    x = torch.randn(2, 2, 8, 8, 3, device="cuda", requires_grad=True)
    y = torch.randn(2, 2, 8, 8, 3, device="cuda")
    loss = tb.MSE()(x, y, None).mean()
    loss.backward()
    assert torch.allclose(x.grad, 2 * (x.detach() - y) / x.numel(), atol=1e-7)


def test_r_evaler_end_to_end_on_device():
    """eval.py's call sequence on the real CUDA model: R_Evaler + the four configured metrics (tante.yaml:67-74)."""
    import tante_b200 as tb
    from gpu_util import make_model
    cfg = O.OracleConfig(n_fields=3, H=32, W=32, taylor_order=1, attn_axes="THW", deg=False)
    sd = O.make_state_dict(cfg, 9, rt_bias=1.3)
    model = make_model(cfg, sd)

    class DM:
        class train_dataset:
            metadata = None

        def test_dataloader(self):
            g = torch.Generator().manual_seed(1)
            return [{"input": torch.randn(2, 4, 32, 32, 3, generator=g), "output": torch.randn(2, 6, 32, 32, 3, generator=g)}
                    for _ in range(3)]
    ev = tb.R_Evaler(model=model, datamodule=DM(), eval_loss_fn1=tb.MSE(), eval_loss_fn2=tb.L2RE(), eval_loss_fn3=tb.NNMSE(),
                     eval_loss_fn4=tb.VRMSE(), device="cuda:0", n_steps_rollout=6, batch_size=2)
    loss, std, RT, Step, t, s_err, s_rt = ev.Eval("common")
    # the same numbers from the oracle rollout + oracle metrics
    want = [[], [], [], []]
    for b in DM().test_dataloader():
        xw = b["input"].permute(0, 1, 4, 2, 3).contiguous()
        with torch.inference_mode():
            yp, _, _ = O.rollout_eval(sd, cfg, xw, 6)
        want[0].append(float(O.mse_eval(yp, b["output"]).mean()))
        want[1].append(float(O.nnmse_eval(yp, b["output"]).mean()))
        want[2].append(float(O.l2re_eval(yp, b["output"]).mean()))
        want[3].append(float(O.vrmse_eval(yp, b["output"]).mean()))
    for got, w in zip(loss, want):
        assert abs(got - sum(w) / 3) < 2e-5 * abs(sum(w) / 3)
    assert 1.0 < RT < 6.0 and Step >= 2 and t > 0
