"""Restorer-level step of the mmedit route: host logic on the CPU, kernels on the GPU."""
import math
import os

import pytest
import torch

from fcvsr_b200 import restorer as R


def test_cosine_restart_schedule_matches_mmcv_formula():
    """CosineRestartLrUpdaterHook with the FCVSR configuration (fcvsr_redsLD_QP22.py:118-127) and with restarts."""
    p = [torch.nn.Parameter(torch.zeros(1))]
    opt = torch.optim.Adam(p, lr=5e-6, betas=(0.9, 0.99))
    sch = R.CosineRestartLR(opt, periods=[600000], restart_weights=[1], min_lr=1e-7)
    assert opt.param_groups[0]["lr"] == pytest.approx(5e-6)
    for t in range(1, 4):
        opt.step()
        sch.step()
        want = 1e-7 + 0.5 * (5e-6 - 1e-7) * (1 + math.cos(math.pi * t / 600000))
        assert opt.param_groups[0]["lr"] == pytest.approx(want, rel=1e-12)
    opt2 = torch.optim.Adam(p, lr=1.0)
    sch2 = R.CosineRestartLR(opt2, periods=[4, 6], restart_weights=[1, 0.5], min_lr=0.1)
    lrs = []
    for _ in range(10):
        lrs.append(opt2.param_groups[0]["lr"])
        opt2.step()
        if len(lrs) < 10:
            sch2.step()
    want = [0.1 + 0.5 * 1.0 * 0.9 * (1 + math.cos(math.pi * t / 4)) for t in range(4)] + \
           [0.1 + 0.5 * 0.5 * 0.9 * (1 + math.cos(math.pi * t / 6)) for t in range(6)]
    assert lrs == pytest.approx(want, rel=1e-12)
    with pytest.raises(ValueError):
        sch2.step()                        # beyond the last period, as mmcv's get_position_from_periods raises


def test_parse_losses_and_checkpoint_layout(tmp_path):
    loss, log_vars = R.BasicVSRRestorer.parse_losses({"loss_pix": torch.tensor([1.0, 3.0]), "aux": torch.tensor(5.0)})
    assert float(loss) == 2.0 and log_vars == {"loss_pix": 2.0, "aux": 5.0, "loss": 2.0}
    gen = torch.nn.Conv2d(3, 3, 1)
    m = R.BasicVSRRestorer(gen, torch.nn.MSELoss(), train_cfg=dict(fix_iter=100))
    assert m.fix_iter == 100 and "step_counter" in m.state_dict() and "generator.weight" in m.state_dict()
    opt = {"generator": torch.optim.Adam(gen.parameters(), lr=1e-3)}
    gen(torch.zeros(1, 3, 4, 4)).sum().backward()
    opt["generator"].step()
    m.step_counter += 7
    path = os.path.join(tmp_path, "iter_7.pth")
    R.save_checkpoint(m, path, optimizer=opt, meta=dict(iter=7))
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"meta", "state_dict", "optimizer"} and set(ck["optimizer"]) == {"generator"}
    m2 = R.BasicVSRRestorer(torch.nn.Conv2d(3, 3, 1), torch.nn.MSELoss(), train_cfg=dict(fix_iter=100))
    opt2 = {"generator": torch.optim.Adam(m2.generator.parameters(), lr=1e-3)}
    meta = R.load_checkpoint(m2, path, optimizer=opt2)
    assert meta["iter"] == 7 and m2._steps == 7 and torch.equal(m2.generator.weight, gen.weight)
    assert opt2["generator"].state_dict()["state"][0]["step"] == opt["generator"].state_dict()["state"][0]["step"]


@pytest.mark.gpu
@pytest.mark.parametrize("cls,kw", [("MSELoss", {}), ("L1Loss", {}), ("CharbonnierLoss", {}), ("MSELoss", dict(reduction="sum", loss_weight=0.5))])
def test_mmedit_pixel_losses_match_torch(cls, kw):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(4)
    x = torch.rand(2, 3, 33, 41, generator=g)
    y = torch.rand(2, 3, 33, 41, generator=g)
    d = x - y
    elem = {"MSELoss": d * d, "L1Loss": d.abs(), "CharbonnierLoss": torch.sqrt(d * d + 1e-12)}[cls]
    want = kw.get("loss_weight", 1.0) * (elem.sum() if kw.get("reduction") == "sum" else elem.mean())
    xr = x.clone().requires_grad_(True)
    ({"MSELoss": (xr - y) ** 2, "L1Loss": (xr - y).abs(), "CharbonnierLoss": torch.sqrt((xr - y) ** 2 + 1e-12)}[cls]).sum().backward()
    gscale = kw.get("loss_weight", 1.0) * (1.0 if kw.get("reduction") == "sum" else 1.0 / x.numel())
    xd = x.to(dev).requires_grad_(True)
    loss = getattr(R, cls)(**kw)(xd, y.to(dev))
    loss.backward()
    assert float(loss) == pytest.approx(float(want), rel=2e-6)
    assert float((xd.grad.cpu() - gscale * xr.grad).abs().max()) <= 2e-6 * gscale * float(xr.grad.abs().max())


@pytest.mark.gpu
def test_restorer_train_step_on_fcvsr_snet(lib):
    """BasicVSR.train_step semantics (basicvsr.py:85-117) around the RGB backbone with this repository's kernels end to end:
    MSELoss(mean), Adam(betas 0.9 / 0.99), CosineRestart, fix_iter, step_counter, checkpoint round trip."""
    from fcvsr_b200 import arch
    from fcvsr_b200.ops.optim import Adam
    dev = torch.device("cuda:0")
    gen = arch.FCVSR_SNet().to(dev)
    gen.load_state_dict(arch.seeded_state_dict("rgb_S", 0))
    model = R.BasicVSRRestorer(gen, R.MSELoss(loss_weight=1.0, reduction="mean"), train_cfg=dict(fix_iter=1)).to(dev)
    opt = {"generator": Adam(gen.parameters(), lr=0.5 * 1e-5, betas=(0.9, 0.99))}
    sch = R.CosineRestartLR(opt["generator"], periods=[600000], restart_weights=[1], min_lr=1e-7)
    g = torch.Generator().manual_seed(8)
    lq = torch.rand(1, 7, 3, 16, 16, generator=g).to(dev)
    gt = torch.rand(1, 7, 3, 64, 64, generator=g).to(dev)
    losses = []
    for _ in range(3):
        out = model.train_step(dict(lq=lq, gt=gt), opt)
        sch.step()
        assert set(out) == {"num_samples", "results", "log_vars"} and out["num_samples"] == 1
        assert out["results"]["output"].shape == (1, 3, 64, 64) and not out["results"]["output"].is_cuda
        assert set(out["log_vars"]) == {"loss_pix", "loss"}
        losses.append(out["log_vars"]["loss"])
    assert float(model.step_counter) == 3.0 and model.is_weight_fixed
    assert all(math.isfinite(v) for v in losses) and losses[2] < losses[0]
