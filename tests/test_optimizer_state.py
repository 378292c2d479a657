"""FusedAdamClip as a torch.optim.Optimizer: the reference's checkpoint layout (train.py:443-454 writes
``optimizer.state_dict()`` / ``aux_optimizer.state_dict()``; coder.py:104-116 reads them back and reads
``optimizer.param_groups[0]['lr']``) and scheduler attachment.  Host logic only (no kernel launch): runs on CPU."""
import io

import pytest
import torch


def _params():
    torch.manual_seed(0)
    return [torch.nn.Parameter(torch.randn(*s)) for s in ((8, 4, 3, 3), (8,), (5, 5), (3, 1, 1))]


def test_state_dict_has_adam_layout_and_round_trips():
    from imagecompression_adversarial_b200.training import FusedAdamClip
    ps = _params()
    opt = FusedAdamClip(ps, lr=1e-4, max_norm=1.0)
    assert opt.state_dict()["state"] == {}                      # like torch.optim.Adam before its first step
    opt.m.copy_(torch.randn_like(opt.m)); opt.v.copy_(torch.rand_like(opt.v)); opt.steps = 7
    sd = opt.state_dict()
    ref = torch.optim.Adam(_params(), lr=1e-4)
    assert set(sd["param_groups"][0]) >= {"lr", "betas", "eps", "weight_decay", "amsgrad", "params"}
    assert sd["param_groups"][0]["params"] == ref.state_dict()["param_groups"][0]["params"]
    for i, p in enumerate(ps):
        assert set(sd["state"][i]) == {"step", "exp_avg", "exp_avg_sq"}
        assert sd["state"][i]["exp_avg"].shape == p.shape and float(sd["state"][i]["step"]) == 7.0
    # the reference's checkpoint dictionary through torch.save / torch.load into a fresh optimiser
    buf = io.BytesIO()
    torch.save({"epoch": 3, "step": 70, "optimizer": sd}, buf)
    buf.seek(0)
    ck = torch.load(buf, weights_only=False)
    opt2 = FusedAdamClip(_params(), lr=5e-3, max_norm=1.0)
    opt2.load_state_dict(ck["optimizer"])
    assert opt2.steps == 7 and opt2.param_groups[0]["lr"] == 1e-4
    assert torch.equal(opt2.m, opt.m) and torch.equal(opt2.v, opt.v)
    # a torch.optim.Adam state dict of the same parameters loads too (resuming a reference-trained run)
    ref_ps = _params()
    ref = torch.optim.Adam(ref_ps, lr=2e-4)
    for p in ref_ps:
        p.grad = torch.randn_like(p)
    ref.step()
    opt3 = FusedAdamClip(_params(), lr=1e-4, max_norm=1.0)
    opt3.load_state_dict(ref.state_dict())
    assert opt3.steps == 1 and opt3.param_groups[0]["lr"] == 2e-4
    off = 0
    for i, p in enumerate(ref_ps):
        assert torch.equal(opt3.m[off:off + p.numel()].view(p.shape), ref.state[p]["exp_avg"])
        off += p.numel()
    with pytest.raises(ValueError):
        FusedAdamClip(_params()[:2], lr=1e-4).load_state_dict(sd)


def test_lr_scheduler_attaches_and_drives_the_live_learning_rate():
    from imagecompression_adversarial_b200.training import FusedAdamClip
    opt = FusedAdamClip(_params(), lr=1e-4, max_norm=1.0)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, "min", factor=0.5, patience=0)
    sched.step(1.0); sched.step(2.0)                             # no improvement -> lr halves
    assert abs(opt.param_groups[0]["lr"] - 5e-5) < 1e-12 and abs(opt.lr - 5e-5) < 1e-12
    ps = _params()
    opt = FusedAdamClip(ps, lr=1e-4)
    # parameters keep their identity but live in the flat buffer; gradients are views of the flat gradient
    assert all(p.data_ptr() >= opt.flat.data_ptr() for p in ps)
    ps[0].grad.fill_(1.0)
    assert float(opt.flat_grad[:ps[0].numel()].sum()) == ps[0].numel()
    opt.zero_grad()
    assert float(opt.flat_grad.abs().sum()) == 0.0
