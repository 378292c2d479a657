"""One tiny invocation of the hot path on cuda:0, checked against the oracle (driver smoke test)."""
import torch


def run():
    from oracle import attack as oatk          # checker only (allowed in smoke, see oracle/__init__.py)
    from oracle import models as om
    from . import models as pm
    from . import ops
    from .engine import AttackEngine
    ops.require_device()
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    onet = om.init_model("hyper", 3, seed=0).to(dev)
    pnet = pm.init_model("hyper", 3, "mse", pretrained=False).to(dev)
    pnet.load_state_dict(onet.state_dict())
    x = oatk.synthetic_image(0, 64, 64).to(dev)
    args = oatk.default_args(model="hyper", quality=3, metric="mse", steps=4)
    output_s, _, _ = oatk.clean_pass(x, onet, args)
    pnet.train()
    eng = AttackEngine(pnet, 1, 64, 64, steps=4, use_graph=False)
    eng.load(x, output_s)
    rec = []
    eng.run(3, record=rec)
    # oracle: same three iterations
    onet.train()
    noise = torch.zeros_like(x, requires_grad=True)
    opt = torch.optim.Adam([noise], lr=0.01)
    sch = torch.optim.lr_scheduler.MultiStepLR(opt, [1, 2, 3], gamma=0.33)
    from oracle.layers import low_bound, up_bound
    for i in range(3):
        nc = up_bound(low_bound(noise, -16 / 255), 16 / 255)
        im_in = up_bound(low_bound(x + nc, 0.0), 1.0)
        loss, loss_i, _, br = oatk.attack_our(x, output_s, im_in, onet, args)
        opt.zero_grad(); loss.backward(); opt.step()
        if i % (4 // 3) == 0:
            sch.step()
        pb, pli, plo = int(rec[i][0][0]), float(rec[i][1][0]), float(rec[i][2][0])
        assert (pb == 1) == (br == "B"), (i, pb, br)
        assert abs(pli - float(loss_i)) <= 1e-3 * float(loss_i) + 1e-9, (i, pli, float(loss_i))
        if pb == 1:
            assert abs(plo - float(loss)) <= 1e-3 * abs(float(loss)), (i, plo, float(loss))
    print("smoke ok:", [(int(r[0][0]), float(r[1][0]), float(r[2][0])) for r in rec])
