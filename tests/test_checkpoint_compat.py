"""CompressAI / zoo checkpoint compatibility of the operator surface: host logic only, runs
on CPU.  Reference: coder.py:104-116 (load), anchors/balle.py:57-72 + anchors/utils.py:46-109 (dynamic buffer resizing),
train.py:443-454 (save_checkpoint round trip)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "reference")


def _zoo_style(sd, model, new_names):
    """What a CompressAI checkpoint of this model looks like: the parameters plus filled entropy-coder tables."""
    out = {}
    for k, v in sd.items():
        if new_names:
            for old, new in (("_matrix", "matrices."), ("_bias", "biases."), ("_factor", "factors.")):
                k = k.replace("entropy_bottleneck." + old, "entropy_bottleneck." + new)
        out[k] = v.clone()
    C = sd["entropy_bottleneck.quantiles"].shape[0]
    out["entropy_bottleneck._quantized_cdf"] = torch.arange(C * 37, dtype=torch.int32).reshape(C, 37)
    out["entropy_bottleneck._offset"] = torch.full((C,), -17, dtype=torch.int32)
    out["entropy_bottleneck._cdf_length"] = torch.full((C,), 37, dtype=torch.int32)
    if model != "factorized":
        out["gaussian_conditional._quantized_cdf"] = torch.arange(64 * 99, dtype=torch.int32).reshape(64, 99)
        out["gaussian_conditional._offset"] = torch.full((64,), -48, dtype=torch.int32)
        out["gaussian_conditional._cdf_length"] = torch.full((64,), 99, dtype=torch.int32)
        out["gaussian_conditional.scale_table"] = torch.exp(torch.linspace(-2.2, 5.5, 64))
    out.update(_compressai_constant_buffers(sd, model))
    return out


def _compressai_constant_buffers(sd, model):
    """The persistent 1-element buffers a CompressAI (1.1.x - 1.2.x) state_dict carries next to the parameters:
    NonNegativeParametrizer.pedestal, LowerBound.bound (GDN reparametrisers, likelihood bounds, scale bound).
    Key list written from CompressAI's source (compressai/ops/parametrizers.py, ops/bound_ops.py, entropy_models.py);
    the package itself is not installable here, so this list -- not a real checkpoint file -- is what pins it."""
    out = {}
    pedestal = (2.0 ** -18) ** 2
    for k in sd:
        if k.endswith(".beta") and k[:-5] + ".gamma" in sd:          # a GDN module
            base = k[:-5]
            for rep, minimum in (("beta_reparam", 1e-6), ("gamma_reparam", 0.0)):
                out[f"{base}.{rep}.pedestal"] = torch.tensor([pedestal])
                out[f"{base}.{rep}.lower_bound.bound"] = torch.tensor([(minimum + pedestal) ** 0.5])
    out["entropy_bottleneck.likelihood_lower_bound.bound"] = torch.tensor([1e-9])
    if model != "factorized":
        out["gaussian_conditional.likelihood_lower_bound.bound"] = torch.tensor([1e-9])
        out["gaussian_conditional.lower_bound_scale.bound"] = torch.tensor([0.11])
    return out


@pytest.mark.parametrize("model,quality", [("factorized", 1), ("hyper", 3), ("context", 4), ("cheng2020", 1)])
@pytest.mark.parametrize("new_names", [False, True])
def test_zoo_style_state_dict_loads_and_round_trips(model, quality, new_names):
    from imagecompression_adversarial_b200 import models as pm
    from oracle import models as om
    onet = om.init_model(model, quality, seed=3)
    zoo = _zoo_style(onet.state_dict(), model, new_names)
    net = pm.init_model(model, quality, "mse", pretrained=False)
    res = net.load_state_dict(zoo, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    sd = net.state_dict()
    # what this package writes has exactly the key set of a CompressAI checkpoint of the model (old-style names)
    canon = _zoo_style(onet.state_dict(), model, False)
    assert set(sd.keys()) == set(canon.keys()), set(sd.keys()) ^ set(canon.keys())
    for k, v in _compressai_constant_buffers(onet.state_dict(), model).items():
        assert sd[k].shape == (1,) and abs(float(sd[k]) - float(v)) <= 1e-6 * abs(float(v)), k
    for k, v in onet.state_dict().items():          # every parameter arrived (old-style names are the canonical ones)
        assert torch.equal(sd[k], v), k
    for k, v in zoo.items():
        if "cdf" in k or "_offset" in k or k.endswith("scale_table"):
            assert sd[k].shape == v.shape and torch.equal(sd[k], v.to(sd[k].dtype)), k
    # save_checkpoint / resume round trip (train.py:443-454, coder.py:104-108) into a FRESH model
    net2 = pm.init_model(model, quality, "mse", pretrained=False)
    res = net2.load_state_dict({"state_dict": sd}["state_dict"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in sd.items():
        assert torch.equal(net2.state_dict()[k], v), k


def test_plain_state_dict_without_tables_still_loads_strictly():
    from imagecompression_adversarial_b200 import models as pm
    from oracle import models as om
    onet = om.init_model("hyper", 3, seed=0)
    net = pm.init_model("hyper", 3, "mse", pretrained=False)
    res = net.load_state_dict(onet.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert net.entropy_bottleneck._quantized_cdf.numel() == 0       # still empty: nothing to resize to
    bad = dict(onet.state_dict())
    bad.pop("g_a.0.weight")
    with pytest.raises(RuntimeError):
        net.load_state_dict(bad, strict=True)                       # real parameters stay mandatory


def test_reference_update_registered_buffers_works_on_the_operator_surface():
    """The reference's own helper (anchors/utils.py:75-109), unmodified, resizes this package's buffers the way
    anchors/balle.py:57-72 calls it."""
    if not os.path.exists(os.path.join(REF, "anchors", "utils.py")):
        pytest.skip("no reference copy under baseline/_ref/reference")
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_anchor_utils", os.path.join(REF, "anchors", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except ImportError as e:                                         # the file imports compressai at module level
        from imagecompression_adversarial_b200 import launch
        launch.install_shims()
        spec.loader.exec_module(mod)
    from imagecompression_adversarial_b200 import models as pm
    from oracle import models as om
    onet = om.init_model("hyper", 3, seed=1)
    zoo = {"net." + k: v for k, v in _zoo_style(onet.state_dict(), "hyper", False).items()}
    net = pm.init_model("hyper", 3, "mse", pretrained=False)
    mod.update_registered_buffers(net.entropy_bottleneck, "net.entropy_bottleneck",
                                  ["_quantized_cdf", "_offset", "_cdf_length"], zoo)
    mod.update_registered_buffers(net.gaussian_conditional, "net.gaussian_conditional",
                                  ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], zoo)
    assert tuple(net.entropy_bottleneck._quantized_cdf.shape) == (128, 37)
    assert tuple(net.gaussian_conditional.scale_table.shape) == (64,)
    wrapper = torch.nn.Module()
    wrapper.net = net
    res = wrapper.load_state_dict(zoo, strict=True)
    assert not res.missing_keys and not res.unexpected_keys


def test_image_folder_stand_in(tmp_path):
    """compressai.datasets.ImageFolder as train.py:114-123 constructs it (root / split / files, transform)."""
    import numpy as np
    from PIL import Image
    from imagecompression_adversarial_b200.shims.compressai.datasets import ImageFolder
    (tmp_path / "train").mkdir()
    for i in range(3):
        Image.fromarray((np.random.RandomState(i).rand(20, 24, 3) * 255).astype("uint8")).save(tmp_path / "train" / f"{i}.png")
    (tmp_path / "train" / "notes.txt").write_text("not an image")
    ds = ImageFolder(str(tmp_path), split="train", transform=lambda im: np.asarray(im).shape)
    assert len(ds) == 3 and ds[1] == (20, 24, 3)
    with pytest.raises(RuntimeError):
        ImageFolder(str(tmp_path), split="test")
