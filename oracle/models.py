"""Oracle codecs: the four CompressAI zoo families the reference instantiates
(``/root/reference/anchors/model.py:60-78``), restated from SURVEY.md Appendix A.2.

Test infrastructure only (see ``oracle/__init__.py``).  Forward passes follow the reference's
own spelled-out versions: ``anchors/model.py:86-108`` (``entropy_estimator``) and
``anchors/balle.py:25-55`` (note ``h_a(abs(y))`` for the hyperprior, balle.py:38).
"""
import torch
import torch.nn as nn

from .layers import (GDN, AttentionBlock, EntropyBottleneck, GaussianConditional, MaskedConv2d, ResidualBlock,
                     ResidualBlockUpsample, ResidualBlockWithStride, conv, conv3x3, deconv,
                     subpel_conv3x3)

# compressai.zoo.image.cfgs (A.0)
ZOO = {
    "factorized": {q: (128, 192) if q <= 5 else (192, 320) for q in range(1, 9)},
    "hyper": {q: (128, 192) if q <= 5 else (192, 320) for q in range(1, 9)},
    "context": {q: (192, 192) if q <= 4 else (192, 320) for q in range(1, 9)},
    "cheng2020": {q: (128,) if q <= 3 else (192,) for q in range(1, 7)},
    "cheng2020_attn": {q: (128,) if q <= 3 else (192,) for q in range(1, 7)},
    "debug": {q: (3, 192) for q in range(1, 9)},      # anchors/model.py:61-68: ae_onelayer(N=3, M=192) at every quality
}

# CompressAI's published parameter counts (A.8) -- structural checksum of this restatement
PARAM_COUNTS = {
    ("factorized", 128, 192): 2_998_147, ("factorized", 192, 320): 7_030_531,
    ("hyper", 128, 192): 5_075_843, ("hyper", 192, 320): 11_816_323,
    ("context", 192, 192): 14_130_467, ("context", 192, 320): 25_504_596,
    ("cheng2020", 128): 11_833_149, ("cheng2020", 192): 26_598_956,
    ("cheng2020_attn", 128): 13_183_293,      # published; N=192 follows from the same layer list (29 631 788)
}


class CompressionModel(nn.Module):
    def __init__(self, entropy_bottleneck_channels):
        super().__init__()
        self.entropy_bottleneck = EntropyBottleneck(entropy_bottleneck_channels)

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def _init_weights(self):
        # CompressAI 1.1.x CompressionModel._initialize_weights
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)


def _g_a(N, M):
    return nn.Sequential(conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))


def _g_s(N, M):
    return nn.Sequential(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
                         deconv(N, N), GDN(N, inverse=True), deconv(N, 3))


class FactorizedPrior(CompressionModel):
    def __init__(self, N, M):
        super().__init__(M)
        self.g_a, self.g_s = _g_a(N, M), _g_s(N, M)
        self.N, self.M = N, M
        self._init_weights()

    def forward(self, x):
        y = self.g_a(x)
        y_hat, y_lik = self.entropy_bottleneck(y)  # anchors/model.py:87-89
        return {"x_hat": self.g_s(y_hat), "likelihoods": {"y": y_lik}}


class ScaleHyperprior(CompressionModel):
    def __init__(self, N, M):
        super().__init__(N)
        self.g_a, self.g_s = _g_a(N, M), _g_s(N, M)
        self.h_a = nn.Sequential(conv(M, N, 3, 1), nn.ReLU(inplace=True), conv(N, N),
                                 nn.ReLU(inplace=True), conv(N, N))
        self.h_s = nn.Sequential(deconv(N, N), nn.ReLU(inplace=True), deconv(N, N),
                                 nn.ReLU(inplace=True), conv(N, M, 3, 1), nn.ReLU(inplace=True))
        self.gaussian_conditional = GaussianConditional()
        self.N, self.M = N, M
        self._init_weights()

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(torch.abs(y))  # anchors/model.py:92, anchors/balle.py:38
        z_hat, z_lik = self.entropy_bottleneck(z)
        scales_hat = self.h_s(z_hat)
        y_hat, y_lik = self.gaussian_conditional(y, scales_hat)
        return {"x_hat": self.g_s(y_hat), "likelihoods": {"y": y_lik, "z": z_lik}}


class JointAutoregressiveHierarchicalPriors(CompressionModel):
    """mbt2018 (widths pinned by InvCompress/ours.py:22-32)."""

    def __init__(self, N, M):
        super().__init__(N)
        self.g_a, self.g_s = _g_a(N, M), _g_s(N, M)
        lr = lambda: nn.LeakyReLU(inplace=True)
        self.h_a = nn.Sequential(conv(M, N, 3, 1), lr(), conv(N, N), lr(), conv(N, N))
        self.h_s = nn.Sequential(deconv(N, M), lr(), deconv(M, M * 3 // 2), lr(),
                                 conv(M * 3 // 2, M * 2, 3, 1))
        self.entropy_parameters = nn.Sequential(
            nn.Conv2d(M * 12 // 3, M * 10 // 3, 1), lr(),
            nn.Conv2d(M * 10 // 3, M * 8 // 3, 1), lr(),
            nn.Conv2d(M * 8 // 3, M * 6 // 3, 1))
        self.context_prediction = MaskedConv2d(M, 2 * M, kernel_size=5, padding=2, stride=1)
        self.gaussian_conditional = GaussianConditional()
        self.N, self.M = N, M
        self._init_weights()

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(y)  # anchors/model.py:98 (no abs)
        z_hat, z_lik = self.entropy_bottleneck(z)
        params = self.h_s(z_hat)
        y_hat = self.gaussian_conditional.quantize(
            y, "noise" if self.training else "dequantize",
            noise=self.gaussian_conditional.noise_override)
        ctx = self.context_prediction(y_hat)
        gp = self.entropy_parameters(torch.cat((params, ctx), dim=1))
        scales_hat, means_hat = gp.chunk(2, 1)
        _, y_lik = self.gaussian_conditional(y, scales_hat, means=means_hat)
        return {"x_hat": self.g_s(y_hat), "likelihoods": {"y": y_lik, "z": z_lik}}


class MeanScaleHyperprior(ScaleHyperprior):
    """compressai MeanScaleHyperprior (mbt2018_mean): the base class of the reference's ``debug`` model."""

    def __init__(self, N, M):
        super().__init__(N, M)
        lr = lambda: nn.LeakyReLU(inplace=True)
        self.h_a = nn.Sequential(conv(M, N, 3, 1), lr(), conv(N, N), lr(), conv(N, N))
        self.h_s = nn.Sequential(deconv(N, M), lr(), deconv(M, M * 3 // 2), lr(), conv(M * 3 // 2, M * 2, 3, 1))
        self._init_weights()

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(y)
        z_hat, z_lik = self.entropy_bottleneck(z)
        scales_hat, means_hat = self.h_s(z_hat).chunk(2, 1)
        y_hat, y_lik = self.gaussian_conditional(y, scales_hat, means=means_hat)
        return {"x_hat": self.g_s(y_hat), "likelihoods": {"y": y_lik, "z": z_lik}}


class AeOneLayer(MeanScaleHyperprior):
    """The reference's ``debug`` model (anchors/model.py:9-35): one 3x3 stride-1 conv as ``g_a``, one 3x3 stride-1
    transposed conv as ``g_s``; its forward decodes the UNQUANTISED latent (``x_hat = g_s(y)``, :31-32) while the
    likelihoods come from the mean-scale hyperprior."""

    def __init__(self, N, M):
        super().__init__(N, M)
        self.g_a = nn.Sequential(conv(3, M, kernel_size=3, stride=1))
        self.g_s = nn.Sequential(deconv(M, 3, kernel_size=3, stride=1))
        self._init_weights()

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(y)
        z_hat, z_lik = self.entropy_bottleneck(z)
        scales_hat, means_hat = self.h_s(z_hat).chunk(2, 1)
        _, y_lik = self.gaussian_conditional(y, scales_hat, means=means_hat)
        return {"x_hat": self.g_s(y), "likelihoods": {"y": y_lik, "z": z_lik}}


class Cheng2020Anchor(JointAutoregressiveHierarchicalPriors):
    """cheng2020_anchor: residual blocks + sub-pixel convs, no attention, single Gaussian
    (anchors/model.py:77; widths pinned by InvCompress/ours.py:33-55)."""

    def __init__(self, N):
        super().__init__(N, N)
        lr = lambda: nn.LeakyReLU(inplace=True)
        self.g_a = nn.Sequential(
            ResidualBlockWithStride(3, N, 2), ResidualBlock(N, N),
            ResidualBlockWithStride(N, N, 2), ResidualBlock(N, N),
            ResidualBlockWithStride(N, N, 2), ResidualBlock(N, N), conv3x3(N, N, 2))
        self.h_a = nn.Sequential(conv3x3(N, N), lr(), conv3x3(N, N), lr(), conv3x3(N, N, 2), lr(),
                                 conv3x3(N, N), lr(), conv3x3(N, N, 2))
        self.h_s = nn.Sequential(conv3x3(N, N), lr(), subpel_conv3x3(N, N, 2), lr(),
                                 conv3x3(N, N * 3 // 2), lr(),
                                 subpel_conv3x3(N * 3 // 2, N * 3 // 2, 2), lr(),
                                 conv3x3(N * 3 // 2, N * 2))
        self.g_s = nn.Sequential(
            ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2), ResidualBlock(N, N),
            ResidualBlockUpsample(N, N, 2), ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2),
            ResidualBlock(N, N), subpel_conv3x3(N, 3, 2))
        self._init_weights()


class Cheng2020Attention(Cheng2020Anchor):
    """cheng2020_attn (the zoo entry BASELINE config 4 names): the anchor plus four AttentionBlocks."""

    def __init__(self, N):
        super().__init__(N)
        self.g_a = nn.Sequential(
            ResidualBlockWithStride(3, N, 2), ResidualBlock(N, N),
            ResidualBlockWithStride(N, N, 2), AttentionBlock(N), ResidualBlock(N, N),
            ResidualBlockWithStride(N, N, 2), ResidualBlock(N, N), conv3x3(N, N, 2), AttentionBlock(N))
        self.g_s = nn.Sequential(
            AttentionBlock(N), ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2), ResidualBlock(N, N),
            ResidualBlockUpsample(N, N, 2), AttentionBlock(N), ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2),
            ResidualBlock(N, N), subpel_conv3x3(N, 3, 2))
        self._init_weights()


def init_model(model, quality, metric="mse", pretrained=False, seed=None):
    """Mirror of ``anchors.model.init_model`` (anchors/model.py:60-78) on the oracle classes.

    ``pretrained`` must be False (no network; the parity protocol shares a seeded state_dict).
    """
    assert not pretrained, "oracle has no zoo weights"
    if seed is not None:
        torch.manual_seed(seed)
    cfg = ZOO[model][quality]
    if model == "factorized":
        return FactorizedPrior(*cfg)
    if model == "hyper":
        return ScaleHyperprior(*cfg)
    if model == "context":
        return JointAutoregressiveHierarchicalPriors(*cfg)
    if model == "cheng2020":
        return Cheng2020Anchor(*cfg)
    if model == "cheng2020_attn":
        return Cheng2020Attention(*cfg)
    if model == "debug":
        return AeOneLayer(*cfg)
    raise ValueError(model)


def count_parameters(net):
    return sum(p.numel() for p in net.parameters())
