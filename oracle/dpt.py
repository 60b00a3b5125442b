"""Oracle DPT neck+head, restating /root/reference/genpercept/models/dpt_head.py
(DPTNeckHeadForUnetAfterUpsampleIdentity :585, forward core :530-546) with the constants of
/root/reference/hf_configs/dpt-sd2.1-unet-after-upsample-general/config.json.
TEST INFRASTRUCTURE ONLY.  Pinned against the reference class itself in tests/test_oracle.py.
"""
import torch.nn as nn
import torch.nn.functional as F

NECK_SIZES = (320, 640, 1280, 1280)
FUSION = 256


class _Up(nn.Module):           # dpt_head.py:92 Upsample2D(use_conv=True): nearest x2 + 3x3 conv
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class PreActResidual(nn.Module):   # dpt_head.py:213-271, no bias, no BN
    def __init__(self):
        super().__init__()
        self.convolution1 = nn.Conv2d(FUSION, FUSION, 3, padding=1, bias=False)
        self.convolution2 = nn.Conv2d(FUSION, FUSION, 3, padding=1, bias=False)

    def forward(self, x):
        return x + self.convolution2(F.relu(self.convolution1(F.relu(x))))


class FusionLayer(nn.Module):      # dpt_head.py:274-309
    def __init__(self, with_residual_1=True):
        super().__init__()
        self.projection = nn.Conv2d(FUSION, FUSION, 1, bias=True)
        if with_residual_1:
            self.residual_layer1 = PreActResidual()
        self.residual_layer2 = PreActResidual()

    def forward(self, x, residual=None):
        if residual is not None:
            if x.shape != residual.shape:
                residual = F.interpolate(residual, size=x.shape[2:], mode="bilinear", align_corners=False)
            x = x + self.residual_layer1(residual)
        x = self.residual_layer2(x)
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        return self.projection(x)


class FusionStage(nn.Module):      # dpt_head.py:312-335
    def __init__(self):
        super().__init__()
        self.layers = nn.ModuleList([FusionLayer(with_residual_1=(i != 0)) for i in range(4)])

    def forward(self, hs):
        hs = hs[::-1]
        out = []
        x = self.layers[0](hs[0])
        out.append(x)
        for h, layer in zip(hs[1:], self.layers[1:]):
            x = layer(x, h)
            out.append(x)
        return out


class Neck(nn.Module):             # dpt_head.py:338-388 (reassemble_stage=None)
    def __init__(self):
        super().__init__()
        self.convs = nn.ModuleList([nn.Conv2d(c, FUSION, 3, padding=1, bias=False) for c in NECK_SIZES])
        self.fusion_stage = FusionStage()

    def forward(self, hs):
        return self.fusion_stage([self.convs[i](h) for i, h in enumerate(hs)])


class HeadIdentity(nn.Module):     # dpt_head.py:564-582 + :80-90
    def __init__(self):
        super().__init__()
        self.projection = nn.Conv2d(FUSION, FUSION, 3, padding=1)
        self.head = nn.Sequential(
            nn.Conv2d(FUSION, FUSION // 2, 3, padding=1),
            nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True),
            nn.Conv2d(FUSION // 2, 32, 3, padding=1),
            nn.ReLU(),
            nn.Conv2d(32, 1, 1),
            nn.Identity(),
        )

    def forward(self, hs):
        x = F.relu(self.projection(hs[-1]))
        return self.head(x).squeeze(1)


class DPTNeckHeadIdentity(nn.Module):
    """Input: [320@h, 640@h, 1280@h/2, 1280@h/4] (already reversed, genpercept_pipeline.py:479)."""

    def __init__(self):
        super().__init__()
        self.feature_upsample_0 = _Up(NECK_SIZES[0])
        self.neck = Neck()
        self.head = HeadIdentity()

    def forward(self, hidden_states):
        hs = list(hidden_states)
        hs[0] = self.feature_upsample_0(hs[0])
        return self.head(self.neck(hs))
