"""Oracle AutoencoderKL (SD-2.1 VAE config).  TEST INFRASTRUCTURE ONLY.

Call sites restated: /root/reference/genpercept/genpercept_pipeline.py:500-504 (encoder,
quant_conv, mean * 0.18215) and :519-525 (/0.18215, post_quant_conv, decoder, channel mean).
"""
import torch.nn as nn
import torch.nn.functional as F

from .blocks import Attention, Downsample2D, ResnetBlock2D, Upsample2D

VAE_BLOCK_OUT = (128, 256, 512, 512)


def _res(cin, cout):
    return ResnetBlock2D(cin, cout, temb_channels=None, eps=1e-6)


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([_res(cin, cout), _res(cout, cout)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=0)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin, cout, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList([_res(cin if i == 0 else cout, cout) for i in range(3)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNetMidBlock2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.resnets = nn.ModuleList([_res(c, c), _res(c, c)])
        self.attentions = nn.ModuleList([
            Attention(c, heads=1, dim_head=c, bias=True, norm_num_groups=32, eps=1e-6,
                      residual_connection=True)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv_in = nn.Conv2d(3, 128, 3, padding=1)
        chans = (128,) + VAE_BLOCK_OUT
        self.down_blocks = nn.ModuleList(
            [DownEncoderBlock2D(chans[i], chans[i + 1], add_downsample=i < 3) for i in range(4)])
        self.mid_block = UNetMidBlock2D(512)
        self.conv_norm_out = nn.GroupNorm(32, 512, eps=1e-6)
        self.conv_out = nn.Conv2d(512, 8, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv_in = nn.Conv2d(4, 512, 3, padding=1)
        self.mid_block = UNetMidBlock2D(512)
        rev = (512, 512, 256, 128)
        prev = (512, 512, 512, 256)
        self.up_blocks = nn.ModuleList(
            [UpDecoderBlock2D(prev[i], rev[i], add_upsample=i < 3) for i in range(4)])
        self.conv_norm_out = nn.GroupNorm(32, 128, eps=1e-6)
        self.conv_out = nn.Conv2d(128, 3, 3, padding=1)

    def forward(self, z):
        x = self.conv_in(z)
        x = self.mid_block(x)
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKL(nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder = Encoder()
        self.decoder = Decoder()
        self.quant_conv = nn.Conv2d(8, 8, 1)
        self.post_quant_conv = nn.Conv2d(4, 4, 1)
