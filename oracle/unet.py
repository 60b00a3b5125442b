"""Oracle UNet2DConditionModel, SD-2.1 config.  TEST INFRASTRUCTURE ONLY.

Dataflow follows /root/reference/genpercept/models/custom_unet.py: time path :146-170,
conv_in :273, down loop :305-327, mid :341-352, up loop :369-400 (with the multi_level_feats tap
at :400), early return for the DPT readout :402-408, conv_norm_out/SiLU/conv_out :411-415.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .blocks import (Downsample2D, ResnetBlock2D, Timesteps, TimestepEmbedding, Transformer2DModel,
                     Upsample2D)

BLOCK_OUT = (320, 640, 1280, 1280)
HEADS = (5, 10, 20, 20)   # config "attention_head_dim" is used as num_heads -> head_dim 64
CROSS_DIM = 1024


class CrossAttnDownBlock2D(nn.Module):
    def __init__(self, cin, cout, heads, add_downsample=True):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin, cout), ResnetBlock2D(cout, cout)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, cout // heads, cout) for _ in range(2)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=1)]) if add_downsample else None

    def forward(self, x, temb, ctx):
        outs = ()
        for r, a in zip(self.resnets, self.attentions):
            x = a(r(x, temb), ctx)
            outs += (x,)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs += (x,)
        return x, outs


class DownBlock2D(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin, cout), ResnetBlock2D(cout, cout)])

    def forward(self, x, temb, ctx=None):
        outs = ()
        for r in self.resnets:
            x = r(x, temb)
            outs += (x,)
        return x, outs


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, c, heads):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c), ResnetBlock2D(c, c)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, c // heads, c)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    """UpBlock2D (heads=None) or CrossAttnUpBlock2D."""

    def __init__(self, cin_prev, cout, skip_channels, heads=None, add_upsample=True):
        super().__init__()
        res = []
        for i in range(3):
            cin = cin_prev if i == 0 else cout
            res.append(ResnetBlock2D(cin + skip_channels[i], cout))
        self.resnets = nn.ModuleList(res)
        if heads is not None:
            self.attentions = nn.ModuleList([Transformer2DModel(heads, cout // heads, cout) for _ in range(3)])
        else:
            self.attentions = None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None

    def forward(self, x, skips, temb, ctx, upsample_size=None):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips[-1 - i]], dim=1)   # pops from the end of the skip stack
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x, upsample_size)
        return x


class UNet2DConditionModel(nn.Module):
    def __init__(self, in_channels=4, out_channels=4):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, 320, 3, padding=1)
        self.time_proj = Timesteps(320)
        self.time_embedding = TimestepEmbedding(320, 1280)
        self.down_blocks = nn.ModuleList([
            CrossAttnDownBlock2D(320, 320, 5),
            CrossAttnDownBlock2D(320, 640, 10),
            CrossAttnDownBlock2D(640, 1280, 20),
            DownBlock2D(1280, 1280),
        ])
        self.mid_block = UNetMidBlock2DCrossAttn(1280, 20)
        # skip stack (bottom->top popping order): see App. A.2
        self.up_blocks = nn.ModuleList([
            UpBlock(1280, 1280, (1280, 1280, 1280), heads=None),
            UpBlock(1280, 1280, (1280, 1280, 640), heads=20),
            UpBlock(1280, 640, (640, 640, 320), heads=10),
            UpBlock(640, 320, (320, 320, 320), heads=5, add_upsample=False),
        ])
        self.conv_norm_out = nn.GroupNorm(32, 320, eps=1e-5)
        self.conv_out = nn.Conv2d(320, out_channels, 3, padding=1)

    def time_embed(self, timestep, batch):
        t = torch.as_tensor(timestep).reshape(-1).expand(batch)
        # custom_unet.py:168: the fp32 sinusoid is cast to the sample dtype before the MLP
        return self.time_embedding(self.time_proj(t).to(self.conv_in.weight.dtype))

    def forward(self, sample, timestep, encoder_hidden_states, return_feature=False):
        b = sample.shape[0]
        # custom_unet.py:105-119 — only forward upsample sizes when not a multiple of 2**3
        forward_upsample_size = any(s % 8 != 0 for s in sample.shape[-2:])
        emb = self.time_embed(timestep, b)
        ctx = encoder_hidden_states
        x = self.conv_in(sample)
        skips = (x,)
        for blk in self.down_blocks:
            x, outs = blk(x, emb, ctx)
            skips += outs
        x = self.mid_block(x, emb, ctx)
        feats = []
        for i, blk in enumerate(self.up_blocks):
            is_final = i == len(self.up_blocks) - 1
            s = skips[-3:]
            skips = skips[:-3]
            up_size = None
            if not is_final and forward_upsample_size:
                up_size = skips[-1].shape[2:]
            x = blk(x, s, emb, ctx, up_size)
            feats.append(x)
        if return_feature:
            return feats      # [1280@h/4, 1280@h/2, 640@h, 320@h]  (custom_unet.py:400-408)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return x
