"""Oracle ``single_infer``: restates /root/reference/genpercept/genpercept_pipeline.py:375-526.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  fp32 on CPU.
"""
import torch

from .dpt import DPTNeckHeadIdentity
from .scheduler import DDIMOneStep
from .unet import UNet2DConditionModel
from .vae import AutoencoderKL

LATENT_SCALE = 0.18215          # genpercept_pipeline.py:96
ONE_CHANNEL_MODES = ("depth", "matting", "dis", "disparity")   # genpercept_pipeline.py:523


class OraclePipeline:
    def __init__(self, state, text_embed, use_dpt=False, dtype=torch.float32):
        """state: {"unet": sd, "vae": sd, "dpt": sd or None} with diffusers keys (fp32 tensors).
        dtype=torch.float16 emulates the reference's ``--half_precision`` run (run.py:273-281:
        every module and activation in fp16) and is only used to calibrate test tolerances."""
        self.dtype = dtype
        self.unet = UNet2DConditionModel().eval()
        self.vae = AutoencoderKL().eval()
        sd_unet = dict(state["unet"])
        if use_dpt:   # run.py:322-331 deletes conv_out / conv_norm_out for the DPT readout
            miss = self.unet.load_state_dict(sd_unet, strict=False)
            assert all(k.startswith(("conv_out", "conv_norm_out")) for k in miss.missing_keys), miss
        else:
            self.unet.load_state_dict(sd_unet, strict=True)
        self.vae.load_state_dict(state["vae"], strict=True)
        self.head = None
        if use_dpt:
            self.head = DPTNeckHeadIdentity().eval()
            self.head.load_state_dict(state["dpt"], strict=True)
        self.text_embed = text_embed.float().reshape(1, -1, 1024).to(dtype)
        self.scheduler = DDIMOneStep()
        if dtype != torch.float32:
            self.unet.to(dtype)
            self.vae.to(dtype)
            if self.head is not None:
                self.head.to(dtype)

    @torch.no_grad()
    def encode_rgb(self, rgb_in):                       # :488-505
        h = self.vae.encoder(rgb_in)
        moments = self.vae.quant_conv(h)
        mean, _ = torch.chunk(moments, 2, dim=1)
        return mean * LATENT_SCALE

    @torch.no_grad()
    def decode_pred(self, pred_latent, mode):           # :507-526
        z = self.vae.post_quant_conv(pred_latent / LATENT_SCALE)
        stacked = self.vae.decoder(z)
        if mode in ONE_CHANNEL_MODES:
            stacked = stacked.mean(dim=1, keepdim=True)
        return stacked

    @torch.no_grad()
    def unet_forward(self, latent, timestep=1, return_feature=False):
        ctx = self.text_embed.repeat(latent.shape[0], 1, 1)
        return self.unet(latent, torch.tensor([timestep]), ctx, return_feature=return_feature)

    @torch.no_grad()
    def single_infer(self, rgb_in, mode="depth", fix_timesteps=None, return_intermediates=False):
        """rgb_in: [B,3,H,W] float in [-1,1].  Returns [B,1|3,H,W] in [0,1]."""
        timesteps = self.scheduler.set_timesteps(1)
        if fix_timesteps:
            timesteps = torch.tensor([fix_timesteps]).long()
        rgb_latent = self.encode_rgb(rgb_in.to(self.dtype))
        pred_latent = rgb_latent
        inter = {"rgb_latent": rgb_latent}
        if self.head is None:
            t = timesteps[0]
            noise_pred = self.unet_forward(pred_latent, int(t))
            _, x0 = self.scheduler.step(noise_pred, 1, pred_latent)   # scheduler indexes its own t=1
            inter["unet_out"] = noise_pred
            inter["pred_latent"] = x0
            pred = self.decode_pred(x0, mode)
            inter["decoded"] = pred
            pred = torch.clip(pred, -1.0, 1.0)
            pred = (pred + 1.0) / 2.0
        else:
            feats = self.unet_forward(pred_latent, int(timesteps[0]), return_feature=True)[::-1]
            inter["feats"] = feats
            pred = self.head(feats)[:, None]
            inter["head_out"] = pred
            # genpercept_pipeline.py:482 normalises over the whole tensor; the reference only ever
            # runs B=1 (F10/F12), so the batched restatement is per image.
            mn = pred.amin(dim=(1, 2, 3), keepdim=True)
            mx = pred.amax(dim=(1, 2, 3), keepdim=True)
            pred = (pred - mn) / (mx - mn)
        return (pred, inter) if return_intermediates else pred
