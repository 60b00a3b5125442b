"""Oracle of the multi-step archs.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates the denoising loop of /root/reference/genpercept/genpercept_pipeline.py:399-472 for
``genpercept_pipeline=False``:
  * marigold    (rgb_blending False): pred_latent = randn; unet_input = cat([rgb_latent, pred_latent]) (8-channel conv_in,
    run.py:59-78); scheduler.step per timestep;
  * rgb_blending (True): pred_latent = rgb_latent; unet_input = pred_latent;
then ``decode_pred(step_output.pred_original_sample)``, clip, shift.  The scheduler is oracle/scheduler.py's restatement of
DDIMSchedulerCustomized + diffusers' DDIM step (eta = 0, v_prediction) with the reference's hf_configs/scheduler_beta_*.
"""
import torch

from .pipeline import LATENT_SCALE, ONE_CHANNEL_MODES
from .scheduler import DDIMOneStep
from .unet import UNet2DConditionModel
from .vae import AutoencoderKL


class OracleMultiStep:
    def __init__(self, state, text_embed, rgb_blending, beta_start=0.00085, beta_end=0.012):
        in_ch = int(state["unet"]["conv_in.weight"].shape[1])
        assert in_ch == (4 if rgb_blending else 8)
        self.rgb_blending = rgb_blending
        self.unet = UNet2DConditionModel(in_channels=in_ch).eval()
        self.unet.load_state_dict(state["unet"], strict=True)
        self.vae = AutoencoderKL().eval()
        self.vae.load_state_dict(state["vae"], strict=True)
        self.text_embed = text_embed.float().reshape(1, -1, 1024)
        self.scheduler = DDIMOneStep(beta_start=beta_start, beta_end=beta_end)

    @torch.no_grad()
    def single_infer(self, rgb_in, num_inference_steps, noise=None, mode="depth", fix_timesteps=None):
        ts = self.scheduler.set_timesteps(num_inference_steps)
        steps = [int(t) for t in ts]
        h = self.vae.encoder(rgb_in.float())
        rgb_latent = torch.chunk(self.vae.quant_conv(h), 2, dim=1)[0] * LATENT_SCALE
        pred_latent = rgb_latent if self.rgb_blending else noise.float()
        ctx = self.text_embed.repeat(rgb_latent.shape[0], 1, 1)
        x0 = None
        for t in steps:
            unet_input = pred_latent if self.rgb_blending else torch.cat([rgb_latent, pred_latent], dim=1)
            t_in = int(fix_timesteps) if fix_timesteps else t
            model_output = self.unet(unet_input, torch.tensor([t_in]), ctx)
            pred_latent, x0 = self.scheduler.step(model_output, t_in, pred_latent)
        z = self.vae.post_quant_conv(x0 / LATENT_SCALE)
        out = self.vae.decoder(z)
        if mode in ONE_CHANNEL_MODES:
            out = out.mean(dim=1, keepdim=True)
        return (torch.clip(out, -1.0, 1.0) + 1.0) / 2.0
