"""Independent cross-check of the oracle's AutoencoderKL ENCODER (oracle/vae.py, restated from SURVEY.md App. A.3 because
diffusers is not installable here): HF transformers ships the original LDM / taming-transformers encoder — the code
diffusers' ``Encoder`` / ``DownEncoderBlock2D`` / ``UNetMidBlock2D`` descends from — as ``ChameleonVQVAEEncoder``.
Configured like the SD VAE (128 base channels, multipliers (1,2,4,4), 2 resnets per level, mid attention, double
latent) and loaded with the oracle's weights under the LDM names it must reproduce the oracle's encoder + nothing else:
GroupNorm(32, eps 1e-6)-swish-conv resnets with 1x1 shortcuts, the (0,1,0,1)-padded stride-2 downsamplers, the 1-head
d=512 mid attention with its 1/sqrt(C) scale, norm_out-swish-conv_out.  (Not the reference — the reference has no
source for these blocks — but an implementation the oracle was not written from.)"""
import pytest
import torch

cham = pytest.importorskip("transformers.models.chameleon.modeling_chameleon")


def _ldm_key(k):
    """diffusers AutoencoderKL.encoder key -> LDM / taming key."""
    k = k.replace("conv_norm_out.", "norm_out.")
    k = k.replace("mid_block.resnets.0.", "mid.block_1.").replace("mid_block.resnets.1.", "mid.block_2.")
    k = k.replace("mid_block.attentions.0.group_norm.", "mid.attn_1.norm.")
    for a, b in (("to_q", "q"), ("to_k", "k"), ("to_v", "v"), ("to_out.0", "proj_out")):
        k = k.replace(f"mid_block.attentions.0.{a}.", f"mid.attn_1.{b}.")
    if k.startswith("down_blocks."):
        p = k.split(".")
        lvl = p[1]
        if p[2] == "resnets":
            k = f"down.{lvl}.block.{p[3]}." + ".".join(p[4:])
        else:                                           # downsamplers.0.conv.*
            k = f"down.{lvl}.downsample." + ".".join(p[4:])
    return k.replace("conv_shortcut.", "nin_shortcut.")


def test_vae_encoder_matches_the_ldm_encoder(synth_state):
    from transformers.models.chameleon.configuration_chameleon import ChameleonVQVAEConfig
    from oracle.vae import AutoencoderKL
    vae = AutoencoderKL().eval()
    vae.load_state_dict(synth_state["vae"], strict=True)
    cfg = ChameleonVQVAEConfig(double_latent=True, latent_channels=4, resolution=64, in_channels=3, base_channels=128,
                               channel_multiplier=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=None, dropout=0.0,
                               attn_type="vanilla")
    ldm = cham.ChameleonVQVAEEncoder(cfg).eval()
    sd = {}
    for k, v in vae.encoder.state_dict().items():
        nk = _ldm_key(k)
        if ".attn_1." in nk and nk.endswith(".weight") and v.dim() == 2:
            v = v[:, :, None, None]                     # linear -> the 1x1 conv the LDM code uses
        sd[nk] = v
    ldm.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(12)
    x = torch.rand((2, 3, 64, 72), generator=g) * 2 - 1
    with torch.no_grad():
        ref = ldm(x.clone())
        mine = vae.encoder(x)
    assert ref.shape == mine.shape == (2, 8, 8, 9)
    err = (ref - mine).abs().max().item()
    print(f"oracle VAE encoder vs LDM encoder: max|delta| = {err:.3e} (max|ref| = {ref.abs().max().item():.3f})")
    assert err < 2e-5 * max(1.0, ref.abs().max().item())


def _ldm_dec_key(k):
    """diffusers AutoencoderKL.decoder key -> LDM / taming decoder key."""
    k = k.replace("conv_norm_out.", "norm_out.")
    k = k.replace("mid_block.resnets.0.", "mid.block_1.").replace("mid_block.resnets.1.", "mid.block_2.")
    k = k.replace("mid_block.attentions.0.group_norm.", "mid.attn_1.norm.")
    for a, b in (("to_q", "q"), ("to_k", "k"), ("to_v", "v"), ("to_out.0", "proj_out")):
        k = k.replace(f"mid_block.attentions.0.{a}.", f"mid.attn_1.{b}.")
    if k.startswith("up_blocks."):
        p = k.split(".")
        lvl = p[1]
        if p[2] == "resnets":
            k = f"up.{lvl}.block.{p[3]}." + ".".join(p[4:])
        else:                                           # upsamplers.0.conv.*
            k = f"up.{lvl}.upsample." + ".".join(p[4:])
    return k.replace("conv_shortcut.", "nin_shortcut.")


def test_vae_decoder_matches_the_ldm_decoder(synth_state):
    """Same cross-check for the DECODER: ``JanusVQVAEDecoder`` is the LDM decoder (conv_in, mid resnet-attn-resnet, per
    level three resnets + nearest-2x-then-conv upsampling, norm_out-swish-conv_out); its extra attention blocks in the
    deepest level, which the SD VAE does not have, are removed before the comparison."""
    janus = pytest.importorskip("transformers.models.janus.modeling_janus")
    import torch.nn as nn
    from transformers.models.janus.configuration_janus import JanusVQVAEConfig
    from oracle.vae import AutoencoderKL
    vae = AutoencoderKL().eval()
    vae.load_state_dict(synth_state["vae"], strict=True)
    cfg = JanusVQVAEConfig(latent_channels=4, in_channels=3, out_channels=3, base_channels=128, channel_multiplier=[1, 2, 4, 4],
                           num_res_blocks=2, dropout=0.0)
    ldm = janus.JanusVQVAEDecoder(cfg).eval()
    ldm.up[0].attn = nn.ModuleList()
    sd = {}
    for k, v in vae.decoder.state_dict().items():
        nk = _ldm_dec_key(k)
        if ".attn_1." in nk and nk.endswith(".weight") and v.dim() == 2:
            v = v[:, :, None, None]
        sd[nk] = v
    ldm.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(13)
    z = torch.randn((2, 4, 8, 9), generator=g)
    with torch.no_grad():
        ref = ldm(z.clone())
        mine = vae.decoder(z)
    assert ref.shape == mine.shape == (2, 3, 64, 72)
    err = (ref - mine).abs().max().item()
    print(f"oracle VAE decoder vs LDM decoder: max|delta| = {err:.3e} (max|ref| = {ref.abs().max().item():.3f})")
    assert err < 2e-5 * max(1.0, ref.abs().max().item())
