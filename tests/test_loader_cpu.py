"""Host-side weight assembly: the directory layouts and the LoRA merge of run.py:283-354."""
import os

import pytest
import torch
from safetensors.torch import save_file

from genpercept_b200 import loader as L


def _t(*shape, seed=0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def _sd21(root):
    os.makedirs(root / "unet"), os.makedirs(root / "vae")
    save_file({"conv_in.weight": _t(4, 4, 3, 3), "conv_out.weight": _t(4, 4, 3, 3, seed=1),
               "conv_norm_out.weight": _t(4, seed=2)}, str(root / "unet" / "diffusion_pytorch_model.safetensors"))
    save_file({"decoder.conv_in.weight": _t(4, 4, 3, 3, seed=3), "post_quant_conv.weight": _t(4, 4, 1, 1, seed=4),
               "encoder.mid_block.attentions.0.query.weight": _t(8, 8, 1, 1, seed=5)},
              str(root / "vae" / "diffusion_pytorch_model.safetensors"))


def test_merge_lora_matches_explicit_product_and_strips_adapter_keys():
    p = "down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q"
    base, a, b = _t(16, 12, seed=1), _t(4, 12, seed=2), _t(16, 4, seed=3)
    sd = {f"{p}.base_layer.weight": base, f"{p}.lora_A.default.weight": a, f"{p}.lora_B.default.weight": b,
          "conv_in.weight": _t(3, 3)}
    out = L.merge_lora(sd)
    assert set(out) == {f"{p}.weight", "conv_in.weight"}
    assert torch.allclose(out[f"{p}.weight"], base + b @ a, atol=1e-6)
    out2 = L.merge_lora(sd, lora_alpha=8)                       # alpha / r = 2
    assert torch.allclose(out2[f"{p}.weight"], base + 2.0 * (b @ a), atol=1e-6)
    with pytest.raises(ValueError):
        L.merge_lora({f"{p}.base_layer.weight": base, f"{p}.lora_A.default.weight": a,
                      f"{p}.lora_B.default.weight": _t(16, 5)})


def test_resolve_unet_checkpoint_layouts(tmp_path):
    sd21 = tmp_path / "sd21"
    _sd21(sd21)
    # training-output layout: <dir>/unet/diffusion_pytorch_model.bin wins over .safetensors (run.py:323-326)
    out = tmp_path / "ckpt"
    os.makedirs(out / "unet")
    torch.save({"conv_in.weight": _t(4, 4, 3, 3, seed=9)}, str(out / "unet" / "diffusion_pytorch_model.bin"))
    save_file({"conv_in.weight": _t(4, 4, 3, 3, seed=8)}, str(out / "unet" / "diffusion_pytorch_model.safetensors"))
    f, dec = L.resolve_unet_checkpoint(out, sd21)
    assert f.endswith("diffusion_pytorch_model.bin") and dec == str(out)
    # hub layout: no sub-folder, a head only for the *_dpt_head_v2 folder (run.py:287-292)
    hub = tmp_path / "genpercept-models" / "unet_depth_v2"
    os.makedirs(hub)
    save_file({"conv_in.weight": _t(4, 4, 3, 3, seed=7)}, str(hub / "diffusion_pytorch_model.safetensors"))
    f, dec = L.resolve_unet_checkpoint(hub, sd21)
    assert f == str(hub / "diffusion_pytorch_model.safetensors") and dec is None
    hub2 = tmp_path / "genpercept-models" / "unet_disparity_dpt_head_v2"
    os.makedirs(hub2)
    f, dec = L.resolve_unet_checkpoint(hub2, sd21)              # no file: falls back to the base UNet
    assert f == os.path.join(str(sd21), "unet", "diffusion_pytorch_model.safetensors") and dec == str(tmp_path / "genpercept-models")


def test_assemble_head_decoder_and_lora(tmp_path):
    sd21 = tmp_path / "sd21"
    _sd21(sd21)
    base = L.assemble(sd21)
    assert base["customized_head"] is None and "conv_out.weight" in base["unet"]
    assert "encoder.mid_block.attentions.0.to_q.weight" in base["vae"]          # legacy attention names remapped
    assert base["vae"]["encoder.mid_block.attentions.0.to_q.weight"].shape == (8, 8)
    # DPT readout: head next to the UNet, conv_out / conv_norm_out dropped
    ck = tmp_path / "ckpt_dpt"
    os.makedirs(ck / "unet"), os.makedirs(ck / "dpt_head_identity")
    save_file({"conv_in.weight": _t(4, 4, 3, 3, seed=11), "conv_out.weight": _t(4, 4, 3, 3), "conv_norm_out.bias": _t(4)},
              str(ck / "unet" / "diffusion_pytorch_model.safetensors"))
    save_file({"head.head.0.weight": _t(2, 2, 3, 3)}, str(ck / "dpt_head_identity" / "model.safetensors"))
    a = L.assemble(sd21, unet=ck)
    assert set(a["unet"]) == {"conv_in.weight"} and "head.head.0.weight" in a["customized_head"]
    # fine-tuned decoder + post_quant_conv on top of the base VAE
    ck2 = tmp_path / "ckpt_dec"
    os.makedirs(ck2 / "unet"), os.makedirs(ck2 / "vae_decoder"), os.makedirs(ck2 / "vae_post_quant_conv")
    save_file({"conv_in.weight": _t(4, 4, 3, 3, seed=12)}, str(ck2 / "unet" / "diffusion_pytorch_model.safetensors"))
    new_dec = _t(4, 4, 3, 3, seed=13)
    save_file({"conv_in.weight": new_dec}, str(ck2 / "vae_decoder" / "model.safetensors"))
    save_file({"weight": _t(4, 4, 1, 1, seed=14)}, str(ck2 / "vae_post_quant_conv" / "model.safetensors"))
    b = L.assemble(sd21, unet=ck2)
    assert torch.equal(b["vae"]["decoder.conv_in.weight"], new_dec)
    assert torch.equal(b["vae"]["post_quant_conv.weight"], _t(4, 4, 1, 1, seed=14))
    assert b["customized_head"] is None
    # LoRA checkpoint needs --lora_rank
    ck3 = tmp_path / "ckpt_lora"
    os.makedirs(ck3 / "unet")
    p = "mid_block.attentions.0.transformer_blocks.0.attn2.to_out.0"
    save_file({f"{p}.base_layer.weight": _t(6, 6), f"{p}.base_layer.bias": _t(6), f"{p}.lora_A.default.weight": _t(2, 6, seed=1),
               f"{p}.lora_B.default.weight": _t(6, 2, seed=2)}, str(ck3 / "unet" / "diffusion_pytorch_model.safetensors"))
    with pytest.raises(ValueError):
        L.assemble(sd21, unet=ck3)
    c = L.assemble(sd21, unet=ck3, lora_rank=2)
    assert set(c["unet"]) == {f"{p}.weight", f"{p}.bias"}
    assert torch.allclose(c["unet"][f"{p}.weight"], _t(6, 6) + _t(6, 2, seed=2) @ _t(2, 6, seed=1), atol=1e-6)
