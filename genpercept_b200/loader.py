"""Weight assembly for the on-disk layouts the reference reads, mirroring the model section of
/root/reference/run.py:273-376 (the same logic is repeated in infer.py:299-405):

* an SD-2.1 folder (``unet/``, ``vae/``, ``text_encoder/``, ``tokenizer/``) as the base;
* ``--unet <dir>``: a fine-tuned UNet, either ``<dir>/unet/diffusion_pytorch_model.{bin,safetensors}`` or — for the
  ``guangkaixu/genpercept-models`` layout, recognised by ``'genpercept-models' in path`` — directly in ``<dir>``;
  falls back to the base UNet with the reference's warning when neither file exists (run.py:323-329);
* next to it, optionally ``dpt_head_identity/model.safetensors`` (DPT readout; ``conv_out`` / ``conv_norm_out`` of the
  UNet are dropped, run.py:331-340) or ``vae_decoder/`` + ``vae_post_quant_conv/`` (fine-tuned decoder on top of the
  base VAE, run.py:306-310);
* ``--lora_rank r``: the checkpoint then holds peft adapter tensors (``….to_q.base_layer.weight``,
  ``….to_q.lora_A.default.weight`` [r, in], ``….to_q.lora_B.default.weight`` [out, r]) for to_q / to_k / to_v /
  to_out.0 (run.py:345-354, ``lora_alpha == r`` so the scale is 1): they are merged into plain weights here, because
  the engine folds and packs weights once at ``gp_finalize``.

Everything here is host-side dictionary work; the result feeds ``GenPerceptPipeline(unet=…, vae=…, customized_head=…)``.
"""
import logging
import os
import os.path as osp
import re

import torch

from . import weights as W

_LORA_A = re.compile(r"^(.*)\.lora_A\.([^.]+)\.weight$")


def load_file_any(path):
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)      # tensors only: no pickled code


def merge_lora(sd, scale=None, lora_alpha=None):
    """Folds peft LoRA tensors into their base weights: W = W_base + (alpha / r) * B @ A.

    `scale` overrides alpha / r; the reference uses lora_alpha == r (run.py:346-348) -> 1.  Keys without adapter
    tensors pass through; ``base_layer.`` is stripped so the result has plain diffusers names."""
    out = {}
    pairs = {}
    for k, v in sd.items():
        m = _LORA_A.match(k)
        if m:
            pairs[m.group(1)] = m.group(2)
    for k, v in sd.items():
        if ".lora_A." in k or ".lora_B." in k:
            continue
        out[k.replace(".base_layer.", ".")] = v
    for prefix, adapter in pairs.items():
        a = sd[f"{prefix}.lora_A.{adapter}.weight"].float()
        b = sd[f"{prefix}.lora_B.{adapter}.weight"].float()
        r = a.shape[0]
        if b.shape[1] != r:
            raise ValueError(f"{prefix}: lora_A {tuple(a.shape)} and lora_B {tuple(b.shape)} disagree on the rank")
        s = scale if scale is not None else ((lora_alpha / r) if lora_alpha is not None else 1.0)
        wk = f"{prefix}.weight"
        if wk not in out:
            raise KeyError(f"{prefix}: adapter tensors without a base weight")
        base = out[wk]
        delta = (b @ a) * s
        if tuple(delta.shape) != tuple(base.shape):
            raise ValueError(f"{prefix}: merged delta {tuple(delta.shape)} does not match the base weight {tuple(base.shape)}")
        out[wk] = (base.float() + delta).to(base.dtype)
    return out


def resolve_unet_checkpoint(unet_dir, checkpoint_path):
    """run.py:287-295,323-329 -> (file with the UNet weights, folder to search for a head / decoder or None)."""
    unet_dir = str(unet_dir)
    if "genpercept-models" in unet_dir:
        sub = ""
        decoder_dir = osp.dirname(unet_dir) if "unet_disparity_dpt_head_v2" in unet_dir else None
    else:
        sub = "unet"
        decoder_dir = unet_dir
    for name in ("diffusion_pytorch_model.bin", "diffusion_pytorch_model.safetensors"):      # .bin wins, as in run.py
        p = osp.join(unet_dir, sub, name)
        if osp.exists(p):
            return p, decoder_dir
    logging.warning("Warning!!! the saved checkpoint does not contain U-Net. Load U-Net from pretrained models...")
    return osp.join(str(checkpoint_path), "unet", "diffusion_pytorch_model.safetensors"), decoder_dir


def assemble(checkpoint_path, unet=None, lora_rank=0, variant=None):
    """-> dict(unet=state_dict, vae=state_dict, customized_head=state_dict | None) for GenPerceptPipeline.

    `checkpoint_path`: SD-2.1 folder; `unet`: the reference's ``--unet`` argument (None = base UNet)."""
    checkpoint_path = str(checkpoint_path)
    head = None
    vae = W.remap_legacy_vae_keys(_load_dir(osp.join(checkpoint_path, "vae"), variant))
    if unet is None:
        unet_sd = _load_dir(osp.join(checkpoint_path, "unet"), variant)
        return {"unet": unet_sd, "vae": vae, "customized_head": None}
    unet_file, decoder_dir = resolve_unet_checkpoint(unet, checkpoint_path)
    if decoder_dir:
        entries = os.listdir(decoder_dir)
        if "dpt_head_identity" in entries:
            head = load_file_any(osp.join(decoder_dir, "dpt_head_identity", "model.safetensors"))
        elif "dpt_head" in entries:
            raise NotImplementedError("DPTNeckHeadForUnetAfterUpsample (non-identity head, run.py:303-308) is not part "
                                      "of the accelerated path; only dpt_head_identity is")
        elif "vae_decoder" in entries and "vae_post_quant_conv" in entries:
            dec = load_file_any(osp.join(decoder_dir, "vae_decoder", "model.safetensors"))
            pq = load_file_any(osp.join(decoder_dir, "vae_post_quant_conv", "model.safetensors"))
            vae = dict(vae)
            vae.update(W.remap_legacy_vae_keys({f"decoder.{k}": v for k, v in dec.items()}))
            vae.update({f"post_quant_conv.{k}": v for k, v in pq.items()})
    unet_sd = load_file_any(unet_file)
    if head is not None:                                            # run.py:331-340
        unet_sd = {k: v for k, v in unet_sd.items() if "conv_out" not in k and "conv_norm_out" not in k}
    if lora_rank and lora_rank > 0:
        unet_sd = merge_lora(unet_sd)
    elif any(".lora_A." in k for k in unet_sd):
        raise ValueError("the checkpoint holds LoRA adapter tensors: pass lora_rank (run.py --lora_rank)")
    return {"unet": unet_sd, "vae": vae, "customized_head": head}


def _load_dir(path, variant=None):
    """`variant` ("fp16"): prefer diffusion_pytorch_model.<variant>.safetensors, like from_pretrained(variant=...) (run.py:374)."""
    names = ["diffusion_pytorch_model.safetensors", "model.safetensors", "diffusion_pytorch_model.bin"]
    if variant:
        names = [f"diffusion_pytorch_model.{variant}.safetensors", f"model.{variant}.safetensors",
                 f"diffusion_pytorch_model.{variant}.bin"] + names
    for n in names:
        p = osp.join(path, n)
        if osp.exists(p):
            return load_file_any(p)
    raise FileNotFoundError(f"no checkpoint file under {path}")
