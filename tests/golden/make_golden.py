"""Regenerates the committed fixtures under tests/golden/ (run in the BUILD container only:
it reads /root/reference, which does not exist on the GPU box).

  python tests/golden/make_golden.py

1. dpt_ref_h8.npz    — the reference's own ``DPTNeckHeadForUnetAfterUpsampleIdentity``
                       (/root/reference/genpercept/models/dpt_head.py:585) imported through a
                       2-symbol ``diffusers`` shim, loaded with the seeded synthetic DPT weights
                       (genpercept_b200.weights.synth_state(1234)['dpt']), run on seeded features.
2. empty_text_embed_2x1024.npy — rows 0-1 of /root/reference/GenPercept_v1/empty_text_embed.npy
                       (= the v2 pipeline's do_not_pad "" embedding).
3. oracle_e2e_64.npz — oracle fp32 single_infer outputs (VAE readout depth+normal, DPT readout) on
                       seeded 64x64 inputs: guards the oracle against silent edits.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def load_reference_dpt():
    import torch.nn as nn
    d = types.ModuleType("diffusers"); dm = types.ModuleType("diffusers.models")
    dl = types.ModuleType("diffusers.models.lora"); du = types.ModuleType("diffusers.utils")
    dl.LoRACompatibleConv = nn.Conv2d
    du.USE_PEFT_BACKEND = True
    d.models = dm; dm.lora = dl; d.utils = du
    for n, m in (("diffusers", d), ("diffusers.models", dm), ("diffusers.models.lora", dl),
                 ("diffusers.utils", du)):
        sys.modules.setdefault(n, m)
    spec = importlib.util.spec_from_file_location("ref_dpt_head", f"{REF}/genpercept/models/dpt_head.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_dpt_head"] = mod
    spec.loader.exec_module(mod)
    from transformers import DPTConfig
    cfg = DPTConfig.from_pretrained(f"{REF}/hf_configs/dpt-sd2.1-unet-after-upsample-general")
    return mod.DPTNeckHeadForUnetAfterUpsampleIdentity(cfg).eval()


def dpt_features(h, seed=7, batch=1):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn((batch, 320, h, h), generator=g), torch.randn((batch, 640, h, h), generator=g),
            torch.randn((batch, 1280, h // 2, h // 2), generator=g),
            torch.randn((batch, 1280, h // 4, h // 4), generator=g)]


def main():
    from genpercept_b200 import weights as W
    state = W.synth_state(1234)
    # 1. DPT reference
    ref = load_reference_dpt()
    ref.load_state_dict(state["dpt"], strict=True)
    feats = dpt_features(8)
    with torch.no_grad():
        out = ref(hidden_states=[f.clone() for f in feats], return_depth_only=True)
    np.savez_compressed(os.path.join(HERE, "dpt_ref_h8.npz"), out=out.numpy())
    print("dpt ref", out.shape, float(out.mean()), float(out.std()))
    # 2. text embed
    e = np.load(f"{REF}/GenPercept_v1/empty_text_embed.npy")
    assert e.shape == (77, 1024) and e.dtype == np.float16
    np.save(os.path.join(HERE, "empty_text_embed_2x1024.npy"), e[:2])
    # 3. oracle e2e
    from oracle.pipeline import OraclePipeline
    te = torch.from_numpy(e[:2].astype(np.float32))[None]
    g = torch.Generator().manual_seed(1001)
    rgb = torch.randint(0, 256, (2, 3, 64, 64), generator=g, dtype=torch.uint8)
    x = rgb.float() / 255.0 * 2.0 - 1.0
    res = {"rgb": rgb.numpy()}
    p = OraclePipeline(state, te, use_dpt=False)
    y, inter = p.single_infer(x, mode="depth", return_intermediates=True)
    res["depth"] = y.numpy(); res["rgb_latent"] = inter["rgb_latent"].numpy()
    res["unet_out"] = inter["unet_out"].numpy()
    res["normal"] = p.single_infer(x, mode="normal").numpy()
    p = OraclePipeline(state, te, use_dpt=True)
    res["dpt"] = p.single_infer(x, mode="depth").numpy()
    np.savez_compressed(os.path.join(HERE, "oracle_e2e_64.npz"), **res)
    for k, v in res.items():
        print(k, v.shape, float(v.mean()), float(v.std()))


if __name__ == "__main__":
    main()
