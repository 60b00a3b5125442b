"""Pins the oracle's orchestration (oracle/pipeline.py) against the REFERENCE'S OWN ``GenPerceptPipeline.single_infer`` /
``encode_rgb`` / ``decode_pred`` (/root/reference/genpercept/genpercept_pipeline.py:375-526), executed here:

the reference module is imported through a minimal ``diffusers`` / ``matplotlib`` shim (base classes and type names only:
diffusers itself is not installable here) and instantiated with the oracle's VAE / UNet modules and one-step scheduler
behind thin adapters, plus the reference's own DPT head class.  Everything between those modules — latent scaling, the
mean half of the moments, the scheduler call and ``pred_original_sample``, ``/ scale`` + post_quant_conv + decoder, the
channel mean, clip and shift, the DPT feature order and min-max — is then the reference's code, not a restatement.
(What stays unpinned is the inside of the diffusers blocks; see oracle/__init__.py.)  Build container only."""
import importlib
import os
import sys
import types

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (build container only)")


def _install_shims():
    import torch.nn as nn
    if "diffusers" not in sys.modules or not hasattr(sys.modules["diffusers"], "DiffusionPipeline"):
        d = sys.modules.get("diffusers") or types.ModuleType("diffusers")

        class DiffusionPipeline:
            def __init__(self):
                self._cfg = {}

            def register_modules(self, **kw):
                for k, v in kw.items():
                    setattr(self, k, v)

            def register_to_config(self, **kw):
                self._cfg.update(kw)

            @property
            def device(self):
                return torch.device("cpu")

            @property
            def dtype(self):
                return torch.float32

        for name in ("AutoencoderKL", "DDIMScheduler", "LCMScheduler", "UNet2DConditionModel"):
            setattr(d, name, type(name, (), {}))
        d.DiffusionPipeline = DiffusionPipeline
        du = sys.modules.get("diffusers.utils") or types.ModuleType("diffusers.utils")
        du.BaseOutput = type("BaseOutput", (), {})
        du.USE_PEFT_BACKEND = True
        dm = sys.modules.get("diffusers.models") or types.ModuleType("diffusers.models")
        dl = sys.modules.get("diffusers.models.lora") or types.ModuleType("diffusers.models.lora")
        dl.LoRACompatibleConv = nn.Conv2d
        d.utils, d.models, dm.lora = du, dm, dl
        sys.modules.update({"diffusers": d, "diffusers.utils": du, "diffusers.models": dm, "diffusers.models.lora": dl})
    if "matplotlib" not in sys.modules:
        m = types.ModuleType("matplotlib")
        mp = types.ModuleType("matplotlib.pyplot")
        m.pyplot = mp
        sys.modules.update({"matplotlib": m, "matplotlib.pyplot": mp})
    if REF not in sys.path:
        sys.path.insert(0, REF)


class _UNetAdapter:
    def __init__(self, unet):
        self.unet = unet

    def __call__(self, x, t, encoder_hidden_states=None, return_feature=False):
        t = torch.as_tensor(t).reshape(-1)[:1]
        out = self.unet(x, t, encoder_hidden_states, return_feature=return_feature)
        return types.SimpleNamespace(multi_level_feats=out) if return_feature else types.SimpleNamespace(sample=out)


class _SchedulerAdapter:
    beta_start = 1
    beta_end = 1

    def __init__(self, s):
        self.s = s

    def set_timesteps(self, n, device=None):
        self.timesteps = self.s.set_timesteps(n)

    def step(self, model_output, t, sample, generator=None):
        prev, x0 = self.s.step(model_output, int(t), sample)
        return types.SimpleNamespace(prev_sample=prev, pred_original_sample=x0)


@pytest.fixture(scope="module")
def ref_pipeline_cls():
    _install_shims()
    return importlib.import_module("genpercept.genpercept_pipeline")


def test_single_infer_glue_matches_the_reference(ref_pipeline_cls, synth_state, text_embed):
    from oracle.pipeline import OraclePipeline
    mod = ref_pipeline_cls
    g = torch.Generator().manual_seed(21)
    rgb = torch.rand((1, 3, 64, 64), generator=g) * 2 - 1
    # VAE readout, 1- and 3-channel modes
    op = OraclePipeline(synth_state, text_embed)
    rp = mod.GenPerceptPipeline(unet=_UNetAdapter(op.unet), vae=op.vae, scheduler=_SchedulerAdapter(op.scheduler),
                                text_encoder=None, tokenizer=None, genpercept_pipeline=True)
    rp.text_embed = op.text_embed
    for mode in ("depth", "normal"):
        rp.mode = mode
        with torch.no_grad():
            ref = rp.single_infer(rgb, 1, None, False)
        mine = op.single_infer(rgb, mode=mode)
        assert ref.shape == mine.shape and torch.allclose(ref, mine, atol=1e-6, rtol=0), float((ref - mine).abs().max())
    with torch.no_grad():
        assert torch.allclose(rp.encode_rgb(rgb), op.encode_rgb(rgb), atol=1e-7)
    # --fix_timesteps: the reference feeds that timestep to the UNet instead of the scheduler's
    with torch.no_grad():
        ref = rp.single_infer(rgb, 1, None, False, fix_timesteps=7)
    assert torch.allclose(ref, op.single_infer(rgb, mode="normal", fix_timesteps=7), atol=1e-6, rtol=0)
    # DPT readout with the reference's own head class (isinstance check at genpercept_pipeline.py:475)
    od = OraclePipeline(synth_state, text_embed, use_dpt=True)
    head_mod = sys.modules["genpercept.models.dpt_head"]
    # transformers >= 4.4x refuses ModelOutput subclasses that are not dataclasses; the reference's output container
    # (dpt_head.py:24-49, written for an older transformers) is replaced by a plain attribute bag — no arithmetic involved
    head_mod.DepthEstimatorOutput = lambda **kw: types.SimpleNamespace(**kw)
    from transformers import DPTConfig
    head = head_mod.DPTNeckHeadForUnetAfterUpsampleIdentity(
        DPTConfig.from_pretrained(f"{REF}/hf_configs/dpt-sd2.1-unet-after-upsample-general")).eval()
    head.load_state_dict(synth_state["dpt"], strict=True)
    rd = mod.GenPerceptPipeline(unet=_UNetAdapter(od.unet), vae=od.vae, scheduler=_SchedulerAdapter(od.scheduler),
                                text_encoder=None, tokenizer=None, customized_head=head, genpercept_pipeline=True)
    rd.text_embed = od.text_embed
    rd.mode = "depth"
    with torch.no_grad():
        ref = rd.single_infer(rgb, 1, None, False)
    mine = od.single_infer(rgb, mode="depth")
    assert ref.shape == mine.shape and torch.allclose(ref, mine, atol=2e-6, rtol=0), float((ref - mine).abs().max())
