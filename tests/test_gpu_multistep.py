"""Multi-step archs (run.py --archs marigold / rgb_blending) on the GPU against the oracle's restatement of
the reference's denoising loop (genpercept_pipeline.py:399-472), and the ensembling tail against golden vectors produced by
the reference's OWN ensemble_depth (tests/golden/make_golden_ensemble.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
SCHED = {"num_train_timesteps": 1000, "beta_start": 0.00085, "beta_end": 0.012, "beta_schedule": "scaled_linear",
         "clip_sample": False, "set_alpha_to_one": False, "steps_offset": 1, "prediction_type": "v_prediction",
         "timestep_spacing": "leading"}          # /root/reference/hf_configs/scheduler_beta_0.00085_0.012/scheduler_config.json


def _max(name, got, ref):
    e = float(np.abs(np.asarray(got, np.float64) - np.asarray(ref, np.float64)).max())
    print(f"{name}: max|err| = {e:.3e}")
    return e


@pytest.mark.parametrize("arch", ["marigold", "rgb_blending"])
def test_multistep_archs_match_the_oracle(arch, text_embed):
    from genpercept_b200 import weights as W
    from genpercept_b200.pipeline import GenPerceptPipeline
    from oracle.multistep import OracleMultiStep
    blending = arch == "rgb_blending"
    state = W.synth_state(4321, with_dpt=False, unet_in_channels=4 if blending else 8)
    pipe = GenPerceptPipeline(unet=state["unet"], vae=state["vae"], scheduler=dict(SCHED), text_embed=text_embed,
                              genpercept_pipeline=False, rgb_blending=blending, torch_dtype=torch.float16)
    try:
        g = torch.Generator().manual_seed(8)
        rgb = torch.randint(0, 256, (2, 3, 64, 64), generator=g, dtype=torch.uint8)
        gen = torch.Generator().manual_seed(77)                              # CPU generator: the same noise on both sides
        got = pipe.single_infer(rgb.cuda(), num_inference_steps=3, generator=gen, mode="depth").cpu().numpy()
        got_fix = pipe.single_infer(rgb.cuda(), num_inference_steps=2, generator=torch.Generator().manual_seed(77), mode="depth",
                                    fix_timesteps=400).cpu().numpy()
    finally:
        pipe._engine.close()
    orc = OracleMultiStep(state, text_embed, rgb_blending=blending, beta_start=SCHED["beta_start"], beta_end=SCHED["beta_end"])
    x = rgb.float() / 255.0 * 2.0 - 1.0
    noise = None if blending else torch.randn((2, 4, 8, 8), generator=torch.Generator().manual_seed(77))
    ref = orc.single_infer(x, 3, noise=noise, mode="depth").numpy()
    ref_fix = orc.single_infer(x, 2, noise=noise, mode="depth", fix_timesteps=400).numpy()
    assert got.shape == ref.shape == (2, 1, 64, 64)
    assert _max(f"{arch}, 3 DDIM steps", got, ref) < 1.5e-2               # three UNet passes of fp16-storage error
    assert _max(f"{arch}, 2 steps, fix_timesteps=400", got_fix, ref_fix) < 1.5e-2


def test_ensemble_depth_matches_the_reference_function(golden_dir):
    from genpercept_b200.ensemble import ensemble_depth
    g = np.load(os.path.join(golden_dir, "ensemble_ref.npz"))
    d = torch.from_numpy(g["depth"]).cuda()
    for tag, kw in (("default", {}), ("pipeline", {"max_res": 50}),
                    ("mean_scale_only", {"reduction": "mean", "shift_invariant": False, "max_res": 50})):
        out, unc = ensemble_depth(d.clone(), **kw)
        assert unc is None and out.is_cuda and tuple(out.shape) == (1, 1, 96, 128)
        assert _max(f"ensemble_depth[{tag}]", out.cpu().numpy(), g[tag]) < 2e-6
    with pytest.raises(ValueError):
        ensemble_depth(d, scale_invariant=False, shift_invariant=True)


def test_call_with_ensembling(text_embed):
    """__call__ of the marigold arch: the image repeated ensemble_size times, 2 DDIM steps each, ensemble_depth, resize back."""
    from PIL import Image
    from genpercept_b200 import weights as W
    from genpercept_b200.pipeline import GenPerceptPipeline
    state = W.synth_state(4321, with_dpt=False, unet_in_channels=8)
    pipe = GenPerceptPipeline(unet=state["unet"], vae=state["vae"], scheduler=dict(SCHED), text_embed=text_embed,
                              genpercept_pipeline=False, rgb_blending=False, torch_dtype=torch.float16)
    try:
        img = Image.fromarray(np.random.default_rng(3).integers(0, 256, (80, 120, 3), dtype=np.uint8))
        out = pipe(img, denoising_steps=2, ensemble_size=3, processing_res=64, batch_size=2, mode="depth",
                   generator=torch.Generator().manual_seed(5))
    finally:
        pipe._engine.close()
    assert out.pred_np.shape == (80, 120) and out.pred_np.min() >= 0 and out.pred_np.max() <= 1 and out.pred_np.std() > 1e-3
    assert out.pred_colored.size == (120, 80)
