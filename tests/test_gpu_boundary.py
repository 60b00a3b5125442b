"""GPU tests of the drop-in boundary beyond single_infer at the canonical sizes: per-call fix_timesteps, device-resident
encode_rgb / decode_pred, input sizes that are not multiples of 8 (VAE) / 64 (DPT), the bounded plan cache, and the
weight loader driving the engine from an on-disk layout."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 8e-3          # fp16-storage engine against the fp32 oracle (tests/test_gpu_e2e.py)


def _max(name, got, ref):
    e = float(np.abs(np.asarray(got, np.float64) - np.asarray(ref, np.float64)).max())
    print(f"{name}: max|err| = {e:.3e}")
    return e


@pytest.fixture(scope="module")
def pipe(synth_state, text_embed):
    from genpercept_b200.pipeline import GenPerceptPipeline
    p = GenPerceptPipeline(unet=synth_state["unet"], vae=synth_state["vae"], text_embed=text_embed, torch_dtype=torch.float16)
    yield p
    p._engine.close()


@pytest.fixture(scope="module")
def oracle(synth_state, text_embed):
    from oracle.pipeline import OraclePipeline
    return OraclePipeline(synth_state, text_embed)


def _rgb(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8)


def test_per_call_fix_timesteps(pipe, oracle):
    """genpercept_pipeline.py:405-408: fix_timesteps replaces the scheduler's [1] for one call; the engine re-folds the
    ResNet time-embedding biases (cached per timestep) and the next plain call is bit-identical to the first."""
    rgb = _rgb(2, 64, 64, 41)
    x = rgb.float() / 255.0 * 2.0 - 1.0
    base = pipe.single_infer(rgb.cuda(), mode="depth").cpu()
    for t in (5, 400):
        got = pipe.single_infer(rgb.cuda(), mode="depth", fix_timesteps=t).cpu().numpy()
        ref = oracle.single_infer(x, mode="depth", fix_timesteps=t).numpy()
        assert _max(f"fix_timesteps={t}", got, ref) < TOL
        assert np.abs(got - base.numpy()).max() > 1e-4          # the timestep matters
    again = pipe.single_infer(rgb.cuda(), mode="depth").cpu()
    assert torch.equal(again, base)


def test_encode_rgb_and_decode_pred_on_device(pipe, oracle):
    """encode_rgb (:488-505) and decode_pred (:507-526) as device-resident calls; decode_pred applies
    post_quant_conv(latent / 0.18215) by default like the reference and returns the map clipped to [-1, 1]."""
    rgb = _rgb(2, 64, 96, 42)
    x = rgb.float() / 255.0 * 2.0 - 1.0
    lat = pipe.encode_rgb(rgb.cuda())
    assert lat.is_cuda and tuple(lat.shape) == (2, 4, 8, 12)
    lat_ref = oracle.encode_rgb(x)
    assert _max("encode_rgb", lat.float().cpu().numpy(), lat_ref.numpy()) < 4e-3
    assert _max("encode_rgb (float input)", pipe.encode_rgb(x.cuda()).float().cpu().numpy(), lat_ref.numpy()) < 4e-3
    g = torch.Generator().manual_seed(43)
    foreign = torch.randn((2, 4, 8, 12), generator=g) * 0.5            # a latent the engine did not produce
    for mode, ch in (("depth", 1), ("normal", 3)):
        pipe.mode = mode
        got = pipe.decode_pred(foreign.cuda())
        ref = torch.clip(oracle.decode_pred(foreign, mode), -1.0, 1.0)
        assert got.is_cuda and tuple(got.shape) == (2, ch, 64, 96)
        assert _max(f"decode_pred {mode}", got.cpu().numpy(), ref.numpy()) < 2 * TOL       # [-1,1] is twice the [0,1] scale


@pytest.mark.parametrize("hw", [(100, 76), (231, 130)])
def test_vae_readout_sizes_that_are_not_multiples_of_8(pipe, oracle, hw):
    """AutoencoderKL's stride-2 stages floor (asymmetric padding): a 100x76 input decodes to 96x72, for the reference
    (the oracle runs its graph unchanged) and for the engine; __call__ then resizes back (match_input_res)."""
    H, W = hw
    rgb = _rgb(1, H, W, H * 7 + W)
    got = pipe.single_infer(rgb.cuda(), mode="depth").cpu().numpy()
    ref = oracle.single_infer(rgb.float() / 255.0 * 2.0 - 1.0, mode="depth").numpy()
    assert got.shape == ref.shape == (1, 1, H // 8 * 8, W // 8 * 8)
    assert _max(f"depth {H}x{W}", got, ref) < TOL


@pytest.mark.parametrize("hw", [(72, 88), (104, 200), (100, 60)])
def test_dpt_readout_sizes_that_are_not_multiples_of_64(synth_state, text_embed, hw):
    """dpt_head.py:297-300: fusion stages resize a skip feature (bilinear, align_corners=False) to the running map when the
    pyramid extents differ; the map comes out as the pyramid dictates (a multiple of 64 covering the input)."""
    from genpercept_b200.engine import Engine
    from oracle.pipeline import OraclePipeline
    H, W = hw
    rgb = _rgb(1, H, W, H * 11 + W)
    ref = OraclePipeline(synth_state, text_embed, use_dpt=True).single_infer(rgb.float() / 255.0 * 2.0 - 1.0).numpy()
    e = Engine(dtype=torch.float16, readout="dpt")
    try:
        e.load_state("unet", synth_state["unet"]); e.load_state("vae", synth_state["vae"]); e.load_state("dpt", synth_state["dpt"])
        e.set_text_embed(text_embed)
        e.finalize()
        got = e.infer(rgb.cuda()).cpu().numpy()
    finally:
        e.close()
    print(f"input {H}x{W} -> map {got.shape[-2]}x{got.shape[-1]}")
    assert got.shape == ref.shape
    assert _max(f"dpt {H}x{W}", got, ref) < 2e-2


def test_call_accepts_any_size_and_restores_the_input_resolution(pipe):
    """1242x375 (KITTI) becomes 768x231 at processing_res=768 — any size must run and come back at the input size."""
    from PIL import Image
    g = np.random.default_rng(7)
    img = Image.fromarray(g.integers(0, 256, (125, 414, 3), dtype=np.uint8))       # same aspect ratio, small
    out = pipe(img, processing_res=256, mode="depth", color_map=None)
    assert out.pred_np.shape == (125, 414) and out.pred_np.min() >= 0 and out.pred_np.max() <= 1


def test_plan_cache_is_bounded(pipe):
    """every distinct (B, H, W) used to keep its arena and graphs forever."""
    e = pipe._engine
    first = pipe.single_infer(_rgb(1, 64, 64, 1).cuda(), mode="depth").cpu()
    for k in range(8):
        pipe.single_infer(_rgb(1, 64, 64 + 8 * (k + 1), 2 + k).cuda(), mode="depth")
        assert e.plan_count() <= 4
    assert torch.equal(pipe.single_infer(_rgb(1, 64, 64, 1).cuda(), mode="depth").cpu(), first)     # rebuilt after eviction


def test_from_run_args_on_disk_layout_drives_the_engine(tmp_path, synth_state, text_embed, oracle):
    """SURVEY.md §8 f2 on the GPU: the training-output layout the reference's writer produces
    (src/trainer/genpercept_trainer.py:411-445; read by run.py:283-343) -> loader.assemble -> engine == oracle."""
    from safetensors.torch import save_file
    from genpercept_b200.pipeline import GenPerceptPipeline
    sd21 = tmp_path / "sd21"
    (sd21 / "vae").mkdir(parents=True)
    (sd21 / "unet").mkdir()
    save_file({k: v.half().contiguous() for k, v in synth_state["vae"].items()}, str(sd21 / "vae" / "diffusion_pytorch_model.fp16.safetensors"))
    stale = {k: torch.zeros_like(v) for k, v in synth_state["unet"].items()}          # the base UNet must be replaced by --unet
    save_file(stale, str(sd21 / "unet" / "diffusion_pytorch_model.safetensors"))
    ft = tmp_path / "ft" / "unet"
    ft.mkdir(parents=True)
    save_file({k: v.contiguous() for k, v in synth_state["unet"].items()}, str(ft / "diffusion_pytorch_model.safetensors"))
    p = GenPerceptPipeline.from_run_args(str(sd21), unet=str(tmp_path / "ft"), text_embed=text_embed, torch_dtype=torch.float16,
                                         variant="fp16")
    try:
        rgb = _rgb(1, 64, 64, 99)
        got = p.single_infer(rgb.cuda(), mode="depth").cpu().numpy()
    finally:
        p._engine.close()
    ref = oracle.single_infer(rgb.float() / 255.0 * 2.0 - 1.0, mode="depth").numpy()
    assert _max("from_run_args", got, ref) < 1e-2          # the VAE weights went through fp16 on disk
