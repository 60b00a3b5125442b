"""GPU parity of the pre/post-processing through the C-ABI (gp_resize_aa / gp_colorize /
gp_quantize) against oracle/imgproc.py (pinned on torchvision in tests/test_oracle_imgproc.py) and against
torchvision itself; then the whole ``GenPerceptPipeline.__call__`` against the oracle pipeline."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [(480, 640, 576, 768), (540, 960, 216, 384), (333, 517, 247, 384), (100, 37, 384, 142), (64, 64, 64, 64),
         (96, 96, 48, 200), (768, 768, 768, 400)]


@pytest.mark.parametrize("mode", ["bilinear", "bicubic"])
@pytest.mark.parametrize("shape", CASES)
def test_resize_u8(shape, mode):
    from genpercept_b200 import engine as E
    from oracle import imgproc as IP
    h, w, oh, ow = shape
    g = torch.Generator().manual_seed(h * 7 + w)
    x = torch.randint(0, 256, (2, 3, h, w), generator=g, dtype=torch.uint8)
    ref = IP.resize_aa(x.numpy(), oh, ow, mode)
    got_dev = E.resize_aa(x.cuda(), oh, ow, mode)
    got_host = E.resize_aa(x, oh, ow, mode)                      # host buffers in and out
    assert got_dev.is_cuda and not got_host.is_cuda
    assert torch.equal(got_dev.cpu(), got_host)
    d = np.abs(ref.astype(np.int32) - got_host.numpy().astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() <= 1e-4               # same float32 arithmetic: ties only
    from torchvision.transforms import InterpolationMode
    from torchvision.transforms.functional import resize
    tv = resize(x, [oh, ow], InterpolationMode.BILINEAR if mode == "bilinear" else InterpolationMode.BICUBIC, antialias=True)
    d = np.abs(tv.numpy().astype(np.int32) - got_host.numpy().astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() <= 5e-4


@pytest.mark.parametrize("mode", ["bilinear", "bicubic"])
def test_resize_f32(mode):
    from genpercept_b200 import engine as E
    from oracle import imgproc as IP
    g = torch.Generator().manual_seed(5)
    x = torch.rand((3, 1, 192, 256), generator=g)
    for oh, ow in [(120, 160), (480, 640), (192, 100), (200, 256)]:
        ref = IP.resize_aa(x.numpy(), oh, ow, mode)
        got = E.resize_aa(x.cuda(), oh, ow, mode).cpu().numpy()
        assert np.abs(ref - got).max() < 2e-6
        tv = torch.nn.functional.interpolate(x, size=(oh, ow), mode=mode, align_corners=False, antialias=True).numpy()
        assert np.abs(tv - got).max() < 5e-6


def test_colorize_and_quantize_are_exact():
    from genpercept_b200 import engine as E
    from oracle import imgproc as IP
    g = torch.Generator().manual_seed(9)
    p = torch.rand((2, 70, 90), generator=g) * 1.2 - 0.1            # some values outside [0,1]
    p[0, 0, :4] = torch.tensor([0.0, 1.0, 0.5, 0.99999])
    lut = IP.spectral_lut_u8()
    got = E.colorize(p.cuda(), lut).numpy()
    assert np.array_equal(got, IP.colorize_u8(p.numpy(), 0.0, 1.0, lut))
    assert np.array_equal(E.colorize(p, lut).numpy(), got)          # host buffer in
    q = p.clamp(0, 1)
    for bits in (8, 16):
        assert np.array_equal(E.quantize(q.cuda(), bits), IP.quantize(q.numpy(), bits))
    with pytest.raises(RuntimeError):
        E.colorize(p.cuda(), lut, 1.0, 1.0)                         # vmax must exceed vmin


def test_call_matches_the_oracle_end_to_end(synth_state, text_embed):
    """PIL image -> resize_max_res -> single_infer -> resize back -> colour map: genpercept_pipeline.py:146-337."""
    from PIL import Image
    from genpercept_b200.pipeline import GenPerceptPipeline
    from oracle import imgproc as IP
    from oracle.pipeline import OraclePipeline
    g = torch.Generator().manual_seed(11)
    base = torch.rand((3, 25, 25), generator=g)
    img = torch.nn.functional.interpolate(base[None], size=(200, 200), mode="bicubic")[0].clamp(0, 1)
    pil = Image.fromarray((img.permute(1, 2, 0).numpy() * 255).astype(np.uint8))
    pipe = GenPerceptPipeline(unet=synth_state["unet"], vae=synth_state["vae"], text_embed=text_embed,
                              torch_dtype=torch.float16)
    out = pipe(pil, processing_res=128, match_input_res=True, mode="depth", color_map="Spectral")
    assert out.pred_np.shape == (200, 200) and out.pred_np.dtype == np.float32
    assert out.pred_colored.size == (200, 200)
    # oracle: the same steps on the CPU
    x = np.asarray(pil).transpose(2, 0, 1)[None]
    xr = IP.resize_aa(x, 128, 128)
    op = OraclePipeline(synth_state, text_embed)
    ref = op.single_infer(torch.from_numpy(xr).float() / 255.0 * 2.0 - 1.0, mode="depth").numpy()
    ref = IP.resize_aa(ref.astype(np.float32), 200, 200).clip(0, 1)[0, 0]
    err = np.abs(out.pred_np - ref)
    print(f"__call__: max|err|={err.max():.3e} mean={err.mean():.3e}")
    assert err.max() < 1e-2
    col = np.asarray(out.pred_colored)
    assert np.array_equal(col, IP.colorize_u8(out.pred_np))         # colour map of the engine's own map: exact
    ref_col = IP.colorize_u8(ref)
    assert (np.abs(col.astype(int) - ref_col.astype(int)).max(-1) > 8).mean() < 0.02
    # native resolution, no colour map, 3-channel mode
    out2 = pipe(pil.resize((128, 64)), processing_res=0, mode="normal", color_map=None)
    assert out2.pred_np.shape == (64, 128, 3) and out2.pred_colored.size == (128, 64)
