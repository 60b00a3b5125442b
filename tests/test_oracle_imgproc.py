"""Pins oracle/imgproc.py (CPU restatement of the reference's host-side pre/post-processing) against
torchvision's own resize in this image; see the oracle's header for what stays unpinned."""
import numpy as np
import pytest
import torch
from torchvision.transforms import InterpolationMode
from torchvision.transforms.functional import resize

from oracle import imgproc as IP

CASES = [(480, 640, 576, 768), (540, 960, 216, 384), (333, 517, 247, 384), (100, 37, 384, 142), (64, 64, 64, 64),
         (96, 96, 48, 200)]


@pytest.mark.parametrize("mode", ["bilinear", "bicubic"])
@pytest.mark.parametrize("shape", CASES)
def test_resize_u8_matches_torchvision(shape, mode):
    h, w, oh, ow = shape
    g = torch.Generator().manual_seed(h * 7 + w)
    x = torch.randint(0, 256, (1, 3, h, w), generator=g, dtype=torch.uint8)
    tv = InterpolationMode.BILINEAR if mode == "bilinear" else InterpolationMode.BICUBIC
    ref = resize(x, [oh, ow], tv, antialias=True).numpy()
    got = IP.resize_aa(x.numpy(), oh, ow, mode)
    d = np.abs(ref.astype(np.int32) - got.astype(np.int32))
    assert d.max() <= 1                                   # rounding ties only
    assert (d > 0).mean() <= 5e-4


@pytest.mark.parametrize("mode", ["bilinear", "bicubic"])
def test_resize_f32_matches_torch_interpolate(mode):
    g = torch.Generator().manual_seed(3)
    x = torch.rand((2, 1, 192, 256), generator=g)
    for oh, ow in [(120, 160), (480, 640), (192, 100)]:
        ref = torch.nn.functional.interpolate(x, size=(oh, ow), mode=mode, align_corners=False, antialias=True).numpy()
        got = IP.resize_aa(x.numpy(), oh, ow, mode)
        assert np.abs(ref - got).max() < 5e-6             # values in [0,1]: float32 accumulation order


def test_resize_max_res_shape_and_identity():
    assert IP.resize_max_res_shape(480, 640, 768) == (576, 768)
    assert IP.resize_max_res_shape(1080, 1920, 768) == (432, 768)
    x = np.arange(64 * 64, dtype=np.uint8).reshape(1, 64, 64)
    assert np.array_equal(IP.resize_aa(x, 64, 64), x)


def test_colorize_and_quantize():
    lut = IP.spectral_lut_u8()
    assert lut.shape == (256, 3) and tuple(lut[0]) == (158, 1, 66) and tuple(lut[255]) == (94, 79, 162)
    p = np.array([[0.0, 0.5, 1.0], [0.99999, 1.5, -1.0]], np.float32)
    c = IP.colorize_u8(p)
    assert c.shape == (2, 3, 3)
    assert tuple(c[0, 0]) == tuple(lut[0]) and tuple(c[0, 1]) == tuple(lut[128]) and tuple(c[0, 2]) == tuple(lut[255])
    assert tuple(c[1, 1]) == tuple(lut[255]) and tuple(c[1, 2]) == tuple(lut[0])
    from genpercept_b200.image_util import colorize_depth_maps      # host mirror of image_util.py:25-63
    m = (colorize_depth_maps(p, 0, 1).squeeze() * 255).astype(np.uint8)
    assert np.array_equal(np.moveaxis(m, 0, -1), c)
    q = np.array([0.0, 0.5, 1.0, 0.9999], np.float32)
    assert IP.quantize(q, 8).tolist() == [0, 127, 255, 254]
    assert IP.quantize(q, 16).tolist() == [0, 32767, 65535, 65528]


def test_resize_matches_the_committed_torchvision_golden(golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "resize_torchvision.npz"))
    for mode in ("bilinear", "bicubic"):
        for oh, ow in ((27, 36), (96, 130)):
            d = np.abs(g[f"u8_{mode}_{oh}x{ow}"].astype(np.int32) - IP.resize_aa(g["x"], oh, ow, mode).astype(np.int32))
            assert d.max() <= 1 and (d > 0).mean() <= 1e-3
            assert np.abs(g[f"f32_{mode}_{oh}x{ow}"] - IP.resize_aa(g["f"], oh, ow, mode)).max() < 5e-6
