"""CPU oracle for the pre/post-processing around the hot path.  TEST INFRASTRUCTURE ONLY.

numpy restatement of what the reference runs on the host around ``single_infer``:

* ``resize_max_res`` (/root/reference/genpercept/util/image_util.py:75-105) and the resize back to the
  input resolution (/root/reference/genpercept/genpercept_pipeline.py:301-307).  Both call
  ``torchvision.transforms.functional.resize(tensor, size, interpolation, antialias=True)``, whose tensor
  path (torchvision ``_functional_tensor.resize``) casts uint8 to float32, runs
  ``torch.nn.functional.interpolate(mode='bilinear'|'bicubic', align_corners=False, antialias=True)`` and,
  for integer inputs, ``torch.round`` (half to even) + cast back.  The separable anti-aliased filter is
  restated from ATen ``UpSampleKernel.cpp`` (``HelperInterpBase::_compute_indices_min_size_weights_aa``):
  per output index a window ``[xmin, xmin + xsize)`` of triangle (bilinear) or Keys a=-0.5 (bicubic)
  weights, stretched by the scale when down-sampling, normalised to sum 1; width pass first, then height.
* ``colorize_depth_maps`` (image_util.py:25-63): ``matplotlib.colormaps[cmap](x)`` = LUT[floor(x * 256)]
  (x == 1 -> 255), then ``(c * 255).astype(uint8)`` at the call site (genpercept_pipeline.py:318-321).
* the uint8 / uint16 quantisation of the saved prediction (/root/reference/run.py:449-455).

PINNED (tests/test_oracle_imgproc.py): resize against torchvision's own ``resize`` in this image (float
results to 1e-4 on the 0..255 scale — torch's vectorised CPU kernel accumulates in a different order —
and the rounded uint8 results equal except for at most 5e-4 of the pixels off by one level at rounding
ties).  UNPINNED: the Spectral LUT — matplotlib is not installed here; the LUT is restated as the linear
interpolation of the 11 ColorBrewer "Spectral" anchors over 256 entries, which is how matplotlib builds
``colormaps['Spectral']`` (LinearSegmentedColormap.from_list, N=256).
"""
import numpy as np

SPECTRAL_ANCHORS = np.array([(158, 1, 66), (213, 62, 79), (244, 109, 67), (253, 174, 97), (254, 224, 139),
                             (255, 255, 191), (230, 245, 152), (171, 221, 164), (102, 194, 165), (50, 136, 189),
                             (94, 79, 162)], dtype=np.float64) / 255.0


def _filter(x, mode):
    x = np.abs(x)
    if mode == "bilinear":
        return np.where(x < 1.0, 1.0 - x, 0.0)
    a = -0.5                                            # bicubic (Keys), ATen uses a = -0.5 for the aa kernels
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0,
                    np.where(x < 2.0, ((a * x - 5.0 * a) * x + 8.0 * a) * x - 4.0 * a, 0.0))


def aa_weights(in_size, out_size, mode="bilinear"):
    """-> xmin[out], xsize[out], W[out, kmax] (float32), exactly the float32 arithmetic of ATen."""
    f = np.float32
    interp = 2 if mode == "bilinear" else 4
    scale = f(in_size) / f(out_size)
    support = f(interp * 0.5) * scale if scale >= 1.0 else f(interp * 0.5)
    invscale = f(1.0) / scale if scale >= 1.0 else f(1.0)
    kmax = int(np.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    xsize = np.zeros(out_size, np.int32)
    W = np.zeros((out_size, kmax), f)
    for i in range(out_size):
        center = scale * f(i + 0.5)
        lo = max(int(center - support + f(0.5)), 0)
        n = min(int(center + support + f(0.5)), in_size) - lo
        n = min(max(n, 0), kmax)
        x = ((np.arange(n, dtype=f) + f(lo) - center + f(0.5)) * invscale).astype(f)
        w = _filter(x, mode).astype(f)
        tot = w.sum(dtype=f)
        if tot != 0:
            w = (w * (f(1.0) / tot)).astype(f)
        xmin[i], xsize[i] = lo, n
        W[i, :n] = w
    return xmin, xsize, W


def _pass(img, axis, out_size, mode):
    in_size = img.shape[axis]
    if in_size == out_size:
        return img
    xmin, xsize, W = aa_weights(in_size, out_size, mode)
    src = np.moveaxis(img, axis, -1)
    out = np.zeros(src.shape[:-1] + (out_size,), np.float32)
    for i in range(out_size):
        acc = (src[..., xmin[i]] * W[i, 0]).astype(np.float32)
        for j in range(1, xsize[i]):                   # fused multiply-add, as the AVX2 build of ATen contracts it
            acc = (acc.astype(np.float64) + src[..., xmin[i] + j].astype(np.float64) * np.float64(W[i, j])).astype(np.float32)
        out[..., i] = acc
    return np.moveaxis(out, -1, axis)


def resize_aa(img, out_h, out_w, mode="bilinear"):
    """img [..., H, W] uint8 or float32 -> same dtype; width pass, then height pass, float32 in between."""
    a = np.asarray(img)
    x = a.astype(np.float32)
    y = _pass(_pass(x, -1, out_w, mode), -2, out_h, mode)
    if a.dtype == np.uint8:
        if mode == "bicubic":
            y = y.clip(0, 255)
        return np.rint(y).astype(np.uint8)             # torch.round: half to even
    return y


def resize_max_res_shape(h, w, max_edge):
    """image_util.py:98-102."""
    f = min(max_edge / w, max_edge / h)
    return int(h * f), int(w * f)


def spectral_lut_u8():
    x = np.linspace(0, 1, 256)
    xp = np.linspace(0, 1, len(SPECTRAL_ANCHORS))
    lut = np.stack([np.interp(x, xp, SPECTRAL_ANCHORS[:, c]) for c in range(3)], axis=1)
    return (lut * 255).astype(np.uint8)


def colorize_u8(pred, vmin=0.0, vmax=1.0, lut=None):
    """pred [..., H, W] float32 -> uint8 [..., H, W, 3]."""
    lut = spectral_lut_u8() if lut is None else lut
    d = ((np.asarray(pred, np.float32) - np.float32(vmin)) / np.float32(vmax - vmin)).clip(0, 1)
    idx = (d * np.float32(256)).astype(np.int64).clip(0, 255)
    return lut[idx]


def quantize(pred, bits):
    """run.py:449-455: (pred * 65535).astype(uint16) / (pred * 255).astype(uint8) on a float32 map in [0,1]."""
    p = np.asarray(pred, np.float32)
    return (p * np.float32(65535.0)).astype(np.uint16) if bits == 16 else (p * np.float32(255.0)).astype(np.uint8)
