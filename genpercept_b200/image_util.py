"""Host-side image helpers mirroring /root/reference/genpercept/util/image_util.py
(resize_max_res :75, get_tv_resample_method :108, colorize_depth_maps :25, chw2hwc :66)."""
import numpy as np
import torch
from torchvision.transforms import InterpolationMode
from torchvision.transforms.functional import resize

# ColorBrewer "Spectral" (11 classes) — the anchors of matplotlib's 'Spectral' colormap; the
# reference calls matplotlib.colormaps['Spectral'] (image_util.py:44) which interpolates them
# linearly into a 256-entry LUT.  matplotlib is not a dependency here.
_SPECTRAL = np.array([(158, 1, 66), (213, 62, 79), (244, 109, 67), (253, 174, 97), (254, 224, 139),
                      (255, 255, 191), (230, 245, 152), (171, 221, 164), (102, 194, 165), (50, 136, 189),
                      (94, 79, 162)], dtype=np.float64) / 255.0


def _lut(cmap):
    try:
        import matplotlib
        cm = matplotlib.colormaps[cmap]
        return cm(np.linspace(0, 1, 256))[:, :3]
    except Exception:
        if cmap != "Spectral":
            raise ValueError(f"colormap {cmap!r} needs matplotlib; only 'Spectral' is built in")
        x = np.linspace(0, 1, 256)
        xp = np.linspace(0, 1, len(_SPECTRAL))
        return np.stack([np.interp(x, xp, _SPECTRAL[:, c]) for c in range(3)], axis=1)


def colorize_depth_maps(depth_map, min_depth, max_depth, cmap="Spectral", valid_mask=None):
    assert len(depth_map.shape) >= 2, "Invalid dimension"
    if isinstance(depth_map, torch.Tensor):
        depth = depth_map.detach().squeeze().cpu().numpy()
    else:
        depth = np.asarray(depth_map).copy().squeeze()
    if depth.ndim < 3:
        depth = depth[np.newaxis, :, :]
    depth = ((depth - min_depth) / (max_depth - min_depth)).clip(0, 1)
    lut = _lut(cmap)
    idx = (depth * 256).astype(np.int64).clip(0, 255)       # matplotlib: floor(x*N), x==1 -> N-1
    img = np.rollaxis(lut[idx], 3, 1)                        # [B,3,H,W], values 0..1
    if valid_mask is not None:
        vm = np.asarray(valid_mask).squeeze()
        vm = vm[np.newaxis, np.newaxis] if vm.ndim < 3 else vm[:, np.newaxis]
        img[~np.repeat(vm, 3, axis=1)] = 0
    return torch.from_numpy(img).float() if isinstance(depth_map, torch.Tensor) else img


def chw2hwc(chw):
    assert 3 == len(chw.shape)
    if isinstance(chw, torch.Tensor):
        return torch.permute(chw, (1, 2, 0))
    return np.moveaxis(chw, 0, -1)


def resize_max_res(img, max_edge_resolution, resample_method=InterpolationMode.BILINEAR):
    assert 4 == img.dim(), f"Invalid input shape {img.shape}"
    h, w = img.shape[-2:]
    f = min(max_edge_resolution / w, max_edge_resolution / h)
    return resize(img, (int(h * f), int(w * f)), resample_method, antialias=True)


def get_tv_resample_method(method_str):
    d = {"bilinear": InterpolationMode.BILINEAR, "bicubic": InterpolationMode.BICUBIC,
         "nearest": InterpolationMode.NEAREST_EXACT, "nearest-exact": InterpolationMode.NEAREST_EXACT}
    m = d.get(method_str, None)
    if m is None:
        raise ValueError(f"Unknown resampling method: {m}")
    return m
