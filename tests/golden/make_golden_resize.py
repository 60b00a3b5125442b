"""Golden vectors for the pre/post-processing oracle: torchvision's own `resize(tensor, antialias=True)` — the call
the reference makes (genpercept/util/image_util.py:104, genpercept_pipeline.py:303) — on a small seeded image.
Generated in the build container (torchvision 0.26.0); run from the repo root:
    python tests/golden/make_golden_resize.py
"""
import os

import numpy as np
import torch
import torchvision
from torchvision.transforms import InterpolationMode
from torchvision.transforms.functional import resize

HERE = os.path.dirname(os.path.abspath(__file__))
g = torch.Generator().manual_seed(77)
x = torch.randint(0, 256, (1, 3, 45, 61), generator=g, dtype=torch.uint8)
f = torch.rand((1, 1, 45, 61), generator=g)
res = {"x": x.numpy(), "f": f.numpy(), "torchvision": np.array(torchvision.__version__)}
for name, tv in (("bilinear", InterpolationMode.BILINEAR), ("bicubic", InterpolationMode.BICUBIC)):
    for oh, ow in ((27, 36), (96, 130)):
        res[f"u8_{name}_{oh}x{ow}"] = resize(x, [oh, ow], tv, antialias=True).numpy()
        res[f"f32_{name}_{oh}x{ow}"] = resize(f, [oh, ow], tv, antialias=True).numpy()
np.savez_compressed(os.path.join(HERE, "resize_torchvision.npz"), **res)
print({k: v.shape for k, v in res.items()})
