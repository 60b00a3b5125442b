"""Golden vectors for genpercept_b200.ensemble.ensemble_depth from the REFERENCE's own function
(/root/reference/genpercept/util/ensemble.py:43-205), run here on CPU tensors.  Run in the build container (the
reference checkout does not exist on the GPU box):   python tests/golden/make_golden_ensemble.py
Writes tests/golden/ensemble_ref.npz (inputs + the reference's outputs for three argument sets)."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/genpercept/util"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load_reference_ensemble():
    pkg = types.ModuleType("refutil")
    pkg.__path__ = [REF]
    sys.modules["refutil"] = pkg
    if "matplotlib" not in sys.modules:                      # image_util imports it for the colour map only
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            sys.modules["matplotlib"] = types.ModuleType("matplotlib")
    for name in ("image_util", "ensemble"):
        spec = importlib.util.spec_from_file_location(f"refutil.{name}", os.path.join(REF, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"refutil.{name}"] = mod
        spec.loader.exec_module(mod)
    return sys.modules["refutil.ensemble"].ensemble_depth


def main():
    ens = _load_reference_ensemble()
    g = torch.Generator().manual_seed(20)
    base = torch.rand((1, 1, 96, 128), generator=g)
    base = torch.nn.functional.avg_pool2d(base, 9, stride=1, padding=4)
    members = []
    for k in range(5):                                       # affine-distorted noisy copies of one map
        s, t = 0.5 + torch.rand(1, generator=g).item(), 0.3 * torch.rand(1, generator=g).item()
        members.append(base * s + t + 0.02 * torch.randn(base.shape, generator=g))
    depth = torch.cat(members, dim=0).clamp(min=0).float()
    out = {"depth": depth.numpy()}
    for tag, kw in (("default", {}), ("pipeline", {"max_res": 50}), ("mean_scale_only", {"reduction": "mean", "shift_invariant": False, "max_res": 50})):
        pred, _ = ens(depth.clone(), scale_invariant=True, **({"shift_invariant": True} | kw))
        out[tag] = pred.numpy()
        print(tag, pred.shape, float(pred.min()), float(pred.max()))
    np.savez_compressed(os.path.join(HERE, "ensemble_ref.npz"), **out)


if __name__ == "__main__":
    main()
