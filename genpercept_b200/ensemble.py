"""Test-time ensembling of the multi-step archs: ``ensemble_depth`` of
/root/reference/genpercept/util/ensemble.py:43-205 with the same arguments, defaults and errors.

Split like the rest of the path: the per-member (scale, shift) come from the reference's solver — scipy BFGS (numerical
gradients, ``max_iter`` steps) on the cost of :158-170 evaluated on maps down-sampled to ``max_res`` (nearest-exact,
:176-184) — which is a 2B-parameter optimisation over <= 50x50 maps and runs on the host (the reference's own torch
expressions on CPU tensors, see ensemble_depth); the part that touches
every pixel (align, per-pixel median / mean over the members, min-max normalisation, :186-203) is one pass of the
engine's ``gp_ensemble_reduce`` kernel on the device.
"""
from functools import partial

import numpy as np
import torch

from . import engine as E
from .image_util import get_tv_resample_method, resize_max_res


def ensemble_depth(depth, scale_invariant=True, shift_invariant=True, output_uncertainty=False, reduction="median",
                   regularizer_strength=0.02, max_iter=2, tol=1e-3, max_res=1024):
    """depth: fp32 [B,1,H,W] (cuda).  Returns (ensembled [1,1,H,W] on the device, None).

    The solver half repeats the reference's torch expressions literally (on CPU tensors): its finite-difference BFGS
    is sensitive to the last bit of the cost, so only the same float32 operations in the same order reproduce its
    (scale, shift)."""
    if depth.dim() != 4 or depth.shape[1] != 1:
        raise ValueError(f"Expecting 4D tensor of shape [B,1,H,W]; got {depth.shape}.")
    if reduction not in ("mean", "median"):
        raise ValueError(f"Unrecognized reduction method: {reduction}.")
    if not scale_invariant and shift_invariant:
        raise ValueError("Pure shift-invariant ensembling is not supported.")
    if output_uncertainty:
        raise NotImplementedError("output_uncertainty is not used by the reference's callers (genpercept_pipeline.py:289-296)")
    if not scale_invariant:
        raise ValueError("Unrecognized alignment.")               # :198-199 (absolute predictions are not normalised there)
    ensemble_size = depth.shape[0]
    depth = depth.to(torch.float32).contiguous()

    def init_param(d):                                            # :97-112
        init_min = d.reshape(ensemble_size, -1).min(dim=1).values
        init_max = d.reshape(ensemble_size, -1).max(dim=1).values
        if shift_invariant:
            init_s = 1.0 / (init_max - init_min).clamp(min=1e-6)
            init_t = -init_s * init_min
            return torch.cat((init_s, init_t)).cpu().numpy()
        return (1.0 / init_max.clamp(min=1e-6)).cpu().numpy()

    def align(d, param):                                          # :114-126
        if shift_invariant:
            s, t = np.split(param, 2)
            s = torch.from_numpy(s).to(d).view(ensemble_size, 1, 1, 1)
            t = torch.from_numpy(t).to(d).view(ensemble_size, 1, 1, 1)
            return d * s + t
        return d * torch.from_numpy(param).to(d).view(ensemble_size, 1, 1, 1)

    def reduce_(d):                                               # :128-145
        if reduction == "mean":
            return torch.mean(d, dim=0, keepdim=True)
        return torch.median(d, dim=0, keepdim=True).values

    def cost_fn(param, d):                                        # :147-162
        cost = 0.0
        da = align(d, param)
        for i, j in torch.combinations(torch.arange(ensemble_size)):
            diff = da[i] - da[j]
            cost += (diff ** 2).mean().sqrt().item()
        if regularizer_strength > 0:
            prediction = reduce_(da)
            err_near = (0.0 - prediction.min()).abs().item()
            err_far = (1.0 - prediction.max()).abs().item()
            cost += (err_near + err_far) * regularizer_strength
        return cost

    import scipy.optimize
    d = depth.cpu()                                               # the members; the solver works on <= max_res maps (:167-177)
    if max_res is not None and max(d.shape[2:]) > max_res:
        d = resize_max_res(d, max_res, get_tv_resample_method("nearest-exact"))
    param = init_param(d)
    res = scipy.optimize.minimize(partial(cost_fn, d=d), param, method="BFGS", tol=tol,
                                  options={"maxiter": max_iter, "disp": False})
    if shift_invariant:
        scale, shift = (a.astype(np.float32) for a in np.split(res.x, 2))
    else:
        scale, shift = res.x.astype(np.float32), np.zeros(ensemble_size, dtype=np.float32)
    # align + per-pixel median / mean + (x - min) / (max - min).clamp(1e-6) on the device (:186-203)
    out = E.ensemble_reduce(depth, scale, shift, median=reduction == "median", normalise=1 if shift_invariant else 2)
    return out, None
