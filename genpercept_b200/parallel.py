"""Batch sharding across ranks: images are independent, so rank r runs images
[r*B/R, (r+1)*B/R) on its own engine replica with no data-path collective; a single all-gather
of the decoded maps happens only when the caller asks for the stacked result.

One process per GPU (torchrun); the collective is torch.distributed (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  The reference has no multi-GPU inference path at all.
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, world_size, rank):
    """Contiguous near-equal split; the first (batch % world_size) ranks get one extra image."""
    base, rem = divmod(batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_infer(infer_fn, rgb, stacked=True, group=None):
    """infer_fn(rgb_shard[b,3,H,W]) -> fp32 [b,C,H,W] on the local device.

    rgb: the FULL batch [B,3,H,W] (every rank passes the same tensor, as a sharded data loader
    would).  Returns the local shard's result, or — stacked=True — the full [B,C,H,W] on every rank.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return infer_fn(rgb)
    ws, rk = dist.get_world_size(group), dist.get_rank(group)
    B = rgb.shape[0]
    lo, hi = shard_bounds(B, ws, rk)
    local = infer_fn(rgb[lo:hi]) if hi > lo else None
    if not stacked:
        return local
    sizes = [shard_bounds(B, ws, r)[1] - shard_bounds(B, ws, r)[0] for r in range(ws)]
    ref = local          # None when there are more ranks than images: shape comes from the others
    meta = torch.tensor(list(ref.shape[1:]) if ref is not None else [0, 0, 0], dtype=torch.int64,
                        device=ref.device if ref is not None else ("cuda" if torch.cuda.is_available() and dist.get_backend(group) == "nccl" else "cpu"))
    dist.all_reduce(meta, op=dist.ReduceOp.MAX, group=group)
    C, H, W = (int(v) for v in meta.tolist())
    dev = meta.device
    if max(sizes) == min(sizes):
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad, C, H, W), dtype=torch.float32, device=dev)
    if local is not None:
        buf[: hi - lo] = local
    parts = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
