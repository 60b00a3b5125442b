"""world_size-2 gloo test of the batch sharding + all-gather logic (spec ⑤) with a stand-in
infer function (the engine itself needs a GPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from genpercept_b200.parallel import shard_bounds, sharded_infer


def test_shard_bounds_cover_batch():
    for B in (1, 2, 7, 8, 64):
        for ws in (1, 2, 4, 8):
            spans = [shard_bounds(B, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for a, b in zip(spans[:-1], spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_infer(x):
    return x.float().mean(dim=1, keepdim=True) / 255.0


def _worker(rank, ws, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    g = torch.Generator().manual_seed(0)
    rgb = torch.randint(0, 256, (B, 3, 8, 16), generator=g, dtype=torch.uint8)
    full = sharded_infer(_fake_infer, rgb, stacked=True)
    local = sharded_infer(_fake_infer, rgb, stacked=False)
    lo, hi = shard_bounds(B, ws, rank)
    ok = torch.equal(full, _fake_infer(rgb)) and torch.equal(local, _fake_infer(rgb[lo:hi]))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
    assert all(ok for _, ok in res), res


def test_sharded_infer_even_split_gloo_ws2():
    _run(4)


def test_sharded_infer_ragged_split_gloo_ws2():
    _run(5)
