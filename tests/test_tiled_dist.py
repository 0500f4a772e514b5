"""Host logic of tiled multi-GPU inference (esrganplus_b200/tiled.py), run on CPU with the gloo backend and
world_size 2: crops are sharded with no data-path collective and the gathered image equals the single-process
result.  The network is a stand-in callable (nearest x4) — what is under test is the sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from esrganplus_b200 import tiled


def _fake_net(x):
    return F.interpolate(x, scale_factor=4, mode="nearest") + 0.5


def test_crop_grid_and_shard_cover_everything():
    crops = tiled.crop_grid(70, 130, 32)
    assert len(crops) == 3 * 5
    area = sum(th * tw for _, _, th, tw in crops)
    assert area == 70 * 130
    parts = [tiled.shard(crops, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == sorted(crops)
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        tiled.shard(crops, 4, 4)


def test_single_process_tiled_equals_per_crop():
    img = torch.rand(1, 3, 70, 130)
    out = tiled.infer_tiled(_fake_net, img, 32)
    assert out.shape == (1, 3, 280, 520)
    assert torch.equal(out, _fake_net(img))  # the stand-in is local, so tiling is exact


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    img = torch.rand(1, 3, 64, 96)
    calls = []

    def net(x):
        calls.append(x.shape[0])
        return _fake_net(x)

    out = tiled.infer_tiled(net, img, 32, rank=rank, world=world)
    if rank == 0:
        q.put((torch.equal(out, _fake_net(img)), sum(calls)))
    else:
        q.put((out is None, sum(calls)))
    dist.destroy_process_group()


def test_world2_gloo_shards_and_gathers():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for ok, _ in res)
    assert sorted(n for _, n in res) == [3, 3]  # 6 crops of 32x32, 3 per rank
