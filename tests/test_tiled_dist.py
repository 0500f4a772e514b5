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


def test_halo_makes_the_interior_converge_to_the_whole_image_forward():
    """crops with zero padding at every crop edge leave seams.  With a stand-in network of receptive radius 3
    (three 3x3 convs, zero padding, then nearest x4) a halo >= 3 reproduces the whole-image forward EXACTLY everywhere;
    halo = 0 differs along the crop borders only; image borders are zero-padded in both (like the reference's forward)."""
    torch.manual_seed(0)
    w1, w2, w3 = (torch.randn(3, 3, 3, 3) * 0.3 for _ in range(3))

    def net(x):
        for wt in (w1, w2, w3):
            x = F.leaky_relu(F.conv2d(x, wt, padding=1), 0.2)
        return F.interpolate(x, scale_factor=4, mode="nearest")

    img = torch.rand(1, 3, 40, 56)
    whole = net(img)
    seams = tiled.infer_tiled(net, img, 16, halo=0)
    exact = tiled.infer_tiled(net, img, 16, halo=3)
    assert torch.allclose(exact, whole, atol=1e-6), (exact - whole).abs().max()
    assert torch.allclose(tiled.infer_tiled(net, img, 16, halo=8), whole, atol=1e-6)
    d = (seams - whole).abs().amax(dim=(0, 1))
    assert d.max() > 1e-3, "halo = 0 must show the seam this test is about"
    interior = d[4 * 4:4 * 12, 4 * 4:4 * 12]      # centre of the first crop, more than 3 LR pixels from its edges
    assert interior.max() < 1e-6
    # read regions never leave the image and keep the crop at the stated offset
    assert tiled.with_halo((0, 16, 16, 16), 40, 56, 3) == (0, 13, 19, 22, 0, 3)
    assert tiled.with_halo((32, 48, 8, 8), 40, 56, 3) == (29, 45, 11, 11, 3, 3)
