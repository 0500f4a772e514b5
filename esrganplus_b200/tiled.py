"""Tiled multi-GPU inference: independent LR crops sharded across ranks, no data-path collective.

The reference processes a whole image in one forward (test_image/test.py:37, SRRaGAN_model.py:188-192) on
one GPU.  BASELINE.json config 3 ("512x512 LR tiled across 8 GPUs") cuts the LR image into crops; crops are
independent units (no BatchNorm in G, noise off in eval), so rank r simply takes crops r, r+W, r+2W, ... and
runs them as one batch.  Results are per crop (zero padding at crop edges, exactly like running the reference
on that crop); the only communication is the optional gather of finished crops for image assembly, outside
the timed path.  All functions here are host logic and work with any callable `net(batch) -> batch`.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch


def crop_grid(h: int, w: int, tile: int) -> List[Tuple[int, int, int, int]]:
    """Row-major list of (y0, x0, th, tw) covering an h x w image with crops of at most tile x tile."""
    if tile < 1:
        raise ValueError("tile must be positive")
    return [(y, x, min(tile, h - y), min(tile, w - x)) for y in range(0, h, tile) for x in range(0, w, tile)]


def shard(items: Sequence, rank: int, world: int) -> List:
    """Round-robin assignment of work units to ranks (unit i -> rank i % world)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return [it for i, it in enumerate(items) if i % world == rank]


def run_crops(net: Callable[[torch.Tensor], torch.Tensor], img: torch.Tensor,
              crops: Sequence[Tuple[int, int, int, int]], scale: int) -> List[torch.Tensor]:
    """Run `net` over the given crops of img [1,C,H,W]; equally-sized crops go through as one batch."""
    outs: List[torch.Tensor] = [None] * len(crops)  # type: ignore[list-item]
    by_shape = {}
    for i, (y, x, th, tw) in enumerate(crops):
        by_shape.setdefault((th, tw), []).append(i)
    for (th, tw), idxs in by_shape.items():
        batch = torch.cat([img[:, :, crops[i][0]:crops[i][0] + th, crops[i][1]:crops[i][1] + tw] for i in idxs], 0)
        y = net(batch.contiguous())
        assert y.shape[-2:] == (th * scale, tw * scale), (y.shape, th, tw, scale)
        for k, i in enumerate(idxs):
            outs[i] = y[k:k + 1]
    return outs


def assemble(outs: Sequence[torch.Tensor], crops: Sequence[Tuple[int, int, int, int]], h: int, w: int,
             scale: int) -> torch.Tensor:
    c = outs[0].shape[1]
    full = torch.empty((1, c, h * scale, w * scale), dtype=outs[0].dtype, device=outs[0].device)
    for o, (y, x, th, tw) in zip(outs, crops):
        full[:, :, y * scale:(y + th) * scale, x * scale:(x + tw) * scale] = o
    return full


def infer_tiled(net: Callable[[torch.Tensor], torch.Tensor], img: torch.Tensor, tile: int, scale: int = 4,
                rank: int = 0, world: int = 1, gather: bool = True):
    """x4-SR of img [1,C,H,W] by independent crops.  With world > 1 (torch.distributed initialised, one
    process per GPU) every rank runs its share; if `gather`, rank 0 returns the assembled image (others None)."""
    _, _, h, w = img.shape
    crops = crop_grid(h, w, tile)
    mine = shard(list(range(len(crops))), rank, world)
    outs = run_crops(net, img, [crops[i] for i in mine], scale)
    if world == 1:
        return assemble(outs, crops, h, w, scale)
    if not gather:
        return dict(zip(mine, outs))
    import torch.distributed as dist
    payload = [(i, o.cpu()) for i, o in zip(mine, outs)]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if rank != 0:
        return None
    flat = dict(kv for part in gathered for kv in part)
    return assemble([flat[i] for i in range(len(crops))], crops, h, w, scale)
