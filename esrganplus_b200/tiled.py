"""Tiled multi-GPU inference: independent LR crops sharded across ranks, no data-path collective.

The reference processes a whole image in one forward (test_image/test.py:37, SRRaGAN_model.py:188-192) on
one GPU.  BASELINE.json config 3 ("512x512 LR tiled across 8 GPUs") cuts the LR image into crops; crops are
independent units (no BatchNorm in G, noise off in eval), so rank r simply takes crops r, r+W, r+2W, ... and
runs them as one batch.  The only communication is the optional gather of finished crops for image assembly,
outside the timed path.  All functions here are host logic and work with any callable `net(batch) -> batch`.

Seams.  With ``halo=0`` every crop is super-resolved on its own: the convolutions zero-pad at the crop's edges,
so pixels near a crop border differ from the whole-image forward the reference runs (the generator's receptive
radius is 347 LR pixels; the influence decays fast, but a seam is visible).  ``halo=k`` reads
each crop with k extra LR pixels on every side that lies inside the image and keeps only the centre: the interior
then converges to the whole-image result as k grows (tests/test_tiled_dist.py measures it), at (1 + 2k/tile)^2
the work.  Image borders are zero-padded by the network in both cases, exactly like the whole-image forward.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch


def crop_grid(h: int, w: int, tile: int) -> List[Tuple[int, int, int, int]]:
    """Row-major list of (y0, x0, th, tw) covering an h x w image with crops of at most tile x tile."""
    if tile < 1:
        raise ValueError("tile must be positive")
    return [(y, x, min(tile, h - y), min(tile, w - x)) for y in range(0, h, tile) for x in range(0, w, tile)]


def with_halo(crop: Tuple[int, int, int, int], h: int, w: int, halo: int) -> Tuple[int, int, int, int, int, int]:
    """The region to READ for a crop: (y0, x0, th, tw, top, left) = the crop grown by `halo` pixels on every side that
    stays inside the h x w image; (top, left) is where the crop itself starts inside that region."""
    y, x, th, tw = crop
    y0, x0 = max(0, y - halo), max(0, x - halo)
    y1, x1 = min(h, y + th + halo), min(w, x + tw + halo)
    return y0, x0, y1 - y0, x1 - x0, y - y0, x - x0


def shard(items: Sequence, rank: int, world: int) -> List:
    """Round-robin assignment of work units to ranks (unit i -> rank i % world)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return [it for i, it in enumerate(items) if i % world == rank]


def run_crops(net: Callable[[torch.Tensor], torch.Tensor], img: torch.Tensor,
              crops: Sequence[Tuple[int, int, int, int]], scale: int, halo: int = 0) -> List[torch.Tensor]:
    """Run `net` over the given crops of img [1,C,H,W]; equally-sized read regions go through as one batch.  With
    halo > 0 each crop is read with its halo (see the module docstring) and only its centre is returned."""
    _, _, h, w = img.shape
    outs: List[torch.Tensor] = [None] * len(crops)  # type: ignore[list-item]
    reads = [with_halo(c, h, w, halo) if halo > 0 else (c[0], c[1], c[2], c[3], 0, 0) for c in crops]
    by_shape = {}
    for i, (_, _, rh, rw, _, _) in enumerate(reads):
        by_shape.setdefault((rh, rw), []).append(i)
    for (rh, rw), idxs in by_shape.items():
        batch = torch.cat([img[:, :, reads[i][0]:reads[i][0] + rh, reads[i][1]:reads[i][1] + rw] for i in idxs], 0)
        y = net(batch.contiguous())
        assert y.shape[-2:] == (rh * scale, rw * scale), (y.shape, rh, rw, scale)
        for k, i in enumerate(idxs):
            _, _, th, tw = crops[i]
            top, left = reads[i][4] * scale, reads[i][5] * scale
            outs[i] = y[k:k + 1, :, top:top + th * scale, left:left + tw * scale]
    return outs


def assemble(outs: Sequence[torch.Tensor], crops: Sequence[Tuple[int, int, int, int]], h: int, w: int,
             scale: int) -> torch.Tensor:
    c = outs[0].shape[1]
    full = torch.empty((1, c, h * scale, w * scale), dtype=outs[0].dtype, device=outs[0].device)
    for o, (y, x, th, tw) in zip(outs, crops):
        full[:, :, y * scale:(y + th) * scale, x * scale:(x + tw) * scale] = o
    return full


def infer_tiled(net: Callable[[torch.Tensor], torch.Tensor], img: torch.Tensor, tile: int, scale: int = 4,
                rank: int = 0, world: int = 1, gather: bool = True, halo: int = 0):
    """x4-SR of img [1,C,H,W] by independent crops (read with `halo` extra LR pixels per side, see the module
    docstring; halo = 0 leaves seams at crop borders).  With world > 1 (torch.distributed initialised, one process
    per GPU) every rank runs its share; if `gather`, rank 0 returns the assembled image (others None)."""
    _, _, h, w = img.shape
    crops = crop_grid(h, w, tile)
    mine = shard(list(range(len(crops))), rank, world)
    outs = run_crops(net, img, [crops[i] for i in mine], scale, halo)
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
