"""Image-tile partition of one frame across the GPUs of a box, and the gather of
the finished tiles to rank 0 (the reference is single-GPU, this
is the new multi-GPU capability BASELINE.json asks for).

Every ray is independent and the octree is read-only during a frame, so the
octree is REPLICATED on every GPU and only the image is partitioned.  The unit
is a band of `band` image rows (a multiple of 4: the kernels walk 8x4 pixel
tiles); bands are dealt round-robin so that sky rows (cheap) and terrain rows
(expensive) are spread evenly.  One exchange step per frame: every rank's bands
go to rank 0 (`gather_bands`: `torch.distributed` gather -- NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests).

Host-side logic only; nothing here renders.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["row_bands", "bands_of", "pack_bands", "unpack_bands", "gather_bands"]


def row_bands(height: int, band: int = 8) -> List[Tuple[int, int]]:
    """All bands [(y0, y1)] of an image, top to bottom; the last may be short."""
    if band <= 0 or band % 4:
        raise ValueError("band must be a positive multiple of 4")
    return [(y, min(y + band, height)) for y in range(0, height, band)]


def bands_of(rank: int, world_size: int, height: int, band: int = 8) -> List[Tuple[int, int]]:
    """The bands rank `rank` renders: every world_size-th band, starting at its rank."""
    return row_bands(height, band)[rank::world_size]


def _slots(world_size: int, height: int, band: int) -> int:
    """Bands per rank, padded so every rank sends the same amount."""
    n = len(row_bands(height, band))
    return (n + world_size - 1) // world_size


def pack_bands(image: torch.Tensor, rank: int, world_size: int, band: int = 8) -> torch.Tensor:
    """Copy this rank's bands of a full-size [H, W, ...] plane into a dense [slots*band, W, ...] send buffer."""
    height = image.shape[0]
    slots = _slots(world_size, height, band)
    out = image.new_zeros((slots * band,) + tuple(image.shape[1:]))
    for i, (y0, y1) in enumerate(bands_of(rank, world_size, height, band)):
        out[i * band:i * band + (y1 - y0)] = image[y0:y1]
    return out


def unpack_bands(gathered: Sequence[torch.Tensor], height: int, band: int = 8) -> torch.Tensor:
    """Inverse of pack_bands over all ranks: assemble the full [H, W, ...] plane (rank 0 side)."""
    world_size = len(gathered)
    first = gathered[0]
    out = first.new_zeros((height,) + tuple(first.shape[1:]))
    for r in range(world_size):
        for i, (y0, y1) in enumerate(bands_of(r, world_size, height, band)):
            out[y0:y1] = gathered[r][i * band:i * band + (y1 - y0)]
    return out


def gather_bands(image: torch.Tensor, band: int = 8, dst: int = 0, group=None):
    """The per-frame exchange step.  `image` is this rank's full-size plane with (at least) its own bands
    rendered.  Returns the assembled plane on rank `dst`, None elsewhere."""
    world_size = dist.get_world_size(group)
    rank = dist.get_rank(group)
    send = pack_bands(image, rank, world_size, band)
    if world_size == 1:
        return unpack_bands([send], image.shape[0], band)
    recv = [torch.empty_like(send) for _ in range(world_size)] if rank == dst else None
    dist.gather(send, recv, dst=dst, group=group)
    if rank != dst:
        return None
    return unpack_bands(recv, image.shape[0], band)
