"""Frame sharding across GPUs: the band map and the gather plumbing.

Pixels are independent (naive_renderer.c:216-236 keeps no state between them), so
the frame is cut into bands of 4 rows (one 8x4 warp tile high) and band b goes to
rank b % world: cyclic, because contiguous blocks are badly unbalanced (sky vs
objects).  A rank stores its bands compactly, band after band, and
every rank's buffer is padded to the size of rank 0's, so one gather with equal
counts moves the frame; rank 0 then de-interleaves with a CUDA kernel
(lolb200_deinterleave_device).

One process per GPU; torch.distributed (NCCL over NVLink on the GPUs, gloo in the
CPU tests of this plumbing) carries the gather.  There is no CPU de-interleave in
the product: tests/ restate the integer map in numpy.
"""
from __future__ import annotations

from typing import List, Optional

BAND_ROWS = 4


def n_bands(h: int) -> int:
    return (h + BAND_ROWS - 1) // BAND_ROWS


def local_bands(h: int, world: int, rank: int) -> int:
    """Bands rank owns: b = rank, rank + world, ..."""
    nb = n_bands(h)
    return (nb - rank + world - 1) // world if nb > rank else 0


def padded_local_bands(h: int, world: int) -> int:
    """Rank 0 owns the most bands; every shard buffer is sized for that."""
    return (n_bands(h) + world - 1) // world


def shard_rows(h: int, world: int, rank: int) -> List[int]:
    """Global row of each local row of rank's compact buffer, in storage order."""
    rows = []
    for lb in range(local_bands(h, world, rank)):
        band = lb * world + rank
        for r in range(BAND_ROWS):
            y = band * BAND_ROWS + r
            rows.append(y if y < h else -1)  # -1: padding row below the frame
    return rows


def row_location(y: int, world: int):
    """(rank, local row) of global row y: the inverse of shard_rows."""
    band = y // BAND_ROWS
    return band % world, (band // world) * BAND_ROWS + y % BAND_ROWS


class FrameGatherer:
    """Moves every rank's compact shard to rank 0 and rebuilds the W x H frame there."""

    def __init__(self, w: int, h: int, world: int, rank: int, device, group=None):
        import torch

        self.w, self.h, self.world, self.rank, self.group = w, h, world, rank, group
        self.shard_px = padded_local_bands(h, world) * BAND_ROWS * w
        self.device = device
        if rank == 0:
            self.gathered = torch.zeros((world, self.shard_px), dtype=torch.int32, device=device)
            self.local = self.gathered[0]  # rank 0 renders straight into its slot
            self.frame = torch.zeros((h, w), dtype=torch.int32, device=device)
        else:
            self.gathered = None
            self.local = torch.zeros((self.shard_px,), dtype=torch.int32, device=device)
            self.frame = None

    def gather(self):
        """Collective: after it, rank 0's `gathered` holds all shards (stream-ordered)."""
        import torch.distributed as dist

        if self.world == 1:
            return self.gathered
        glist = [self.gathered[i] for i in range(self.world)] if self.rank == 0 else None
        dist.gather(self.local, glist, dst=0, group=self.group)
        return self.gathered

    def assemble(self, stream: int = 0):
        """Rank 0, CUDA only: de-interleave `gathered` into `frame`."""
        from . import api

        if self.rank != 0:
            return None
        if self.frame.device.type != "cuda":
            raise api.LolB200Error(-3, "FrameGatherer.assemble needs CUDA tensors: the de-interleave "
                                       "is a CUDA kernel and there is no CPU fallback")
        api.deinterleave(self.gathered.data_ptr(), self.frame.data_ptr(), self.w, self.h, self.world,
                         self.shard_px, stream=stream)
        return self.frame
