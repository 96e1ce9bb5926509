"""Image-space sharding across GPUs.

The frame is cut into row tiles of 8..64 rows (64 = the reference's tile height, RayTracer.cpp:13;
smaller tiles balance the ranks better); tile t is rendered by rank t % world.  Pixels are independent, so there is no exchange while tracing: the only
collective is one gather of the finished RGB8 bands to rank 0 per frame (NCCL over NVLink on GPUs,
gloo in the CPU tests).  Everything here is plumbing on torch tensors; the kernels live in csrc/.
"""
import torch
import torch.distributed as dist


def bands_of(rank: int, world: int, height: int, tile_rows: int = 64):
    """Row tiles rendered by `rank`: tile t (tile_rows rows) belongs to rank t % world.  Only the
    floor(H/64)*64 rendered rows are tiled, the rest of the frame stays 127."""
    return list(range(rank, (height // 64) * 64 // tile_rows, world))


class FrameGather:
    """Pre-allocated buffers + the per-frame gather of one rank's bands to rank 0."""

    def __init__(self, width: int, height: int, rank: int, world: int, device, tile_rows: int = 64):
        self.w, self.h, self.rank, self.world, self.tile_rows = width, height, rank, world, tile_rows
        self.blk_h = (height // 64) * 64 // tile_rows      # number of row tiles in the frame
        self.band_bytes = tile_rows * width * 3
        self.max_bands = (self.blk_h + world - 1) // world
        self.mine = torch.zeros((max(self.max_bands, 1), self.band_bytes), dtype=torch.uint8, device=device)
        self.parts = [torch.zeros_like(self.mine) for _ in range(world)] if rank == 0 else None
        self.full = torch.full((height, width, 3), 127, dtype=torch.uint8, device=device) if rank == 0 else None

    def _bands_view(self, frame):
        return frame.view(-1)[: self.blk_h * self.band_bytes].view(self.blk_h, self.band_bytes)

    def gather(self, frame: torch.Tensor):
        """frame: this rank's (H, W, 3) uint8 framebuffer (only its own bands are valid).
        Returns the assembled frame on rank 0, None elsewhere."""
        n_mine = len(bands_of(self.rank, self.world, self.h, self.tile_rows))
        if n_mine:
            self.mine[:n_mine].copy_(self._bands_view(frame)[self.rank::self.world])
        if self.world > 1:
            dist.gather(self.mine, self.parts, dst=0)
        elif self.rank == 0:
            self.parts[0].copy_(self.mine)
        if self.rank != 0:
            return None
        out = self._bands_view(self.full)
        for r in range(self.world):
            nb = len(bands_of(r, self.world, self.h, self.tile_rows))
            if nb:
                out[r::self.world].copy_(self.parts[r][:nb])
        return self.full
