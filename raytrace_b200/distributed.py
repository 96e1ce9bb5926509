"""Image-space sharding across GPUs.

The frame is cut into row tiles of 8..64 rows (64 = the reference's tile height, RayTracer.cpp:13;
smaller tiles balance the ranks better); tile t is rendered by rank t % world (or, RT_FLAG_SERPENTINE, in
boustrophedon order).  Pixels are independent, so there is no exchange while tracing: the only
collective is one gather of the finished RGB8 bands to rank 0 per frame (NCCL over NVLink on GPUs,
gloo in the CPU tests).  Everything here is plumbing on torch tensors; the kernels live in csrc/.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist


def bands_of(rank: int, world: int, height: int, tile_rows: int = 64, serpentine: bool = False):
    """Row tiles rendered by `rank`: tile t (tile_rows rows) belongs to rank t % world; with `serpentine`
    (RT_FLAG_SERPENTINE) the odd groups of `world` tiles are dealt in reverse order, which evens out a
    ray-cost gradient down the image.  Only the floor(H/64)*64 rendered rows are tiled, the rest stays 127."""
    n = (height // 64) * 64 // tile_rows
    if not serpentine or world <= 1:
        return list(range(rank, n, world))
    out, k = [], 0
    while True:
        t = k * world + (world - 1 - rank if k & 1 else rank)
        if t >= n:
            return out
        out.append(t)
        k += 1


class FrameGather:
    """Pre-allocated buffers + the per-frame gather of one rank's bands to rank 0."""

    def __init__(self, width: int, height: int, rank: int, world: int, device, tile_rows: int = 64, serpentine: bool = False):
        self.w, self.h, self.rank, self.world, self.tile_rows = width, height, rank, world, tile_rows
        self.bands = [torch.tensor(bands_of(r, world, height, tile_rows, serpentine), dtype=torch.long, device=device) for r in range(world)]
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
        n_mine = len(self.bands[self.rank])
        if n_mine:
            torch.index_select(self._bands_view(frame), 0, self.bands[self.rank], out=self.mine[:n_mine])
        if self.world > 1:
            dist.gather(self.mine, self.parts, dst=0)
        elif self.rank == 0:
            self.parts[0].copy_(self.mine)
        if self.rank != 0:
            return None
        out = self._bands_view(self.full)
        for r in range(self.world):
            nb = len(self.bands[r])
            if nb:
                out.index_copy_(0, self.bands[r], self.parts[r][:nb])
        return self.full


class FrameLanding:
    """One-sided gather over NVLink (include/rt_b200.h rt_landing_*): the destination rank owns a
    frame-sized device buffer that every other rank maps over CUDA IPC; after a frame each rank
    copies its row tiles straight to their final place in it (one strided peer copy on the copy
    engines) and signals with a sequence number.  torch.distributed only ships the 64-byte handle."""

    def __init__(self, ctx, width: int, height: int, rank: int, world: int, dst: int = 0, backpressure: bool = False):
        from ._capi import rt

        def _check(rc, what):
            if rc != 0:
                raise RuntimeError(f"{what} failed ({rc}): {rt.rt_last_error().decode()}")
        self._rt, self._check = rt, _check
        self.rank, self.world, self.dst, self.seq = rank, world, dst, 0
        self._h = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        if rank == dst:
            _check(rt.rt_landing_create(ctx, width, height, C.byref(self._h), handle), "rt_landing_create")
            if backpressure:   # armed before anyone can push: push k then waits for the release of frame k - 1
                _check(rt.rt_landing_release(ctx, self._h, 0, None), "rt_landing_release")
        box = [bytes(handle) if rank == dst else None]
        if world > 1:
            dist.broadcast_object_list(box, src=dst)
        if rank != dst:
            buf = (C.c_ubyte * 64).from_buffer_copy(box[0])
            _check(rt.rt_landing_open(ctx, width, height, buf, C.byref(self._h)), "rt_landing_open")

    def device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._rt.rt_landing_ptr(self._h, C.byref(p), C.byref(n)), "rt_landing_ptr")
        return p.value, n.value

    def push(self, ctx, consumer_stream=None, frame=None, release=False):
        """Enqueue, behind the frame just rendered on `ctx`, the copy of this rank's rows + the signal;
        on the destination rank also the wait for every rank's signal -- on `consumer_stream` (a
        torch.cuda.Stream: whoever reads the assembled frame) or, if None, on the pipeline's own stream.
        release=True: the destination hands the buffer back right behind that wait (a consumer that reads the frame
        calls release() itself after its read); the other ranks' next push into this buffer waits for it."""
        self.seq += 1
        if frame is None:
            self._check(self._rt.rt_push_rows(ctx, self._h, self.seq), "rt_push_rows")
        else:   # frame `frame` of the batch just enqueued on ctx (rt_render_batch_async)
            self._check(self._rt.rt_push_batch_rows(ctx, frame, self._h, self.seq), "rt_push_batch_rows")
        if self.rank == self.dst and not os.environ.get("RT_DIAG_NO_LANDING_WAIT"):
            cs = C.c_void_p(consumer_stream.cuda_stream) if consumer_stream is not None else None
            self._check(self._rt.rt_landing_wait(ctx, self._h, self.seq, self.world, cs), "rt_landing_wait")
            if release:
                self._check(self._rt.rt_landing_release(ctx, self._h, self.seq, cs), "rt_landing_release")

    def release(self, ctx, consumer_stream=None):
        """Destination rank: the consumer is done with the frame of the last push (stream-ordered on `consumer_stream`)."""
        if self.rank == self.dst:
            cs = C.c_void_p(consumer_stream.cuda_stream) if consumer_stream is not None else None
            self._check(self._rt.rt_landing_release(ctx, self._h, self.seq, cs), "rt_landing_release")

    def close(self):
        if self._h:
            self._rt.rt_landing_close(self._h)
            self._h = C.c_void_p()
