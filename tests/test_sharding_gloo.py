"""world_size-2 (and 3) gloo test of the multi-GPU plumbing on CPU: every rank produces its
interleaved 64-row bands (here with the CPU oracle, on the GPU box the CUDA path does), the bands
are gathered to rank 0 and must assemble to the single-process frame bit-exactly."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, w, h, level, out_path, tile_rows, serpentine):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import raytrace_b200 as R
    from parity_util import oracle_render
    from raytrace_b200.distributed import FrameGather, bands_of

    sc = R.Scene("t_mixed", w, h)
    part, _, cnt = oracle_render(sc, level, want_ids=False, threads=2, rank=rank, world=world, tile_rows=tile_rows,
                                 flags=R.RT_FLAG_SERPENTINE if serpentine else 0)
    g = FrameGather(w, h, rank, world, torch.device("cpu"), tile_rows, serpentine)
    full = g.gather(torch.from_numpy(part.copy()))
    rays = torch.tensor([cnt.primary + cnt.shadow + cnt.reflect + cnt.refract], dtype=torch.int64)
    dist.all_reduce(rays)
    if rank == 0:
        np.save(out_path, full.numpy())
        np.save(out_path + ".rays.npy", rays.numpy())
    if not serpentine:
        assert bands_of(rank, world, h, tile_rows) == [t for t in range((h // 64) * 64 // tile_rows) if t % world == rank]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,tile_rows,serpentine", [(2, 64, False), (3, 64, False), (3, 8, False), (2, 8, True), (3, 16, True)])
def test_bands_gather_to_the_full_frame(tmp_path, world, tile_rows, serpentine):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import raytrace_b200 as R
    from parity_util import oracle_render

    w, h, level = 256, 330, 3   # 5 bands of 64 rows + 10 never-rendered rows: uneven split on purpose
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, _free_port(), w, h, level, out, tile_rows, serpentine), nprocs=world, join=True)
    full = np.load(out)
    sc = R.Scene("t_mixed", w, h)
    ref, _, cnt = oracle_render(sc, level, want_ids=False, threads=2)
    assert np.array_equal(full, ref)
    assert int(np.load(out + ".rays.npy")[0]) == cnt.primary + cnt.shadow + cnt.reflect + cnt.refract
