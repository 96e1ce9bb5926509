"""torchrun helper: N ranks render their row tiles of one frame; the tiles reach rank 0 (a) through
the one-sided NVLink landing buffer (rt_push_rows) and (b) through the NCCL gather; rank 0 compares
both with its own full-frame render.  Exit code 0 = identical.  Used by tests/test_gpu_multi.py."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytrace_b200 as R
from raytrace_b200.distributed import FrameGather, FrameLanding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
w, h, level, tile_rows = 640, 448, 4, 8
os.makedirs(f"/tmp/rt_p2p_{rank}", exist_ok=True)
sc = R.Scene("c3", w, h, 96, 6, tmpdir=f"/tmp/rt_p2p_{rank}")


def ck(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: {R.rt.rt_last_error().decode()}")


owner = C.c_void_p()
ck(R.rt.rt_create(local, C.byref(owner)), "rt_create")
ck(R.rt.rt_upload_scene(owner, sc.flatten()), "rt_upload_scene")
ok = True
full = None
if rank == 0:
    ck(R.rt.rt_render_async(owner, C.byref(R.RenderParams(R.MY_MODEL_RAYTRACE, level, 0, 1, 0, 64))), "render full")
    full = np.empty((h, w, 3), dtype=np.uint8)
    ck(R.rt.rt_wait(owner, None), "rt_wait")
    ck(R.rt.rt_read_output(owner, full.ctypes.data_as(C.c_void_p), w * 3), "rt_read_output")
pipes = []
for i in range(2):
    p = C.c_void_p()
    ck(R.rt.rt_create_shared(owner, C.byref(p)), "rt_create_shared")
    st = torch.cuda.Stream(dev)
    ck(R.rt.rt_set_stream(p, C.c_void_p(st.cuda_stream)), "rt_set_stream")
    landing = FrameLanding(p, w, h, rank, world)
    frame = None
    if rank == 0:
        ptr, nbytes = landing.device_ptr()
        ck(R.rt.rt_set_output(p, C.c_void_p(ptr), nbytes), "rt_set_output")
    else:
        frame = torch.full((h, w, 3), 127, dtype=torch.uint8, device=dev)
        ck(R.rt.rt_set_output(p, C.c_void_p(frame.data_ptr()), frame.numel()), "rt_set_output")
    pipes.append((p, st, landing, frame))
SERP = os.environ.get("RT_P2P_SERPENTINE", "1") != "0"       # boustrophedon shard order (RT_FLAG_SERPENTINE), as bench.py uses it
params = R.RenderParams(R.MY_MODEL_RAYTRACE, level, rank, world, R.RT_FLAG_SERPENTINE if SERP else 0, tile_rows)
def check_landings(label):
    """rank 0: every pipeline's landing buffer (its own rows + the peers' pushes) against the full-frame render"""
    global ok
    for p, st, landing, frame in pipes:
        ck(R.rt.rt_wait(p, None), "rt_wait")
        st.synchronize()
    torch.cuda.synchronize(dev)
    dist.barrier()
    if rank == 0:
        for p, st, landing, frame in pipes:
            got = np.empty((h, w, 3), dtype=np.uint8)
            ck(R.rt.rt_read_output(p, got.ctypes.data_as(C.c_void_p), w * 3), "rt_read_output")   # the pipeline's output IS the landing buffer
            same = bool(np.array_equal(got, full))
            print(f"{label} == full frame:", same, flush=True)
            ok = ok and same
    dist.barrier()


for k in range(6):                      # 3 frames per pipeline, 2 in flight
    if k == 0 and rank == 0:
        import time
        time.sleep(0.3)                 # the peers' rows of the FIRST frame land before rank 0 has enqueued anything: nothing
                                        # rank 0 does for its first frame (no grey fill of a landing buffer) may wipe them
    p, st, landing, frame = pipes[k % 2]
    ck(R.rt.rt_render_async(p, C.byref(params)), "rt_render_async")
    landing.push(p)
    if k == 1:
        check_landings("first p2p landing")
check_landings("p2p landing")
# (a2) a BATCH of two frames in one launch (rt_render_batch_async), each frame pushed into its own landing buffer
p0 = pipes[0][0]
outs = (C.c_void_p * 2)()
for f in range(2):
    outs[f] = pipes[f][2].device_ptr()[0] if rank == 0 else pipes[f][3].data_ptr()
for k in range(2):
    ck(R.rt.rt_render_batch_async(p0, C.byref(params), 2, None, outs), "rt_render_batch_async")
    for f in range(2):
        pipes[f][2].push(p0, frame=f)
ck(R.rt.rt_wait(p0, None), "rt_wait")
pipes[0][1].synchronize()
torch.cuda.synchronize(dev)
dist.barrier()
if rank == 0:
    for f in range(2):
        got = np.empty((h, w, 3), dtype=np.uint8)
        ck(R.rt.rt_read_batch_output(p0, f, got.ctypes.data_as(C.c_void_p), w * 3, 0), "rt_read_batch_output")
        same = bool(np.array_equal(got, full))
        print(f"batched p2p landing {f} == full frame:", same, flush=True)
        ok = ok and same
dist.barrier()
# (a3) back-pressure: ONE landing buffer, four frames with four cameras pushed back to back, a consumer on rank 0 that is
# slow (it sleeps on its stream before it snapshots the assembled frame, then releases the buffer).  A peer's push of
# frame k + 1 must wait for the release of frame k, or the snapshot of frame k shows rows of frame k + 1.
class _DevView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}


bp, bst = pipes[0][0], pipes[0][1]
bl = FrameLanding(bp, w, h, rank, world, backpressure=True)
bframe = None
if rank == 0:
    bptr, bn = bl.device_ptr()
    bview = torch.as_tensor(_DevView(bptr, bn), device=dev)
else:
    bframe = torch.full((h, w, 3), 127, dtype=torch.uint8, device=dev)
cons = torch.cuda.Stream(dev)
moves = ((0.3, -0.2, 0.5), (-0.6, 0.3, 0.2), (0.4, 0.4, -0.3), (-0.2, -0.5, 0.6))
cams = (R.Camera * 4)()
for k, mv in enumerate(moves):
    sc.camera_move(*mv)
    cams[k] = sc.flatten().contents.camera
fulls, snaps, ev = [], [], None
if rank == 0:
    fp = R.RenderParams(R.MY_MODEL_RAYTRACE, level, 0, 1, 0, 64)
    for k in range(4):
        ck(R.rt.rt_render_batch_async(owner, C.byref(fp), 1, C.cast(C.byref(cams, k * C.sizeof(R.Camera)), C.POINTER(R.Camera)), None), "render full k")
        ck(R.rt.rt_wait(owner, None), "rt_wait")
        g = np.empty((h, w, 3), dtype=np.uint8)
        ck(R.rt.rt_read_batch_output(owner, 0, g.ctypes.data_as(C.c_void_p), w * 3, 0), "rt_read_batch_output")
        fulls.append(g)
dist.barrier()
one = (C.c_void_p * 1)()
one[0] = bl.device_ptr()[0] if rank == 0 else bframe.data_ptr()
for k in range(4):
    if rank == 0 and ev is not None:
        bst.wait_event(ev)            # rank 0's own rows of frame k + 1 also wait for the snapshot of frame k
    ck(R.rt.rt_render_batch_async(bp, C.byref(params), 1, C.cast(C.byref(cams, k * C.sizeof(R.Camera)), C.POINTER(R.Camera)), one), "rt_render_batch_async")
    bl.push(bp, cons, frame=0)
    if rank == 0:
        with torch.cuda.stream(cons):
            torch.cuda._sleep(60_000_000)          # ~30 ms: long enough for the peers to have enqueued every later frame
            snaps.append(bview.clone())
        bl.release(bp, cons)
        ev = torch.cuda.Event()
        ev.record(cons)
ck(R.rt.rt_wait(bp, None), "rt_wait")
bst.synchronize(), cons.synchronize()
torch.cuda.synchronize(dev)
dist.barrier()
if rank == 0:
    for k in range(4):
        same = bool(np.array_equal(snaps[k].cpu().numpy().reshape(h, w, 3), fulls[k]))
        print(f"back-pressure: snapshot of frame {k} == full frame of camera {k}:", same, flush=True)
        ok = ok and same
dist.barrier()
bl.close()
# (b) NCCL gather of the same shards
for _ in range(4):
    pass
sc = R.Scene("c3", w, h, 96, 6, tmpdir=f"/tmp/rt_p2p_{rank}")       # back to the configuration's own camera
ck(R.rt.rt_upload_scene(owner, sc.flatten()), "rt_upload_scene")
p, st, landing, frame = pipes[0]
if rank == 0:
    frame = torch.full((h, w, 3), 127, dtype=torch.uint8, device=dev)
    ck(R.rt.rt_set_output(p, C.c_void_p(frame.data_ptr()), frame.numel()), "rt_set_output")
g = FrameGather(w, h, rank, world, dev, tile_rows, SERP)
ck(R.rt.rt_render_async(p, C.byref(params)), "rt_render_async")
with torch.cuda.stream(st):
    out = g.gather(frame)
st.synchronize()
if rank == 0:
    same = bool(np.array_equal(out.cpu().numpy(), full))
    print("nccl gather == full frame:", same, flush=True)
    ok = ok and same
dist.barrier()
for p, st, landing, frame in pipes:
    landing.close()
    R.rt.rt_destroy(p)
R.rt.rt_destroy(owner)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
