"""Frames in flight on one GPU: M pipelines (rt_create_shared) over one resident scene (development aid).

    python tools/pipe_probe.py [config] [world]     env: RT_PIPE_M="1 2 3 4", RT_PIPE_SHARE="0 4 2", RT_PIPE_STEPS=60
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytrace_b200 as R

CFG = {"c2": ("c2", 1920, 1080, 5), "c3": ("c3", 1920, 1080, 5), "c4": ("c4", 3840, 2160, 8)}
name, w, h, level = CFG[sys.argv[1] if len(sys.argv) > 1 else "c3"]
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(os.environ.get("RT_PIPE_STEPS", "60"))
BATCH = int(os.environ.get("RT_PIPE_BATCH", "1"))   # frames per launch (rt_render_batch_async)


def ck(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: {R.rt.rt_last_error().decode()}")


sc = R.Scene(name, w, h, 0, 0)
rt = R.RayTracer(sc)
rt.maxLevel = level
shard = dict(rank=0, world=world, tile_rows=8) if world > 1 else {}
ref = rt.render(R.MY_MODEL_RAYTRACE, **shard)
c0 = rt.counters()
rays = c0.primary + c0.shadow + c0.reflect + c0.refract
parent = C.c_void_p(rt.context())
params = R.RenderParams(R.MY_MODEL_RAYTRACE, level, 0, world, 0, 8 if world > 1 else 64)
pipes = [parent]
for M in [int(x) for x in os.environ.get("RT_PIPE_M", "1 2 3 4").split()]:
    while len(pipes) < M:
        hnd = C.c_void_p()
        ck(R.rt.rt_create_shared(parent, C.byref(hnd)), "rt_create_shared")
        pipes.append(hnd)
    for share in [int(x) for x in os.environ.get("RT_PIPE_SHARE", "0").split()]:
        for p in pipes[:M]:
            ck(R.rt.rt_set_sm_share(p, share), "rt_set_sm_share")
        def enqueue(p):
            if BATCH > 1:
                ck(R.rt.rt_render_batch_async(p, C.byref(params), BATCH, None, None), "rt_render_batch_async")
            else:
                ck(R.rt.rt_render_async(p, C.byref(params)), "rt_render_async")
        for p in pipes[:M]:
            enqueue(p)
        same = True
        for p in pipes[:M]:
            ck(R.rt.rt_wait(p, None), "rt_wait")
            for f in range(BATCH):
                out = np.empty((h, w, 3), dtype=np.uint8)
                ck(R.rt.rt_read_batch_output(p, f, out.ctypes.data_as(C.c_void_p), w * 3, 0), "rt_read_batch_output")
                same = same and bool((out == ref).all())
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            for k in range(steps):
                enqueue(pipes[k % M])
            for p in pipes[:M]:
                ck(R.rt.rt_wait(p, None), "rt_wait")
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        print(json.dumps({"cfg": name, "world": world, "batch": BATCH, "pipelines": M, "ctas_per_sm": share, "sched": os.environ.get("RT_B200_SCHED", "auto"), "frames_identical": same,
                          "ms_per_frame": round(best / (steps * BATCH) * 1e3, 4), "mrays_s": round(rays * steps * BATCH / best / 1e6, 1)}), flush=True)
