"""Per-stage device times of frames in steady state with M pipelines in flight (development aid).
    python tools/stage_probe.py [world] [M] [share]"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytrace_b200 as R

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
M = int(sys.argv[2]) if len(sys.argv) > 2 else 8
share = int(sys.argv[3]) if len(sys.argv) > 3 else 1


def ck(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: {R.rt.rt_last_error().decode()}")


sc = R.Scene("c3", 1920, 1080, 0, 0)
rt = R.RayTracer(sc)
rt.maxLevel = 5
shard = dict(rank=0, world=world, tile_rows=8) if world > 1 else {}
rt.render(R.MY_MODEL_RAYTRACE, **shard)
parent = C.c_void_p(rt.context())
params = R.RenderParams(R.MY_MODEL_RAYTRACE, 5, 0, world, 0, 8 if world > 1 else 64)
pipes = []
for _ in range(M):
    h = C.c_void_p()
    ck(R.rt.rt_create_shared(parent, C.byref(h)), "rt_create_shared")
    ck(R.rt.rt_set_sm_share(h, share), "rt_set_sm_share")
    pipes.append(h)
rows = []
cnt = R.Counters()
import time
t0 = None
for k in range(40 * M):
    p = pipes[k % M]
    if k >= M:
        ck(R.rt.rt_wait(p, None), "rt_wait")
        ck(R.rt.rt_read_counters(p, C.byref(cnt)), "rt_read_counters")
        if k >= 8 * M:
            rows.append((cnt.render_ms, cnt.trace_ms, cnt.shade_ms, cnt.other_ms))
    if k == 8 * M:
        t0 = time.perf_counter()
    ck(R.rt.rt_render_async(p, C.byref(params)), "rt_render_async")
for p in pipes:
    ck(R.rt.rt_wait(p, None), "rt_wait")
dt = time.perf_counter() - t0
a = np.array(rows)
print(json.dumps({"world": world, "M": M, "share": share, "ms_per_frame": round(dt / (32 * M) * 1e3, 4),
                  "render_ms": [round(float(x), 4) for x in (a[:, 0].mean(), a[:, 0].min(), a[:, 0].max())],
                  "trace_ms_mean": round(float(a[:, 1].mean()), 4), "shade_ms_mean": round(float(a[:, 2].mean()), 4),
                  "other_ms_mean": round(float(a[:, 3].mean()), 4)}))
