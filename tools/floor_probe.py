import sys, os, time
sys.path.insert(0, os.getcwd())
import raytrace_b200 as R
sc = R.Scene("c3", 1920, 1080)
rt = R.RayTracer(sc); rt.maxLevel = 5
for world in (1, 2, 4, 8, 16):
    best = None
    for _ in range(6):
        rt.start(R.MY_MODEL_RAYTRACE, rank=world // 2 if world > 1 else 0, world=world); rt.wait()
        c = rt.counters()
        if best is None or c.render_ms < best[0]: best = (c.render_ms, c.primary + c.shadow + c.reflect + c.refract, c.trace_ms)
    print(os.environ.get("RT_B200_SCHED", "auto"), "world", world, "render_ms %.3f traverse_ms %.3f rays %d" % (best[0], best[2], best[1]), flush=True)
