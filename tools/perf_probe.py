"""Per-stage timing / traversal statistics of one configuration (development aid)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytrace_b200 as R

CFG = {
    "c1": ("c1", 1088, 576, 1, 0, 0),
    "c2": ("c2", 1920, 1080, 5, 0, 0),
    "c3": ("c3", 1920, 1080, 5, 0, 0),
    "c4": ("c4", 3840, 2160, 8, 0, 0),
    "c3s": ("c3", 1920, 1080, 5, 240, 15),
}
names = sys.argv[1:] or ["c2", "c3"]
WORLD = int(os.environ.get("RT_PROBE_WORLD", "1"))   # render only shard 0 of WORLD (8-row tiles): the per-GPU share of a multi-GPU frame
SHARD = dict(rank=0, world=WORLD, tile_rows=8) if WORLD > 1 else {}
for nm in names:
    name, w, h, level, n, parts = CFG[nm]
    t0 = time.time()
    sc = R.Scene(name, w, h, n, parts)
    t_build = time.time() - t0
    rt = R.RayTracer(sc)
    rt.maxLevel = level
    img = rt.render(R.MY_MODEL_RAYTRACE, flags=R.RT_FLAG_STATS, **SHARD)
    import hashlib
    frame_hash = hashlib.sha1(img.tobytes()).hexdigest()[:16]
    cs = rt.counters()
    best = None
    for _ in range(5):
        t0 = time.time()
        rt.start(R.MY_MODEL_RAYTRACE, **SHARD)
        rt.wait()
        wall = time.time() - t0
        c = rt.counters()
        if best is None or c.render_ms < best[0].render_ms:
            best = (c, wall)
    c, wall = best
    total = c.primary + c.shadow + c.reflect + c.refract
    flops = cs.nodes_visited * 4 * 22 + cs.tri_tests * 47 + cs.prim_tests * 23
    print(json.dumps({"cfg": nm, "lib": os.path.basename(os.environ.get("RT_B200_LIBDIR", "lib")), "frame_sha1": frame_hash, "world": WORLD, "leaf": os.environ.get("RT_B200_LEAF_SIZE", "2"), "scene_s": round(t_build, 1), "rays": total,
                      "rays_per_px": round(total / max(c.primary, 1), 2),
                      "render_ms": round(c.render_ms, 3), "start_to_finish_ms": round(wall * 1e3, 3), "mrays_s": round(total / c.render_ms / 1e3, 1),
                      "trace_ms": round(c.trace_ms, 3), "shadow_ms": round(c.shadow_ms, 3), "shade_ms": round(c.shade_ms, 3), "other_ms": round(c.other_ms, 3),
                      "upload_ms": round(cs.upload_ms, 2), "build_ms": round(cs.build_ms, 2), "bvh_nodes": c.bvh_nodes, "bvh_depth": c.bvh_depth,
                      "nodes_per_ray": round(cs.nodes_visited / total, 1), "tris_per_ray": round(cs.tri_tests / total, 2),
                      "prims_per_ray": round(cs.prim_tests / total, 2),
                      "alg_tflops": round(flops / ((c.trace_ms + c.shadow_ms) * 1e-3) / 1e12, 2), "launches": c.launches}), flush=True)
