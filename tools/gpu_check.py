"""Ad-hoc GPU-vs-oracle report (development aid; the graded checks live in tests/)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import raytrace_b200 as R
from parity_util import compare_ids, compare_images, oracle_render

CASES = [
    ("c1", 1088, 576, 1, 0, 0, 0x80),
    ("c1", 1088, 576, 5, 0, 0, 0x80),
    ("t_mixed", 640, 384, 4, 0, 0, 0x80),
    ("t_mixed", 640, 384, 3, 0, 0, 7),
    ("t_ballplane", 640, 384, 4, 0, 0, 0x80),
    ("c2", 640, 384, 3, 8, 0, 0x80),
    ("c2", 960, 576, 5, 16, 0, 0x80),
    ("t_mesh", 640, 384, 3, 0, 0, 0x80),
    ("t_twomesh", 640, 384, 3, 0, 0, 0x80),
    ("c3", 640, 384, 5, 96, 6, 0x80),
    ("c4", 640, 384, 6, 96, 6, 0x80),
]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if c[0] in sys.argv[1:]]

for name, w, h, level, n, parts, typ in CASES:
    sc = R.Scene(name, w, h, n, parts)
    t0 = time.time()
    oimg, oids, ocnt = oracle_render(sc, level, typ)
    t_or = time.time() - t0
    rt = R.RayTracer(sc)
    rt.maxLevel = level
    gimg = rt.render(typ, flags=R.RT_FLAG_STATS)
    c = rt.counters()
    rt.start(typ)
    rt.wait()
    c2 = rt.counters()
    gids = None
    if typ == 0x80:
        rt.start(typ, flags=R.RT_FLAG_HIT_IDS)
        gids = rt.hit_ids()
    res = compare_images(gimg, oimg)
    idd = compare_ids(gids, oids) if gids is not None else None
    rays_g = (c.primary, c.shadow, c.reflect, c.refract)
    rays_o = (ocnt.primary, ocnt.shadow, ocnt.reflect, ocnt.refract)
    if typ == 7:
        rays_o = (ocnt.primary, ocnt.shadow - 0, ocnt.reflect, ocnt.refract)
    total = sum(rays_g)
    print(json.dumps({"case": [name, w, h, level, n, parts, typ], "hash_equal": R.fnv1a64(gimg) == R.fnv1a64(oimg), "img": res,
                      "id_diff": idd, "rays_gpu": rays_g, "rays_oracle": rays_o, "render_ms": round(c2.render_ms, 3),
                      "mrays_s": round(total / c2.render_ms / 1e3, 1), "oracle_s": round(t_or, 2), "build_ms": round(c.build_ms, 2),
                      "bvh_depth": c.bvh_depth, "nodes_per_ray": round(c.nodes_visited / max(total, 1), 1),
                      "tris_per_ray": round(c.tri_tests / max(total, 1), 2), "prims_per_ray": round(c.prim_tests / max(total, 1), 2)}), flush=True)
