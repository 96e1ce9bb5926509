import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import raytrace_b200 as R
from parity_util import oracle_render
for scene, level, n, parts in [("t_mesh", 5, 0, 0), ("t_mesh", 2, 0, 0), ("c4", 6, 48, 3)]:
    sc = R.Scene(scene, 384, 256, n, parts)
    o, _, _ = oracle_render(sc, level, want_ids=False)
    a = R.RayTracer(sc); a.maxLevel = level
    one = a.render(R.MY_MODEL_RAYTRACE)
    lv = a.render(R.MY_MODEL_RAYTRACE, flags=R.RT_FLAG_COMBINE_LEVELS)
    one2 = a.render(R.MY_MODEL_RAYTRACE)
    b = R.RayTracer(sc); b.maxLevel = level
    lvb = b.render(R.MY_MODEL_RAYTRACE, flags=R.RT_FLAG_COMBINE_LEVELS)
    print(scene, level, "resolve==oracle", np.array_equal(one, o), "levels==oracle", np.array_equal(lv, o), "resolve again", np.array_equal(one2, o),
          "fresh levels==oracle", np.array_equal(lvb, o), "ndiff", int((lv != o).any(axis=2).sum()), int((lvb != o).any(axis=2).sum()))
    d = np.argwhere((lv != o).any(axis=2))
    if len(d): print(" first diffs", d[:5].tolist(), lv[tuple(d[0])], o[tuple(d[0])])
