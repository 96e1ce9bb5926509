"""GPU parity tests of the kernel combination bench.py actually times:
per-level wave kernels (RT_B200_SCHED=waves) over a Model BVH, primary rays made inside k_wave(0) or by
k_raygen, single frames and batches of frames with DISTINCT cameras -- every frame against the CPU oracle and,
where the golden set holds the case, against the unmodified reference's frame hash.  Plus the full-size
C4 band (3840x2160, 4 147 200 triangles, glass + mirror spheres, depth 8)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import raytrace_b200 as R
from parity_util import compare_ids, oracle_render

pytestmark = pytest.mark.gpu
rt = R.rt
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = {(c["scene"], c["w"], c["h"], c["level"], c["n"], c["parts"], c["type"]): c
          for c in json.load(open(os.path.join(HERE, "golden", "golden.json")))["cases"]}

# (scene, w, h, level, n, parts): the mesh goldens + the deep glass case
MESH_CASES = [("t_mesh", 640, 384, 3, 0, 0), ("t_twomesh", 640, 384, 3, 0, 0), ("c3", 640, 384, 5, 96, 6),
              ("c4", 640, 384, 6, 96, 6), ("t_mixed", 320, 192, 8, 0, 0)]


def ck(rc):
    assert rc == 0, rt.rt_last_error().decode()


def counts(c):
    return (c.primary, c.shadow, c.reflect, c.refract)


@pytest.mark.parametrize("sched", ["waves", "frame"])
@pytest.mark.parametrize("genprimary", ["0", "1"])
@pytest.mark.parametrize("case", MESH_CASES, ids=lambda c: f"{c[0]}-{c[1]}x{c[2]}-l{c[3]}")
def test_forced_scheduler_matches_reference_golden(gpu_present, monkeypatch, case, genprimary, sched):
    if sched == "frame" and genprimary == "0":
        pytest.skip("k_frame always makes its own primary rays")
    scene, w, h, level, n, parts = case
    monkeypatch.setenv("RT_B200_SCHED", sched)
    monkeypatch.setenv("RT_B200_WAVE_GENPRIMARY", genprimary)
    sc = R.Scene(scene, w, h, n, parts)
    t = R.RayTracer(sc)
    t.maxLevel = level
    img = t.render(R.MY_MODEL_RAYTRACE, flags=R.RT_FLAG_HIT_IDS)
    ids, c = t.hit_ids(), t.counters()
    assert c.frame_sched == (1 if sched == "frame" else 0)       # the forced scheduler really ran
    oimg, oids, oc = oracle_render(sc, level)
    assert np.array_equal(img, oimg)
    assert compare_ids(ids, oids) == (0, 0)
    assert counts(c) == counts(oc)
    g = GOLDEN[(scene, w, h, level, n, parts, 0x80)]
    assert R.fnv1a64(img) == g["hash"] and R.fnv1a64(ids) == g["ids_hash"]      # == the unmodified reference


@pytest.mark.parametrize("genprimary", ["0", "1"])
@pytest.mark.parametrize("case", [("t_mesh", 448, 320, 4, 0, 0), ("c3", 448, 320, 5, 96, 6), ("c4", 448, 320, 6, 96, 6)],
                         ids=lambda c: f"{c[0]}-l{c[3]}")
def test_wave_scheduler_batches_with_distinct_cameras(gpu_present, monkeypatch, case, genprimary):
    # what bench.py times: rt_render_batch_async under the wave kernels over a Model BVH, one camera per frame
    scene, w, h, level, n, parts = case
    monkeypatch.setenv("RT_B200_SCHED", "waves")
    monkeypatch.setenv("RT_B200_WAVE_GENPRIMARY", genprimary)
    sc = R.Scene(scene, w, h, n, parts)
    cams = (R.Camera * 3)()
    singles, rays = [], 0
    for f, mv in enumerate(((0.0, 0.0, 0.0), (0.5, -0.3, 0.8), (-0.7, 0.4, 0.4))):
        sc.camera_move(*mv)
        cams[f] = sc.flatten().contents.camera
        oimg, _, oc = oracle_render(sc, level, want_ids=False)
        singles.append(oimg)
        rays += sum(counts(oc))
    assert not np.array_equal(singles[0], singles[1]) and not np.array_equal(singles[1], singles[2])
    ctx = C.c_void_p()
    ck(rt.rt_create(0, C.byref(ctx)))
    ck(rt.rt_upload_scene(ctx, sc.flatten()))
    for params in (R.RenderParams(R.MY_MODEL_RAYTRACE, level, 0, 1, 0, 64),):
        for rep in range(2):
            ck(rt.rt_render_batch_async(ctx, C.byref(params), 3, cams, None))
            ck(rt.rt_wait(ctx, None))
            c = R.Counters()
            ck(rt.rt_read_counters(ctx, C.byref(c)))
            assert c.frame_sched == 0
            assert sum(counts(c)) == rays
            for f in range(3):
                out = np.empty((h, w, 3), dtype=np.uint8)
                ck(rt.rt_read_batch_output(ctx, f, out.ctypes.data_as(C.c_void_p), w * 3, 0))
                assert np.array_equal(out, singles[f]), (rep, f)
    # sharded batch (what every rank of a multi-GPU bench run renders): serpentine 8-row tiles, rank 2 of 4
    from raytrace_b200.distributed import bands_of
    sparams = R.RenderParams(R.MY_MODEL_RAYTRACE, level, 2, 4, R.RT_FLAG_SERPENTINE, 8)
    ck(rt.rt_render_batch_async(ctx, C.byref(sparams), 3, cams, None))
    tiles = bands_of(2, 4, h, 8, serpentine=True)
    rows = [y for y in range(h // 64 * 64) if y // 8 in tiles]
    for f in range(3):
        out = np.empty((h, w, 3), dtype=np.uint8)
        ck(rt.rt_read_batch_output(ctx, f, out.ctypes.data_as(C.c_void_p), w * 3, 0))
        assert np.array_equal(out[rows], singles[f][rows])
    rt.rt_destroy(ctx)


def test_full_size_c4_band_against_oracle(gpu_present):
    # BASELINE config 4 at full size: 3840x2160, 4 147 200-triangle Model + 64 glass + 6 mirror spheres, depth 8
    # (binary ray trees, Beer's law, k_combine over ~74 M rays).  Four 8-row tiles spread down the frame
    # (rank 29 of 66) through the oracle, bit-exact; the whole frame through size-independent properties.
    sc = R.Scene("c4", 3840, 2160)
    t = R.RayTracer(sc)
    t.maxLevel = 8
    full = t.render(R.MY_MODEL_RAYTRACE)
    cf = t.counters()
    assert (full[2112:] == 127).all()
    assert cf.refract > 0 and cf.reflect > 0 and cf.frame_sched == 0      # 8 M pixels: the wave kernels
    part = t.render(R.MY_MODEL_RAYTRACE, flags=R.RT_FLAG_HIT_IDS, rank=29, world=66, tile_rows=8)
    pids, cp = t.hit_ids(), t.counters()
    opart, oids, oc = oracle_render(sc, 8, rank=29, world=66, tile_rows=8)
    assert np.array_equal(part, opart)
    assert compare_ids(pids, oids) == (0, 0)
    assert counts(cp) == counts(oc)
    rows = [y for y in range(2112) if (y // 8) % 66 == 29]
    assert len(rows) == 32 and np.array_equal(full[rows], part[rows])
    again = t.render(R.MY_MODEL_RAYTRACE)
    assert np.array_equal(again, full)                                       # deterministic at 74 M rays
