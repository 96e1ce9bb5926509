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
              ("c4", 640, 384, 6, 96, 6), ("t_mixed", 320, 192, 8, 0, 0), ("t_textured", 640, 384, 3, 0, 0)]


def ck(rc):
    assert rc == 0, rt.rt_last_error().decode()


def counts(c):
    return (c.primary, c.shadow, c.reflect, c.refract)


@pytest.mark.parametrize("sched", ["waves", "waves-defer", "waves-steal", "waves-voted", "waves-bin", "frame"])
@pytest.mark.parametrize("genprimary", ["0", "1"])
@pytest.mark.parametrize("case", MESH_CASES, ids=lambda c: f"{c[0]}-{c[1]}x{c[2]}-l{c[3]}")
def test_forced_scheduler_matches_reference_golden(gpu_present, monkeypatch, case, genprimary, sched):
    if sched == "frame" and genprimary == "0":
        pytest.skip("k_frame always makes its own primary rays")
    scene, w, h, level, n, parts = case
    if sched.startswith("waves-"):
        # the Model walk of the wave kernels: deferred triangle tests (rt_defer.cuh) or the voted walk, whichever is not the default
        if genprimary == "0":
            pytest.skip("one primary-ray mode is enough for the non-default walk")
        if sched == "waves-bin":
            monkeypatch.setenv("RT_B200_BIN", "4")      # coherence binning of the secondary rays (rtk_bin_rays)
        else:
            monkeypatch.setenv("RT_B200_TRAV", sched[6:])
        sched = "waves"
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


@pytest.mark.parametrize("genprimary", ["0", "1", "1-defer", "1-steal", "1-voted", "1-bin"])
@pytest.mark.parametrize("case", [("t_mesh", 448, 320, 4, 0, 0), ("c3", 448, 320, 5, 96, 6), ("c4", 448, 320, 6, 96, 6)],
                         ids=lambda c: f"{c[0]}-l{c[3]}")
def test_wave_scheduler_batches_with_distinct_cameras(gpu_present, monkeypatch, case, genprimary):
    # what bench.py times: rt_render_batch_async under the wave kernels over a Model BVH, one camera per frame
    scene, w, h, level, n, parts = case
    if "-" in genprimary:
        genprimary, walk = genprimary.split("-")
        if walk == "bin":
            monkeypatch.setenv("RT_B200_BIN", "5")
        else:
            monkeypatch.setenv("RT_B200_TRAV", walk)
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


# ---- BASELINE configs[4]: jittered supersampling accumulated on the device ------------------------------

def oracle_supersampled(sc, level, table, **kw):
    """integer mean of len(table) oracle renders through the jittered sample cameras == the definition of an spp frame"""
    n0 = sc.camera_n()
    acc = None
    try:
        for dx, dy in table:
            sc.set_camera_n(n0)
            sc.camera_jitter(dx, dy)
            f = oracle_render(sc, level, want_ids=False, **kw)[0].astype(np.uint32)
            acc = f if acc is None else acc + f
    finally:
        sc.set_camera_n(n0)
    return (acc // len(table)).astype(np.uint8)


@pytest.mark.parametrize("scene,w,h,level,side", [("t_mesh", 320, 192, 3, 2), ("c4", 256, 192, 4, 4)])
def test_device_side_supersampling_equals_n_oracle_renders(gpu_present, monkeypatch, scene, w, h, level, side):
    from raytrace_b200.supersample import device_table
    table = device_table(side, seed=0)
    assert len(table) == side * side and all(0.0 <= dx < 1.0 and 0.0 <= dy < 1.0 for dx, dy in table)
    sc = R.Scene(scene, w, h, 48 if scene == "c4" else 0, 3 if scene == "c4" else 0)
    want = oracle_supersampled(sc, level, table)
    single = oracle_render(sc, level, want_ids=False)[0]
    assert not np.array_equal(want, single)                       # the jitter really moves the samples
    t = R.RayTracer(sc)
    t.maxLevel = level
    t.set_samples(table)
    got = t.render(R.MY_MODEL_RAYTRACE)
    assert np.array_equal(got, want)
    c = t.counters()
    assert c.primary == side * side * (w // 64 * 64) * (h // 64 * 64)
    # band by band (what an 8K frame does): the budget forces several launches per frame, same pixels, same ray totals
    monkeypatch.setenv("RT_B200_SS_PIXELS", str(side * side * (w // 64 * 64) * 64))
    sc2 = R.Scene(scene, w, h, 48 if scene == "c4" else 0, 3 if scene == "c4" else 0)
    t2 = R.RayTracer(sc2)
    t2.maxLevel = level
    t2.set_samples(table)
    assert np.array_equal(t2.render(R.MY_MODEL_RAYTRACE), want)
    c2 = t2.counters()
    assert (c2.primary, c2.shadow, c2.reflect, c2.refract) == (c.primary, c.shadow, c.reflect, c.refract)
    assert c2.launches > c.launches
    # a shard of the supersampled frame (what each GPU of a multi-GPU run renders): serpentine 8-row tiles, rank 1 of 3
    part = t2.render(R.MY_MODEL_RAYTRACE, flags=R.RT_FLAG_SERPENTINE, rank=1, world=3, tile_rows=8)
    from raytrace_b200.distributed import bands_of
    tiles = bands_of(1, 3, h, 8, serpentine=True)
    rows = [y for y in range(h // 64 * 64) if y // 8 in tiles]
    other = [y for y in range(h) if y // 8 not in tiles or y >= h // 64 * 64]
    assert np.array_equal(part[rows], want[rows]) and (part[other] == 127).all()
    # back to one sample per pixel on the same tracer
    t2.set_samples([])
    assert np.array_equal(t2.render(R.MY_MODEL_RAYTRACE), single)


def test_tile_window_renders_a_band_of_the_shard(gpu_present):
    # rt_render_params::tile_first / tile_count: the band mechanism under the supersampling loop, against the oracle
    w, h, level = 448, 320, 3
    sc = R.Scene("t_mesh", w, h)
    full = oracle_render(sc, level, want_ids=False)[0]
    ctx = R.Context()
    ctx.upload(sc.flatten())
    from raytrace_b200.distributed import bands_of
    for rank, world, tr, serp, first, count in ((0, 1, 64, False, 1, 2), (1, 3, 8, True, 2, 5), (2, 4, 16, False, 0, 1), (0, 2, 8, True, 17, 0)):
        flags = R.RT_FLAG_SERPENTINE if serp else 0
        ctx.render_async(R.MY_MODEL_RAYTRACE, level, rank, world, flags | R.RT_FLAG_HIT_IDS, tr, first, count)
        ctx.wait()
        img = ctx.read_output(w, h)
        ids = ctx.hit_ids(w, h)
        p = R.RenderParams(R.MY_MODEL_RAYTRACE, level, rank, world, flags, tr, first, count)
        import ctypes as C
        from parity_util import oracle_lib
        oimg = np.empty((h, w, 3), np.uint8)
        oids = np.zeros(w * h, R.HIT_DTYPE)
        oc = R.Counters()
        assert oracle_lib().rto_render(sc.flatten(), C.byref(p), oimg.ctypes.data, oids.ctypes.data, C.byref(oc), 8) == 0
        mine = bands_of(rank, world, h, tr, serpentine=serp)
        win = mine[first:first + count] if count else mine[first:]
        rows = [y for y in range(h) if y // tr in win]
        assert len(rows) == len(win) * tr
        assert np.array_equal(img[rows], full[rows]) and np.array_equal(img[rows], oimg[rows])
        assert compare_ids(ids[rows], oids.reshape(h, w)[rows]) == (0, 0)
        c = ctx.counters()
        assert (c.primary, c.shadow, c.reflect, c.refract) == (oc.primary, oc.shadow, oc.reflect, oc.refract)
        assert c.primary == len(rows) * (w // 64 * 64)
