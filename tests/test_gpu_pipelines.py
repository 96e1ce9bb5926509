"""GPU tests of the frame pipelines (several frames in flight over one resident scene) and of the
fused frame kernels (ray generation inside k_frame, one-launch combine)."""
import ctypes as C

import numpy as np
import pytest

import raytrace_b200 as R
from parity_util import oracle_render

pytestmark = pytest.mark.gpu
rt = R.rt


def ck(rc):
    assert rc == 0, rt.rt_last_error().decode()


def read(ctx, w, h):
    out = np.empty((h, w, 3), dtype=np.uint8)
    ck(rt.rt_wait(ctx, None))
    ck(rt.rt_read_output(ctx, out.ctypes.data_as(C.c_void_p), w * 3))
    return out


def test_shared_pipelines_render_the_parents_scene_concurrently(gpu_present):
    w, h, level = 512, 384, 4
    sc = R.Scene("t_mesh", w, h)
    oimg, _, oc = oracle_render(sc, level, want_ids=False)
    parent = C.c_void_p()
    ck(rt.rt_create(0, C.byref(parent)))
    ck(rt.rt_upload_scene(parent, sc.flatten()))
    pipes = []
    for share in (4, 2, 1, 0):
        p = C.c_void_p()
        ck(rt.rt_create_shared(parent, C.byref(p)))
        ck(rt.rt_set_sm_share(p, share))
        pipes.append(p)
    params = R.RenderParams(R.MY_MODEL_RAYTRACE, level, 0, 1, 0, 64)
    for rep in range(3):                       # 12 frames, 4 in flight at any time
        for p in pipes:
            ck(rt.rt_render_async(p, C.byref(params)))
    for p in pipes:
        assert np.array_equal(read(p, w, h), oimg)
        c = R.Counters()
        ck(rt.rt_read_counters(p, C.byref(c)))
        assert c.primary + c.shadow + c.reflect + c.refract == oc.primary + oc.shadow + oc.reflect + oc.refract
    # a scene edit uploaded through ANY pipeline reaches all of them (they alias the parent's tables)
    sc.move(R.MY_MODEL_OBJECT, 2, 0.4, 0.0, -0.6)       # moves the mesh: clTri + BVH rebuilt on the device
    ck(rt.rt_upload_scene(pipes[1], sc.flatten()))
    oimg2, _, _ = oracle_render(sc, level, want_ids=False)
    assert not np.array_equal(oimg2, oimg)
    for p in pipes:
        ck(rt.rt_render_async(p, C.byref(params)))
    for p in pipes:
        assert np.array_equal(read(p, w, h), oimg2)
    assert rt.rt_set_sm_share(pipes[0], 9) != 0
    for p in pipes:
        rt.rt_destroy(p)
    rt.rt_destroy(parent)


def test_tracers_of_one_scene_share_residency_and_overlap(gpu_present):
    # the reference's idiom for several views: one RayTracer per view over the same Scene
    w, h, level = 448, 320, 3
    sc = R.Scene("t_mixed", w, h)
    oimg, _, _ = oracle_render(sc, level, want_ids=False)
    tracers = [R.RayTracer(sc) for _ in range(3)]
    for t in tracers:
        t.maxLevel = level
        t.smShare = 2
    for rep in range(2):
        for t in tracers:
            t.wait()
            t.start(R.MY_MODEL_RAYTRACE)
    for t in tracers:
        assert np.array_equal(t.output(), oimg)
    # only the first start() uploaded the scene; the others found it resident
    assert tracers[0].counters().bvh_nodes == tracers[2].counters().bvh_nodes


@pytest.mark.parametrize("scene,level", [("t_mesh", 5), ("c4", 6), ("c2", 3), ("c3", 5), ("t_mixed", 4)])
def test_tree_walk_combine_equals_level_by_level_combine(gpu_present, scene, level):
    w, h = 384, 256
    sc = R.Scene(scene, w, h, {"c2": 6, "c4": 48, "c3": 48}.get(scene, 0), 3 if scene in ("c3", "c4") else 0)
    a = R.RayTracer(sc)
    a.maxLevel = level
    one = a.render(R.MY_MODEL_RAYTRACE)
    lv = a.render(R.MY_MODEL_RAYTRACE, flags=R.RT_FLAG_COMBINE_LEVELS)
    assert np.array_equal(one, lv)
    assert a.counters().launches > 3                   # the diagnostic flag really took the level-by-level path


def test_whole_frame_scheduler_launches_three_kernels(gpu_present, monkeypatch):
    monkeypatch.setenv("RT_B200_SCHED", "frame")
    sc = R.Scene("c3", 320, 256, 48, 3)                # mirror mesh + plane: no refraction, every ray tree is a chain
    t = R.RayTracer(sc)
    t.maxLevel = 5
    img = t.render(R.MY_MODEL_RAYTRACE)
    c = t.counters()
    assert c.frame_sched == 1 and c.launches == 3      # k_frame (makes its own primary rays), k_shade, k_resolve
    oimg, _, _ = oracle_render(sc, 5, want_ids=False)
    assert np.array_equal(img, oimg)


@pytest.mark.parametrize("scene,level", [("t_mesh", 4), ("t_mixed", 3)])
def test_frame_batches_equal_single_frames(gpu_present, scene, level):
    # rt_render_batch_async: B frames with B cameras share one launch and one set of ray queues; every frame must
    # equal the frame rendered alone through that camera (== the oracle), whole and as a shard
    w, h = 448, 320
    sc = R.Scene(scene, w, h)
    cams = (R.Camera * 3)()
    singles, rays = [], 0
    for f, mv in enumerate(((0.0, 0.0, 0.0), (0.6, -0.3, 1.0), (-0.8, 0.4, 0.5))):
        sc.camera_move(*mv)
        cams[f] = sc.flatten().contents.camera
        oimg, _, oc = oracle_render(sc, level, want_ids=False)
        singles.append(oimg)
        rays += oc.primary + oc.shadow + oc.reflect + oc.refract
    assert not np.array_equal(singles[0], singles[1])
    ctx = C.c_void_p()
    ck(rt.rt_create(0, C.byref(ctx)))
    ck(rt.rt_upload_scene(ctx, sc.flatten()))
    params = R.RenderParams(R.MY_MODEL_RAYTRACE, level, 0, 1, 0, 64)
    for rep in range(2):                     # twice: the second batch reuses greyed buffers and zeroed hit lists
        ck(rt.rt_render_batch_async(ctx, C.byref(params), 3, cams, None))
        ck(rt.rt_wait(ctx, None))
        for f in range(3):
            out = np.empty((h, w, 3), dtype=np.uint8)
            ck(rt.rt_read_batch_output(ctx, f, out.ctypes.data_as(C.c_void_p), w * 3, 0))
            assert np.array_equal(out, singles[f]), (rep, f)
        c = R.Counters()
        ck(rt.rt_read_counters(ctx, C.byref(c)))
        assert c.primary + c.shadow + c.reflect + c.refract == rays
    # a batch of shards (serpentine 8-row tiles, rank 1 of 3): rows of the shard equal the single frames, the rest stays 127
    from raytrace_b200.distributed import bands_of
    sparams = R.RenderParams(R.MY_MODEL_RAYTRACE, level, 1, 3, R.RT_FLAG_SERPENTINE, 8)
    ck(rt.rt_render_batch_async(ctx, C.byref(sparams), 3, cams, None))
    tiles = bands_of(1, 3, h, 8, serpentine=True)
    rows = [y for y in range(h) if y // 8 in tiles]
    other = [y for y in range(h) if y // 8 not in tiles]
    for f in range(3):
        out = np.empty((h, w, 3), dtype=np.uint8)
        ck(rt.rt_read_batch_output(ctx, f, out.ctypes.data_as(C.c_void_p), w * 3, 0))
        assert np.array_equal(out[rows], singles[f][rows]) and (out[other] == 127).all()
        part = np.full((h, w, 3), 99, np.uint8)
        ck(rt.rt_read_batch_output(ctx, f, part.ctypes.data_as(C.c_void_p), w * 3, 1))
        assert np.array_equal(part[rows], singles[f][rows]) and (part[other] == 99).all()
    # a single frame after a batch still works (framebuffer and fill bookkeeping)
    ck(rt.rt_render_async(ctx, C.byref(params)))
    assert np.array_equal(read(ctx, w, h), singles[2])
    # limits
    assert rt.rt_render_batch_async(ctx, C.byref(params), 65, None, None) != 0
    hid = R.RenderParams(R.MY_MODEL_RAYTRACE, level, 0, 1, R.RT_FLAG_HIT_IDS, 64)
    assert rt.rt_render_batch_async(ctx, C.byref(hid), 2, None, None) != 0
    rt.rt_destroy(ctx)


def test_coalescing_tracers_render_their_own_frames(gpu_present):
    # RayTracer::coalesce: start() calls of several tracers of one Scene are rendered together by the Scene's
    # batch workers (one launch, shared ray queues), each through the camera its start() saw; every tracer
    # still gets exactly its own frame, and start()/isFinish/output behave as before
    w, h, level = 384, 256, 3
    sc = R.Scene("t_mesh", w, h)
    twin = R.Scene("t_mesh", w, h)            # the same camera path, walked ahead of time for the oracle frames
    moves = [(0.1 * (k % 6 + 1), -0.05 * (k % 6 + 1), 0.2) for k in range(18)]
    expect = []
    for mv in moves:
        twin.camera_move(*mv)
        expect.append(oracle_render(twin, level, want_ids=False)[0])
    tracers = []
    for _ in range(6):
        t = R.RayTracer(sc)
        t.maxLevel = level
        t.coalesce = True
        tracers.append(t)
    for k, mv in enumerate(moves):            # 18 start() calls back to back: frames pile up and share launches
        t = tracers[k % 6]
        t.wait()
        if k >= 6:
            assert np.array_equal(t.output(), expect[k - 6])       # the tracer's previous frame, before it is overwritten
        sc.camera_move(*mv)                   # every start() sees another camera
        t.start(R.MY_MODEL_RAYTRACE)
    for k, t in enumerate(tracers):
        t.wait()
        assert t.isFinish and t.useTime > 0
        assert np.array_equal(t.output(), expect[12 + k])
    # a shard through the coalescing path (rows-only read-back from the second frame on) and a tracer that opts out
    t = tracers[0]
    for rep in range(2):
        img = t.render(R.MY_MODEL_RAYTRACE, flags=R.RT_FLAG_SERPENTINE, rank=1, world=3, tile_rows=8)
        oimg, _, _ = oracle_render(sc, level, want_ids=False, rank=1, world=3, tile_rows=8, flags=R.RT_FLAG_SERPENTINE)
        assert np.array_equal(img, oimg)
    t.coalesce = False
    img = t.render(R.MY_MODEL_RAYTRACE)
    assert np.array_equal(img, oracle_render(sc, level, want_ids=False)[0])
    ids_img = tracers[1].render(R.MY_MODEL_RAYTRACE, flags=R.RT_FLAG_HIT_IDS)     # taps take the tracer's own pipeline
    assert np.array_equal(ids_img, img) and tracers[1].hit_ids().shape == (h, w)


@pytest.mark.parametrize("sched", ["waves", "frame"])
def test_overflowing_ray_levels_regrow_and_rerender(gpu_present, monkeypatch, sched):
    # with refraction a level can hold more rays than levelFactor x pixels.  The queues are deliberately
    # undersized here (factor 0.2: the camera sits inside a glass sphere, every pixel reflects AND refracts); the
    # library must regrow them from the counts of the overflowing frame and render it again, not fail or abort.
    monkeypatch.setenv("RT_B200_LEVEL_FACTOR", "0.2")
    monkeypatch.setenv("RT_B200_SCHED", sched)
    sc = R.Scene("t_inside", 384, 256)
    t = R.RayTracer(sc)
    t.maxLevel = 6
    img = t.render(R.MY_MODEL_RAYTRACE)
    assert not t.failed
    oimg, _, oc = oracle_render(sc, 6, want_ids=False)
    assert np.array_equal(img, oimg)
    c = t.counters()
    assert (c.primary, c.shadow, c.reflect, c.refract) == (oc.primary, oc.shadow, oc.reflect, oc.refract)
    assert oc.reflect + oc.refract > 0.2 * oc.primary * 2          # the undersized queues really overflowed
    assert np.array_equal(t.render(R.MY_MODEL_RAYTRACE), oimg)      # the learnt capacities stay: no second regrow needed


def test_stop_of_a_coalescing_tracer_leaves_the_others_alone(gpu_present):
    # stop() of one coalescing tracer must cancel only its own request -- never a launch that also
    # holds other tracers' frames
    w, h, level = 384, 256, 4
    sc = R.Scene("t_mesh", w, h)
    oimg = oracle_render(sc, level, want_ids=False)[0]
    tracers = []
    for _ in range(6):
        t = R.RayTracer(sc)
        t.maxLevel = level
        t.coalesce = True
        tracers.append(t)
    for rep in range(4):
        for t in tracers:
            t.start(R.MY_MODEL_RAYTRACE)
        R.rth.rth_tracer_stop(tracers[rep % 6]._h)           # cancelled while waiting, or delivered if already in a launch
        R.rth.rth_tracer_stop(tracers[(rep + 3) % 6]._h)
        for k, t in enumerate(tracers):
            t.wait()
            assert t.isFinish and not t.failed
            if k not in (rep % 6, (rep + 3) % 6):
                assert np.array_equal(t.output(), oimg), (rep, k)
    for t in tracers:                                        # every tracer, the stopped ones too, renders correctly afterwards
        assert np.array_equal(t.render(R.MY_MODEL_RAYTRACE), oimg)


def test_moving_a_model_refits_its_bvh_instead_of_rebuilding(gpu_present):
    # f-3: Scene::MovePos of a Model only re-translates bounds in the reference (Model.cpp:404,418-419); here the 4-wide
    # tree keeps its topology and is refitted level by level from the new triangle boxes -- same frames as a rebuild
    sc = R.Scene("t_twomesh", 448, 320)
    t = R.RayTracer(sc)
    t.maxLevel = 3
    img = t.render(R.MY_MODEL_RAYTRACE)
    c0 = t.counters()
    assert c0.bvh_refit == 0 and np.array_equal(img, oracle_render(sc, 3, want_ids=False)[0])
    for k, (obj, mv) in enumerate(((1, (0.7, 0.3, -1.1)), (3, (2.5, -1.0, 3.0)), (1, (-6.0, 0.2, 4.0)), (3, (0.001, 0.0, 0.0)))):
        sc.move(R.MY_MODEL_OBJECT, obj, *mv)               # objects 1 and 3 are the two meshes
        img = t.render(R.MY_MODEL_RAYTRACE, flags=R.RT_FLAG_HIT_IDS)
        ids, c = t.hit_ids(), t.counters()
        oimg, oids, oc = oracle_render(sc, 3)
        assert np.array_equal(img, oimg), k
        from parity_util import compare_ids
        assert compare_ids(ids, oids) == (0, 0)
        assert (c.primary, c.shadow, c.reflect, c.refract) == (oc.primary, oc.shadow, oc.reflect, oc.refract)
        assert c.bvh_refit == 1 and c.bvh_nodes == c0.bvh_nodes, k
    # the refitted tree against brute force (every primitive, no BVH): nothing was lost
    b = t.render(R.MY_MODEL_RAYTRACE, flags=R.RT_FLAG_BRUTE)
    assert np.array_equal(b, img)
    # an edit that is not a pure move rebuilds: hiding the box changes the object list
    sc.switch(R.MY_MODEL_OBJECT, 2, False)
    img = t.render(R.MY_MODEL_RAYTRACE)
    assert t.counters().bvh_refit == 0 and np.array_equal(img, oracle_render(sc, 3, want_ids=False)[0])


def test_progressive_output_fills_in_band_by_band(gpu_present):
    # f-3: main.cpp:207-208,246-254 blits RayTracer::output while !isFinish.  progressiveBands = k renders the frame as k
    # bands of its row tiles; a band's rows are in `output` -- final -- as soon as bandsDone counts it
    w, h, level = 1920, 1088, 5
    sc = R.Scene("c3", w, h, 240, 15)
    t = R.RayTracer(sc)
    t.maxLevel = level
    final = t.render(R.MY_MODEL_RAYTRACE)                      # one launch
    t.progressiveBands = 17                                   # 17 tiles of 64 rows: one per band
    seen = []
    t.start(R.MY_MODEL_RAYTRACE)
    while not t.isFinish:
        b = t.bandsDone
        snap = t.peek_output()
        seen.append((b, snap))
        if len(seen) > 4000:
            break
    t.wait()
    assert not t.failed and t.bandsDone == 17
    assert np.array_equal(t.output(), final)                  # same pixels as the single launch
    partial = [(b, s) for b, s in seen if 0 < b < 17]
    assert partial, "the frame finished before a single partial state could be observed"
    for b, snap in partial:
        assert np.array_equal(snap[:b * 64], final[:b * 64])  # completed bands are final ...
    b, snap = seen[0]
    assert (snap[(b + 1) * 64:] == 127).all()                 # ... the bands not started yet are still grey
    # a sharded progressive frame (serpentine 8-row tiles): still equal to the oracle's shard
    t.progressiveBands = 5
    part = t.render(R.MY_MODEL_RAYTRACE, flags=R.RT_FLAG_SERPENTINE, rank=1, world=3, tile_rows=8)
    opart = oracle_render(sc, level, want_ids=False, rank=1, world=3, tile_rows=8, flags=R.RT_FLAG_SERPENTINE)[0]
    assert np.array_equal(part, opart)
    t.progressiveBands = 0
    assert np.array_equal(t.render(R.MY_MODEL_RAYTRACE), final)
