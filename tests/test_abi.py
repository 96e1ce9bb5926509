"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/rt_b200.h
declares, fails loudly without a GPU, and the host object model flattens the reference scenes."""
import ctypes as C
import os
import re

import pytest

import raytrace_b200 as R
from raytrace_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 14
    lib = C.CDLL(os.path.join(_capi.LIB_DIR, "librt_b200.so"))
    for s in syms:
        assert hasattr(lib, s), f"librt_b200.so does not export {s}"
    assert set(syms) == set(_capi.RT_SYMBOLS), "ctypes table and header disagree"
    assert _capi.rt.rt_abi_version() == 2


def test_struct_layouts_match_the_header():
    # sizes the C compiler gives the ABI structs (checked against ctypes mirrors)
    assert C.sizeof(_capi.Vec4) == 16 and C.sizeof(_capi.Material) == 80 and C.sizeof(_capi.Light) == 96
    assert C.sizeof(_capi.Camera) == 96 and C.sizeof(_capi.Prim) == 96 and C.sizeof(_capi.Model) == 64
    assert C.sizeof(_capi.Part) == 48 and C.sizeof(_capi.HitId) == 20 and C.sizeof(_capi.RenderParams) == 32


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="box has a GPU")
def test_no_gpu_means_loud_failure_not_fallback():
    h = C.c_void_p()
    rc = _capi.rt.rt_create(0, C.byref(h))
    assert rc == -3 and not h.value          # RT_E_NODEVICE
    assert b"no CPU fallback" in _capi.rt.rt_last_error()
    sc = R.Scene("c1", 128, 128)
    rt = R.RayTracer(sc)
    with pytest.raises(R.RtError):
        rt.start()


def test_default_scene_flattens_like_the_reference():
    # numbers probed from the reference in SURVEY.md section 3.1
    sc = R.Scene("c1", 1088, 576)
    d = sc.flatten().contents
    assert (d.n_lights, d.n_prims, d.n_models, d.n_tris) == (2, 2, 0, 0)
    par, pt = d.lights[0], d.lights[1]
    assert par.type == 1 and abs(par.position.x - 8) < 1e-5 and abs(par.position.y - 11.3137083) < 1e-5 and abs(par.position.z - 8) < 1e-5
    assert pt.type == 2 and pt.position.w == 1.0 and abs(pt.position.z - 16) < 1e-6 and 0 < pt.position.y < 1e-14
    assert abs(pt.ambient.x - 38.4) < 1e-4 and abs(pt.diffuse.x - 140.8) < 1e-4 and abs(pt.specular.x - 76.8) < 1e-4
    plane, sphere = d.prims[0], d.prims[1]
    assert plane.kind == 4 and plane.texture == 0 and plane.a.y == 1.0 and abs(plane.a.z) < 2e-16
    assert sphere.kind == 1 and sphere.radius == 1.0 and (sphere.position.x, sphere.position.y, sphere.position.z) == (0.0, 1.0, 0.0)
    assert abs(d.materials[plane.material].reflect - 0.6) < 1e-7 and abs(d.materials[sphere.material].reflect - 0.35) < 1e-7
    assert d.camera.width == 1088 and d.camera.zFar == 100.0


def test_mesh_scene_flattens_with_parts_and_limits():
    sc = R.Scene("t_twomesh", 256, 128)
    d = sc.flatten().contents
    assert d.n_models == 2 and d.n_parts == 2 * 2 + 1
    assert d.n_tris == 32 * 32 * 2 + 10 * 16 * 2
    objs = [d.prims[i].object for i in range(d.n_prims)]
    assert objs == sorted(objs)
    for p in range(d.n_parts):
        assert d.parts[p].tri_count <= 32767
    # loader rescales every model to largest extent 8 (Model.cpp:185-197)
    m = d.models[1]
    ext = max(m.ver_max.x - m.ver_min.x, m.ver_max.y - m.ver_min.y, m.ver_max.z - m.ver_min.z)
    assert abs(ext - 8.0) < 1e-4
    # geometry epoch is stable while nothing changes, so re-uploads can skip the triangles
    e1 = d.geometry_epoch
    sc.move(R.MY_MODEL_OBJECT, 1, 0.5, 0, 0)
    assert sc.flatten().contents.geometry_epoch == e1


def test_ballplane_expands_to_16_spheres():
    sc = R.Scene("t_ballplane", 128, 128)
    d = sc.flatten().contents
    subs = [d.prims[i].sub for i in range(d.n_prims) if d.prims[i].object == 1]
    assert subs == list(range(1, 17))


def test_batch_and_shard_entry_points_reject_bad_arguments_without_touching_a_gpu():
    # the entry points added for frame batches and shard read-back validate before any CUDA call
    rt = _capi.rt
    p = R.RenderParams(R.MY_MODEL_RAYTRACE, 1, 0, 1, 0, 64)
    assert rt.rt_render_batch_async(None, C.byref(p), 2, None, None) != 0
    assert rt.rt_render_batch_async(None, C.byref(p), 0, None, None) != 0 and b"frames" in rt.rt_last_error()
    assert rt.rt_render_batch_async(None, C.byref(p), 65, None, None) != 0 and b"frames" in rt.rt_last_error()
    buf = (C.c_uint8 * 16)()
    assert rt.rt_read_batch_output(None, 0, buf, 12, 0) != 0
    assert rt.rt_read_output_rows(None, buf, 12) != 0
    assert rt.rt_push_batch_rows(None, 0, None, 1) != 0


def test_serpentine_tile_order_is_a_balanced_partition():
    # RT_FLAG_SERPENTINE as the host side states it (distributed.bands_of): every tile exactly once, tile counts
    # differ by at most one, and the tile-index sums (a linear cost gradient down the image) are equal
    # whenever the tile count is a multiple of 2 * world
    from raytrace_b200.distributed import bands_of
    for world, tile_rows, h in ((8, 8, 1080), (4, 8, 1080), (3, 16, 330), (2, 64, 576), (5, 8, 320)):
        n = (h // 64) * 64 // tile_rows
        parts = [bands_of(r, world, h, tile_rows, serpentine=True) for r in range(world)]
        assert sorted(t for p in parts for t in p) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
        if n % (2 * world) == 0:
            assert len({sum(p) for p in parts}) == 1
        assert [t // world for t in parts[0]] == list(range(len(parts[0])))      # one tile per group of `world`
