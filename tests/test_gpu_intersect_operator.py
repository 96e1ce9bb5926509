"""GPU tests of boundary B2: the per-primitive operator DrawObject::intersect(ray, hr, min)
(/root/reference/3DElement.h:201) evaluated on the device (rt_intersect_object) against the oracle's
restatement of Sphere/Box/Plane/BallPlane/Model::intersect, field by field, bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

import raytrace_b200 as R
from raytrace_b200._capi import Hit, HitId, Ray, Vec4, rt, rth
from parity_util import oracle_lib

pytestmark = pytest.mark.gpu


def _oracle_intersect():
    lib = oracle_lib()
    lib.rto_intersect_object.restype = C.c_int
    lib.rto_intersect_object.argtypes = [C.POINTER(R.SceneDesc), C.c_uint32, C.POINTER(Ray), C.POINTER(Hit), C.c_float, C.POINTER(Hit), C.c_uint32]
    return lib.rto_intersect_object


def _rays(n, seed, origin=(0.0, 4.0, 15.0), spread=1.0):
    rng = np.random.RandomState(seed)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[:, 2] = -np.abs(d[:, 2]) - spread          # mostly towards the scene
    d[:, 1] -= 0.3
    d /= np.sqrt((d.astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(np.float32)
    rays = (Ray * n)()
    for i in range(n):
        o = np.float32(origin) + rng.normal(size=3).astype(np.float32) * np.float32(0.5)
        rays[i].origin = Vec4(*[float(x) for x in o], 0.0)
        rays[i].direction = Vec4(float(d[i, 0]), float(d[i, 1]), float(d[i, 2]), 0.0)
        rays[i].mtlrfr, rays[i].type, rays[i].is_inside = 1.0, 1, 0
    return rays


def _fresh_hits(n, distance=1e20):
    hits = (Hit * n)()
    for i in range(n):
        hits[i].material = hits[i].texture = -1
        hits[i].id = HitId(-1, -1, -1, -1, distance)
        hits[i].rfr = 1.0
    return hits


def _same(a, b):
    # bit-exact, except that a NaN is a NaN (x86 0/0 gives 0xFFC00000, CUDA 0x7FFFFFFF): Box normals on
    # edge hits are 0/0 in the reference (Basic3DObject.cpp:287-293)
    ua, ub = np.frombuffer(bytes(a), np.uint32).copy(), np.frombuffer(bytes(b), np.uint32).copy()
    fa, fb = ua.view(np.float32), ub.view(np.float32)
    both = np.isnan(fa) & np.isnan(fb)
    # only the float fields can hold NaN patterns that matter; integer fields never alias a NaN here
    ua[both] = 0
    ub[both] = 0
    return np.array_equal(ua, ub)


@pytest.mark.parametrize("scene,n,parts", [("t_mixed", 0, 0), ("t_ballplane", 0, 0), ("t_mesh", 0, 0), ("t_twomesh", 0, 0), ("t_inside", 0, 0)])
def test_device_operator_matches_oracle(gpu_present, scene, n, parts):
    sc = R.Scene(scene, 256, 192, n, parts)
    desc = sc.flatten()
    ctx = R.Context()
    ctx.upload(desc)
    ora = _oracle_intersect()
    N = 3000
    nobj = sc.object_count
    for obj in range(nobj):
        # (a) fresh HitRes, closest-hit semantics
        rays, hin = _rays(N, 100 + obj), _fresh_hits(N)
        g, o = (Hit * N)(), (Hit * N)()
        assert rt.rt_intersect_object(ctx._h, obj, rays, hin, 0.0, g, N) == 0, rt.rt_last_error()
        assert ora(desc, obj, rays, hin, 0.0, o, N) == 0
        assert _same(g, o), f"object {obj}: primary operator differs"
        nhit = sum(1 for h in g if h.id.object >= 0)
        # (b) second generation: leave every hit point again with hr.obj = that primitive (self-skip,
        # octant copies, inside-sphere exit), as reflect / refract / shadow rays do
        rng = np.random.RandomState(7 + obj)
        rays2, hin2 = (Ray * N)(), _fresh_hits(N)
        for i in range(N):
            h = g[i]
            src = h if h.id.object >= 0 else None
            d = rng.normal(size=3).astype(np.float32)
            d /= np.float32(np.sqrt(float((d.astype(np.float64) ** 2).sum())))
            if src is not None:
                rays2[i].origin = src.position
                hin2[i].id = HitId(src.id.object, src.id.sub, src.id.index, src.id.octant, 1e20)
            else:
                rays2[i].origin = rays[i].origin
            rays2[i].direction = Vec4(float(d[0]), float(d[1]), float(d[2]), 0.0)
            rays2[i].mtlrfr = 1.5 if i % 3 == 0 else 1.0
            rays2[i].type = 4 if i % 3 == 0 else (3 if i % 3 == 1 else 2)
            rays2[i].is_inside = 0xFF if (i % 3 == 0 and src is not None) else 0
        for target in range(nobj):
            g2, o2 = (Hit * N)(), (Hit * N)()
            assert rt.rt_intersect_object(ctx._h, target, rays2, hin2, 0.0, g2, N) == 0
            assert ora(desc, target, rays2, hin2, 0.0, o2, N) == 0
            assert _same(g2, o2), f"object {obj} -> {target}: secondary operator differs"
        # (c) shadow-style call: hr.distance = light distance, min = the same (any-hit early exit)
        hin3 = _fresh_hits(N, 6.0)
        g3, o3 = (Hit * N)(), (Hit * N)()
        assert rt.rt_intersect_object(ctx._h, obj, rays, hin3, 6.0, g3, N) == 0
        assert ora(desc, obj, rays, hin3, 6.0, o3, N) == 0
        assert _same(g3, o3), f"object {obj}: any-hit operator differs"
    assert nhit >= 0


def test_host_class_operator_forwards_to_the_device(gpu_present):
    # scene.Objects[i]->intersect(ray, hr) through the C++ classes == the ABI call
    sc = R.Scene("t_mixed", 256, 192)
    tr = R.RayTracer(sc)              # the Scene must be attached to a RayTracer (its device context)
    ctx_h = C.c_void_p(tr.context())
    desc = sc.flatten()
    assert rt.rt_upload_scene(ctx_h, desc) == 0
    N = 200
    rays = _rays(N, 5)
    for obj in range(sc.object_count):
        hin = _fresh_hits(N)
        g = (Hit * N)()
        assert rt.rt_intersect_object(ctx_h, obj, rays, hin, 0.0, g, N) == 0
        for i in range(0, N, 7):
            h = Hit()
            h.material = h.texture = -1
            h.id = HitId(-1, -1, -1, -1, 1e20)
            assert rth.rth_object_intersect(sc._h, obj, C.byref(rays[i]), C.byref(h), 0.0) == 0, rth.rth_last_error()
            assert h.id.distance == g[i].id.distance
            if g[i].id.object >= 0:
                assert bytes(h.position) == bytes(g[i].position) and bytes(h.normal) == bytes(g[i].normal)
                assert (h.tu, h.tv, h.rfr, h.is_inside) == (g[i].tu, g[i].tv, g[i].rfr, g[i].is_inside)


def test_operator_without_a_tracer_fails_loudly(gpu_present):
    sc = R.Scene("c1", 128, 128)      # no RayTracer attached: nowhere to evaluate the operator
    rays = _rays(1, 1)
    h = Hit()
    h.id = HitId(-1, -1, -1, -1, 1e20)
    assert rth.rth_object_intersect(sc._h, 1, C.byref(rays[0]), C.byref(h), 0.0) != 0
    assert b"no CPU implementation" in rth.rth_last_error()
