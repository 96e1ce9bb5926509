"""raytrace_b200 -- B200-native trace-and-shade path behind the XZiar/RayTrace object model.

Python is only a thin driver for tests and benchmarks: scenes are built by the C++ object model
(host/, same API as the reference's Scene/RayTracer), flattened to include/rt_b200.h and rendered
by hand-written sm_100a kernels (csrc/).  Nothing here computes a ray on the CPU.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import Camera, Counters, HitId, RenderParams, SceneDesc, rt, rth

# RayTracer::start `type` codes (RayTracer.h:5-13 of the reference)
MY_MODEL_CHECK, MY_MODEL_DEPTHTEST, MY_MODEL_NORMALTEST, MY_MODEL_TEXTURETEST = 1, 2, 3, 4
MY_MODEL_MATERIALTEST, MY_MODEL_SHADOWTEST, MY_MODEL_REFLECTTEST, MY_MODEL_REFRACTTEST = 5, 6, 7, 8
MY_MODEL_RAYTRACE = 0x80
MY_MODEL_LIGHT, MY_MODEL_OBJECT = 1, 2
RT_FLAG_HIT_IDS, RT_FLAG_STATS, RT_FLAG_BRUTE, RT_FLAG_COMBINE_LEVELS, RT_FLAG_SERPENTINE = 1, 2, 4, 8, 16

HIT_DTYPE = np.dtype([("object", "<i4"), ("sub", "<i4"), ("index", "<i4"), ("octant", "<i4"), ("distance", "<f4")])


class RtError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise RtError(f"{what} failed ({rc}): {rt.rt_last_error().decode(errors='replace')}")


def fnv1a64(buf) -> str:
    """FNV-1a-64 of a byte buffer -- the golden-hash convention of BASELINE.md."""
    a = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf).reshape(-1).view(np.uint8)
    return f"{rth.rth_fnv1a64(a.ctypes.data, a.size):016x}"


class Scene:
    """A Scene of the C++ object model (host/Scene.h), built by a named synthetic builder."""

    def __init__(self, name="c1", width=1088, height=576, n=0, parts=0, tmpdir="/tmp"):
        self._h = rth.rth_scene_new()
        self.name, self.width, self.height = name, width, height
        rc = rth.rth_scene_build(self._h, name.encode(), n, parts, width, height, tmpdir.encode())
        if rc != 0:
            raise RtError(rth.rth_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None) and rth is not None:   # (module globals are already gone at interpreter exit)
            rth.rth_scene_free(self._h)
            self._h = None

    def flatten(self) -> "C.POINTER(SceneDesc)":
        d = rth.rth_scene_flatten(self._h)
        if not d:
            raise RtError(rth.rth_last_error().decode())
        return d

    def resize(self, w, h):
        self.width, self.height = w, h
        rth.rth_scene_resize(self._h, w, h)

    def move(self, kind, num, x, y, z):
        return rth.rth_scene_move(self._h, kind, num, x, y, z)

    def set_object_position(self, num, x, y, z):
        return rth.rth_scene_set_object_position(self._h, num, x, y, z)

    def set_light_position(self, num, x, y, z, w=1.0):
        return rth.rth_scene_set_light_position(self._h, num, x, y, z, w)

    def switch(self, kind, num, show):
        return rth.rth_scene_switch(self._h, kind, num, int(show))

    def chgmtl(self, num, lib_index):
        return rth.rth_scene_chgmtl(self._h, num, lib_index)

    def camera_jitter(self, dx, dy):
        return rth.rth_scene_camera_jitter(self._h, dx, dy)

    def camera_n(self):
        v = (C.c_float * 4)()
        rth.rth_scene_camera_get_n(self._h, v)
        return tuple(v)

    def set_camera_n(self, xyzw):
        v = (C.c_float * 4)(*xyzw)
        rth.rth_scene_camera_set_n(self._h, v)

    def orbit_camera(self, k, K) -> Camera:
        """camera k of the K-camera orbit around the scene's current camera (scenes.h orbit_camera); camera 0 is the
        scene's own camera.  The Scene is not changed."""
        c = Camera()
        rth.rth_scene_orbit_camera(self._h, k, K, C.byref(c))
        return c

    def camera(self) -> Camera:
        c = Camera()
        rth.rth_scene_camera_get(self._h, C.byref(c))
        return c

    def set_camera_position(self, x, y, z):
        return rth.rth_scene_camera_set_position(self._h, x, y, z)

    def jittered_camera(self, dx, dy) -> Camera:
        """the sample camera of sub-pixel offset (dx, dy) (scenes.h jittered_camera); the Scene is not changed"""
        c = Camera()
        rth.rth_scene_jittered_camera(self._h, dx, dy, C.byref(c))
        return c

    def camera_move(self, x, y, z):
        return rth.rth_scene_camera_move(self._h, x, y, z)

    @property
    def object_count(self):
        return rth.rth_scene_object_count(self._h)


class RayTracer:
    """The drop-in render surface (host/RayTracer.h): start()/isFinish/useTime/output."""

    def __init__(self, scene: Scene, device=0):
        self.scene = scene
        self._h = rth.rth_tracer_new(scene._h, device)
        self.maxLevel = 1
        self.smShare = 0   # resident traversal CTAs per SM (0 = all 8); set when several tracers of one Scene run concurrently
        self.progressiveBands = 0   # display mode: the frame is rendered band by band, `output` fills in while !isFinish (host/RayTracer.h)
        self.coalesce = False   # throughput mode: frames of this Scene's tracers that wait together are rendered in one launch (host/RayTracer.h)

    def __del__(self):
        if getattr(self, "_h", None) and rth is not None:
            rth.rth_tracer_free(self._h)
            self._h = None

    def set_samples(self, table):
        """jittered supersampling: `table` = [(dx, dy), ...] sub-pixel offsets (host/RayTracer.h `samples`); [] = one sample"""
        flat = (C.c_float * (2 * len(table)))(*[v for dxy in table for v in dxy])
        rth.rth_tracer_set_samples(self._h, len(table), flat)

    def start(self, type=MY_MODEL_RAYTRACE, tnum=1, flags=0, rank=0, world=1, tile_rows=64):
        rth.rth_tracer_set_max_level(self._h, self.maxLevel)
        rth.rth_tracer_set_sm_share(self._h, self.smShare)
        rth.rth_tracer_set_flags(self._h, flags)
        rth.rth_tracer_set_coalesce(self._h, 1 if self.coalesce else 0)
        rth.rth_tracer_set_progressive(self._h, self.progressiveBands)
        rth.rth_tracer_set_shard(self._h, rank, world, tile_rows)
        if rth.rth_tracer_start(self._h, type, tnum) != 0:
            raise RtError(rth.rth_last_error().decode())

    @property
    def isFinish(self):
        return bool(rth.rth_tracer_is_finished(self._h))

    @property
    def useTime(self):
        return rth.rth_tracer_use_time(self._h)

    def wait(self):
        rth.rth_tracer_wait(self._h)

    @property
    def bandsDone(self):
        return rth.rth_tracer_bands_done(self._h)

    def peek_output(self) -> np.ndarray:
        """RayTracer::output as it stands right now, without waiting (progressive display reads it while !isFinish)"""
        w, h = rth.rth_tracer_width(self._h), rth.rth_tracer_height(self._h)
        return np.ctypeslib.as_array(rth.rth_tracer_output(self._h), shape=(h, w, 3)).copy()

    @property
    def failed(self):
        """the last frame ended with an error (isFinish is true, output keeps the previous frame); see lastError"""
        return bool(rth.rth_tracer_failed(self._h))

    @property
    def lastError(self):
        return rth.rth_tracer_last_error(self._h).decode(errors="replace")

    def output(self) -> np.ndarray:
        """RayTracer::output as an (H, W, 3) uint8 array, row 0 = bottom."""
        self.wait()
        if self.failed:
            raise RtError(f"frame failed: {self.lastError}")
        w, h = rth.rth_tracer_width(self._h), rth.rth_tracer_height(self._h)
        ptr = rth.rth_tracer_output(self._h)
        return np.ctypeslib.as_array(ptr, shape=(h, w, 3)).copy()

    def render(self, type=MY_MODEL_RAYTRACE, **kw) -> np.ndarray:
        self.start(type, **kw)
        return self.output()

    def hit_ids(self) -> np.ndarray:
        w, h = rth.rth_tracer_width(self._h), rth.rth_tracer_height(self._h)
        out = np.zeros(w * h, dtype=HIT_DTYPE)
        if rth.rth_tracer_read_hit_ids(self._h, out.ctypes.data_as(C.POINTER(HitId))) != 0:
            raise RtError(rt.rt_last_error().decode())
        return out.reshape(h, w)

    def counters(self) -> Counters:
        c = Counters()
        if rth.rth_tracer_read_counters(self._h, C.byref(c)) != 0:
            raise RtError(rt.rt_last_error().decode())
        return c

    def context(self):
        c = rth.rth_tracer_context(self._h)
        if not c:
            raise RtError(rth.rth_last_error().decode())
        return c


class Context:
    """Direct use of the C ABI (include/rt_b200.h) with a flattened scene description."""

    def __init__(self, device=0):
        h = C.c_void_p()
        _check(rt.rt_create(device, C.byref(h)), "rt_create")
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None) and rt is not None:
            rt.rt_destroy(self._h)
            self._h = None

    def set_stream(self, cuda_stream: int):
        _check(rt.rt_set_stream(self._h, C.c_void_p(cuda_stream)), "rt_set_stream")

    def upload(self, desc):
        _check(rt.rt_upload_scene(self._h, desc), "rt_upload_scene")

    def render_async(self, type=MY_MODEL_RAYTRACE, max_level=1, rank=0, world=1, flags=0, tile_rows=64, tile_first=0, tile_count=0):
        p = RenderParams(type, max_level, rank, world, flags, tile_rows, tile_first, tile_count)
        _check(rt.rt_render_async(self._h, C.byref(p)), "rt_render_async")

    def wait(self) -> float:
        s = C.c_double()
        _check(rt.rt_wait(self._h, C.byref(s)), "rt_wait")
        return s.value

    def read_output(self, width, height, out=None) -> np.ndarray:
        if out is None:
            out = np.empty((height, width, 3), dtype=np.uint8)
        _check(rt.rt_read_output(self._h, out.ctypes.data_as(C.c_void_p), width * 3), "rt_read_output")
        return out

    def read_output_into(self, host_ptr: int, stride: int):
        _check(rt.rt_read_output(self._h, C.c_void_p(host_ptr), stride), "rt_read_output")

    def output_device(self):
        p, n = C.c_void_p(), C.c_size_t()
        _check(rt.rt_output_device(self._h, C.byref(p), C.byref(n)), "rt_output_device")
        return p.value, n.value

    def set_output(self, device_ptr: int, nbytes: int):
        _check(rt.rt_set_output(self._h, C.c_void_p(device_ptr), nbytes), "rt_set_output")

    def hit_ids(self, width, height) -> np.ndarray:
        out = np.zeros(width * height, dtype=HIT_DTYPE)
        _check(rt.rt_read_hit_ids(self._h, out.ctypes.data_as(C.POINTER(HitId))), "rt_read_hit_ids")
        return out.reshape(height, width)

    def counters(self) -> Counters:
        c = Counters()
        _check(rt.rt_read_counters(self._h, C.byref(c)), "rt_read_counters")
        return c
