"""ctypes bindings of the two in-tree shared libraries.

  librt_b200.so  the C ABI of include/rt_b200.h (CUDA kernels, sm_100a)
  librt_host.so  the C++ object model (Scene/Model/RayTracer, mirror of the reference API)
                 behind the small rth_* C shim of host/capi.cpp

There is no Python implementation of any of this and no CPU fallback: if the libraries are
missing the import fails with instructions, and on a box without a B200 ``rt_create`` fails.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.environ.get("RT_B200_LIBDIR") or os.path.join(_HERE, "lib")   # RT_B200_LIBDIR: A/B builds during development


class Vec4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class Material(C.Structure):
    _fields_ = [("ambient", Vec4), ("diffuse", Vec4), ("specular", Vec4), ("emission", Vec4),
                ("shiness", C.c_float), ("reflect", C.c_float), ("refract", C.c_float), ("rfr", C.c_float)]


class Light(C.Structure):
    _fields_ = [("position", Vec4), ("ambient", Vec4), ("diffuse", Vec4), ("specular", Vec4), ("attenuation", Vec4),
                ("type", C.c_uint32), ("enabled", C.c_uint32), ("pad0", C.c_uint32), ("pad1", C.c_uint32)]


class Camera(C.Structure):
    _fields_ = [("u", Vec4), ("v", Vec4), ("n", Vec4), ("position", Vec4),
                ("width", C.c_int32), ("height", C.c_int32),
                ("fovy", C.c_float), ("zNear", C.c_float), ("zFar", C.c_float),
                ("pad0", C.c_uint32), ("pad1", C.c_uint32), ("pad2", C.c_uint32)]


class Texture(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("offset", C.c_uint32), ("pad0", C.c_uint32)]


class Prim(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("object", C.c_uint32), ("sub", C.c_uint32), ("material", C.c_uint32),
                ("texture", C.c_int32), ("radius", C.c_float), ("radius_sqr", C.c_float), ("pad0", C.c_uint32),
                ("position", Vec4), ("a", Vec4), ("b", Vec4), ("c", Vec4)]


class Model(C.Structure):
    _fields_ = [("object", C.c_uint32), ("part_begin", C.c_uint32), ("part_count", C.c_uint32), ("pad0", C.c_uint32),
                ("position", Vec4), ("ver_min", Vec4), ("ver_max", Vec4)]


class Part(C.Structure):
    _fields_ = [("border_min", Vec4), ("border_max", Vec4), ("tri_begin", C.c_uint32), ("tri_count", C.c_uint32),
                ("material", C.c_uint32), ("texture", C.c_int32)]


class SceneDesc(C.Structure):
    _fields_ = [("camera", Camera), ("env_light", Vec4),
                ("n_lights", C.c_uint32), ("lights", C.POINTER(Light)),
                ("n_materials", C.c_uint32), ("materials", C.POINTER(Material)),
                ("n_textures", C.c_uint32), ("textures", C.POINTER(Texture)),
                ("texel_bytes", C.c_size_t), ("texels", C.POINTER(C.c_uint8)),
                ("n_prims", C.c_uint32), ("prims", C.POINTER(Prim)),
                ("n_models", C.c_uint32), ("models", C.POINTER(Model)),
                ("n_parts", C.c_uint32), ("parts", C.POINTER(Part)),
                ("n_tris", C.c_uint32),
                ("tri_points", C.POINTER(Vec4)), ("tri_norms", C.POINTER(Vec4)), ("tri_tcoords", C.POINTER(C.c_float)),
                ("geometry_epoch", C.c_uint64)]


class RenderParams(C.Structure):
    _fields_ = [("type", C.c_uint32), ("max_level", C.c_uint32), ("rank", C.c_uint32), ("world", C.c_uint32),
                ("flags", C.c_uint32), ("tile_rows", C.c_uint32), ("tile_first", C.c_uint32), ("tile_count", C.c_uint32)]


class HitId(C.Structure):
    _fields_ = [("object", C.c_int32), ("sub", C.c_int32), ("index", C.c_int32), ("octant", C.c_int32),
                ("distance", C.c_float)]


class Ray(C.Structure):
    _fields_ = [("origin", Vec4), ("direction", Vec4), ("mtlrfr", C.c_float), ("type", C.c_uint32),
                ("is_inside", C.c_uint32), ("pad0", C.c_uint32)]


class Hit(C.Structure):
    _fields_ = [("position", Vec4), ("normal", Vec4), ("tu", C.c_float), ("tv", C.c_float),
                ("material", C.c_int32), ("texture", C.c_int32), ("id", HitId), ("rfr", C.c_float),
                ("is_inside", C.c_uint32), ("pad0", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("shadow", C.c_uint64), ("reflect", C.c_uint64), ("refract", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("tri_tests", C.c_uint64), ("prim_tests", C.c_uint64),
                ("render_ms", C.c_double), ("trace_ms", C.c_double), ("shadow_ms", C.c_double),
                ("shade_ms", C.c_double), ("other_ms", C.c_double), ("upload_ms", C.c_double), ("build_ms", C.c_double),
                ("launches", C.c_uint32), ("bvh_nodes", C.c_uint32), ("bvh_depth", C.c_uint32), ("frame_sched", C.c_uint32),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("bvh_refit", C.c_uint32), ("pad0", C.c_uint32)]


# every symbol include/rt_b200.h declares: name -> (restype, argtypes)
RT_SYMBOLS = {
    "rt_last_error": (C.c_char_p, []),
    "rt_abi_version": (C.c_int, []),
    "rt_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "rt_destroy": (None, [C.c_void_p]),
    "rt_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rt_create_shared": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "rt_set_sm_share": (C.c_int, [C.c_void_p, C.c_int]),
    "rt_reserve_batch": (C.c_int, [C.c_void_p, C.c_uint32]),
    "rt_upload_scene": (C.c_int, [C.c_void_p, C.POINTER(SceneDesc)]),
    "rt_render_async": (C.c_int, [C.c_void_p, C.POINTER(RenderParams)]),
    "rt_render_supersampled": (C.c_int, [C.c_void_p, C.POINTER(RenderParams), C.c_uint32, C.POINTER(Camera)]),
    "rt_poll": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "rt_wait": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "rt_stop": (C.c_int, [C.c_void_p]),
    "rt_read_output": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "rt_read_output_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "rt_render_batch_async": (C.c_int, [C.c_void_p, C.POINTER(RenderParams), C.c_uint32, C.POINTER(Camera), C.POINTER(C.c_void_p)]),
    "rt_read_batch_output": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_int]),
    "rt_push_batch_rows": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]),
    "rt_output_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "rt_set_output": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "rt_landing_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p]),
    "rt_landing_open": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "rt_landing_close": (None, [C.c_void_p]),
    "rt_landing_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "rt_push_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "rt_landing_wait": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]),
    "rt_landing_release": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "rt_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "rt_host_free": (C.c_int, [C.c_void_p]),
    "rt_intersect_object": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(Ray), C.POINTER(Hit), C.c_float, C.POINTER(Hit), C.c_uint32]),
    "rt_read_hit_ids": (C.c_int, [C.c_void_p, C.POINTER(HitId)]),
    "rt_read_counters": (C.c_int, [C.c_void_p, C.POINTER(Counters)]),
    "rt_transfer_totals": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
}

RTH_SYMBOLS = {
    "rth_last_error": (C.c_char_p, []),
    "rth_fnv1a64": (C.c_ulonglong, [C.c_void_p, C.c_size_t]),
    "rth_scene_new": (C.c_void_p, []),
    "rth_scene_free": (None, [C.c_void_p]),
    "rth_scene_build": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]),
    "rth_scene_resize": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "rth_scene_object_count": (C.c_int, [C.c_void_p]),
    "rth_scene_light_count": (C.c_int, [C.c_void_p]),
    "rth_scene_move": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]),
    "rth_scene_switch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "rth_scene_chgmtl": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "rth_scene_set_object_position": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]),
    "rth_scene_set_light_position": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]),
    "rth_scene_camera_move": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float]),
    "rth_scene_camera_yaw": (C.c_int, [C.c_void_p, C.c_float]),
    "rth_scene_camera_pitch": (C.c_int, [C.c_void_p, C.c_float]),
    "rth_scene_camera_jitter": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "rth_scene_orbit_camera": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(Camera)]),
    "rth_scene_camera_get": (C.c_int, [C.c_void_p, C.POINTER(Camera)]),
    "rth_scene_camera_set_position": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float]),
    "rth_scene_camera_get_n": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "rth_scene_camera_set_n": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "rth_scene_flatten": (C.POINTER(SceneDesc), [C.c_void_p]),
    "rth_tracer_new": (C.c_void_p, [C.c_void_p, C.c_int]),
    "rth_tracer_free": (None, [C.c_void_p]),
    "rth_tracer_start": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "rth_tracer_stop": (None, [C.c_void_p]),
    "rth_tracer_failed": (C.c_int, [C.c_void_p]),
    "rth_tracer_last_error": (C.c_char_p, [C.c_void_p]),
    "rth_tracer_is_finished": (C.c_int, [C.c_void_p]),
    "rth_tracer_wait": (None, [C.c_void_p]),
    "rth_tracer_use_time": (C.c_double, [C.c_void_p]),
    "rth_tracer_output": (C.POINTER(C.c_uint8), [C.c_void_p]),
    "rth_tracer_width": (C.c_int, [C.c_void_p]),
    "rth_tracer_height": (C.c_int, [C.c_void_p]),
    "rth_tracer_set_max_level": (None, [C.c_void_p, C.c_int]),
    "rth_tracer_set_shard": (None, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "rth_tracer_set_flags": (None, [C.c_void_p, C.c_uint]),
    "rth_tracer_set_samples": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "rth_stratified_table": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "rth_scene_jittered_camera": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.POINTER(Camera)]),
    "rth_tracer_set_progressive": (None, [C.c_void_p, C.c_int]),
    "rth_tracer_bands_done": (C.c_int, [C.c_void_p]),
    "rth_tracer_set_coalesce": (None, [C.c_void_p, C.c_int]),
    "rth_tracer_set_sm_share": (None, [C.c_void_p, C.c_int]),
    "rth_tracer_read_hit_ids": (C.c_int, [C.c_void_p, C.POINTER(HitId)]),
    "rth_tracer_read_counters": (C.c_int, [C.c_void_p, C.POINTER(Counters)]),
    "rth_object_intersect": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Ray), C.POINTER(Hit), C.c_float]),
    "rth_tracer_context": (C.c_void_p, [C.c_void_p]),
}


def _load(name, symbols):
    path = os.path.join(LIB_DIR, name)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build the native libraries first (python -c 'import __graft_entry__ as g; g.build()' "
            f"or `make`). raytrace_b200 has no pure-Python or CPU fallback.")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for sym, (res, args) in symbols.items():
        fn = getattr(lib, sym)   # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


rt = _load("librt_b200.so", RT_SYMBOLS)
rth = _load("librt_host.so", RTH_SYMBOLS)
