"""Jittered supersampling (BASELINE.json config 5): N sub-pixel samples per pixel, each sample a full
frame quantised by Color::put, averaged in integer -- so the oracle of an N-spp frame is simply N
reference renders with the same camera offsets (SURVEY.md 8d "C5").

The jitter moves the camera's forward vector exactly like the survey's harness does:
n' = n + u*(dx*dp) + v*(dy*dp), dp = tan(fovy*pi/360)/(H/2)  (host/capi.cpp rth_scene_camera_jitter).
"""
import numpy as np


def device_table(side=4, seed=0):
    """the same table as the C++ side computes it (scenes.h stratified_table, float offsets) -- what RayTracer.samples
    and bench.py's c5 use"""
    import ctypes as C

    from ._capi import rth
    buf = (C.c_float * (2 * side * side))()
    n = rth.rth_stratified_table(side, seed, buf)
    return [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]


def stratified_table(side=4, seed=0):
    """side x side stratified sub-pixel offsets in [0,1)^2, fixed by `seed` (PCG-free LCG, no numpy RNG
    so the table is identical everywhere)."""
    s = (0x9E3779B97F4A7C15 ^ seed) & ((1 << 64) - 1)
    out = []
    for j in range(side):
        for i in range(side):
            s = (s * 6364136223846793005 + 1442695040888963407) & ((1 << 64) - 1)
            a = ((s >> 40) & 0xFFFFFF) / float(1 << 24)
            s = (s * 6364136223846793005 + 1442695040888963407) & ((1 << 64) - 1)
            b = ((s >> 40) & 0xFFFFFF) / float(1 << 24)
            out.append(((i + a) / side, (j + b) / side))
    return out


def render_supersampled(scene, render_one, table):
    """render_one() -> (H, W, 3) uint8 frame of `scene` as it currently stands.
    Returns the integer average over the jitter table (floor division, like summing bytes)."""
    n0 = scene.camera_n()
    acc = None
    try:
        for dx, dy in table:
            scene.set_camera_n(n0)
            scene.camera_jitter(dx, dy)
            f = render_one().astype(np.uint32)
            acc = f if acc is None else acc + f
    finally:
        scene.set_camera_n(n0)
    return (acc // len(table)).astype(np.uint8)
