"""Regenerates tests/golden/golden.json (+ two small raw frames) from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and oracle/_ref/ref_render, built by
oracle/build_ref.sh).  The fixtures pin the CPU restatement (oracle/rt_oracle.cpp) and, through it,
the CUDA path: image hash (FNV-1a-64 of RayTracer::output), ray counts from the counting proxy and a
hash of the primary closest-hit identities.
"""
import gzip
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_render")
HIT = np.dtype([("object", "<i4"), ("sub", "<i4"), ("index", "<i4"), ("octant", "<i4"), ("distance", "<f4")])

# (scene, width, height, level, n, parts, type)
CASES = [
    ("c1", 1088, 576, 1, 0, 0, 0x80),      # the reference's default scene at its default size/depth
    ("c1", 1088, 576, 5, 0, 0, 0x80),
    ("c1", 1920, 1080, 5, 0, 0, 0x80),
    ("c1", 320, 200, 1, 0, 0, 0x80),       # ragged: 200 rows -> only 192 rendered, rest stays 127
    ("c1", 100, 50, 1, 0, 0, 0x80),        # smaller than one 64x64 tile: nothing rendered
    ("t_mixed", 640, 384, 4, 0, 0, 0x80),
    ("t_mixed", 640, 384, 0, 0, 0, 0x80),  # depth 0
    ("t_mixed", 320, 192, 8, 0, 0, 0x80),  # deep recursion through glass
    ("t_mixed", 320, 192, 3, 0, 0, 0x07),  # RTflec
    ("t_mixed", 320, 192, 3, 0, 0, 0x08),  # MY_MODEL_REFRACTTEST == RTfrac
    ("t_ballplane", 640, 384, 4, 0, 0, 0x80),
    ("c2", 640, 384, 3, 8, 0, 0x80),
    ("c2", 960, 576, 5, 16, 0, 0x80),
    ("t_mesh", 640, 384, 3, 0, 0, 0x80),
    ("t_twomesh", 640, 384, 3, 0, 0, 0x80),
    ("t_textured", 640, 384, 3, 0, 0, 0x80),   # map_Kd + 24-bit BMP on a mesh part, a zRotate()d model, both again in a mirror
    ("t_textured", 320, 192, 1, 0, 0, 0x04),   # RTtex: the texture lookup alone
    ("c3", 640, 384, 5, 96, 6, 0x80),
    ("c4", 640, 384, 6, 96, 6, 0x80),
    ("t_empty", 256, 128, 2, 0, 0, 0x80),
    ("t_nolight", 320, 192, 2, 0, 0, 0x80),
    ("t_lights", 640, 384, 4, 0, 0, 0x80),
    ("t_lights", 320, 192, 3, 0, 0, 0x07),
    ("t_inside", 384, 256, 6, 0, 0, 0x80),
    # the staged debug shaders (RayTracer.cpp:48-325), MY_MODEL_CHECK .. MY_MODEL_SHADOWTEST
    ("t_mixed", 320, 192, 1, 0, 0, 0x01),
    ("t_mesh", 320, 192, 1, 0, 0, 0x02),
    ("t_mesh", 320, 192, 1, 0, 0, 0x03),
    ("t_mesh", 320, 192, 1, 0, 0, 0x04),
    ("t_lights", 320, 192, 1, 0, 0, 0x05),
    ("t_lights", 320, 192, 1, 0, 0, 0x06),
    ("t_mesh", 320, 192, 1, 0, 0, 0x06),
]
SAVE_FRAMES = {("c1", 320, 200, 1), ("t_mixed", 320, 192, 8)}


def fnv(b):
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & ((1 << 64) - 1)
    return f"{h:016x}"


def main():
    out = []
    for name, w, h, level, n, parts, typ in CASES:
        ids_path, rgb_path = "/tmp/golden_ids.bin", "/tmp/golden.rgb"
        cmd = [REF, "--scene", name, "--width", str(w), "--height", str(h), "--level", str(level), "--n", str(n),
               "--parts", str(parts), "--type", str(typ), "--counts", "--out", rgb_path]
        if typ == 0x80:
            cmd += ["--ids", ids_path]
        j = json.loads(subprocess.check_output(cmd).decode())
        rec = {"scene": name, "w": w, "h": h, "level": level, "n": n, "parts": parts, "type": typ,
               "hash": j["hash"], "rays": j["rays"]}
        if typ == 0x80:
            ids = np.fromfile(ids_path, HIT)
            rec["ids_hash"] = fnv(ids.tobytes())
            rec["hit_pixels"] = int((ids["object"] >= 0).sum())
        if (name, w, h, level) in SAVE_FRAMES:
            fn = f"{name}_{w}x{h}_l{level}.rgb.gz"
            with open(rgb_path, "rb") as f, gzip.open(os.path.join(HERE, fn), "wb", 9) as g:
                g.write(f.read())
            rec["frame"] = fn
        out.append(rec)
        print(rec, file=sys.stderr)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py over oracle/_ref/ref_render (g++ -O2 -mavx2 -mfma -ffp-contract=off)",
                   "cases": out}, f, indent=1)


if __name__ == "__main__":
    main()
