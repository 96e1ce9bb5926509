"""Regenerates tests/golden/fullsize.json: frame hash (FNV-1a-64 of RayTracer::output) and ray counts of the
FULL-SIZE benchmark configurations, rendered by the UNMODIFIED reference (oracle/_ref/ref_render, built by
oracle/build_ref.sh from /root/reference).  bench.py compares the frame the GPU path renders inside its own run
with these hashes ("frame_check"); tests/test_gpu_timed_path.py does the same under pytest.

Run in the build container only.  C4 (3840x2160, 4 147 200 triangles, depth 8) takes about an hour on 8 cores:
    python tests/golden/make_fullsize.py c1 c2 c3 [c4]
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_render")
OUT = os.path.join(HERE, "fullsize.json")
# the configurations of bench.py
CONFIGS = {"c1": (1088, 576, 1), "c2": (1920, 1080, 5), "c3": (1920, 1080, 5), "c4": (3840, 2160, 8)}


def main():
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in sys.argv[1:] or ["c1", "c2", "c3"]:
        w, h, level = CONFIGS[name]
        out = subprocess.check_output([REF, "--scene", name, "--width", str(w), "--height", str(h), "--level", str(level),
                                       "--threads", str(min(32, os.cpu_count() or 1)), "--repeat", "1", "--counts", "--tmpdir", "/tmp"]).decode()
        j = json.loads(out.strip().splitlines()[-1])
        res[name] = {"w": w, "h": h, "level": level, "hash": j["hash"], "rays": j["rays"],
                     "reference_wall_s_in_build_container": j["wall_s"][0], "threads": j["threads"]}
        print(name, res[name], flush=True)
        json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
