"""profiles/ncu_traffic.json from an `ncu --set full` report of ONE launch (= one batch of frames) of bench.py:
the DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the traversal kernels of that launch, which bench.py
reports as roofline.traffic for the matching configuration.  Usage:
    python tools/ncu_traffic.py <report.ncu-rep> <config> <world> <frames_per_launch> [kernel-regex]
The entry records the report, the HEAD it was captured at and a hash of the kernel sources; bench.py drops the entry when the
sources have changed since (a stale capture must not be quoted)."""
import csv
import hashlib
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEVICE_SOURCES = ["rt_device.cuh", "rt_intersect.cuh", "rt_traverse.cuh", "rt_defer.cuh", "rt_steal.cuh", "rt_kernels.cu", "rt_kernels.h"]   # what the traced kernels are compiled from


def kernel_source_hash():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "raytrace_b200", "csrc")
    for f in DEVICE_SOURCES:
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def main():
    rep, config, world, B = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    rx = re.compile(sys.argv[5] if len(sys.argv) > 5 else r"k_wave|k_bin")
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    total, n, ms = 0.0, 0, 0.0
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        if not rx.search(d["Kernel Name"]):
            continue
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]
            total += float(d[k]) * scale
        t = float(d["gpu__time_duration.sum"]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}[u["gpu__time_duration.sum"]]
        ms += t
        n += 1
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        table = json.load(open(path))
    except Exception:
        table = {}
    head = subprocess.check_output(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"]).decode().strip()
    table[f"{config}:n{world}:b{B}"] = {
        "bytes_per_launch": int(total), "kernels": n, "kernel_ms_under_ncu": round(ms, 4),
        "source": f"{os.path.relpath(rep, ROOT)} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum over the {n} traversal kernel launches of one batch of {B} frames), captured at {head}",
        "kernel_source_hash": kernel_source_hash()}
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    print(path, table[f"{config}:n{world}:b{B}"])


if __name__ == "__main__":
    main()
