#!/usr/bin/env python
"""bench.py -- Mrays/s and ms/frame of the trace-and-shade path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--impl b200|reference]

One "step" = one pass of the hot path over one BATCH of synthetic input: the `step_frames` frames (32 for c3) of a
camera orbit around the named configuration's scene (default c3: 1920x1080, 1 036 800-triangle Model + ground
plane, 2 lights, shadows + reflection depth 5 -- the configuration BASELINE.json's target is quoted on).  Every
frame of a step has its own camera (scenes.h orbit_camera; camera 0 is the configuration's own camera) and its
own framebuffer.  The step is the same whatever N is ("scaling": "strong"): N GPUs render 1/N of every frame
each.  A ray = one closest-hit or one shadow any-hit query.

  value    whole-job Mrays/s with the scene resident in HBM: K steps enqueued through the C ABI, B frames per
           launch (rt_render_batch_async: the frames of a launch share the ray queues, so every warp serves
           every frame and the thin tail of the ray trees is paid once per launch; B is chosen so that a launch
           holds ~8 M pixels per GPU: 4 frames on one GPU, 32 eighth-frame shards on eight) and M = 2 launches
           in flight per GPU (rt_create_shared pipelines over one resident scene, each on its own stream), timed
           with CUDA events that bracket all streams, max over ranks.
  latency  `ms_per_frame_alone`: ONE frame (camera 0) with nothing else in flight, per N (the whole-frame
           scheduler k_frame); `ms_per_launch_alone`: one launch of B frames alone.
  e2e      same metric through the reference-facing call RayTracer::start() with HOST buffers: every frame of
           every step re-flattens the Scene with that frame's camera, uploads the per-frame tables (H2D) and
           reads the RGB8 frame back into RayTracer::output (D2H) inside the timed region.
  frame_check  after the timed region the first launch of a step is rendered once more through the very same
           call sequence and frame 0 (camera 0; at N > 1 the frame assembled on rank 0) is hashed and compared
           with the hash of the UNMODIFIED reference's full frame (tests/golden/fullsize.json); every frame of that
           launch is also compared with the same camera rendered alone (other scheduler, no batch).
  roofline FP32-issue roofline of the traversal kernels of one launch: algorithmic FLOPs from device counters
           (DESIGN.md "flop model") / the kernels' CUDA-event time in a launch run alone right after the timed
           region, against 148 SMs x 128 lanes x sm_max_mhz of MEASURED_PEAKS.json.
  cpu_baseline  the reference's own CPU tracer (oracle/_ref/ref_render) on the box's host cores: ONE FULL FRAME of
           the configuration through RayTracer::start() (wall and the reference's own useTime); its frame hash is
           compared with the GPU frame of the same run.

N > 1 (torchrun): interleaved 8-row tiles, boustrophedon order, scene replicated; every frame's rows are delivered
into rank 0's frame by one-sided NVLink peer copies (rt_push_batch_rows) or, with --gather nccl, an NCCL gather.
`--impl reference` times the reference CPU tracer itself (rank 0 only) on a stratified tile sample per step.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (scene, width, height, maxLevel, n, parts, frames per step, description)
    "c1": ("c1", 1088, 576, 1, 0, 0, 32, "reference default scene (plane + sphere, 2 lights), 1088x576, depth 1"),
    "c2": ("c2", 1920, 1080, 5, 0, 0, 32, "1024 spheres + plane, 4 point lights, 1920x1080, depth 5"),
    "c3": ("c3", 1920, 1080, 5, 0, 0, 32, "1036800-triangle Model + plane, 2 lights, 1920x1080, depth 5, GPU LBVH"),
    "c4": ("c4", 3840, 2160, 8, 0, 0, 4, "4147200-triangle Model + 64 glass + 6 mirror spheres + plane, 3840x2160, depth 8"),
    # BASELINE configs[4]: the c4 scene at 8K, 16 jittered samples per pixel accumulated on the GPU; a step is ONE frame
    "c5": ("c5", 7680, 4320, 8, 0, 0, 1, "c4 scene at 7680x4320, 16 jittered samples per pixel (4x4 stratified table, seed 0), integer mean of the quantised samples, depth 8"),
}
SPP_SIDE = {"c5": 4}


def workload_config(name):
    """The `config` object of the JSON line -- identical for both arms (the driver compares them)."""
    scene, w, h, level, n, parts, sframes, desc = CONFIGS[name]
    if name in SPP_SIDE:
        return {"workload": f"{name}: {desc}", "step": "one supersampled frame", "step_frames": 1, "samples_per_pixel": SPP_SIDE[name] ** 2,
                "pixels_per_frame": (w // 64 * 64) * (h // 64 * 64),
                "l2_policy": "every band launch holds ~8 M sample rays per level (> 1 GB of queues) plus a 640 MB BVH: far beyond the 126 MB L2; no flush needed"}
    return {"workload": f"{name}: {desc}", "step": f"{sframes} frames of a {sframes}-camera orbit around the scene (camera 0 = the configuration's own camera), one framebuffer per frame",
            "step_frames": sframes, "pixels_per_frame": (w // 64 * 64) * (h // 64 * 64),
            "l2_policy": "per-step working set (ray/hit queues of ~8 M pixels per launch + BVH + triangles, > 2 GB touched per launch) exceeds the 126 MB L2; no flush needed"}


BACKPRESSURE = os.environ.get("RT_BENCH_BACKPRESSURE", "1") != "0"   # landing buffers armed + released (diagnostic switch)


def golden_fullsize(name):
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize.json"))).get(name)
    except Exception:
        return None


def ncu_traffic(name, world, B):
    """DRAM bytes of the traversal kernels of one launch from the committed ncu capture of THIS source state
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from an `ncu --set full` report); None when no capture
    matches the configuration."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = t.get(f"{name}:n{world}:b{B}")
        if not e:
            return (None, None)
        # a capture of other kernel sources is not quoted
        import hashlib
        h = hashlib.sha256()
        d = os.path.join(ROOT, "raytrace_b200", "csrc")
        for f in ("rt_device.cuh", "rt_intersect.cuh", "rt_traverse.cuh", "rt_defer.cuh", "rt_steal.cuh", "rt_kernels.cu", "rt_kernels.h"):   # tools/ncu_traffic.py DEVICE_SOURCES
            h.update(open(os.path.join(d, f), "rb").read())
        if h.hexdigest()[:16] != e.get("kernel_source_hash"):
            return (None, "profiles/ncu_traffic.json holds a capture of older kernel sources: not quoted")
        return (e["bytes_per_launch"], e["source"])
    except Exception:
        return (None, None)


def sm_peak_fp32_tflops():
    """148 SMs x 128 FP32 lanes x max SM clock -> T lane-instr/s."""
    mhz, src = 1965.0, "fallback 1965 MHz"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        mhz, src = float(mp["sm_max_mhz"]), "MEASURED_PEAKS.json sm_max_mhz"
        hbm = float(mp["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    return 148 * 128 * mhz * 1e6 / 1e12, src, hbm


class ClockSampler:
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md clocks line).  Sampled
    through NVML in a thread (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*`
    prints, without a second process polling the GPU); falls back to an `nvidia-smi -lms` loop."""

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread, self.stop_flag, self.src = index, [], None, None, False, None

    def _nvml_loop(self, nv, h):
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), float(mx), int(rs)))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.src = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.src = "nvidia-smi"
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) >= 7 and r[0].replace(".", "").isdigit() and r[1].replace(".", "").isdigit():
                bits = sum(b for b, i in ((0x8, 3), (0x40, 4), (0x20, 5), (0x4, 6)) if r[i].lower().startswith("active"))
                self.rows.append((float(r[0]), float(r[1]), bits))

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"], "samples": 0, "source": self.src}
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(r[1] for r in self.rows),
                "reasons": sorted(n for b, n in self.NAMES.items() if bits & b), "samples": len(sm), "source": self.src}


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_render")


def ref_cmd(name, threads, *extra):
    scene, w, h, level, n, parts, sframes, desc = CONFIGS[name]
    return [REF_BIN, "--scene", scene, "--width", str(w), "--height", str(h), "--level", str(level), "--n", str(n), "--parts", str(parts),
            "--threads", str(threads), "--tmpdir", "/tmp", *[str(x) for x in extra]]


def run_reference(args):
    """--impl reference: the reference's own CPU tracer on the host cores (rank 0 only).

    Each step traces a STRATIFIED sample of 64x64 tiles (every k-th tile of the frame in row-major tile order, so
    all tile rows are covered) of one frame of the orbit -- the K timed steps are spread evenly over the orbit's cameras -- through the
    reference's per-pixel entry RTfrac on all host threads (64-pixel rows handed out one at a time, so every
    thread works).  The reference's single-threaded per-frame RTPrepare is timed apart and charged in proportion:
    step seconds = trace_s + prepare_s * sample_tiles / frame_tiles, which is what a full frame costs per tile.
    The sample size is calibrated so that the whole --steps/--warmup run takes about 2.5 minutes."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    name = args.config
    scene, w, h, level, n, parts, sframes, desc = CONFIGS[name]
    cores = min(32, os.cpu_count() or 1)   # the reference hard-caps at 32 threads (RayTracer.h:22)
    frame_tiles = (w // 64) * (h // 64)
    passes = max(args.steps + args.warmup, 1)
    if os.path.exists(REF_BIN):
        kind = "reference"
        cal_tiles = min(frame_tiles, max(4, frame_tiles // 40))
        cal = json.loads(subprocess.check_output(ref_cmd(name, cores, "--tiles", cal_tiles, "--stratified", "--repeat", 1)).decode().strip().splitlines()[-1])
        per_tile = max(cal["step_s"][0] / cal_tiles, 1e-6)
        tiles = int(max(4, min(frame_tiles, round(150.0 / passes / per_tile))))
        j = json.loads(subprocess.check_output(ref_cmd(name, cores, "--tiles", tiles, "--stratified", "--counts", "--orbit", sframes,
                                                       "--repeat", args.steps, "--warmup", args.warmup)).decode().strip().splitlines()[-1])
        frac = tiles / frame_tiles
        secs = [t + p * frac for t, p in zip(j["step_s"], j["prepare_s"])]
        rays, px = sum(j["rays_s"]) / len(j["rays_s"]), j["pixels"]      # every timed pass counts its own rays
        sample = (f"{tiles} of {frame_tiles} 64x64 tiles, stratified over the {w}x{h} frame ({px} px, {rays:.0f} rays per step on average), one orbit camera per step (spread evenly over the orbit), "
                  f"{cores} threads; step seconds = trace + RTPrepare ({sum(j['prepare_s']) / len(j['prepare_s']):.3f} s per frame) x {frac:.3f}")
    else:
        kind = "port"
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import raytrace_b200 as R
        from parity_util import oracle_render
        sc = R.Scene(scene, w, h, n, parts)
        world = 16
        secs = []
        for s in range(args.warmup + args.steps):
            t0 = time.time()
            _, _, c = oracle_render(sc, level, want_ids=False, threads=cores, rank=s % world, world=world, tile_rows=8)
            if s >= args.warmup:
                secs.append(time.time() - t0)
        rays, px = c.primary + c.shadow + c.reflect + c.refract, c.primary
        sample = f"oracle port (oracle/rt_oracle.cpp), interleaved 8-row tiles rank s%16 of 16 ({px} px, {rays} rays) of the {w}x{h} frame per step, {cores} threads"
    total = sum(secs)
    mrays = rays * len(secs) / total / 1e6
    line = {"impl": "reference", "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total / len(secs) * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(name),
            "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(args, gpu_hash):
    """One FULL frame of the configuration (camera 0) through the reference's RayTracer::start() on the host cores
    (rank 0, N=1 only): wall start()->isFinish, the reference's own useTime, and the frame hash."""
    name = args.config
    scene, w, h, level, n, parts, sframes, desc = CONFIGS[name]
    cores = min(32, os.cpu_count() or 1)
    rays_frame = (golden_fullsize(name) or {}).get("rays", {}).get("total")
    try:
        if os.path.exists(REF_BIN) and name not in ("c4", "c5"):
            j = json.loads(subprocess.check_output(ref_cmd(name, cores, "--repeat", 1), timeout=1500).decode().strip().splitlines()[-1])
            wall, use = j["wall_s"][0], j["useTime_s"][0]
            return {"value": rays_frame / wall / 1e6 if rays_frame else None, "unit": "Mrays/s", "cores": cores, "kind": "reference",
                    "sample": f"one full {w}x{h} frame through the reference's RayTracer::start() ({rays_frame} rays): wall {wall:.2f} s, useTime {use:.2f} s",
                    "wall_s": wall, "useTime_s": use, "frame_hash": j["hash"], "same_frame_as_gpu": (j["hash"] == gpu_hash) if gpu_hash else None}
        if os.path.exists(REF_BIN):
            # c4 / c5: a full frame is hours of CPU -- stratified tiles through the per-pixel entry, labelled as such (SURVEY
            # 8d); c5's samples are reference frames of their own, so one sample per pixel gives the same rays/s
            frame_tiles, tiles = (w // 64) * (h // 64), 60
            j = json.loads(subprocess.check_output(ref_cmd(name, cores, "--tiles", tiles, "--stratified", "--counts", "--repeat", 1), timeout=1500).decode().strip().splitlines()[-1])
            secs = j["step_s"][0] + j["prepare_s"][0] * tiles / frame_tiles
            return {"value": j["rays_per_step"] / secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference",
                    "sample": f"{tiles} of {frame_tiles} 64x64 tiles stratified over the {w}x{h} frame ({j['rays_per_step']} rays), trace {j['step_s'][0]:.2f} s + RTPrepare {j['prepare_s'][0]:.2f} s x {tiles / frame_tiles:.3f}"}
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import raytrace_b200 as R
        from parity_util import oracle_render
        sc = R.Scene(scene, w, h, n, parts)
        t0 = time.time()
        _, _, c = oracle_render(sc, level, want_ids=False, threads=cores, rank=3, world=16, tile_rows=8)
        dt = time.time() - t0
        rays = c.primary + c.shadow + c.reflect + c.refract
        return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                "sample": f"oracle port, interleaved 8-row tiles rank 3 of 16 ({c.primary} px, {rays} rays) of the {w}x{h} frame, 1 pass"}
    except Exception as e:   # a baseline failure must not hide the GPU numbers
        return {"value": None, "unit": "Mrays/s", "cores": cores, "kind": "unavailable", "sample": repr(e)[:200]}


def bench_supersampled(args, sc, rank, world, local, dev, dist):
    """c5: 7680x4320, 16 jittered samples per pixel, image-space shards over N GPUs.  One step = one frame through
    rt_render_supersampled (the samples of a band of row tiles are one frame batch + an integer-mean kernel, band after
    band, nothing leaves the GPU), then this rank's rows go to rank 0's frame over NVLink (rt_push_rows)."""
    import numpy as np
    import torch

    import raytrace_b200 as R
    from raytrace_b200.distributed import FrameLanding, bands_of
    from raytrace_b200.supersample import device_table
    scene, w, h, level, n, parts, S, desc = CONFIGS[args.config]
    table = device_table(SPP_SIDE[args.config], 0)
    ns = len(table)
    cams = (R.Camera * ns)()
    for k, (dx, dy) in enumerate(table):
        cams[k] = sc.jittered_camera(dx, dy)

    def ck(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {R.rt.rt_last_error().decode()}")

    main_stream = torch.cuda.current_stream(dev)
    ctx = C.c_void_p()
    ck(R.rt.rt_create(local, C.byref(ctx)), "rt_create")
    ck(R.rt.rt_set_stream(ctx, C.c_void_p(main_stream.cuda_stream)), "rt_set_stream")
    ck(R.rt.rt_upload_scene(ctx, sc.flatten()), "rt_upload_scene")
    tile_rows = 8 if world > 1 else 64
    serp = world > 1 and args.shard_order == "serpentine"
    flags = R.RT_FLAG_SERPENTINE if serp else 0
    params = R.RenderParams(R.MY_MODEL_RAYTRACE, level, rank, world, flags, tile_rows, 0, 0)
    landing = FrameLanding(ctx, w, h, rank, world) if world > 1 else None
    if landing is not None and rank == 0:
        ptr, nbytes = landing.device_ptr()
        ck(R.rt.rt_set_output(ctx, C.c_void_p(ptr), nbytes), "rt_set_output")

    def frame():
        ck(R.rt.rt_render_supersampled(ctx, C.byref(params), ns, cams), "rt_render_supersampled")
        if landing is not None:
            landing.push(ctx)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(1, min(args.warmup, 2))):     # a frame is seconds of GPU time: two warm-up frames at most
        frame()
    sync_all()
    cnt = R.Counters()
    ck(R.rt.rt_read_counters(ctx, C.byref(cnt)), "rt_read_counters")
    rays_local = cnt.primary + cnt.shadow + cnt.reflect + cnt.refract
    launches = cnt.launches
    rays_t = torch.tensor([rays_local], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(rays_t)
    rays_frame = int(rays_t.item())
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main_stream)
    for _ in range(args.steps):
        frame()
    e1.record(main_stream)
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    mine_t = torch.tensor([e0.elapsed_time(e1) / args.steps, float(rays_local)], dtype=torch.float64, device=dev)
    per_rank = [torch.zeros_like(mine_t) for _ in range(world)]
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_gather(per_rank, mine_t)
    else:
        per_rank = [mine_t]
    per_rank = [{"rank": r, "ms_per_step": float(t[0].item()), "rays_per_step": int(t[1].item())} for r, t in enumerate(per_rank)]
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    # ---- frame check: the first band of this rank's shard, sample by sample (tile window, one frame per launch), meaned on
    #      the host, against the same rows of the frame the device accumulated
    got = np.empty((h, w, 3), dtype=np.uint8)
    ck(R.rt.rt_read_output(ctx, got.ctypes.data_as(C.c_void_p), w * 3), "rt_read_output")
    frame_hash = R.fnv1a64(got) if rank == 0 else None      # rank 0: the assembled frame (its landing buffer)
    mine = bands_of(rank, world, h, tile_rows, serp)
    ntile = max(1, 64 // tile_rows)
    rows = np.array([y for t in mine[:ntile] for y in range(t * tile_rows, (t + 1) * tile_rows)])
    probe = C.c_void_p()
    ck(R.rt.rt_create_shared(ctx, C.byref(probe)), "rt_create_shared")
    acc = np.zeros((len(rows), w, 3), dtype=np.uint32)
    one = np.empty((h, w, 3), dtype=np.uint8)
    win = R.RenderParams(R.MY_MODEL_RAYTRACE, level, rank, world, flags, tile_rows, 0, ntile)
    for k in range(ns):
        ck(R.rt.rt_render_batch_async(probe, C.byref(win), 1, C.cast(C.byref(cams, k * C.sizeof(R.Camera)), C.POINTER(R.Camera)), None), "rt_render_batch_async")
        ck(R.rt.rt_read_batch_output(probe, 0, one.ctypes.data_as(C.c_void_p), w * 3, 0), "rt_read_batch_output")
        acc += one[rows]
    same = bool(np.array_equal((acc // ns).astype(np.uint8), got[rows]))
    R.rt.rt_destroy(probe)
    ok_t = torch.tensor([1 if same else 0], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    # ---- roofline of one counted frame (stats kernels), untimed
    pstats = R.RenderParams(R.MY_MODEL_RAYTRACE, level, rank, world, flags | R.RT_FLAG_STATS, tile_rows, 0, 0)
    ck(R.rt.rt_render_supersampled(ctx, C.byref(params), ns, cams), "rt_render_supersampled")
    ck(R.rt.rt_read_counters(ctx, C.byref(cnt)), "rt_read_counters")
    stage = {"traverse": cnt.trace_ms, "shade": cnt.shade_ms, "other": cnt.other_ms, "render": cnt.render_ms}
    ck(R.rt.rt_render_supersampled(ctx, C.byref(pstats), ns, cams), "rt_render_supersampled(stats)")
    cs = R.Counters()
    ck(R.rt.rt_read_counters(ctx, C.byref(cs)), "rt_read_counters")
    sync_all()
    if landing is not None:
        landing.close()
    R.rt.rt_destroy(ctx)
    # ---- e2e: RayTracer::start() with `samples` set, host frame buffer (D2H of this rank's rows every step)
    e2e = None
    if not args.no_e2e:
        t = R.RayTracer(sc, device=local)
        t.maxLevel = level
        t.set_samples(table)
        kw = dict(flags=flags, rank=rank, world=world, tile_rows=tile_rows)
        t.render(R.MY_MODEL_RAYTRACE, **kw)
        t.render(R.MY_MODEL_RAYTRACE, **kw)
        sync_all()
        up0, dn0, up1, dn1 = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        R.rt.rt_transfer_totals(C.byref(up0), C.byref(dn0))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            t.start(R.MY_MODEL_RAYTRACE, **kw)
            t.wait()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        R.rt.rt_transfer_totals(C.byref(up1), C.byref(dn1))
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e = {"value": rays_frame * args.steps / float(e2e_s.item()) / 1e6, "unit": "Mrays/s", "ms_per_step": float(e2e_s.item()) / args.steps * 1e3,
               "h2d_bytes_per_step": (up1.value - up0.value) // args.steps, "d2h_bytes_per_step": (dn1.value - dn0.value) // args.steps,
               "bytes_counted_on": "rank 0" if world > 1 else "the one GPU", "calls_per_step": 1}
        del t
    if rank == 0:
        peak, peak_src, hbm_peak = sm_peak_fp32_tflops()
        flops = cs.nodes_visited * 4 * 22 + cs.tri_tests * 47 + cs.prim_tests * 23
        achieved = flops / (stage["traverse"] * 1e-3) / 1e12 if stage["traverse"] > 0 else 0.0
        slow = max(per_rank, key=lambda r: r["ms_per_step"])
        line = {"metric": "Mrays/s", "value": rays_frame * args.steps / (ms_total * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(args.config), "ms_per_frame": ms_total / args.steps,
                "run": {"rays_per_step": rays_frame, "sample_rays_per_pixel": rays_frame / ((w // 64 * 64) * (h // 64 * 64)),
                        "parallelism": f"image-space: interleaved {tile_rows}-row tiles over {world} GPUs, rows pushed into rank 0's frame over NVLink P2P" if world > 1 else "single GPU",
                        "kernel_launches_per_frame": launches, "slowest_rank": slow["rank"], "slowest_rank_ms_per_step": slow["ms_per_step"]},
                "frame_check": {"what": "rows of the first band of every rank's shard: device-accumulated mean == host mean of the 16 samples rendered one per launch",
                                "ok": bool(ok_t.item()), "status": "ok" if ok_t.item() else "MISMATCH", "frame_hash": frame_hash,
                                "expected": None, "expected_source": "no CPU frame at this size (16 x 33 M pixels at depth 8 is days of reference time); the definition is pinned at test size by tests/test_gpu_timed_path.py"},
                "gpu_launches": launches * args.steps,
                "roofline": {"bound": "fp32_issue", "kernel": "k_wave (the band launches of one supersampled frame)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": achieved / peak, "peak_source": f"148 SMs x 128 lanes x {peak_src} (of measured)", "traffic": None,
                             "flops_per_launch": flops, "nodes_per_ray": cs.nodes_visited / max(rays_local, 1), "tri_tests_per_ray": cs.tri_tests / max(rays_local, 1),
                             "stage_ms_one_frame": stage},
                "clocks": clocks, "per_rank": per_rank,
                "build": {"upload_ms": cs.upload_ms, "lbvh_build_ms": cs.build_ms, "bvh_nodes": cs.bvh_nodes, "bvh_depth": cs.bvh_depth}}
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, None)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)    # a step is a batch of `step_frames` frames (32 for c3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the RayTracer::start() leg (profiling runs)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: how the row tiles reach rank 0 -- one-sided NVLink peer copies on the copy engines (rt_push_rows) or an NCCL gather")
    ap.add_argument("--shard-order", default="serpentine", choices=["serpentine", "modulo"],
                    help="N>1: which rank renders row tile t -- boustrophedon (RT_FLAG_SERPENTINE, evens out the ray-cost gradient down the image) or t %% N")
    ap.add_argument("--pipelines", type=int, default=0, help="launches in flight per GPU (0 = auto)")
    ap.add_argument("--sm-share", type=int, default=-1, help="resident traversal CTAs per SM per pipeline (-1 = auto)")
    ap.add_argument("--batch", type=int, default=0, help="frames per launch (0 = about 8 M pixels per launch and GPU)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    import raytrace_b200 as R

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    scene, w, h, level, n, parts, S, desc = CONFIGS[args.config]

    tmpdir = f"/tmp/rt_bench_{rank}"
    os.makedirs(tmpdir, exist_ok=True)
    sc = R.Scene(scene, w, h, n, parts, tmpdir=tmpdir)
    if args.config in SPP_SIDE:
        return bench_supersampled(args, sc, rank, world, local, dev, dist)
    cams = (R.Camera * S)()
    for k in range(S):
        cams[k] = sc.orbit_camera(k, S)

    def ck(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {R.rt.rt_last_error().decode()}")

    # ---- frame pipelines: one resident scene, M launches in flight, B frames per launch -----------------
    # about 16 M pixels per launch and GPU (sweeps in profiles/r1i_batch_sweep_c3.txt and r2ac_batch_sweep_c3.txt: launches of that size run the
    # per-level wave kernels with queues long enough that their tails do not matter, and two launches in flight
    # cover each other's level boundaries).  Scenes of analytic primitives only (c1, c2) trace 12-17 Grays/s: their
    # launches are bound by the ray-queue traffic, which stays closer to the L2 with one frame per launch; they keep
    # three single frames in flight.  A frame beyond ~3 M pixels per GPU (c4) is a launch of its own.
    big = w * h // world > 3_000_000
    pix_rank = (w // 64 * 64) * (h // 64 * 64) // world
    mesh = args.config in ("c3", "c4")
    B = args.batch if args.batch > 0 else (1 if big or not mesh else max(1, min(64, S, round(16_000_000 / max(pix_rank, 1)))))
    B_e2e = B if args.batch > 0 else (1 if big or not mesh else max(1, min(64, S, round(8_000_000 / max(pix_rank, 1)))))   # tracers of the RayTracer::start() leg: 4 x this
    M = args.pipelines if args.pipelines > 0 else (1 if big else 2 if B > 1 else 3)
    share = args.sm_share if args.sm_share >= 0 else (0 if M == 1 or B > 1 else 4)
    L = (S + B - 1) // B                         # launches per step
    main_stream = torch.cuda.current_stream(dev)
    owner = C.c_void_p()
    ck(R.rt.rt_create(local, C.byref(owner)), "rt_create")
    ck(R.rt.rt_set_stream(owner, C.c_void_p(main_stream.cuda_stream)), "rt_set_stream")
    ck(R.rt.rt_upload_scene(owner, sc.flatten()), "rt_upload_scene")      # H2D of the scene + LBVH build, once
    tile_rows = 8 if world > 1 else 64           # fine interleave balances the ranks (sky rows are cheap)
    serp = world > 1 and args.shard_order == "serpentine"
    shard_flags = R.RT_FLAG_SERPENTINE if serp else 0
    params = R.RenderParams(R.MY_MODEL_RAYTRACE, level, rank, world, shard_flags, tile_rows)
    from raytrace_b200.distributed import FrameGather, FrameLanding
    p2p = world > 1 and args.gather == "p2p"
    pipes = []
    for _ in range(M):
        hnd = C.c_void_p()
        ck(R.rt.rt_create_shared(owner, C.byref(hnd)), "rt_create_shared")
        st = torch.cuda.Stream(dev)
        ck(R.rt.rt_set_stream(hnd, C.c_void_p(st.cuda_stream)), "rt_set_stream")
        ck(R.rt.rt_set_sm_share(hnd, share if M > 1 else 0), "rt_set_sm_share")
        frames, landings, outs = [], [], (C.c_void_p * B)()
        for f in range(B):
            frame, landing = None, None
            if p2p:
                # rank 0 renders straight into the landing buffer the other ranks push their rows to
                landing = FrameLanding(hnd, w, h, rank, world, backpressure=BACKPRESSURE)
                if rank == 0:
                    outs[f] = landing.device_ptr()[0]
            if not (p2p and rank == 0):
                frame = torch.full((h, w, 3), 127, dtype=torch.uint8, device=dev)
                outs[f] = frame.data_ptr()
            frames.append(frame), landings.append(landing)
        pipes.append({"ctx": hnd, "stream": st, "frames": frames, "landings": landings, "outs": outs, "assembled": [None] * B,
                      "consumer": torch.cuda.Stream(dev) if p2p and rank == 0 else None,   # where the assembled frames become visible
                      "gather": [FrameGather(w, h, rank, world, dev, tile_rows, serp) for _ in range(B)] if world > 1 and not p2p else None})   # one per frame of a launch: a FrameGather assembles into its own pre-allocated frame

    def cam_ptr(first):
        return C.cast(C.byref(cams, first * C.sizeof(R.Camera)), C.POINTER(R.Camera))

    def launch(gj, j):
        """launch j of a step (frames j*B .. of the orbit) as global launch gj, on pipeline gj % M: ONE launch, then
        every frame's rows go to rank 0"""
        p = pipes[gj % M]
        first = j * B
        nb = min(B, S - first)
        ck(R.rt.rt_render_batch_async(p["ctx"], C.byref(params), nb, cam_ptr(first), p["outs"]), "rt_render_batch_async")
        for f in range(nb):
            if p["landings"][f] is not None:
                p["landings"][f].push(p["ctx"], p["consumer"], frame=f, release=BACKPRESSURE)   # (back-pressure: a push waits until rank 0's consumer has released the frame pushed into this buffer before) NVLink P2P: this rank's row tiles -> their place in rank 0's frame (copy engines) + signal
            elif p["gather"] is not None:
                with torch.cuda.stream(p["stream"]):
                    p["assembled"][f] = p["gather"][f].gather(p["frames"][f])   # NCCL: this rank's row tiles -> rank 0, de-interleaved there
        return p, nb

    state = {"gj": 0}

    def run(nsteps):
        for _ in range(nsteps):
            for j in range(L):
                launch(state["gj"], j)
                state["gj"] += 1

    def fork():
        ev = torch.cuda.Event()
        ev.record(main_stream)
        for p in pipes:
            p["stream"].wait_event(ev)

    def join():
        for p in pipes:
            for st in (p["stream"], p["consumer"]):
                if st is not None:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    main_stream.wait_event(ev)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- warm-up, then one counted step (ray totals are deterministic per configuration and shard) --------
    fork()
    run(args.warmup)
    join()
    sync_all()
    cnt = R.Counters()
    rays_step_local, launches_per_launch = 0, 0
    for j in range(L):
        p, nb = launch(state["gj"], j)
        state["gj"] += 1
        ck(R.rt.rt_read_counters(p["ctx"], C.byref(cnt)), "rt_read_counters")      # waits for that launch
        rays_step_local += cnt.primary + cnt.shadow + cnt.reflect + cnt.refract
        launches_per_launch = cnt.launches
    sync_all()
    rays_t = torch.tensor([rays_step_local], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(rays_t)
    rays_step = int(rays_t.item())

    # ---- timed region: exactly K steps, CUDA events bracketing every pipeline stream, max over ranks ------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main_stream)
    fork()
    run(args.steps)
    join()
    e1.record(main_stream)
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    # per-rank view (load balance of the image-space shards)
    mine_t = torch.tensor([e0.elapsed_time(e1) / args.steps, float(rays_step_local)], dtype=torch.float64, device=dev)
    per_rank = [torch.zeros_like(mine_t) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine_t)
    else:
        per_rank = [mine_t]
    per_rank = [{"rank": r, "ms_per_step": float(t[0].item()), "rays_per_step": int(t[1].item())} for r, t in enumerate(per_rank)]
    slow = max(per_rank, key=lambda r: r["ms_per_step"])
    clocks = sampler.stop() if rank == 0 else None
    value = rays_step * args.steps / (ms_total * 1e-3) / 1e6

    # ---- frame check: launch 0 of a step once more through the same call sequence; frame 0 (camera 0, assembled on
    #      rank 0) against the unmodified reference's full-frame hash, every frame of the launch against the same
    #      camera rendered alone ------------------------------------------------------------------------------
    fork()
    pchk, nb0 = launch(0, 0)                      # pipeline 0
    join()
    sync_all()
    ck(R.rt.rt_wait(pchk["ctx"], None), "rt_wait")
    gold = golden_fullsize(args.config)
    frame_check = {"frame": "camera 0 of the orbit = the configuration's own camera; rendered after the timed region by the same call sequence as every timed launch"}
    batch_frames = []
    for f in range(nb0):
        host = np.empty((h, w, 3), dtype=np.uint8)
        if world > 1 and rank == 0 and not p2p:
            host = pchk["assembled"][f].cpu().numpy()
        else:
            ck(R.rt.rt_read_batch_output(pchk["ctx"], f, host.ctypes.data_as(C.c_void_p), w * 3, 0), "rt_read_batch_output")
        batch_frames.append(host)
    if rank == 0:
        frame_check["hash"] = R.fnv1a64(batch_frames[0])
        frame_check["expected"] = gold["hash"] if gold else None
        frame_check["expected_source"] = "tests/golden/fullsize.json: the unmodified reference's full frame (tests/golden/make_fullsize.py)" if gold else "no golden hash for this configuration"
        frame_check["ok"] = bool(gold and frame_check["hash"] == gold["hash"])
    # the same cameras rendered alone (one frame per launch: other scheduler for mesh frames <= 3 M pixels), this rank's rows
    p0 = pipes[0]["ctx"]
    ck(R.rt.rt_set_sm_share(p0, 0), "rt_set_sm_share")
    from raytrace_b200.distributed import bands_of
    rows = np.array([y for y in range(h // 64 * 64) if (y // tile_rows) in set(bands_of(rank, world, h, tile_rows, serp))]) if world > 1 else np.arange(h)
    same_alone = True
    c1 = R.Counters()
    for f in range(nb0):
        ck(R.rt.rt_render_batch_async(p0, C.byref(params), 1, cam_ptr(f), None), "rt_render_batch_async(1)")
        alone = np.empty((h, w, 3), dtype=np.uint8)
        ck(R.rt.rt_read_batch_output(p0, 0, alone.ctypes.data_as(C.c_void_p), w * 3, 0), "rt_read_batch_output")
        same_alone = same_alone and bool(np.array_equal(alone[rows], batch_frames[f][rows]))
    ck(R.rt.rt_read_counters(p0, C.byref(c1)), "rt_read_counters")
    alone_sched = "k_frame" if c1.frame_sched else "k_wave"
    ok_t = torch.tensor([1 if same_alone else 0], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    if rank == 0:
        frame_check["batch_frames_equal_frames_rendered_alone"] = bool(ok_t.item())
        frame_check["alone_scheduler"] = alone_sched
        frame_check["status"] = "ok" if frame_check["ok"] and frame_check["batch_frames_equal_frames_rendered_alone"] else "MISMATCH"

    # ---- one launch alone (per-stage split from the library's own CUDA events), one counted launch for the
    #      roofline, and one single frame alone (latency); all untimed ----------------------------------
    for _ in range(2):
        ck(R.rt.rt_render_batch_async(p0, C.byref(params), min(B, S), cam_ptr(0), pipes[0]["outs"]), "rt_render_batch_async")
        ck(R.rt.rt_read_counters(p0, C.byref(cnt)), "rt_read_counters")
    stage = {"traverse": cnt.trace_ms, "shade": cnt.shade_ms, "other": cnt.other_ms, "render": cnt.render_ms}
    rays_launch0 = cnt.primary + cnt.shadow + cnt.reflect + cnt.refract
    wave_sched = not cnt.frame_sched
    pstats = R.RenderParams(R.MY_MODEL_RAYTRACE, level, rank, world, R.RT_FLAG_STATS | shard_flags, tile_rows)
    ck(R.rt.rt_render_batch_async(p0, C.byref(pstats), min(B, S), cam_ptr(0), pipes[0]["outs"]), "rt_render_batch_async(stats)")
    cs = R.Counters()
    ck(R.rt.rt_read_counters(p0, C.byref(cs)), "rt_read_counters")
    for _ in range(3):
        ck(R.rt.rt_render_batch_async(p0, C.byref(params), 1, cam_ptr(0), None), "rt_render_batch_async(1)")
        ck(R.rt.rt_read_counters(p0, C.byref(c1)), "rt_read_counters")
    alone_t = torch.tensor([c1.render_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(alone_t, op=dist.ReduceOp.MAX)
    ms_frame_alone = float(alone_t.item())
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()                           # nobody unmaps a landing buffer another rank may still push to
    for p in pipes:
        for landing in p["landings"]:
            if landing is not None:
                landing.close()
        R.rt.rt_destroy(p["ctx"])
    R.rt.rt_destroy(owner)

    # ---- e2e: RayTracer::start() with host buffers (flatten + H2D tables + render + D2H frame) ----
    # Tracers of the one Scene (the reference's idiom for several views).  With frame batches on one GPU they run in
    # throughput mode (RayTracer::coalesce): the Scene's batch workers render whatever start() calls are waiting in ONE
    # launch, each frame through the camera its start() saw.  Every start() of step s, frame f is preceded by putting
    # orbit camera f into the Scene (Scene::cam.position, as the reference's UI moves its camera between frames).
    e2e = None
    if not args.no_e2e:
        env_c = os.environ.get("RT_BENCH_COALESCE")
        coalesce = B_e2e > 1 and (env_c != "0")
        M_e2e = 1 if big else (min(64, int(env_c or "8") * B_e2e) if coalesce else 3 if world == 1 else 4 if world < 8 else 8)
        share_e2e = 0 if M_e2e == 1 or coalesce else 4 if world == 1 else 2 if world < 8 else 1
        tracers = []
        for _ in range(M_e2e):
            t = R.RayTracer(sc, device=local)        # the drop-in surface; tracers of one Scene share its residency
            t.maxLevel = level
            t.smShare = share_e2e if M_e2e > 1 else 0
            t.coalesce = coalesce
            tracers.append(t)

        host = {"wait": 0.0, "start": 0.0, "calls": 0}

        def e2e_frames(nframes, k0=0):
            for k in range(k0, k0 + nframes):
                t = tracers[k % M_e2e]
                a = time.perf_counter()
                t.wait()                             # that tracer's previous frame is in RayTracer::output
                b = time.perf_counter()
                cpos = cams[k % S].position
                sc.set_camera_position(cpos.x, cpos.y, cpos.z)
                t.start(R.MY_MODEL_RAYTRACE, flags=shard_flags, rank=rank, world=world, tile_rows=tile_rows)
                host["wait"] += b - a
                host["start"] += time.perf_counter() - b
                host["calls"] += 1
            for t in tracers:
                t.wait()

        e2e_frames(max(2 * M_e2e, S))
        sync_all()
        host.update(wait=0.0, start=0.0, calls=0)
        up0, dn0, up1, dn1 = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        R.rt.rt_transfer_totals(C.byref(up0), C.byref(dn0))      # bytes the library itself copies, counted where it enqueues them
        t0 = time.perf_counter()
        e2e_frames(args.steps * S)
        torch.cuda.synchronize(dev)
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        R.rt.rt_transfer_totals(C.byref(up1), C.byref(dn1))
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e = {"value": rays_step * args.steps / float(e2e_s.item()) / 1e6, "unit": "Mrays/s", "ms_per_step": float(e2e_s.item()) / args.steps * 1e3,
               "h2d_bytes_per_step": (up1.value - up0.value) // args.steps, "d2h_bytes_per_step": (dn1.value - dn0.value) // args.steps,
               "bytes_counted_on": "rank 0" if world > 1 else "the one GPU",
               "calls_per_step": S, "tracers_in_flight": M_e2e, "coalesced_starts": coalesce,
               "host_us_per_start_call": host["start"] / max(host["calls"], 1) * 1e6, "host_us_waiting_per_call": host["wait"] / max(host["calls"], 1) * 1e6}
        del tracers

    if rank == 0:
        peak, peak_src, hbm_peak = sm_peak_fp32_tflops()
        # flop model: 22 per child box (4 per 4-wide node), 47 per triangle test, 23 per analytic primitive
        flops = cs.nodes_visited * 4 * 22 + cs.tri_tests * 47 + cs.prim_tests * 23
        trav_ms = stage["traverse"]
        trav_launches = level + 2 if wave_sched else 1   # whole-frame scheduler: one traversal launch per frame
        achieved = flops / (trav_ms * 1e-3) / 1e12 if trav_ms > 0 else 0.0
        queue_bytes = rays_launch0 * 100   # ~100 B of ray/hit/node records written+read per ray
        traffic, traffic_src = ncu_traffic(args.config, world, min(B, S))
        cfg = workload_config(args.config)
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "ms_per_frame": ms_total / args.steps / S,
            "run": {"rays_per_step": rays_step, "rays_per_frame_mean": rays_step / S,
                    "parallelism": (f"image-space: interleaved {tile_rows}-row tiles over {world} GPUs ({'boustrophedon' if serp else 'modulo'} order), " + ("row tiles pushed into rank 0's frame over NVLink P2P (copy engines, put with signal; NCCL only ships the IPC handles)" if p2p else "NCCL gather of RGB8 tiles to rank 0")) if world > 1 else "single GPU",
                    "frames_per_launch": B, "launches_per_step": L, "launches_in_flight": M, "traversal_ctas_per_sm_per_pipeline": (share if M > 1 and share else 8),
                    "slowest_rank": slow["rank"], "slowest_rank_ms_per_step": slow["ms_per_step"]},
            "latency": {"ms_per_frame_alone": ms_frame_alone, "scheduler": alone_sched, "ms_per_launch_alone": stage["render"], "frames_per_launch": min(B, S),
                        "note": "one frame (camera 0) with nothing else in flight, max over ranks; the stream figure above overlaps launches"},
            "frame_check": frame_check,
            "gpu_launches": launches_per_launch * L * args.steps,
            "roofline": {"bound": "fp32_issue", "kernel": "k_wave (closest-hit level l fused with shadow any-hit level l-1; the level+2 launches of one batch of frames)" if wave_sched else "k_frame (one persistent launch: closest-hit + shadow traversal of all levels)", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak, "peak_source": f"148 SMs x 128 lanes x {peak_src} (of measured)",
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel_launches_per_batch": trav_launches, "frames_per_batch": min(B, S), "avg_launch_ms": trav_ms / trav_launches,
                         "flops_per_launch": flops, "nodes_per_ray": cs.nodes_visited / max(rays_launch0, 1), "tri_tests_per_ray": cs.tri_tests / max(rays_launch0, 1),
                         "stage_ms_one_launch_alone": stage,
                         "hbm_secondary": {"queue_bytes_per_launch": queue_bytes, "achieved_gbs": queue_bytes / (stage["render"] * 1e-3) / 1e9,
                                           "peak_gbs": hbm_peak}},
            "clocks": clocks, "per_rank": per_rank,
            "build": {"upload_ms": cs.upload_ms, "lbvh_build_ms": cs.build_ms, "bvh_nodes": cs.bvh_nodes, "bvh_depth": cs.bvh_depth},
        }
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, frame_check.get("hash"))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
