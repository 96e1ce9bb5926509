#!/usr/bin/env python
"""bench.py -- Mrays/s and ms/frame of the trace-and-shade path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--impl b200|reference]

One "step" = one frame of the named configuration (default c3: 1920x1080, 1 036 800-triangle Model
+ ground plane, 2 lights, shadows + reflection depth 5 -- the configuration BASELINE.json's target
is quoted on).  A ray = one closest-hit or one shadow any-hit query.

  value    whole-job Mrays/s with the scene resident in HBM: K frames enqueued through the C ABI, B frames
           per launch (rt_render_batch_async: the frames of a launch share the ray queues, so every warp
           serves every frame and the thin tail of the ray trees is paid once per launch; B is chosen so
           that a launch holds ~8 M pixels per GPU: 4 frames on one GPU, 32 eighth-frame shards on eight)
           and M = 2 launches in flight per GPU (rt_create_shared pipelines over one resident scene, each on
           its own stream), timed with CUDA events that bracket all streams, max over ranks.  Every frame is
           rendered completely and lands in its own framebuffer; `config.ms_per_frame_alone` is the
           latency of ONE frame with nothing else in flight, `ms_per_launch_alone` that of one batch.
  e2e      same metric through the reference-facing call RayTracer::start() with HOST buffers:
           every step re-flattens the Scene, uploads the per-frame tables (H2D) and reads the
           RGB8 frame back into RayTracer::output (D2H) inside the timed region.  M RayTracer
           objects over the one Scene (the reference's idiom for several views) keep M frames in
           flight; step k waits for step k-M on the same tracer before it starts.
  roofline FP32-issue roofline of the traversal kernels of one launch (the level+2 k_wave launches of a
           batch, or one k_frame): algorithmic FLOPs from device counters (DESIGN.md "flop model") / the
           kernels' CUDA-event time in a launch run alone right after the timed region (inside it the
           launches of the pipelines overlap, so a per-launch time is not defined there), against
           148 SMs x 128 lanes x sm_max_mhz of MEASURED_PEAKS.json (1 lane-instr = 1 flop because the
           parity path is unfused); HBM figures are reported beside it as the secondary bound.
  cpu_baseline  the reference's own CPU tracer (oracle/_ref/ref_render, else the oracle port) on
           the box's host cores, on a bounded tile sample of the same workload.

N > 1 (torchrun): the frame is split into interleaved 8-row tiles (tile % N == rank), the scene is
replicated (boustrophedon tile order by default, --shard-order), and every step each rank's rows are delivered into rank 0's frame ("strong" scaling):
by default with one-sided NVLink peer copies on the copy engines (rt_push_batch_rows; torch.distributed /
NCCL only carries the IPC handles, barriers and timing reductions), with --gather nccl by an NCCL gather.
`--impl reference` times the reference CPU tracer itself (rank 0 only).
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
    # name: (scene, width, height, maxLevel, n, parts, description)
    "c1": ("c1", 1088, 576, 1, 0, 0, "reference default scene (plane + sphere, 2 lights), 1088x576, depth 1"),
    "c2": ("c2", 1920, 1080, 5, 0, 0, "1024 spheres + plane, 4 point lights, 1920x1080, depth 5"),
    "c3": ("c3", 1920, 1080, 5, 0, 0, "1036800-triangle Model + plane, 2 lights, 1920x1080, depth 5, GPU LBVH"),
    "c4": ("c4", 3840, 2160, 8, 0, 0, "4147200-triangle Model + 64 glass + 6 mirror spheres + plane, 3840x2160, depth 8"),
}
# DRAM bytes of the traversal kernels of one launch (k_wave x level+2, or one k_frame) measured once with
# `ncu --set full` (profiles/), keyed by (config, gpus, frames per launch)
NCU_TRAFFIC = {("c3", 1, 4): (3312490751, "profiles/r1i_ncu_full_k_wave_batch_c3.md (dram__bytes_read.sum + dram__bytes_write.sum over the 7 k_wave launches of one batch of 4 frames)"),
               ("c3", 1, 1): (736713216, "profiles/r1h_ncu_full_k_frame_c3.md (dram__bytes_read.sum + dram__bytes_write.sum of one k_frame launch)")}
REF_TILES = {"c1": 153, "c2": 12, "c3": 6, "c4": 2}   # 64x64 tiles per reference step (bounded sample)


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


def run_reference(args, cfg):
    """--impl reference: the reference's own CPU tracer on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene, w, h, level, n, parts, desc = cfg
    cores = min(32, os.cpu_count() or 1)   # the reference hard-caps at 32 threads (RayTracer.h:22)
    # bounded sample: REF_TILES tiles per step at 20 steps, fewer tiles per step for longer runs, so that the
    # whole run stays within a few minutes whatever K is
    tiles = max(1, min(REF_TILES[args.config], round(REF_TILES[args.config] * 20 / max(args.steps + args.warmup, 1))))
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_render")
    if os.path.exists(ref):
        kind = "reference"
        out = subprocess.check_output([ref, "--scene", scene, "--width", str(w), "--height", str(h), "--level", str(level), "--n", str(n),
                                       "--parts", str(parts), "--threads", str(cores), "--tiles", str(tiles), "--counts",
                                       "--repeat", str(args.steps), "--warmup", str(args.warmup)]).decode()
        j = json.loads(out.strip().splitlines()[-1])
        secs, rays, px = j["step_s"], j["rays_per_step"], j["pixels"]
    else:
        kind = "port"
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import raytrace_b200 as R
        from parity_util import oracle_render
        sc = R.Scene(scene, w, h, n, parts)
        world = max(1, (h // 64) * (w // 64) // tiles)
        world = min(world, h // 64)
        secs = []
        for s in range(args.warmup + args.steps):
            t0 = time.time()
            _, _, c = oracle_render(sc, level, want_ids=False, threads=cores, rank=0, world=world)
            if s >= args.warmup:
                secs.append(time.time() - t0)
        rays, px = c.primary + c.shadow + c.reflect + c.refract, c.primary
    total = sum(secs)
    mrays = rays * len(secs) / total / 1e6
    sample = f"{tiles} seeded 64x64 tiles ({px} px, {rays} rays) of the {w}x{h} frame per step, {cores} threads"
    line = {"impl": "reference", "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total / len(secs) * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {desc}", "sample": sample},
            "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(args, cfg):
    """Bounded reference sample timed beside the GPU numbers (rank 0, N=1 only)."""
    scene, w, h, level, n, parts, desc = cfg
    cores = min(32, os.cpu_count() or 1)
    tiles = REF_TILES[args.config]
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_render")
    try:
        if os.path.exists(ref):
            out = subprocess.check_output([ref, "--scene", scene, "--width", str(w), "--height", str(h), "--level", str(level), "--n", str(n),
                                           "--parts", str(parts), "--threads", str(cores), "--tiles", str(tiles), "--counts",
                                           "--repeat", "2", "--warmup", "1"], timeout=900).decode()
            j = json.loads(out.strip().splitlines()[-1])
            v = j["rays_per_step"] * len(j["step_s"]) / sum(j["step_s"]) / 1e6
            return {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "reference",
                    "sample": f"{tiles} seeded 64x64 tiles ({j['pixels']} px, {j['rays_per_step']} rays) of the {w}x{h} frame, 2 timed passes after 1 warm-up"}
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import raytrace_b200 as R
        from parity_util import oracle_render
        sc = R.Scene(scene, w, h, n, parts)
        world = min(h // 64, max(1, (h // 64) * (w // 64) // tiles))
        oracle_render(sc, level, want_ids=False, threads=cores, rank=0, world=world)
        t0 = time.time()
        _, _, c = oracle_render(sc, level, want_ids=False, threads=cores, rank=0, world=world)
        dt = time.time() - t0
        rays = c.primary + c.shadow + c.reflect + c.refract
        return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                "sample": f"row bands rank 0 of {world} ({c.primary} px, {rays} rays) of the {w}x{h} frame, 1 timed pass after 1 warm-up"}
    except Exception as e:   # a baseline failure must not hide the GPU numbers
        return {"value": None, "unit": "Mrays/s", "cores": cores, "kind": "unavailable", "sample": repr(e)[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)   # frames; with M frames in flight a short run is dominated by pipeline fill and drain
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: how the row tiles reach rank 0 -- one-sided NVLink peer copies on the copy engines (rt_push_rows) or an NCCL gather")
    ap.add_argument("--shard-order", default="serpentine", choices=["serpentine", "modulo"],
                    help="N>1: which rank renders row tile t -- boustrophedon (RT_FLAG_SERPENTINE, evens out the ray-cost gradient down the image) or t %% N")
    ap.add_argument("--pipelines", type=int, default=0, help="launches in flight per GPU (0 = 3; 1 for frames that run the wave kernels)")
    ap.add_argument("--sm-share", type=int, default=-1, help="resident traversal CTAs per SM per pipeline (-1 = 4 when several pipelines share the GPU)")
    ap.add_argument("--batch", type=int, default=0,
                    help="frames per launch (rt_render_batch_async: the frames of a batch share the ray queues); 0 = N, so that a launch always traces one full frame's worth of pixels per GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, cfg)

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
    scene, w, h, level, n, parts, desc = cfg

    tmpdir = f"/tmp/rt_bench_{rank}"
    os.makedirs(tmpdir, exist_ok=True)
    sc = R.Scene(scene, w, h, n, parts, tmpdir=tmpdir)

    def ck(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {R.rt.rt_last_error().decode()}")

    # ---- frame pipelines: one resident scene, M frames in flight ---------------------------------
    # defaults from the sweeps in profiles/r1g_pipelines_sweep_c3.txt: the smaller a GPU's share of the frame,
    # the more frames have to be in flight to keep its SMs busy; frames beyond ~3 M pixels per GPU run
    # the per-level wave kernels, which want the whole GPU each (one pipeline)
    big = w * h // world > 3_000_000
    # frames per launch: about 8 M pixels per launch and GPU (sweeps in profiles/r1i_batch_sweep_c3.txt: launches of
    # that size run the per-level wave kernels with queues long enough that their tails do not matter, and two
    # launches in flight cover each other's level boundaries)
    # Scenes of analytic primitives only (c1, c2) trace 12-17 Grays/s: their launches are bound by the ray-queue
    # traffic, which stays closer to the L2 with one frame per launch; they keep three single frames in flight.
    pix_rank = (w // 64 * 64) * (h // 64 * 64) // world
    mesh = args.config in ("c3", "c4")
    B = args.batch if args.batch > 0 else (1 if big or not mesh else max(1, min(64, round(8_000_000 / max(pix_rank, 1)))))
    M = args.pipelines if args.pipelines > 0 else (1 if big else 2 if B > 1 else 3)
    share = args.sm_share if args.sm_share >= 0 else (0 if M == 1 or B > 1 else 4)
    # e2e leg: RayTracer objects over the one Scene, one frame each per start().  With frame batches (B > 1) on one
    # GPU the tracers run in throughput mode (RayTracer::coalesce): 4 x B of them feed the Scene's three batch
    # workers, which render what is waiting -- B frames per launch, a quarter of the tracers always waiting so
    # that the next launch is ready when one is delivered -- and hand every tracer its frame (C3: 4 191 against
    # 3 616 Mrays/s with one pipeline per tracer).  N > 1 keeps one pipeline per tracer (4 / 4 / 8 in flight at
    # N = 2 / 4 / 8): the one coalesced run on two GPUs (32 tracers, 8 shards per launch) came out at 666 against
    # 6 689 Mrays/s and could not be investigated in this round.  RT_BENCH_COALESCE=0 / k forces it off / on
    # with k x B tracers.
    env_c = os.environ.get("RT_BENCH_COALESCE")
    coalesce = B > 1 and (env_c != "0") and (world == 1 or bool(env_c))
    M_e2e = 1 if big else (min(64, int(env_c or "4") * B) if coalesce else 3 if world == 1 else 4 if world < 8 else 8)
    share_e2e = 0 if M_e2e == 1 or coalesce else 4 if world == 1 else 2 if world < 8 else 1
    main = torch.cuda.current_stream(dev)
    owner = C.c_void_p()
    ck(R.rt.rt_create(local, C.byref(owner)), "rt_create")
    ck(R.rt.rt_set_stream(owner, C.c_void_p(main.cuda_stream)), "rt_set_stream")
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
                landing = FrameLanding(hnd, w, h, rank, world)
                if rank == 0:
                    outs[f] = landing.device_ptr()[0]
            if not (p2p and rank == 0):
                frame = torch.full((h, w, 3), 127, dtype=torch.uint8, device=dev)
                outs[f] = frame.data_ptr()
            frames.append(frame), landings.append(landing)
        pipes.append({"ctx": hnd, "stream": st, "frames": frames, "landings": landings, "outs": outs,
                      "consumer": torch.cuda.Stream(dev) if p2p and rank == 0 else None,   # where the assembled frames become visible
                      "gather": FrameGather(w, h, rank, world, dev, tile_rows, serp) if world > 1 and not p2p else None})

    def launch(j, nb):
        """batch j: nb <= B frames in ONE launch on pipeline j % M, then every frame's rows go to rank 0"""
        p = pipes[j % M]
        ck(R.rt.rt_render_batch_async(p["ctx"], C.byref(params), nb, None, p["outs"]), "rt_render_batch_async")
        for f in range(nb):
            if p["landings"][f] is not None:
                p["landings"][f].push(p["ctx"], p["consumer"], frame=f)   # NVLink P2P: this rank's row tiles -> their place in rank 0's frame (copy engines) + signal
            elif p["gather"] is not None:
                with torch.cuda.stream(p["stream"]):
                    p["gather"].gather(p["frames"][f])   # NCCL: this rank's row tiles -> rank 0, de-interleaved there

    def run(nframes):
        """nframes frames, B per launch (the last launch may hold fewer), round-robin over the M pipelines"""
        j, left = 0, nframes
        while left > 0:
            nb = min(B, left)
            launch(j, nb)
            j, left = j + 1, left - nb
        return j

    def fork():
        ev = torch.cuda.Event()
        ev.record(main)
        for p in pipes:
            p["stream"].wait_event(ev)

    def join():
        for p in pipes:
            for st in (p["stream"], p["consumer"]):
                if st is not None:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    main.wait_event(ev)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- warm-up + one counted launch (ray totals are deterministic per configuration) ----------
    fork()
    run(max(args.warmup, M) * B)                 # whole batches only, every pipeline at least once
    join()
    sync_all()
    cnt = R.Counters()
    ck(R.rt.rt_read_counters(pipes[0]["ctx"], C.byref(cnt)), "rt_read_counters")
    rays_local = (cnt.primary + cnt.shadow + cnt.reflect + cnt.refract) // B    # the frames of a batch are identical here
    launches_per_batch = cnt.launches
    rays_t = torch.tensor([rays_local], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(rays_t)
    rays_total = int(rays_t.item())

    # ---- timed region: exactly K steps, CUDA events bracketing every pipeline stream, max over ranks
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    fork()
    n_launches = run(args.steps)
    join()
    e1.record(main)
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    # per-rank view (load balance of the image-space shards)
    mine_t = torch.tensor([e0.elapsed_time(e1) / args.steps, float(rays_local)], dtype=torch.float64, device=dev)
    per_rank = [torch.zeros_like(mine_t) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine_t)
    else:
        per_rank = [mine_t]
    per_rank = [{"rank": r, "ms_per_step": float(t[0].item()), "rays_per_frame": int(t[1].item())} for r, t in enumerate(per_rank)]
    clocks = sampler.stop() if rank == 0 else None
    value = rays_total * args.steps / (ms_total * 1e-3) / 1e6

    # ---- one launch alone (per-stage split from the library's own CUDA events), one counted launch for the
    #      roofline, and one single frame alone (latency); all untimed ----------------------------------
    p0 = pipes[0]["ctx"]
    ck(R.rt.rt_set_sm_share(p0, 0), "rt_set_sm_share")
    for _ in range(2):
        ck(R.rt.rt_render_batch_async(p0, C.byref(params), B, None, pipes[0]["outs"]), "rt_render_batch_async")
        ck(R.rt.rt_read_counters(p0, C.byref(cnt)), "rt_read_counters")
    stage = {"traverse": cnt.trace_ms, "shade": cnt.shade_ms, "other": cnt.other_ms, "render": cnt.render_ms}
    pstats = R.RenderParams(R.MY_MODEL_RAYTRACE, level, rank, world, R.RT_FLAG_STATS | shard_flags, tile_rows)
    ck(R.rt.rt_render_batch_async(p0, C.byref(pstats), B, None, pipes[0]["outs"]), "rt_render_batch_async(stats)")
    cs = R.Counters()
    ck(R.rt.rt_read_counters(p0, C.byref(cs)), "rt_read_counters")
    c1 = R.Counters()
    for _ in range(2):
        ck(R.rt.rt_render_async(p0, C.byref(params)), "rt_render_async")
        ck(R.rt.rt_read_counters(p0, C.byref(c1)), "rt_read_counters")
    ms_frame_alone = c1.render_ms
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
    tracers = []
    for _ in range(M_e2e):
        t = R.RayTracer(sc, device=local)        # the drop-in surface; tracers of one Scene share its residency
        t.maxLevel = level
        t.smShare = share_e2e if M_e2e > 1 else 0
        t.coalesce = coalesce
        tracers.append(t)
    for k in range(2 * M_e2e):
        tracers[k % M_e2e].wait()
        tracers[k % M_e2e].start(R.MY_MODEL_RAYTRACE, flags=shard_flags, rank=rank, world=world, tile_rows=tile_rows)
    for t in tracers:
        t.wait()
    sync_all()
    t0 = time.perf_counter()
    for k in range(args.steps):
        t = tracers[k % M_e2e]
        t.wait()                                 # frame k-M is in RayTracer::output
        t.start(R.MY_MODEL_RAYTRACE, flags=shard_flags, rank=rank, world=world, tile_rows=tile_rows)
    for t in tracers:
        t.wait()
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = rays_total * args.steps / float(e2e_s.item()) / 1e6
    ce = tracers[0].counters()                   # bytes the library itself copied for the last start()
    h2d, d2h = int(ce.h2d_bytes), int(ce.d2h_bytes)

    if rank == 0:
        peak, peak_src, hbm_peak = sm_peak_fp32_tflops()
        # flop model: 22 per child box (4 per 4-wide node), 47 per triangle test, 23 per analytic primitive
        flops = cs.nodes_visited * 4 * 22 + cs.tri_tests * 47 + cs.prim_tests * 23
        trav_ms = stage["traverse"]
        trav_launches = 1 if cnt.frame_sched else level + 2   # whole-frame scheduler: one traversal launch per frame
        achieved = flops / (trav_ms * 1e-3) / 1e12 if trav_ms > 0 else 0.0
        queue_bytes = rays_local * B * 100   # ~100 B of ray/hit/node records written+read per ray, B frames per launch
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {desc}", "rays_per_frame": rays_total, "pixels": w * (h // 64 * 64) if w % 64 == 0 else (w // 64 * 64) * (h // 64 * 64),
                       "l2_policy": "per-frame working set (ray/hit/node queues + BVH + triangles, > 500 MB touched per frame) exceeds the 126 MB L2; no flush needed",
                       "parallelism": (f"image-space: interleaved {tile_rows}-row tiles over {world} GPUs ({'boustrophedon' if serp else 'modulo'} order), " + ("row tiles pushed into rank 0's frame over NVLink P2P (copy engines, put with signal; NCCL only ships the IPC handles)" if p2p else "NCCL gather of RGB8 tiles to rank 0")) if world > 1 else "single GPU",
                       "frames_per_launch": B, "launches_in_flight": M, "frames_in_flight": B * M, "traversal_ctas_per_sm_per_pipeline": (share if M > 1 and share else 8),
                       "e2e_frames_in_flight": M_e2e, "e2e_coalesced_starts": coalesce,
                       "ms_per_launch_alone": stage["render"], "ms_per_frame_alone": ms_frame_alone},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_step": float(e2e_s.item()) / args.steps * 1e3,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_batch * n_launches,
            "roofline": {"bound": "fp32_issue", "kernel": "k_frame (one persistent launch per frame: closest-hit + shadow traversal of all levels)" if trav_launches == 1 else "k_wave (closest-hit level l fused with shadow any-hit level l-1; the level+2 launches of one batch of frames)", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak, "peak_source": f"148 SMs x 128 lanes x {peak_src} (of measured)",
                         "traffic": NCU_TRAFFIC.get((args.config, world, B), (None, None))[0], "traffic_source": NCU_TRAFFIC.get((args.config, world, B), (None, None))[1],
                         "kernel_launches_per_batch": trav_launches, "frames_per_batch": B, "avg_launch_ms": trav_ms / trav_launches,
                         "flops_per_step": flops / B, "flops_per_launch": flops, "nodes_per_ray": cs.nodes_visited / max(rays_local * B, 1), "tri_tests_per_ray": cs.tri_tests / max(rays_local * B, 1),
                         "stage_ms_one_launch_alone": stage,
                         "hbm_secondary": {"queue_bytes_per_step": queue_bytes, "achieved_gbs": queue_bytes / (stage["render"] * 1e-3) / 1e9,
                                           "peak_gbs": hbm_peak}},
            "clocks": clocks, "per_rank": per_rank,
            "build": {"upload_ms": cs.upload_ms, "lbvh_build_ms": cs.build_ms, "bvh_nodes": cs.bvh_nodes, "bvh_depth": cs.bvh_depth},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, cfg)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
