#!/usr/bin/env python3
"""bench.py — headline benchmark: ReSTIR PT (GRIS, hybrid shift, temporal + spatial reuse) frames/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one frame of the reference's per-frame pipeline (G-buffer -> GRIS path trace -> temporal reuse ->
spatial reuse + shade -> post-process), driven through the C++ host Renderer (reference src/Renderer.cpp drawFrame
sequence) over the C ABI of include/restirpt.h.

Workload (BASELINE.json config 3): VeachAjar (the reference's shipped scene; a synthetic 380 k-triangle stand-in
room when the asset is absent), 1920x1080 per GPU, indirect = ResampledPT {Hybrid, rrScale 1, temporal 1,
spatial 1, cap 20}, direct = None, per-frame seed hash2(frame + 1), static camera.
N > 1: weak scaling — the film shows the same view with N x (1920x1080) pixels (2712x1526, 3840x2160 = the film of config 4,
5432x3056) and is split into N horizontal strips, one process per GPU, scene + BVH replicated, temporal-pass reservoirs of the 21
boundary rows pushed into the neighbours' halo rows over NVLink peer memory each frame.  `value` is in
1080p-equivalent frames/s summed over the GPUs (= film frames/s x film pixels / 1080p pixels).

`--impl reference` times the CPU oracle (oracle/liboracle.so: the C++ restatement of the reference shaders, all
host threads) on a bounded sample of the same workload — the reference itself cannot run here (Windows-only build,
Vulkan ray tracing).
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
for p in (ROOT, os.path.join(ROOT, "vulkan-restir-pt_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

TILE_W, TILE_H = 1920, 1080
HALO = 21          # ResampleRadius 20 px (gris_resample_spatial.glsl:66) + 1 bilinear tap
METRIC = "ReSTIR PT (GRIS hybrid shift, temporal+spatial) frames/s, 1920x1080 per GPU"
UNIT = "frames/s (1080p-equivalent)"


def film_for(n_gpus):
    """The SAME view at N x the pixels of 1920x1080 (16:9 kept, so every GPU count renders the same picture and a pixel costs the
    same at every N): 1 -> 1920x1080, 2 -> 2712x1526, 4 -> 3840x2160 (the 4K film of config 4), 8 -> 5432x3056 (width = multiple of
    8 nearest to 1920 sqrt(N)).  The pixel count is N x 1080p to within 0.7 %; `value` scales by the exact ratio."""
    if n_gpus < 1:
        raise SystemExit("--gpus must be positive")
    w = 8 * int(round(TILE_W * n_gpus ** 0.5 / 8.0))
    h = 2 * int(round(w * TILE_H / float(TILE_W) / 2.0))
    return w, h


def load_scene():
    import restirpt
    import prepare_assets
    xml = prepare_assets.ajar_xml()
    if xml:
        return restirpt.HostScene.xml(xml), "VeachAjar (reference res/model/VeachAjar.zip, 382690 triangles)"
    return restirpt.HostScene.room(380000, 1), "synthetic ajar-like room (380k triangles; VeachAjar asset absent)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------------
# CPU oracle legs (test infrastructure used only as the reported baseline / reference arm)
# ---------------------------------------------------------------------------------------------------------------
WORKLOAD_FMT = ("{scene}, film {fw}x{fh} ({strips}), direct None, indirect ResampledPT {{Hybrid, rrScale 1, temporal 1, spatial 1, cap 20}}, "
                "seed hash2(frame+1), static camera")


def oracle_fps(scene, width, height, steps, warmup, threads=0, budget_s=None):
    """frames/s of the CPU oracle on a width x height film of the same scene / settings, all host threads.
    Returns (fps from the MEDIAN frame time, threads, median seconds per frame).  budget_s: give up (return None) when the first
    frame says the whole run would take longer."""
    from restirpt import GRISSettings, P
    from common import FrameDriver
    from oracle import binding
    lib = binding.oracle_lib()
    cores = lib.orc_set_threads(threads)
    osc = P(lib.orc_scene_create(C.byref(scene.desc)))
    fr = P(lib.orc_frame_create(width, height))
    gs = GRISSettings(2, 1.0, 1, 1, 20)
    drv = FrameDriver(scene.camera(width, height))
    times = []
    result = None
    for i in range(warmup + steps):
        cur, prev = drv.begin_frame()
        t0 = time.perf_counter()
        lib.orc_set_camera(fr, C.byref(cur), C.byref(prev))
        lib.orc_gbuffer(fr, osc)
        lib.orc_gris_pathtrace(fr, osc, C.byref(gs))
        lib.orc_gris_temporal(fr, osc, C.byref(gs))
        lib.orc_gris_spatial(fr, osc, C.byref(gs))
        lib.orc_frame_flip(fr)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        elif i == 0 and budget_s is not None and dt * (warmup + steps) > budget_s:
            break
    else:
        med = float(np.median(times))
        result = (1.0 / med, (threads or cores), med)
    lib.orc_frame_destroy(fr)
    lib.orc_scene_destroy(osc)
    return result


def oracle_sample(scene, steps, warmup, budget_s):
    """The reference arm's measurement: the CPU oracle on the bench workload itself (the 1920x1080 film) when the host cores
    finish `warmup + steps` frames of it inside the budget, else on a 1/4- or 1/16-area film of the same scene / camera /
    settings, scaled to 1080p-equivalent frames/s.  Returns (value, cores, seconds per sample frame, description)."""
    for div in (1, 2, 4):
        sw, sh = TILE_W // div, TILE_H // div
        got = oracle_fps(scene, sw, sh, steps, warmup, budget_s=None if div == 4 else budget_s)
        if got is not None:
            fps, cores, sec = got
            share = (sw * sh) / float(TILE_W * TILE_H)
            what = (f"{steps} frames of the {sw}x{sh} film itself" if div == 1 else
                    f"{steps} frames of a {sw}x{sh} film (1/{div * div} of the 1080p film, same scene / camera / settings), scaled by the pixel count")
            return fps * share, cores, sec, what + f" after {warmup} warm-up frames; median frame time"


def cpu_baseline_leg(scene):
    """The oracle timed on the host cores for the main arm's `cpu_baseline`: the reference arm of this file in a process of
    its own (inside the GPU process the same loop measured 2-3 times slower than alone: profiles/r1_19_final_bench_*.json),
    or, should that fail, the same loop in this process."""
    import subprocess
    try:
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "6", "--warmup", "2", "--ref-budget", "30"],
                             capture_output=True, text=True, timeout=300, env=env).stdout
        for ln in reversed(out.splitlines()):
            if ln.startswith("{"):
                base = json.loads(ln)["cpu_baseline"]
                if base.get("value", 0) > 0:
                    base["sample"] += ", in a process of its own"
                    return base
    except Exception:   # noqa: BLE001 — any failure of the child falls back to the in-process measurement
        pass
    value, cores, _, what = oracle_sample(scene, 6, 2, 30.0)
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": what}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene, scene_name = load_scene()
    # the bench workload itself — VeachAjar, the 1920x1080 film, the same settings and seeds — when the host cores finish the
    # asked number of frames within ~4 minutes (16 cores: ~1.5 s per frame), else a smaller film of it, scaled
    value, cores, sec, what = oracle_sample(scene, args.steps, max(args.warmup, 1), args.ref_budget)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * sec, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32",
        "data": "reference asset (VeachAjar, res/model/VeachAjar.zip of the reference repository)" if "VeachAjar" in scene_name and "absent" not in scene_name else "synthetic",
        "config": {"workload": WORKLOAD_FMT.format(scene=scene_name, fw=TILE_W, fh=TILE_H, strips="1 strip(s)"),
                   "implementation": "CPU oracle (C++ restatement of the reference shaders, oracle/), all host threads; the reference "
                                     "itself cannot be built or run here (Win32 host + Vulkan ray queries)",
                   "sample": what},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": what},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------------------------
class Arm:
    """One process = one GPU = one strip of the film.  Holds the libraries, the scene and the torch.distributed plumbing."""

    def __init__(self, args):
        import torch
        import restirpt
        self.torch, self.restirpt, self.args = torch, restirpt, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if self.world != args.gpus:
            if self.world == 1 and args.gpus > 1:
                raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
            args.gpus = self.world
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU oracle)")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.host, self.dev_lib = restirpt.host_lib(), restirpt.device_lib()
        self.scene, self.scene_name = load_scene()
        self.gs = restirpt.GRISSettings(2, 1.0, 1, 1, 20)

    # ---- plumbing -------------------------------------------------------------------------------------------------
    def open_strip(self, fw, fh, row0, row1, connect=True):
        from restirpt import multigpu, P
        halo = HALO if self.world > 1 else 0
        r = self.host.rh_renderer_create(self.scene.handle, fw, fh, self.local_rank, row0, row1, halo)
        if not r:
            raise SystemExit("renderer creation failed: " + self.host.rh_last_error().decode())
        self.host.rh_renderer_set_methods(r, 0, 3, 1, 1, 0)   # direct None, indirect ResampledPT, filmic, gamma, no accumulation
        self.host.rh_renderer_set_gris(r, C.byref(self.gs))
        frame = P(self.host.rh_renderer_frame(r))
        link = multigpu.connect_strips(r, frame, self.rank, self.world) if (self.world > 1 and connect) else None
        return r, frame, link

    def close_strip(self, r, link):
        if link:
            if link.error():
                print(f"[rank {self.rank}] WARNING: a device-side strip hand-over timed out", file=sys.stderr)
            link.close()
        self.host.rh_renderer_destroy(r)
        if self.dist:
            self.dist.barrier()

    def barrier(self, frame):
        self.dev_lib.rpt_sync(frame)
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if not self.dist:
            return ms
        t = self.torch.tensor([ms], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    # ---- one measurement of a film --------------------------------------------------------------------------------
    def measure(self, fw, fh, steps, warmup, want_clocks=False, bounds=None):
        """Cuts the fw x fh film into `world` cost-balanced strips, warms up, then times `steps` device-resident frames and
        `steps` end-to-end frames (camera upload + RGBA8 strip read back to pinned memory every frame).  Times are CUDA-event
        times on the frame's stream, max over ranks."""
        torch, restirpt, host, dev_lib = self.torch, self.restirpt, self.host, self.dev_lib
        from restirpt import multigpu, PassStats, P
        world, rank = self.world, self.rank
        # N > 1: the strips start equal and are re-cut so that every GPU has the same amount of work (the pots and the door
        # cost several times more per row than floor and ceiling).  Calibration = a few untimed frames per round with per-pass
        # device timing (time spent waiting for a neighbour is outside the pass timers), costs all-gathered, boundaries moved
        # to the equal-cost points (multigpu.balanced_partition), strips re-created.  Done before the warm-up, never timed.
        balance_log = []
        if bounds is None:
            bounds = multigpu.partition(fh, world)
            rounds = 0 if world == 1 or self.args.no_balance else 4
            for _ in range(rounds):
                # The cost of a strip = its frame time on a GPU of its own, frames pipelined as in the timed regions, NOT connected to
                # its neighbours (their halo rows stay empty: a few pixels' work changes, the cost does not) — what the strip would
                # do if it never had to wait.  (The sum of the per-pass device times of a connected strip, used before, weighs the
                # latency chains of the reuse passes in full although the pipeline hides them: strips balanced that way differed by
                # 12 % in their pipelined frame times, profiles/r2_23_*.)
                r, frame, link = self.open_strip(fw, fh, *bounds[rank], connect=False)
                cal_stream = torch.cuda.ExternalStream(dev_lib.rpt_frame_stream(frame), device=torch.device("cuda", self.local_rank))
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                cal_frames = 8
                for i in range(3 + cal_frames):
                    if i == 3:
                        dev_lib.rpt_frame_join(frame)
                        c0.record(cal_stream)
                    if host.rh_renderer_draw_frame(r, restirpt.hash2(1000 + i), None) != 0:
                        raise SystemExit("draw_frame failed: " + host.rh_last_error().decode())
                dev_lib.rpt_frame_join(frame)
                c1.record(cal_stream)
                dev_lib.rpt_sync(frame)
                cost = c0.elapsed_time(c1) / cal_frames
                costs = [None] * world
                self.dist.all_gather_object(costs, cost)
                balance_log.append({"rows": [b[1] - b[0] for b in bounds], "ms_per_frame": [round(c, 3) for c in costs]})
                bounds = multigpu.balanced_partition(bounds, costs, min_rows=max(2 * HALO, 64))
                self.close_strip(r, link)
        r, frame, link = self.open_strip(fw, fh, *bounds[rank])
        row0, row1 = bounds[rank]
        rows = row1 - row0
        stream = torch.cuda.ExternalStream(dev_lib.rpt_frame_stream(frame), device=torch.device("cuda", self.local_rank))
        frame_no = [0]

        def draw(out_ptr):
            frame_no[0] += 1
            if host.rh_renderer_draw_frame(r, restirpt.hash2(frame_no[0]), out_ptr) != 0:
                raise SystemExit("draw_frame failed: " + host.rh_last_error().decode())

        for _ in range(max(warmup, 3)):
            draw(None)
        self.barrier(frame)

        # ---- timed region 1: device-resident frames (no read-back) ----------------------------------------------------
        clocks = ClockSampler(self.local_rank) if (want_clocks and rank == 0) else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier(frame)
        e0.record(stream)
        for _ in range(steps):
            draw(None)
        dev_lib.rpt_frame_join(frame)      # (the last frame's spatial + post-process passes run on the frame's second stream set)
        e1.record(stream)
        self.barrier(frame)
        dev_ms = self.max_over_ranks(e0.elapsed_time(e1))
        # (the clocks are sampled during the region above only: every nvidia-smi query holds a driver lock for a few milliseconds, which
        # the device-resident region does not feel — the host runs frames ahead — but the end-to-end regions below, where the host
        # waits for an image every frame, lost 5-10 % in one run out of two to it, profiles/r2_25_*)
        clock_info = clocks.stop() if clocks else None
        # the per-pass / per-kernel breakdown comes from a second run of the same frames with the library's event timing on (about
        # 40 timing events per frame, which cost a little and are therefore kept out of the region above)
        # and ONE FRAME AT A TIME (rpt_frame_join after every frame: the next frame's G-buffer and path tracer then wait for this frame's
        # reuse passes instead of running next to them), so that a kernel's time is its own and not that of whatever shared the GPU
        dev_lib.rpt_frame_timing(frame, 1)
        for _ in range(steps):
            draw(None)
            dev_lib.rpt_frame_join(frame)
        self.barrier(frame)
        stats = PassStats()
        dev_lib.rpt_frame_pass_stats(frame, C.byref(stats))
        dev_lib.rpt_frame_timing(frame, 0)

        # ---- timed region 2: end to end through the host Renderer with the RGBA8 strip read back every frame -----------
        # Every frame: camera upload (2 x 352 B, H2D) and the tone-mapped RGBA8 strip read back to pinned host memory (D2H).
        # The read-back is pipelined (Renderer::drawFrameAsync): frame i's copy runs on a copy stream while frame i+1 renders;
        # the host collects frame i-3's image before it issues frame i (whose image goes to the same host buffer), and the LAST images before the region ends.
        strip_bytes = fw * rows * 4
        pinned = [torch.empty(strip_bytes, dtype=torch.uint8).pin_memory() for _ in range(3)]
        ticket = C.c_uint64()

        def draw_async(i):
            frame_no[0] += 1
            if host.rh_renderer_draw_frame_async(r, restirpt.hash2(frame_no[0]), P(pinned[i % 3].data_ptr()), C.byref(ticket)) != 0:
                raise SystemExit("draw_frame_async failed: " + host.rh_last_error().decode())
            return ticket.value

        def collect(t):
            if host.rh_renderer_wait_readback(r, t) != 0:
                raise SystemExit("wait_readback failed: " + host.rh_last_error().decode())

        collect(draw_async(0))     # untimed: the first pipelined read-back creates the copy stream and the device images
        self.barrier(frame)
        e0.record(stream)
        tickets = []
        for i in range(steps):
            if i >= 3:
                collect(tickets[i - 3])      # (frame i is about to reuse that image's host buffer; the host stays two frames ahead of the
            tickets.append(draw_async(i))    #  image it waits for: the reuse passes of frame i-1 run behind the path tracer of frame i)
        for t in tickets[-3:]:
            collect(t)         # the last images are in host memory before the end event is recorded
        e1.record(stream)
        self.barrier(frame)
        e2e_ms = self.max_over_ranks(e0.elapsed_time(e1))
        # the same with the blocking call (rpt_postprocess with a host pointer: copy on the frame's stream + synchronise)
        self.barrier(frame)
        e0.record(stream)
        for _ in range(steps):
            draw(P(pinned[0].data_ptr()))
        dev_lib.rpt_frame_join(frame)
        e1.record(stream)
        self.barrier(frame)
        e2e_blocking_ms = self.max_over_ranks(e0.elapsed_time(e1))
        return {"r": r, "frame": frame, "link": link, "bounds": bounds, "rows": rows, "dev_ms": dev_ms, "e2e_ms": e2e_ms, "e2e_blocking_ms": e2e_blocking_ms,
                "stats": stats, "strip_bytes": strip_bytes, "clocks": clock_info, "balance_log": balance_log, "film": (fw, fh)}

    # ---- N > 1: is the image the strips make the image one GPU makes? ------------------------------------------------
    def check_strips(self, fw, fh, bounds):
        """"N-GPU image == 1-GPU image bit-for-bit" on the real thing — N processes, N GPUs, CUDA-IPC peer
        stores, device-side epoch flags.  Fresh strips render 2 frames with a static camera and 2 with a dolly (temporal reuse
        then follows motion vectors across the cuts: the mirrored final reservoirs are exercised).  Every frame's RGBA8 rows are
        gathered on rank 0's GPU (rpt_frame_gather_*: the post-process kernels store straight into the film image there) and
        compared with rank 0's own render of the uncut film; the float output and the final reservoirs are compared through
        SHA-256 digests of each strip's rows."""
        import hashlib
        restirpt, host, dev_lib, dist = self.restirpt, self.host, self.dev_lib, self.dist
        from restirpt import GatherInfo, P
        rank, world = self.rank, self.world
        moves = [None, None, (0.0, 0.0, 0.002), (0.001, 0.0005, 0.002)]

        def digests(frame, rows_lo, rows_hi, store_lo):
            out = []
            for name in ("INDIRECT_OUTPUT", "GRIS_PREV"):
                b, e = C.c_uint32(), C.c_uint32()
                dev_lib.rpt_frame_rows(frame, C.byref(b), C.byref(e))
                a = restirpt.read_buffer(dev_lib, frame, restirpt.BUF[name], fw, e.value - b.value)
                out.append(hashlib.sha256(np.ascontiguousarray(a[rows_lo - store_lo: rows_hi - store_lo]).tobytes()).hexdigest())
            return out

        r, frame, link = self.open_strip(fw, fh, *bounds[rank])
        info = GatherInfo()
        if rank == 0:
            st = dev_lib.rpt_frame_gather_create(frame, world, C.byref(info))
            if st != 0:
                raise SystemExit("rpt_frame_gather_create failed: " + dev_lib.rpt_last_error(None).decode())
        blobs = [bytes(info)]
        dist.broadcast_object_list(blobs, src=0)
        info = GatherInfo.from_buffer_copy(blobs[0])
        if dev_lib.rpt_frame_gather_connect(frame, C.byref(info), rank) != 0:
            raise SystemExit("rpt_frame_gather_connect failed: " + dev_lib.rpt_last_error(None).decode())
        dist.barrier()
        films, strip_digests = [], []
        film = np.zeros((fh, fw, 4), dtype=np.uint8)
        for i, mv in enumerate(moves):
            if mv is not None:
                host.rh_renderer_camera_move(r, (C.c_float * 3)(*mv))
            if host.rh_renderer_draw_frame(r, restirpt.hash2(7000 + i), None) != 0:
                raise SystemExit("draw_frame failed: " + host.rh_last_error().decode())
            if rank == 0:
                if dev_lib.rpt_gather_output(frame, film.ctypes.data_as(P)) != 0:
                    raise SystemExit("rpt_gather_output failed: " + dev_lib.rpt_last_error(None).decode())
                films.append(film.copy())
            b, e = C.c_uint32(), C.c_uint32()
            dev_lib.rpt_frame_rows(frame, C.byref(b), C.byref(e))
            strip_digests.append(digests(frame, bounds[rank][0], bounds[rank][1], b.value))
        all_digests = [None] * world
        dist.all_gather_object(all_digests, strip_digests)
        dist.barrier()
        dev_lib.rpt_frame_gather_disconnect(frame)
        dist.barrier()
        self.close_strip(r, link)

        result = None
        if rank == 0:
            # the same four frames on the uncut film, on this rank's GPU alone
            r1 = host.rh_renderer_create(self.scene.handle, fw, fh, self.local_rank, 0, fh, 0)
            if not r1:
                raise SystemExit("renderer creation failed: " + host.rh_last_error().decode())
            host.rh_renderer_set_methods(r1, 0, 3, 1, 1, 0)
            host.rh_renderer_set_gris(r1, C.byref(self.gs))
            f1 = P(host.rh_renderer_frame(r1))
            img = np.zeros((fh, fw, 4), dtype=np.uint8)
            same_img, same_buf, diff_px = [], [], []
            for i, mv in enumerate(moves):
                if mv is not None:
                    host.rh_renderer_camera_move(r1, (C.c_float * 3)(*mv))
                if host.rh_renderer_draw_frame(r1, restirpt.hash2(7000 + i), img.ctypes.data_as(P)) != 0:
                    raise SystemExit("draw_frame failed: " + host.rh_last_error().decode())
                same_img.append(bool(np.array_equal(img, films[i])))
                diff_px.append(int(np.any(img != films[i], axis=-1).sum()))
                same_buf.append(all(digests(f1, lo, hi, 0) == all_digests[k][i] for k, (lo, hi) in enumerate(bounds)))
            host.rh_renderer_destroy(r1)
            result = {"static_camera": bool(all(same_img[:2]) and all(same_buf[:2])),
                      "moving_camera": bool(all(same_img[2:]) and all(same_buf[2:])),
                      "frames": len(moves), "rgba8_pixels_differing_per_frame": diff_px,
                      "float_output_and_reservoir_digests_equal_per_frame": same_buf,
                      "how": "N strips on N GPUs (CUDA-IPC peer stores, device-side epoch flags), film gathered on GPU 0 by "
                             "rpt_gather_output, compared bit for bit with the uncut film rendered on GPU 0 alone"}
        if self.dist:
            self.dist.barrier()
        return result


def ncu_reference(kernel):
    """ncu figures of one launch of `kernel`, regenerated by tools/ncu_traffic.py from the round's .ncu-rep captures
    (profiles/ncu_traffic.json; the capture command is in its header)"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(kernel)
    except (OSError, ValueError):
        return None


def run_cuda(args):
    arm = Arm(args)
    torch, restirpt, host, dev_lib = arm.torch, arm.restirpt, arm.host, arm.dev_lib
    from restirpt import PassStats, Counters, PASS_NAMES, KERNEL_NAMES, P
    rank, world, gs, scene = arm.rank, arm.world, arm.gs, arm.scene
    fw, fh = film_for(world)
    strong = bool(args.film)
    if strong:   # a fixed film cut into N strips as the headline (the default run reports config 4 as config.strong_4k instead)
        fw, fh = (int(v) for v in args.film.lower().split("x"))
    halo = HALO if world > 1 else 0

    m = arm.measure(fw, fh, args.steps, args.warmup, want_clocks=True)
    r, frame, link, bounds, rows = m["r"], m["frame"], m["link"], m["bounds"], m["rows"]
    dev_ms, e2e_ms, stats, strip_bytes, clock_info, balance_log = m["dev_ms"], m["e2e_ms"], m["stats"], m["strip_bytes"], m["clocks"], m["balance_log"]
    ctx = P(host.rh_renderer_ctx(r))

    # ---- instrumented (untimed) frame: ray / node / triangle counters per pass for the algorithmic bytes ----------
    counters = {}
    drv_passes = [("gbuffer", None), ("gris_pathtrace", gs), ("gris_temporal", gs), ("gris_spatial", gs)]
    scene_h = P(host.rh_renderer_scene(r))
    dev_lib.rpt_counters_enable(ctx, 1)   # (every rank issues the same passes, so connected strips stay in lock step)
    for name, st in drv_passes:
        dev_lib.rpt_counters_reset(ctx)
        fn = getattr(dev_lib, "rpt_" + name)
        fn(frame, scene_h) if st is None else fn(frame, scene_h, C.byref(st))
        dev_lib.rpt_sync(frame)
        c = Counters()
        dev_lib.rpt_counters_read(ctx, C.byref(c))
        counters[name] = c
    dev_lib.rpt_counters_enable(ctx, 0)

    # memory roofs measured live on this GPU: L2 (a 24 MB buffer = the size of VeachAjar's BVH + triangles) and HBM reads (4 GB)
    l2_gbs, l2_ns, hbm_read_gbs, hbm_ns = C.c_float(), C.c_float(), C.c_float(), C.c_float()
    if rank == 0:
        dev_lib.rpt_membench(ctx, 24 << 20, 200, C.byref(l2_gbs), C.byref(l2_ns))
        dev_lib.rpt_membench(ctx, 4 << 30, 2, C.byref(hbm_read_gbs), C.byref(hbm_ns))
    arm.barrier(frame)   # neighbours may still be storing into this rank's halo rows
    arm.close_strip(r, link)

    # ---- BASELINE.json config 4: the fixed 3840x2160 film cut into N strips (strong scaling), every N ------------------
    strong_4k = None
    if not strong and not args.no_4k:
        if (fw, fh) == (3840, 2160):
            m4 = m
        else:
            m4 = arm.measure(3840, 2160, max(args.steps // 2, 10), max(args.warmup // 2, 5))
            arm.barrier(m4["frame"])
            arm.close_strip(m4["r"], m4["link"])
        steps4 = args.steps if m4 is m else max(args.steps // 2, 10)
        strong_4k = {"film": "3840x2160", "frames_per_s": 1000.0 * steps4 / m4["dev_ms"], "ms_per_frame": m4["dev_ms"] / steps4,
                     "e2e_frames_per_s": 1000.0 * steps4 / m4["e2e_ms"], "steps": steps4,
                     "strip_rows": [b[1] - b[0] for b in m4["bounds"]],
                     "note": "BASELINE.json config 4 (fixed film, N cost-balanced strips): divide by the N=1 line's figure for the "
                             "strong-scaling efficiency"}

    # ---- N > 1: strips == uncut film, on the real transport --------------------------------------------------------
    strip_check = None
    if world > 1 and not args.no_check:
        cw, ch = (fw, fh)
        strip_check = arm.check_strips(cw, ch, bounds)

    rc = 0
    if rank == 0:
        px = fw * rows
        steps = args.steps
        per_pass_ms = {PASS_NAMES[i]: stats.ms[i] / max(stats.launches[i], 1) for i in range(12) if stats.launches[i]}
        # kernels: single-kernel passes as they are; the wavefront path-tracing pass split into its kernels (events between
        # the launches on the frame's stream; its tail, which runs concurrently on a second stream, is not in these figures)
        kernels = {}
        for name in ("gbuffer", "postprocess"):
            i = PASS_NAMES.index(name)
            if stats.launches[i]:
                kernels[name] = {"ms_per_frame": stats.ms[i] / steps, "launches_per_frame": stats.launches[i] / steps}
        for k, name in enumerate(KERNEL_NAMES):
            if stats.kernelLaunches[k]:
                kernels[name] = {"ms_per_frame": stats.kernelMs[k] / steps, "launches_per_frame": stats.kernelLaunches[k] / steps}
        tail_wait = kernels.pop("tail_wait", None)   # not a kernel: the frame's stream waiting for the path tracer's tail
        # The path tracer's traversal launches: bounce 1's extension rays alone ("trace_closest" spans), then per bounce the
        # extension rays of bounce b and the shadow rays of vertex b-1 side by side on two streams ("trace_pair" spans: two
        # launches each).  Reported as one entry, "trace_paths"; "trace_any" is then the reuse passes' visibility launches.
        pair = kernels.pop("trace_pair", None)
        if pair:
            first = kernels.pop("trace_closest")
            kernels = {"trace_paths": {"ms_per_frame": first["ms_per_frame"] + pair["ms_per_frame"],
                                       "launches_per_frame": first["launches_per_frame"] + 2 * pair["launches_per_frame"]}, **kernels}
        # (the spatial pass's pick / shift / shift-list kernels share the reuse_gen span and its merge / shade-list / redo-list
        # kernels the reuse_merge span: + 4 launches per frame)
        # (+ the temporal pass's replay-list kernel and the tail's temporal step: 2; + the path tracer's tail as wavefront rounds on
        # frames of 1.5 M pixels and more: bounces 7..15 = 26 launches instead of the one counted above)
        wavefront_tail = fw * (m["rows"]) >= 1500000
        launches = int((sum(v["launches_per_frame"] for v in kernels.values()) + 4 + 2 + (25 if wavefront_tail else 0)) * steps)
        timed = {k: v for k, v in kernels.items() if k != "gris_tail"}
        dom = max(timed, key=lambda k: timed[k]["ms_per_frame"])
        tail = kernels.pop("gris_tail", None)
        roofline = None
        total_rays = sum(v.closestRays + v.shadowRays for v in counters.values())
        if True:
            # algorithmic bytes: 80 B per CWBVH node + 48 B per triangle a ray must fetch, 48 B ray record in/out,
            # 272 B per shaded hit, plus the kernel's per-pixel stream traffic; counts from the instrumented frame above
            cp, cs, ct, cg = (counters[k] for k in ("gris_pathtrace", "gris_spatial", "gris_temporal", "gbuffer"))

            def ray_bytes(c, kind):
                if kind == "closest":
                    return 80 * (c.nodeVisits - c.shadowNodeVisits) + 48 * (c.triTests - c.shadowTriTests) + 48 * c.closestRays
                if kind == "any":
                    return 80 * c.shadowNodeVisits + 48 * c.shadowTriTests + 33 * c.shadowRays
                return 80 * c.nodeVisits + 48 * c.triTests + 48 * (c.closestRays + c.shadowRays)

            state_bytes = 11 * 16 * 2 + 16 + 8 + 1 + 64   # path state planes in + out, hit, pixel ids, visibility byte, two ray records

            def closest_bytes(c):
                return ray_bytes(c, "closest") - 48 * c.closestRays   # in-line rays: no ray record in memory

            alg = {
                "gbuffer": ray_bytes(cg, "all") + 272 * cg.shadedHits + px * (28 + 16),
                "postprocess": px * 36,
                "trace_closest": ray_bytes(cp, "closest"),
                "trace_any": (0 if pair else ray_bytes(cp, "any")) + ray_bytes(ct, "any") + ray_bytes(cs, "any"),
                "trace_paths": ray_bytes(cp, "closest") + ray_bytes(cp, "any"),
                "gris_begin": px * (24 + state_bytes // 2 + 36),
                "gris_bounce": 272 * cp.shadedHits + cp.closestRays * state_bytes + px * 96,
                # gen: G-buffer + candidate reservoirs in, shift task (7 x 16 B) + visibility ray (32 B) out per candidate, the
                # in-line replay rays and their surface fetches; merge: tasks + reservoirs in, reservoir (+ radiance RMW) out
                "reuse_gen": closest_bytes(ct) + closest_bytes(cs) + 272 * (ct.shadedHits + cs.shadedHits)
                             + px * ((24 + 4 + 24 + 96 + 144) + (24 + 3 * (24 + 96) + 3 * 144)),
                "reuse_merge": px * ((112 + 96 + 96 + 1 + 96) + (96 + 3 * (112 + 96 + 1) + 96 + 32)),
            }
            nrays = {"gbuffer": cg.closestRays + cg.shadowRays, "trace_closest": cp.closestRays,
                     "trace_any": (0 if pair else cp.shadowRays) + ct.shadowRays + cs.shadowRays,
                     "trace_paths": cp.closestRays + cp.shadowRays}
            for name, v in kernels.items():
                v["algorithmic_gb_per_frame"] = alg[name] / 1e9
                v["achieved_gbs"] = alg[name] / (v["ms_per_frame"] * 1e-3) / 1e9
                if name in nrays:
                    v["mrays_per_s"] = nrays[name] / 1e6 / (v["ms_per_frame"] * 1e-3)
            peak, peak_src = measured_peak_gbs()
            c = {"gbuffer": cg}.get(dom, cp)
            rays = c.closestRays + c.shadowRays
            lpf = kernels[dom]["launches_per_frame"]
            achieved = kernels[dom]["achieved_gbs"]
            kernel_ms = kernels[dom]["ms_per_frame"] / lpf
            # ncu figures of the dominant kernel's largest launch (profiles/ncu_traffic.json, regenerated by tools/ncu_traffic.py
            # from this round's captures): DRAM and L2 bytes, issue-slot utilisation, active lanes per instruction
            ref = ncu_reference(dom) or {}
            l2_frac = issue_frac = None
            if ref.get("lts_bytes_per_launch") and ref.get("duration_us") and l2_gbs.value > 0:
                l2_frac = ref["lts_bytes_per_launch"] / (ref["duration_us"] * 1e-6) / 1e9 / l2_gbs.value
            if ref.get("issue_active_pct") and ref.get("threads_per_inst"):
                issue_frac = ref["issue_active_pct"] / 100.0 * ref["threads_per_inst"] / 32.0
            roofline = {
                "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ref.get("dram_bytes_per_launch"), "traffic_source": ref.get("source"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom] / lpf, "kernel_ms": kernel_ms, "launches_per_frame": lpf,
                "kernel_ms_note": "span of the kernel's launches in the frame / launches (launches on two streams overlap); "
                                  "the ncu duration of the largest launch is in `ncu`",
                "nodes_per_ray": c.nodeVisits / max(rays, 1), "tris_per_ray": c.triTests / max(rays, 1),
                # roofs that apply to an L2-resident scene: measured live on this GPU by rpt_membench
                "l2": {"peak_gbs": l2_gbs.value, "peak_how": "rpt_membench: all SMs read a 24 MB buffer 200 times with 16-byte loads (ld.global.cg)",
                       "latency_ns": l2_ns.value, "algorithmic_frac": achieved / l2_gbs.value if l2_gbs.value > 0 else None,
                       "ncu_lts_frac": l2_frac},
                "hbm_read_gbs": hbm_read_gbs.value, "hbm_latency_ns": hbm_ns.value,
                "issue": {"frac": issue_frac, "how": "ncu issue-slot utilisation x active lanes per instruction / 32: the share of the "
                                                       "SMs' lane-issue capacity doing work — the roof this kernel is actually under"},
                "ncu": ref or None,
                "note": "VeachAjar's BVH + triangles (23 MB) are L2-resident and mostly L1-hit: the traversal kernels are bound by "
                        "instruction issue, not by HBM or L2 bandwidth; `frac` is the nominal algorithmic-bytes figure the contract asks for",
            }
        if tail:
            kernels["gris_tail (concurrent stream)"] = tail
        fps = 1000.0 * args.steps / dev_ms
        equiv = (fw * fh) / float(TILE_W * TILE_H)   # 1080p-equivalents per film frame (= N in the default weak-scaling mode)
        line = {
            "metric": METRIC, "value": fps * equiv, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32",
            "data": "reference asset (VeachAjar, res/model/VeachAjar.zip of the reference repository)" if "VeachAjar" in arm.scene_name and "absent" not in arm.scene_name else "synthetic",
            "config": {"workload": WORKLOAD_FMT.format(scene=arm.scene_name, fw=fw, fh=fh, strips=f"{world} strip(s)" + ("" if world == 1 else ", cost-balanced heights " + str([b[1] - b[0] for b in bounds]))),
                       "film_frames_per_s": fps, "halo_rows": halo, "strip_balance_rounds": balance_log,
                       "halo_exchange": ("temporal kernels store boundary rows of the temp reservoirs, spatial kernels those of the final "
                                         "reservoirs, into the neighbours' halo rows through CUDA-IPC peer memory (NVLink); hand-over ordered "
                                         "by device-side epoch flags; no per-frame collective") if world > 1 else "none (single GPU)",
                       "l2_policy": "inputs larger than L2: per-frame working set (G-buffer + 3 reservoir buffers + path state + outputs "
                                    "~1.8 GB at 1080p) exceeds the 126 MB L2; every frame uses a new seed",
                       "mrays_per_s_per_gpu": total_rays / 1e6 * fps,
                       "rays_per_pixel": total_rays / px,
                       "pass_ms": per_pass_ms, "kernels": kernels,
                       "kernel_timing": "pass_ms / kernels / roofline.kernel_ms: CUDA events around every launch in a second run of the same frames, ONE FRAME AT "
                                        "A TIME (rpt_frame_join after each), so a kernel's time is its own; the headline regions run two frames in flight "
                                        "(reuse passes of frame k next to G-buffer + path tracer of frame k+1) and are therefore shorter than the sum of pass_ms",
                       "tail_wait_ms_per_frame": (tail_wait["ms_per_frame"] if tail_wait else 0.0),
                       "strong_4k": strong_4k, "strip_image_equal": strip_check},
            "e2e": {"value": 1000.0 * args.steps / e2e_ms * equiv, "unit": UNIT,
                    "h2d_bytes_per_step": 2 * 352, "d2h_bytes_per_step": strip_bytes,
                    "how": "host Renderer, camera upload + RGBA8 strip read back to pinned host memory every frame; the read-back is "
                           "pipelined (copy stream, three device images, three host buffers): frame i-3's image is collected before frame i is "
                           "issued, the last ones inside the timed region",
                    "blocking_readback_value": 1000.0 * args.steps / m["e2e_blocking_ms"] * equiv},
            "gpu_launches": launches,
            "roofline": roofline,
            "clocks": dict(clock_info or {}, region="sampled every 50 ms during the device-resident timed region (`value`); the end-to-end regions run right after it"),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(scene)
        print(json.dumps(line), flush=True)
        if strip_check and not (strip_check["static_camera"] and strip_check["moving_camera"]):
            print("bench.py: the image made by the strips differs from the uncut film (config.strip_image_equal)", file=sys.stderr)
            rc = 3
    if arm.dist:
        arm.dist.barrier()
        arm.dist.destroy_process_group()
    if rc:
        sys.exit(rc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget", type=float, default=240.0, help="--impl reference: seconds the whole run may take on the full 1920x1080 "
                    "film before a smaller sample film is used instead")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: keep equal strips (skip the cost calibration)")
    ap.add_argument("--no-4k", action="store_true", help="skip the second timed region on the fixed 3840x2160 film (config.strong_4k)")
    ap.add_argument("--no-check", action="store_true", help="N > 1: skip the strips-vs-uncut-film image comparison")
    ap.add_argument("--film", default="", help="WxH: fixed film cut into N strips (strong scaling, e.g. 3840x2160 = config 4); "
                                               "default: N x 1920x1080 pixels (weak scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
