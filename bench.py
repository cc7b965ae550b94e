#!/usr/bin/env python3
"""bench.py — headline benchmark: ReSTIR PT (GRIS, hybrid shift, temporal + spatial reuse) frames/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one frame of the reference's per-frame pipeline (G-buffer -> GRIS path trace -> temporal reuse ->
spatial reuse + shade -> post-process), driven through the C++ host Renderer (reference src/Renderer.cpp drawFrame
sequence) over the C ABI of include/restirpt.h.

Workload (BASELINE.json config 3): VeachAjar (the reference's shipped scene; a synthetic 380 k-triangle stand-in
room when the asset is absent), 1920x1080 per GPU, indirect = ResampledPT {Hybrid, rrScale 1, temporal 1,
spatial 1, cap 20}, direct = None, per-frame seed hash2(frame + 1), static camera.
N > 1: weak scaling — the film grows to N x (1920x1080) pixels (N = 4 is exactly the 3840x2160 film of config 4) and is
split into N horizontal strips, one process per GPU, scene + BVH replicated, temporal-pass reservoirs of the 21
boundary rows pushed into the neighbours' halo rows over NVLink peer memory each frame.  `value` is in
1080p-equivalent frames/s summed over the GPUs (= film frames/s x N).

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
    """N x 1080p pixels: 1 -> 1920x1080, 2 -> 1920x2160, 4 -> 3840x2160 (4K), 8 -> 3840x4320"""
    w, h, k = TILE_W, TILE_H, n_gpus
    grow_h = True
    while k > 1:
        if k % 2:
            raise SystemExit("--gpus must be a power of two")
        if grow_h:
            h *= 2
        else:
            w *= 2
        grow_h = not grow_h
        k //= 2
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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
def oracle_fps(scene, width, height, steps, warmup, threads=0):
    """frames/s of the CPU oracle on a width x height film of the same scene / settings; returns (fps, cores)"""
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
    for i in range(warmup + steps):
        cur, prev = drv.begin_frame()
        t0 = time.perf_counter()
        lib.orc_set_camera(fr, C.byref(cur), C.byref(prev))
        lib.orc_gbuffer(fr, osc)
        lib.orc_gris_pathtrace(fr, osc, C.byref(gs))
        lib.orc_gris_temporal(fr, osc, C.byref(gs))
        lib.orc_gris_spatial(fr, osc, C.byref(gs))
        lib.orc_frame_flip(fr)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    lib.orc_frame_destroy(fr)
    lib.orc_scene_destroy(osc)
    return len(times) / sum(times), (threads or cores), sum(times) / len(times)


def cpu_baseline_leg(scene):
    """The oracle timed on the host cores for the main arm's `cpu_baseline`: the reference arm of this file in a process of
    its own (inside the GPU process the same loop measured 2-3 times slower than alone: profiles/r1_19_final_bench_*.json),
    or, should that fail, the same loop in this process."""
    import subprocess
    try:
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "6", "--warmup", "2"],
                             capture_output=True, text=True, timeout=300, env=env).stdout
        for ln in reversed(out.splitlines()):
            if ln.startswith("{"):
                base = json.loads(ln)["cpu_baseline"]
                if base.get("value", 0) > 0:
                    base["sample"] += ", in a process of its own"
                    return base
    except Exception:   # noqa: BLE001 — any failure of the child falls back to the in-process measurement
        pass
    sw, sh = TILE_W // 4, TILE_H // 4
    cfps, cores, _ = oracle_fps(scene, sw, sh, 6, 2)
    return {"value": cfps * (sw * sh) / (TILE_W * TILE_H), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"6 frames of {sw}x{sh} (1/16 of the 1080p film, same scene/camera/settings) after 2 warm-up frames"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene, scene_name = load_scene()
    # bounded sample: the same scene / camera / settings on a 1/16-area film (480x270); throughput is reported in
    # 1080p-equivalent frames/s = sample frames/s x (480*270)/(1920*1080)
    sw, sh = TILE_W // 4, TILE_H // 4
    fps, cores, sec = oracle_fps(scene, sw, sh, args.steps, args.warmup)
    value = fps * (sw * sh) / (TILE_W * TILE_H)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * sec, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{scene_name}, ReSTIR PT hybrid shift temporal+spatial cap 20, CPU oracle on a {sw}x{sh} "
                               "sample film (1/16 of 1920x1080), value scaled to 1080p-equivalent frames/s"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} frames of {sw}x{sh} (1/16 of the 1080p film) after {args.warmup} warm-up frames"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import restirpt
    from restirpt import GRISSettings, PassStats, Counters, PASS_NAMES, KERNEL_NAMES, P
    from restirpt import multigpu

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    host, dev_lib = restirpt.host_lib(), restirpt.device_lib()
    scene, scene_name = load_scene()
    fw, fh = film_for(world)
    strong = bool(args.film)
    if strong:   # BASELINE.json config 4: a fixed film (3840x2160) cut into N strips
        fw, fh = (int(v) for v in args.film.lower().split("x"))
    halo = HALO if world > 1 else 0
    gs = GRISSettings(2, 1.0, 1, 1, 20)

    def open_strip(row0, row1):
        r = host.rh_renderer_create(scene.handle, fw, fh, local_rank, row0, row1, halo)
        if not r:
            raise SystemExit("renderer creation failed: " + host.rh_last_error().decode())
        host.rh_renderer_set_methods(r, 0, 3, 1, 1, 0)   # direct None, indirect ResampledPT, filmic, gamma, no accumulation
        host.rh_renderer_set_gris(r, C.byref(gs))
        frame = P(host.rh_renderer_frame(r))
        link = multigpu.connect_strips(r, frame, rank, world) if world > 1 else None
        return r, frame, link

    # N > 1: the strips start equal and are re-cut so that every GPU has the same amount of work (the pots and the door
    # cost several times more per row than floor and ceiling).  Calibration = a few untimed frames per round with per-pass
    # device timing (time spent waiting for a neighbour is outside the pass timers), costs all-gathered, boundaries moved
    # to the equal-cost points (multigpu.balanced_partition), strips re-created.  Done before the warm-up, never timed.
    bounds = multigpu.partition(fh, world)
    balance_rounds = 0 if world == 1 or args.no_balance else 3
    balance_log = []
    for round_no in range(balance_rounds + 1):
        r, frame, link = open_strip(*bounds[rank])
        if round_no == balance_rounds:
            break
        import torch.distributed as dist
        for i in range(6):
            if i == 2:
                dev_lib.rpt_sync(frame)
                dev_lib.rpt_frame_timing(frame, 1)
            if host.rh_renderer_draw_frame(r, restirpt.hash2(1000 + i), None) != 0:
                raise SystemExit("draw_frame failed: " + host.rh_last_error().decode())
        st = PassStats()
        dev_lib.rpt_frame_pass_stats(frame, C.byref(st))
        cost = sum(st.ms[i] for i in range(12))
        costs = [None] * world
        dist.all_gather_object(costs, cost)
        balance_log.append({"rows": [b[1] - b[0] for b in bounds], "ms_per_frame": [round(c / 4, 3) for c in costs]})
        bounds = multigpu.balanced_partition(bounds, costs, min_rows=max(2 * HALO, 64))
        link.close()
        host.rh_renderer_destroy(r)
        dist.barrier()
    row0, row1 = bounds[rank]
    rows = row1 - row0
    ctx = P(host.rh_renderer_ctx(r))
    stream = torch.cuda.ExternalStream(dev_lib.rpt_frame_stream(frame), device=torch.device("cuda", local_rank))

    frame_no = [0]

    def draw(out_ptr):
        frame_no[0] += 1
        if host.rh_renderer_draw_frame(r, restirpt.hash2(frame_no[0]), out_ptr) != 0:
            raise SystemExit("draw_frame failed: " + host.rh_last_error().decode())

    def barrier():
        dev_lib.rpt_sync(frame)
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        draw(None)
    barrier()

    # ---- timed region 1: device-resident frames (no read-back), per-pass events enabled -------------------------
    dev_lib.rpt_frame_timing(frame, 1)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        draw(None)
    e1.record(stream)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    stats = PassStats()
    dev_lib.rpt_frame_pass_stats(frame, C.byref(stats))
    dev_lib.rpt_frame_timing(frame, 0)

    # ---- timed region 2: end to end through the host Renderer with the RGBA8 strip read back every frame ---------
    strip_bytes = fw * rows * 4
    pinned = torch.empty(strip_bytes, dtype=torch.uint8).pin_memory()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        draw(P(pinned.data_ptr()))
    e1.record(stream)
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    clock_info = clocks.stop() if clocks else None

    # ---- instrumented (untimed) frame: ray / node / triangle counters per pass for the algorithmic bytes ----------
    counters = {}
    drv_passes = [("gbuffer", None), ("gris_pathtrace", gs), ("gris_temporal", gs), ("gris_spatial", gs)]
    scene_h = P(host.rh_renderer_scene(r))
    dev_lib.rpt_counters_enable(ctx, 1)
    for name, st in drv_passes:
        dev_lib.rpt_counters_reset(ctx)
        fn = getattr(dev_lib, "rpt_" + name)
        fn(frame, scene_h) if st is None else fn(frame, scene_h, C.byref(st))
        dev_lib.rpt_sync(frame)
        c = Counters()
        dev_lib.rpt_counters_read(ctx, C.byref(c))
        counters[name] = c
    dev_lib.rpt_counters_enable(ctx, 0)

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
        launches = int((sum(v["launches_per_frame"] for v in kernels.values()) + 4) * steps)
        timed = {k: v for k, v in kernels.items() if k != "gris_tail"}
        dom = max(timed, key=lambda k: timed[k]["ms_per_frame"])
        # algorithmic bytes: 80 B per CWBVH node + 48 B per triangle a ray must fetch, 48 B ray record in/out,
        # 272 B per shaded hit, plus the kernel's per-pixel stream traffic; counts from the instrumented frame below
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
        # the path tracer's tail (paths alive after bounce 6, run in line by one kernel on a second stream concurrently with the
        # temporal pass) is latency-bound and tiny: reported with its time only
        tail = kernels.pop("gris_tail", None)
        for name, v in kernels.items():
            v["algorithmic_gb_per_frame"] = alg[name] / 1e9
            v["achieved_gbs"] = alg[name] / (v["ms_per_frame"] * 1e-3) / 1e9
            if name in nrays:
                v["mrays_per_s"] = nrays[name] / 1e6 / (v["ms_per_frame"] * 1e-3)
        if tail:
            kernels["gris_tail (concurrent stream)"] = tail
        peak, peak_src = measured_peak_gbs()
        c = {"gbuffer": cg}.get(dom, cp)
        rays = c.closestRays + c.shadowRays
        lpf = kernels[dom]["launches_per_frame"]
        achieved = kernels[dom]["achieved_gbs"]
        total_rays = sum(v.closestRays + v.shadowRays for v in counters.values())
        # DRAM bytes of one launch of the dominant kernel, measured once under ncu --set full (profiles/ncu_traffic.json)
        traffic, traffic_src = None, None
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dom)
            if t:
                traffic, traffic_src = t["bytes_per_launch"], t["source"] + f", a launch of {t['rays_in_launch']} rays"
        except (OSError, ValueError, KeyError):
            pass
        fps = 1000.0 * args.steps / dev_ms
        equiv = (fw * fh) / float(TILE_W * TILE_H)   # 1080p-equivalents per film frame (= N in the default weak-scaling mode)
        line = {
            "metric": METRIC, "value": fps * equiv, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{scene_name}, film {fw}x{fh} ({world} strip(s){'' if world == 1 else ', cost-balanced heights ' + str([b[1] - b[0] for b in bounds])}), direct None, indirect "
                                   "ResampledPT {Hybrid, rrScale 1, temporal 1, spatial 1, cap 20}, seed hash2(frame+1), static camera",
                       "film_frames_per_s": fps, "halo_rows": halo, "strip_balance_rounds": balance_log,
                       "halo_exchange": (link.describe() if link else "none (single GPU)"),
                       "l2_policy": "inputs larger than L2: per-frame working set (G-buffer + 3 reservoir buffers + path state + outputs "
                                    "~1.8 GB at 1080p) exceeds the 126 MB L2; every frame uses a new seed",
                       "mrays_per_s_per_gpu": total_rays / 1e6 * fps,
                       "rays_per_pixel": total_rays / px,
                       "pass_ms": per_pass_ms, "kernels": kernels,
                       "tail_wait_ms_per_frame": (tail_wait["ms_per_frame"] if tail_wait else 0.0)},
            "e2e": {"value": 1000.0 * args.steps / e2e_ms * equiv, "unit": UNIT,
                    "h2d_bytes_per_step": 2 * 352, "d2h_bytes_per_step": strip_bytes},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg[dom] / lpf, "kernel_ms": kernels[dom]["ms_per_frame"] / lpf,
                         "launches_per_frame": lpf,
                         "nodes_per_ray": c.nodeVisits / max(rays, 1), "tris_per_ray": c.triTests / max(rays, 1),
                         "note": "VeachAjar's BVH + triangles (23 MB) are L2-resident: the traversal kernels are instruction-issue / "
                                 "latency-bound, not HBM-bound (profiles/); the fraction is reported against the HBM copy peak as the contract asks"},
            "clocks": clock_info,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(scene)
        print(json.dumps(line), flush=True)

    barrier()   # neighbours may still be storing into this rank's halo rows
    if link:
        if link.error():
            print(f"[rank {rank}] WARNING: a device-side strip hand-over timed out", file=sys.stderr)
        link.close()
    host.rh_renderer_destroy(r)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: keep equal strips (skip the cost calibration)")
    ap.add_argument("--film", default="", help="WxH: fixed film cut into N strips (strong scaling, e.g. 3840x2160 = config 4); "
                                               "default: N x 1920x1080 pixels (weak scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
