#!/usr/bin/env python3
"""What does ONE strip of the 8-GPU 4K film cost on a GPU of its own, with no neighbours to wait for?  Each strip of the balanced
partition of the N = 8 run is rendered standalone (unconnected: its halo rows of the neighbours' reservoirs stay empty, which
changes a few pixels' work but not the cost) on one GPU, frames pipelined as in the bench, and timed — the gap between this and
the 8-GPU frame time is what the hand-over coupling costs; the gap between this and film time / 8 is what small strips cost."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vulkan-restir-pt_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np, torch, restirpt
from restirpt import GRISSettings, PostSettings, P
from common import FrameDriver
import prepare_assets

W, H = 3840, 2160
rows = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] != "one" else "288,276,260,188,160,192,328,468").split(",")]
assert sum(rows) == H
sc = restirpt.HostScene.xml(prepare_assets.ajar_xml())
dev = restirpt.Device(0)
scene = dev.scene(sc.desc)
gs = GRISSettings(2, 1.0, 1, 1, 20)
ps = PostSettings(1, 1, 0, 0)
lib = dev.lib
def run(r0, r1, halo, frames=40, warm=10):
    f = dev.frame(W, H, r0, r1, halo)
    stream = torch.cuda.ExternalStream(lib.rpt_frame_stream(f))
    drv = FrameDriver(sc.camera(W, H))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(warm + frames):
        if i == warm:
            lib.rpt_frame_join(f); e0.record(stream)
        cur, prev = drv.begin_frame()
        restirpt.check(dev.ctx, lib.rpt_set_camera(f, C.byref(cur), C.byref(prev)), "set_camera")
        restirpt.check(dev.ctx, lib.rpt_gbuffer(f, scene), "gbuffer")
        for fn in (lib.rpt_gris_pathtrace, lib.rpt_gris_temporal, lib.rpt_gris_spatial):
            restirpt.check(dev.ctx, fn(f, scene, C.byref(gs)), "gris")
        restirpt.check(dev.ctx, lib.rpt_postprocess(f, C.byref(ps), None), "post")
        lib.rpt_frame_flip(f)
    lib.rpt_frame_join(f); e1.record(stream)
    lib.rpt_sync(f)
    ms = e0.elapsed_time(e1) / frames
    rc = (C.c_uint32 * 16)()
    lib.rpt_reuse_counters(f, rc)
    run.last_lists = (rc[3], rc[5], rc[0])
    lib.rpt_frame_destroy(f)
    return ms
def run_passes(r0, r1, halo, frames=30, warm=8):
    """per-pass device times, one frame at a time (run with RPT_NO_FRAME_OVERLAP=1: every pass then ends on the frame's stream)"""
    f = dev.frame(W, H, r0, r1, halo)
    stream = torch.cuda.ExternalStream(lib.rpt_frame_stream(f))
    drv = FrameDriver(sc.camera(W, H))
    rows_ev = []
    for i in range(warm + frames):
        cur, prev = drv.begin_frame()
        lib.rpt_set_camera(f, C.byref(cur), C.byref(prev))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record(stream)
        lib.rpt_gbuffer(f, scene); ev[1].record(stream)
        lib.rpt_gris_pathtrace(f, scene, C.byref(gs)); ev[2].record(stream)
        lib.rpt_gris_temporal(f, scene, C.byref(gs)); ev[3].record(stream)
        lib.rpt_gris_spatial(f, scene, C.byref(gs)); ev[4].record(stream)
        lib.rpt_postprocess(f, C.byref(ps), None); lib.rpt_frame_join(f); ev[5].record(stream)
        lib.rpt_frame_flip(f)
        if i >= warm: rows_ev.append(ev)
    lib.rpt_sync(f)
    t = np.array([[e[k].elapsed_time(e[k + 1]) for k in range(5)] + [e[0].elapsed_time(e[5])] for e in rows_ev]).mean(0)
    lib.rpt_frame_destroy(f)
    return t
if len(sys.argv) > 3 and sys.argv[1] == "one":   # one strip, a few frames (for a launch list under ncu)
    a_, b_ = int(sys.argv[2]), int(sys.argv[3])
    print(run_passes(a_, b_, 21 if (a_, b_) != (0, H) else 0, frames=3, warm=3))
    sys.exit(0)
if os.environ.get("RPT_NO_FRAME_OVERLAP"):
    for (a, b, h) in ((0, H, 0), (945, 1215, 21), (0, 288, 21), (1692, 2160, 21)):
        t = run_passes(a, b, h)
        print(f"rows {a}..{b} one frame at a time: gbuffer {t[0]:.3f} pathtrace(bounces 0-6) {t[1]:.3f} temporal {t[2]:.3f} spatial {t[3]:.3f} post {t[4]:.3f} frame {t[5]:.3f} ms")
    sys.exit(0)
full = run(0, H, 0, frames=12, warm=4)
print(f"uncut 3840x2160 film on one GPU: {full:.3f} ms per frame; / 8 = {full / 8:.3f} ms   replay pairs in-line / wavefront / shade list: {run.last_lists}")
W, H = 1920, 1080
hd = run(0, H, 0, frames=30, warm=10)
print(f"1920x1080 film: {hd:.3f} ms per frame   replay pairs in-line / wavefront / shade list: {run.last_lists}")
W, H = 3840, 2160
r0 = 0
out = []
for n in rows:
    ms = run(r0, r0 + n, 21)
    out.append(ms)
    print(f"strip rows {r0:4d}..{r0 + n:4d} ({n:3d} rows) standalone: {ms:.3f} ms per frame   replay pairs in-line / wavefront / shade list: {run.last_lists}")
    r0 += n
print(f"slowest strip standalone {max(out):.3f} ms -> {1000 / max(out):.1f} 4K frames/s if nothing else were lost; mean {np.mean(out):.3f} ms")
