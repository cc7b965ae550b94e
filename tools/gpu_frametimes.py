#!/usr/bin/env python3
"""Per-frame path-trace kernel time over many frames (development aid)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vulkan-restir-pt_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np, torch, restirpt
from restirpt import GRISSettings, P
from common import Backend, FrameDriver
import prepare_assets

mode = sys.argv[1] if len(sys.argv) > 1 else "seq"
sc = restirpt.HostScene.xml(prepare_assets.ajar_xml())
dev = restirpt.Device(0)
w, h = 1920, 1080
b = Backend("cuda", sc, w, h, dev)
stream = torch.cuda.ExternalStream(dev.lib.rpt_frame_stream(b.frame))
drv = FrameDriver(sc.camera(w, h))
gs = GRISSettings(2, 1.0, 1, 1, 20)
evs = []
for i in range(80):
    cur, prev = drv.begin_frame(seed=(12345 if mode == "fixedseed" else None))
    b.set_camera(cur, prev)
    b.run("gbuffer")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    b.run("gris_pathtrace", gs)
    e1.record(stream)
    if mode != "ptonly":
        b.run("gris_temporal", gs); b.run("gris_spatial", gs)
    if mode == "sync":
        dev.lib.rpt_sync(b.frame)
    b.flip()
    evs.append((e0, e1))
dev.lib.rpt_sync(b.frame)
t = [a.elapsed_time(c) for a, c in evs]
print(mode, " ".join(f"{x:.1f}" for x in t))
