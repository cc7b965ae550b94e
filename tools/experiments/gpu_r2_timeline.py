#!/usr/bin/env python3
"""Where does the frame's time go when the library's event timing is OFF?  Events on the frame's stream at the pass boundaries
(serial mode: run with RPT_NO_FRAME_OVERLAP=1, where every pass ends on the frame's stream), averaged over the frames."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vulkan-restir-pt_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np, torch, restirpt
from restirpt import GRISSettings, PostSettings, P
from common import Backend, FrameDriver
import prepare_assets

timed = len(sys.argv) > 1 and sys.argv[1] == "timed"
sc = restirpt.HostScene.xml(prepare_assets.ajar_xml())
dev = restirpt.Device(0)
w, h = 1920, 1080
b = Backend("cuda", sc, w, h, dev)
stream = torch.cuda.ExternalStream(dev.lib.rpt_frame_stream(b.frame))
drv = FrameDriver(sc.camera(w, h))
gs = GRISSettings(2, 1.0, 1, 1, 20)
ps = PostSettings()
names = ["gbuffer", "gris_pathtrace", "gris_temporal", "gris_spatial", "postprocess", "frame"]
if timed: dev.lib.rpt_frame_timing(b.frame, 1)
rows = []
N = 60
for i in range(N):
    cur, prev = drv.begin_frame()
    b.set_camera(cur, prev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record(stream)
    b.run("gbuffer"); ev[1].record(stream)
    b.run("gris_pathtrace", gs); ev[2].record(stream)
    b.run("gris_temporal", gs); ev[3].record(stream)
    b.run("gris_spatial", gs); ev[4].record(stream)
    dev.lib.rpt_postprocess(b.frame, C.byref(ps), None); dev.lib.rpt_frame_join(b.frame); ev[5].record(stream)
    b.flip()
    rows.append(ev)
dev.lib.rpt_sync(b.frame)
t = np.array([[r[k].elapsed_time(r[k + 1]) for k in range(5)] + [r[0].elapsed_time(r[5])] for r in rows[20:]])
gaps = np.array([rows[i][5].elapsed_time(rows[i + 1][0]) for i in range(20, N - 1)])
print("timed" if timed else "untimed", os.environ.get("RPT_NO_FRAME_OVERLAP", "overlap"), " ".join(f"{n} {v:.3f}" for n, v in zip(names, t.mean(0))), f"gap {gaps.mean():.3f}",
      f"first-to-last {rows[20][0].elapsed_time(rows[N-1][5]) / (N - 20):.3f} ms/frame")
