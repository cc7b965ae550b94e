#!/usr/bin/env python3
"""Quick per-pass timing of the GRIS frame on one GPU (development aid; bench.py is the contract)."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vulkan-restir-pt_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np
import torch
import restirpt
from restirpt import GRISSettings, PostSettings, DISettings, Counters, BvhStats, P
from common import Backend, FrameDriver
import prepare_assets


def main():
    w, h = int(sys.argv[1]) if len(sys.argv) > 1 else 1920, int(sys.argv[2]) if len(sys.argv) > 2 else 1080
    scene_name = sys.argv[3] if len(sys.argv) > 3 else "ajar"
    frames = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    xml = prepare_assets.ajar_xml()
    if scene_name == "ajar" and xml:
        sc = restirpt.HostScene.xml(xml)
    elif scene_name == "cornell":
        sc = restirpt.HostScene.cornell()
    else:
        sc = restirpt.HostScene.room(380000, 1)
    print("scene", scene_name, "tris", sc.num_triangles)
    dev = restirpt.Device(0)
    t0 = time.time()
    b = Backend("cuda", sc, w, h, dev)
    st = BvhStats()
    dev.lib.rpt_scene_bvh_stats(b.scene, C.byref(st))
    print(f"scene_create {time.time()-t0:.3f}s  bvh build {st.buildMs:.2f} ms nodes {st.numNodes} nodeMB {st.nodeBytes/1e6:.1f} triMB {st.triBytes/1e6:.1f}")
    stream = torch.cuda.ExternalStream(dev.lib.rpt_frame_stream(b.frame))
    cam = sc.camera(w, h)
    drv = FrameDriver(cam)
    gs = GRISSettings(2, 1.0, 1, 1, 20)
    passes = [("gbuffer", None), ("gris_pathtrace", gs), ("gris_temporal", gs), ("gris_spatial", gs)]
    acc = {p[0]: 0.0 for p in passes}
    total = 0.0
    warm = 5
    for i in range(frames + warm):
        cur, prev = drv.begin_frame()
        b.set_camera(cur, prev)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(passes) + 1)]
        with torch.cuda.stream(stream):
            evs[0].record(stream)
            for k, (name, s) in enumerate(passes):
                b.run(name, s)
                evs[k + 1].record(stream)
        dev.lib.rpt_sync(b.frame)
        b.flip()
        if i >= warm:
            for k, (name, s) in enumerate(passes):
                acc[name] += evs[k].elapsed_time(evs[k + 1])
            total += evs[0].elapsed_time(evs[-1])
    print(f"{w}x{h} GRIS: {total/frames:.3f} ms/frame = {1000*frames/total:.1f} fps")
    for k, v in acc.items():
        print(f"  {k:16s} {v/frames:8.3f} ms")
    # ray counters on one more frame
    dev.lib.rpt_counters_enable(dev.ctx, 1)
    dev.lib.rpt_counters_reset(dev.ctx)
    cur, prev = drv.begin_frame()
    b.set_camera(cur, prev)
    for name, s in passes:
        b.run(name, s)
    dev.lib.rpt_sync(b.frame)
    c = Counters()
    dev.lib.rpt_counters_read(dev.ctx, C.byref(c))
    dev.lib.rpt_counters_enable(dev.ctx, 0)
    rays = c.closestRays + c.shadowRays
    print(f"rays/frame {rays/1e6:.2f} M (closest {c.closestRays/1e6:.2f} shadow {c.shadowRays/1e6:.2f}) rays/px {rays/(w*h):.2f} "
          f"nodes/ray {c.nodeVisits/max(rays,1):.1f} tris/ray {c.triTests/max(rays,1):.1f} hits {c.shadedHits/1e6:.2f} M  max nodes of one queued ray {c.maxNodeVisits}")
    wc = (C.c_uint32 * 64)()
    dev.lib.rpt_wavefront_counters(b.frame, wc)
    print("wavefront queue slots per bounce (each = path state + extension ray + the previous vertex's shadow ray):",
          " ".join(f"{b_}:{wc[4*b_]}" for b_ in range(1, 15)))
    print(f"Mrays/s {rays/1e6/(total/frames/1000):.1f}")
    img = b.postprocess(PostSettings(1, 1, 1, 0))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    restirpt.host_lib().rh_write_png(os.path.join(ROOT, "gpurun_out", f"quick_{scene_name}.png").encode(), img.ctypes.data_as(P), w, h)


if __name__ == "__main__":
    main()
