#!/usr/bin/env python3
"""Feasibility experiment for "two half-films in flight" (DESIGN.md §9): the 1920x1080 bench film as K strips ON ONE GPU, each strip
a frame with its own streams (rh_draw_strips enqueues them stage by stage), against the same film as one frame.  The strips pay a
21-row halo of G-buffer each side and the hand-over flags; what they gain is that one strip's traversal drain / shading kernel can
overlap another strip's traversal.  Prints ms per frame for K = 1, 2, 3, 4."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vulkan-restir-pt_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import restirpt
from restirpt import GRISSettings, PeerInfo, P
from restirpt.multigpu import partition
import prepare_assets

host, lib = restirpt.host_lib(), restirpt.device_lib()
xml = prepare_assets.ajar_xml()
sc = restirpt.HostScene.xml(xml) if xml else restirpt.HostScene.room(380000, 1)
W, H = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1920x1080").split("x"))
gs = GRISSettings(2, 1.0, 1, 1, 20)


def make(r0, r1, halo):
    r = host.rh_renderer_create(sc.handle, W, H, 0, r0, r1, halo)
    assert r, host.rh_last_error()
    host.rh_renderer_set_methods(r, 0, 3, 1, 1, 0)
    host.rh_renderer_set_gris(r, C.byref(gs))
    return r


for k in (1, 2, 3, 4, 1):
    bounds = partition(H, k)
    rs = [make(r0, r1, 21 if k > 1 else 0) for r0, r1 in bounds]
    frames = [P(host.rh_renderer_frame(r)) for r in rs]
    if k > 1:
        infos = []
        for f in frames:
            info = PeerInfo()
            assert lib.rpt_frame_export_peer(f, C.byref(info)) == 0
            infos.append(info)
        for i, f in enumerate(frames):
            up = C.byref(infos[i - 1]) if i > 0 else None
            down = C.byref(infos[i + 1]) if i + 1 < k else None
            assert lib.rpt_frame_connect_peers(f, up, down) == 0
    arr = (P * k)(*rs)

    def draw(i):
        if k == 1:
            assert host.rh_renderer_draw_frame(rs[0], restirpt.hash2(i + 1), None) == 0, host.rh_last_error()
        else:
            assert host.rh_draw_strips(arr, k, restirpt.hash2(i + 1), None) == 0, host.rh_last_error()

    for i in range(20):
        draw(i)
    for f in frames:
        lib.rpt_sync(f)
    t0 = time.perf_counter()
    n = 60
    for i in range(n):
        draw(20 + i)
    for f in frames:
        lib.rpt_sync(f)
    ms = (time.perf_counter() - t0) / n * 1e3
    print(f"{k} strip(s) of {W}x{H} on one GPU: {ms:.3f} ms per frame", flush=True)
    for f in frames:
        if k > 1:
            lib.rpt_frame_disconnect_peers(f)
    for r in rs:
        host.rh_renderer_destroy(r)
