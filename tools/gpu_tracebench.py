#!/usr/bin/env python3
"""Traversal microbenchmark: Mrays/s of the two traversal kernels on secondary rays of the bench scene.

Rays: for every pixel of a 1920x1080 G-buffer of VeachAjar, one cosine-distributed bounce ray from the primary hit
(closest-hit test) and one shadow ray towards a uniformly sampled point of a light triangle (any-hit test); in pixel
order (what a warp of the per-pixel kernels sees at bounce 1) and shuffled (deeper bounces)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vulkan-restir-pt_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np
import restirpt
from restirpt import Counters, P
from common import Backend, FrameDriver
import prepare_assets


def secondary_rays(b, sc, cam, w, h, rng):
    dn = b.read("DEPTH_NORMAL")
    depth = dn["f"][..., 0] if dn.dtype.names else dn[..., 0]
    nrm = dn["f"][..., 1:4] if dn.dtype.names else dn[..., 1:4]
    pos, front, right, up = (np.array(getattr(cam, k)[:3], dtype=np.float64) for k in ("pos", "front", "right", "up"))
    ys, xs = np.mgrid[0:h, 0:w]
    u, v = (xs + 0.5) / w, 1.0 - (ys + 0.5) / h
    t = np.tan(np.radians(cam.FOV * 0.5))
    px, py = (u * 2 - 1) * (w / h) * t, (v * 2 - 1) * t
    d = px[..., None] * right + py[..., None] * up + front
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    P0 = pos + d * depth[..., None]
    valid = depth > 0
    n = nrm.astype(np.float64)
    # cosine hemisphere about n
    r1, r2 = rng.random((h, w)), rng.random((h, w))
    phi, st_, ct = 2 * np.pi * r1, np.sqrt(r2), np.sqrt(1 - r2)
    a = np.where(np.abs(n[..., 2:3]) < 0.9, np.array([0, 0, 1.0]), np.array([1.0, 0, 0]))
    tx = np.cross(a, n); tx /= np.linalg.norm(tx, axis=-1, keepdims=True) + 1e-30
    ty = np.cross(n, tx)
    dirs = tx * (st_ * np.cos(phi))[..., None] + ty * (st_ * np.sin(phi))[..., None] + n * ct[..., None]
    rays = np.zeros((h, w, 8), dtype=np.float32)
    rays[..., 0:3] = P0 + dirs * 1e-4
    rays[..., 3] = 1e-4
    rays[..., 4:7] = dirs
    rays[..., 7] = 1e7
    # shadow rays to the lights
    lights = np.ctypeslib.as_array(C.cast(sc.desc.triangleLights, C.POINTER(C.c_float)), shape=(sc.desc.numTriangleLights, 16))
    li = rng.integers(0, lights.shape[0], size=(h, w))
    b1, b2 = rng.random((h, w)), rng.random((h, w))
    flip = b1 + b2 > 1
    b1, b2 = np.where(flip, 1 - b1, b1), np.where(flip, 1 - b2, b2)
    L = lights[li]
    lp = L[..., 0:3] * (1 - b1 - b2)[..., None] + L[..., 4:7] * b1[..., None] + L[..., 8:11] * b2[..., None]
    sd = lp - P0
    dist = np.linalg.norm(sd, axis=-1)
    sh = np.zeros((h, w, 8), dtype=np.float32)
    sh[..., 0:3] = P0
    sh[..., 3] = 1e-4
    sh[..., 4:7] = sd / dist[..., None]
    sh[..., 7] = dist - 1e-4
    return rays[valid], sh[valid]


def main():
    w, h = 1920, 1080
    xml = prepare_assets.ajar_xml()
    sc = restirpt.HostScene.xml(xml) if xml else restirpt.HostScene.room(380000, 1)
    dev = restirpt.Device(0)
    b = Backend("cuda", sc, w, h, dev)
    cam = sc.camera(w, h)
    drv = FrameDriver(cam)
    cur, prev = drv.begin_frame()
    b.set_camera(cur, prev)
    b.run("gbuffer")
    rng = np.random.default_rng(7)
    bounce, shadow = secondary_rays(b, sc, cur, w, h, rng)
    perm = rng.permutation(bounce.shape[0])
    sets = {"bounce/pixel-order": (bounce, 0), "bounce/shuffled": (bounce[perm], 0),
            "shadow/pixel-order": (shadow, 1), "shadow/shuffled": (shadow[perm], 1)}
    if len(sys.argv) > 1 and sys.argv[1] == "sorted":
        # what would sorting the queues buy?  (octant = the three direction signs; cell = Morton code of the origin on a 2^k grid)
        def octant(r):
            return ((r[:, 4] < 0).astype(np.int64) | ((r[:, 5] < 0).astype(np.int64) << 1) | ((r[:, 6] < 0).astype(np.int64) << 2))
        def morton(r, bits):
            o = r[:, 0:3].astype(np.float64)
            lo, hi = o.min(0), o.max(0)
            q = np.minimum(((o - lo) / (hi - lo + 1e-9) * (1 << bits)).astype(np.int64), (1 << bits) - 1)
            m = np.zeros(len(r), dtype=np.int64)
            for i in range(bits):
                for a in range(3):
                    m |= ((q[:, a] >> i) & 1) << (3 * i + a)
            return m
        for tag, rays, any_hit in (("bounce", bounce, 0), ("shadow", shadow, 1), ("bounce-shuffled", bounce[perm], 0)):
            oc = octant(rays)
            sets[f"{tag}/by-octant(stable)"] = (rays[np.argsort(oc, kind="stable")], any_hit)
            sets[f"{tag}/by-octant+cell5"] = (rays[np.lexsort((morton(rays, 5), oc))], any_hit)
            sets[f"{tag}/by-cell5+octant"] = (rays[np.lexsort((oc, morton(rays, 5)))], any_hit)
            sets[f"{tag}/by-cell7+octant"] = (rays[np.lexsort((oc, morton(rays, 7)))], any_hit)
    lib = dev.lib
    for name, (rays, any_hit) in sets.items():
        rays = np.ascontiguousarray(rays)
        n = rays.shape[0]
        res = {}
        for kernel in (0, 1):
            ms = C.c_float(0)
            out = np.zeros(n, dtype=restirpt.ISEC_DTYPE)
            occ = np.zeros(n, dtype=np.uint8)
            restirpt.check(dev.ctx, lib.rpt_trace_bench(dev.ctx, b.scene, rays.ctypes.data_as(P), n, any_hit, kernel, 10, C.byref(ms),
                                                        out.ctypes.data_as(P), occ.ctypes.data_as(P)), "rpt_trace_bench")
            res[kernel] = (ms.value, out, occ)
        same = bool((res[0][1] == res[1][1]).all() and (res[0][2] == res[1][2]).all())
        lib.rpt_counters_enable(dev.ctx, 1); lib.rpt_counters_reset(dev.ctx)
        ms = C.c_float(0)
        lib.rpt_trace_bench(dev.ctx, b.scene, rays.ctypes.data_as(P), n, any_hit, 1, 1, C.byref(ms), None, None)
        c = Counters(); lib.rpt_counters_read(dev.ctx, C.byref(c)); lib.rpt_counters_enable(dev.ctx, 0)
        nr = (c.closestRays + c.shadowRays) or 1
        print(f"{name:32s} n={n/1e6:.2f}M  per-thread {res[0][0]:7.3f} ms = {n/res[0][0]/1e3:7.1f} Mrays/s | queue {res[1][0]:7.3f} ms = "
              f"{n/res[1][0]/1e3:7.1f} Mrays/s | x{res[0][0]/res[1][0]:.2f} | identical={same} | nodes/ray {c.nodeVisits/nr:.1f} tris/ray {c.triTests/nr:.1f}")


if __name__ == "__main__":
    main()
