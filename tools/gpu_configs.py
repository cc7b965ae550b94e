#!/usr/bin/env python3
"""Secondary BASELINE.json configurations on one GPU (development aid; bench.py is the contract for config 3):

    python tools/gpu_configs.py di     VeachAjar 1280x720, ReSTIR DI {Reconnection, Light, temporal 1, spatial 1}      (config 2)
    python tools/gpu_configs.py field [subdiv grid]
                                      instanced field (~50 M world-space triangles), 1920x1080, ReSTIR GI temporal   (config 5)
                                      reports BVH build ms, BVH bytes / triangle, node + triangle bytes / ray
    python tools/gpu_configs.py field_tlas [subdiv grid]
                                      config 5's instanced variant: ONE copy of the mesh, BLAS + TLAS (RPT_SCENE_TWO_LEVEL), then the
                                      same scene description flattened; also times rpt_scene_update_instances on both

Prints one JSON line per configuration (frames/s from CUDA events on the frame's stream, counters from one
instrumented, untimed frame)."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vulkan-restir-pt_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import torch
import restirpt
from restirpt import DISettings, GRISSettings, Counters, BvhStats
from common import Backend, FrameDriver, METHOD_PASSES
import prepare_assets


def measure(sc, scene_name, w, h, method, frames, warm, moves=None, time_update=False):
    dev = restirpt.Device(0)
    t0 = time.time()
    b = Backend("cuda", sc, w, h, dev)
    create_s = time.time() - t0
    st = BvhStats()
    dev.lib.rpt_scene_bvh_stats(b.scene, C.byref(st))
    stream = torch.cuda.ExternalStream(dev.lib.rpt_frame_stream(b.frame))
    drv = FrameDriver(sc.camera(w, h))
    settings = {"di": DISettings(0, 0, 1, 1), "gris": GRISSettings(2, 1.0, 1, 1, 20)}
    passes = [("gbuffer", None)] + [(n, settings[k] if k else None) for n, k in METHOD_PASSES[method]]

    def frame(i):
        cur, prev = drv.begin_frame(move=None if moves is None else moves(i))
        b.set_camera(cur, prev)
        for name, s in passes:
            b.run(name, s)
        b.flip()

    for i in range(warm):
        frame(i)
    dev.lib.rpt_sync(b.frame)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(frames):
        frame(warm + i)
    e1.record(stream)
    dev.lib.rpt_sync(b.frame)
    ms = e0.elapsed_time(e1) / frames
    dev.lib.rpt_counters_enable(dev.ctx, 1)
    dev.lib.rpt_counters_reset(dev.ctx)
    frame(warm + frames)
    dev.lib.rpt_sync(b.frame)
    c = Counters()
    dev.lib.rpt_counters_read(dev.ctx, C.byref(c))
    dev.lib.rpt_counters_enable(dev.ctx, 0)
    rays = c.closestRays + c.shadowRays
    alg = 80 * c.nodeVisits + 48 * c.triTests + 48 * rays + 272 * c.shadedHits
    inst = (restirpt.ObjectInstance * sc.desc.numInstances).from_address(sc.desc.instances)
    instanced = sum(i.indexCount // 3 for i in inst) + sc.desc.numTriangleLights
    update_ms = None
    if time_update:
        dev.lib.rpt_sync(b.frame)
        update_ms = []
        for _ in range(3):
            t0 = time.time()
            restirpt.check(dev.ctx, dev.lib.rpt_scene_update_instances(b.scene, sc.desc.instances, sc.desc.numInstances), "rpt_scene_update_instances")
            update_ms.append(round((time.time() - t0) * 1e3, 3))
            dev.lib.rpt_scene_bvh_stats(b.scene, C.byref(st))
            update_ms.append(round(float(st.buildMs), 3))
    line = {
        "scene": scene_name, "film": [w, h], "method": method, "two_level": int(st.twoLevel), "instanced_triangles": instanced,
        "meshes": int(st.numMeshes), "tlas_nodes": int(st.numTlasNodes), "tlas_build_ms": float(st.tlasBuildMs),
        "update_instances_wall_ms": update_ms, "bvh_bytes_per_instanced_triangle": (st.nodeBytes + st.triBytes) / max(instanced, 1),
        "bvh_bytes": int(st.nodeBytes + st.triBytes),
        "triangles": int(st.numTriangles), "bvh_nodes": int(st.numNodes),
        "bvh_build_ms": float(st.buildMs), "scene_create_s": create_s,
        "bvh_bytes_per_triangle": (st.nodeBytes + st.triBytes) / max(st.numTriangles, 1),
        "ms_per_frame": ms, "frames_per_s": 1000.0 / ms, "rays_per_pixel": rays / (w * h), "mrays_per_s": rays / 1e6 / (ms * 1e-3),
        "nodes_per_ray": c.nodeVisits / max(rays, 1), "tris_per_ray": c.triTests / max(rays, 1),
        "node_bytes_per_ray": 80 * c.nodeVisits / max(rays, 1), "tri_bytes_per_ray": 48 * c.triTests / max(rays, 1),
        "algorithmic_gbs": alg / 1e9 / (ms * 1e-3), "frames": frames, "warmup": warm,
    }
    print(json.dumps(line), flush=True)
    b.close()


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "di"
    if which == "di":
        xml = prepare_assets.ajar_xml()
        sc = restirpt.HostScene.xml(xml) if xml else restirpt.HostScene.room(380000, 1)
        measure(sc, "VeachAjar" if xml else "ajar-like room", 1280, 720, "di", 100, 20)
        # scripted dolly: pos += 0.002 * front per frame
        cam = sc.camera(1280, 720)
        front = [0.002 * v for v in cam.front[:3]]
        measure(sc, ("VeachAjar" if xml else "ajar-like room") + " + dolly", 1280, 720, "di", 64, 20, moves=lambda i: front)
    elif which == "field":
        subdiv = int(sys.argv[2]) if len(sys.argv) > 2 else 5
        grid = int(sys.argv[3]) if len(sys.argv) > 3 else 28
        t0 = time.time()
        sc = restirpt.HostScene.field(subdiv, grid, 42)
        print(f"host scene: {sc.num_triangles} triangles in {time.time() - t0:.1f} s", flush=True)
        measure(sc, f"instanced field subdiv {subdiv} grid {grid}", 1920, 1080, "gi", 30, 10)
        measure(sc, f"instanced field subdiv {subdiv} grid {grid}", 1920, 1080, "gris", 30, 10)
    elif which == "field_tlas":
        subdiv = int(sys.argv[2]) if len(sys.argv) > 2 else 5
        grid = int(sys.argv[3]) if len(sys.argv) > 3 else 28
        sc = restirpt.HostScene.field(subdiv, grid, 42, shared=True, two_level=True)
        name = f"instanced field subdiv {subdiv} grid {grid}, shared mesh"
        measure(sc, name + ", BLAS + TLAS", 1920, 1080, "gi", 30, 10, time_update=True)
        measure(sc, name + ", BLAS + TLAS", 1920, 1080, "gris", 30, 10)
        sc.set_two_level(False)
        measure(sc, name + ", flattened", 1920, 1080, "gi", 30, 10, time_update=True)
    else:
        raise SystemExit("usage: gpu_configs.py di | field [subdiv grid] | field_tlas [subdiv grid]")


if __name__ == "__main__":
    main()
