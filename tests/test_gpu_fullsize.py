"""Full-size (BASELINE.json configs 3 / 4 shape) checks through size-independent properties — the oracle is far too slow at
1920x1080 — on the bench scene (VeachAjar when the asset is prepared, else the synthetic 380 k-triangle room):
  * the two traversal kernels (one ray per thread, persistent queue with dynamic fetch) agree bit for bit on 2 M rays;
  * a 1920x1080 ReSTIR PT film cut into 2 strips (halo exchange through peer pointers) equals the uncut film bit for bit;
  * reservoir invariants: finite non-negative weights, M <= cap, valid samples carry a reconnection vertex id >= 1."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import prepare_assets
import restirpt
from restirpt import GRISSettings, P
from restirpt.multigpu import partition
from common import Backend, FrameDriver, bitwise_mismatch, camera_rays
from test_gpu_strips import _run, HALO

pytestmark = pytest.mark.gpu
W, H = 1920, 1080


@pytest.fixture(scope="module")
def bench_scene(built):
    xml = prepare_assets.ajar_xml()
    return restirpt.HostScene.xml(xml) if xml else restirpt.HostScene.room(380000, 1)


def test_traversal_kernels_agree_on_two_million_rays(bench_scene):
    dev = restirpt.Device(0)
    scene_h = dev.scene(bench_scene.desc)
    cam = bench_scene.camera(W, H)
    o, d = camera_rays(cam, W, H)
    rng = np.random.default_rng(5)
    rays = np.zeros((W * H, 8), dtype=np.float32)
    rays[:, 0:3] = o
    rays[:, 3] = 1e-4
    dirs = d.reshape(-1, 3) + rng.normal(scale=0.3, size=(W * H, 3))    # perturbed primary rays: incoherent within a warp
    rays[:, 4:7] = dirs / np.linalg.norm(dirs, axis=1, keepdims=True)
    rays[:, 7] = 1e7
    res = {}
    for any_hit in (0, 1):
        if any_hit:
            rays[:, 7] = rng.uniform(0.5, 6.0, size=W * H)
        for kernel in (0, 1):
            ms = C.c_float()
            out = np.zeros(W * H, dtype=restirpt.ISEC_DTYPE)
            occ = np.zeros(W * H, dtype=np.uint8)
            restirpt.check(dev.ctx, dev.lib.rpt_trace_bench(dev.ctx, scene_h, rays.ctypes.data_as(P), W * H, any_hit, kernel, 1, C.byref(ms),
                                                            out.ctypes.data_as(P), occ.ctypes.data_as(P)), "rpt_trace_bench")
            res[(any_hit, kernel)] = (out, occ)
    a, b = res[(0, 0)][0], res[(0, 1)][0]
    assert np.array_equal(a["instanceIdx"], b["instanceIdx"]) and np.array_equal(a["triangleIdx"], b["triangleIdx"])
    assert np.array_equal(a["bary"].view(np.uint32), b["bary"].view(np.uint32))
    assert (a["instanceIdx"] != 0xffffffff).mean() > 0.5
    assert np.array_equal(res[(1, 0)][1], res[(1, 1)][1])
    assert 0.02 < res[(1, 0)][1].mean() < 0.98
    dev.lib.rpt_scene_destroy(scene_h)


def test_full_hd_film_strips_and_reservoir_invariants(bench_scene):
    dev = restirpt.Device(0)
    scene_h = dev.scene(bench_scene.desc)
    cam = bench_scene.camera(W, H)
    full = _run(dev, scene_h, [(0, H, 0)], W, H, cam, 2, "gris")[0]
    parts = _run(dev, scene_h, [(r0, r1, HALO) for r0, r1 in partition(H, 2)], W, H, cam, 2, "gris")
    img = np.concatenate([p[0] for p in parts], axis=0)
    res = np.concatenate([p[1] for p in parts], axis=0)
    assert bitwise_mismatch(img, full[0]) == 0
    assert bitwise_mismatch(res, full[1]) == 0
    img, res = full
    assert np.isfinite(img).all() and (img[..., :3] >= 0).all() and (img[..., :3] <= 1e4).all() and img[..., :3].mean() > 0
    w, m = res["resampleWeight"], res["sampleCount"]
    assert np.isfinite(w).all() and (w >= 0).all()
    assert (m <= 20.0).all() and (m >= 0).all()
    valid = res["rcIsec"]["instanceIdx"] != 0xffffffff
    lit = valid & (w > 0)
    assert lit.mean() > 0.2
    assert ((res["flags"][lit] & 0xff) >= 1).all()          # rcVertexId
    assert (((res["flags"][lit] >> 8) & 0xff) <= 15).all()   # pathLength
    dev.lib.rpt_scene_destroy(scene_h)


def test_inline_tail_kernel_equals_wavefront_tail(bench_scene, monkeypatch):
    """The paths alive after bounce 6 run through nine more wavefront rounds on the tail stream (RPT_WAVEFRONT_TAIL=1; the default
    for frames of 1.5 M pixels and more) or are finished by one kernel with in-line traversal (grisTailKernel, RPT_INLINE_TAIL=1;
    the default for smaller frames, where the rounds' latency is on the critical path).  Same stage functions, so the reservoirs must agree bit for bit — at full size, where the tail holds tens
    of thousands of paths."""
    w, h = 960, 540
    gs = GRISSettings(2, 1.0, 1, 1, 20)
    out = {}
    for mode in ("inline", "wavefront"):
        if mode == "inline":
            monkeypatch.setenv("RPT_INLINE_TAIL", "1")
            monkeypatch.delenv("RPT_WAVEFRONT_TAIL", raising=False)
        else:
            monkeypatch.delenv("RPT_INLINE_TAIL", raising=False)
            monkeypatch.setenv("RPT_WAVEFRONT_TAIL", "1")     # (by default the form follows the frame's size)
        dev = restirpt.Device(0)      # (the switches are read when the context is created)
        b = Backend("cuda", bench_scene, w, h, dev)
        drv = FrameDriver(bench_scene.camera(w, h))
        for _ in range(2):
            cur, prev = drv.begin_frame()
            b.set_camera(cur, prev)
            for name in ("gbuffer", "gris_pathtrace", "gris_temporal", "gris_spatial"):
                b.run(name, None if name == "gbuffer" else gs)
            b.flip()
        wc = (C.c_uint32 * 64)()
        dev.lib.rpt_wavefront_counters(b.frame, wc)
        out[mode] = (b.read("GRIS_PREV"), b.read("INDIRECT_OUTPUT"), wc[4 * 7])
        b.close()
        dev.close()
    assert out["inline"][2] > 1000, "the scene must have a tail for this test to mean anything"
    assert bitwise_mismatch(out["inline"][0], out["wavefront"][0]) == 0
    assert bitwise_mismatch(out["inline"][1], out["wavefront"][1]) == 0


@pytest.mark.parametrize("method", ["gris", "gi"])
def test_two_stream_frame_equals_one_stream_frame(bench_scene, monkeypatch, method):
    """By default the frame forks work onto a second stream (the any-hit launch of vertex b-1's shadow rays next to the
    closest-hit launch of bounce b; the spatial pass's replay-list kernel next to its dense shift kernel).
    RPT_TRACE_ONE_STREAM / RPT_SPATIAL_ONE_STREAM put everything back on the frame's stream.  The fork / join events
    carry every dependency, so the two schedules must produce the same bits — at a film size where the kernels really
    run side by side."""
    w, h = 960, 540
    gs = GRISSettings(2, 1.0, 1, 1, 20)
    passes = {"gris": ("gbuffer", "gris_pathtrace", "gris_temporal", "gris_spatial"), "gi": ("gbuffer", "gi_restir")}[method]
    out = {}
    for mode in ("two", "one"):
        for var in ("RPT_TRACE_ONE_STREAM", "RPT_SPATIAL_ONE_STREAM"):
            if mode == "one":
                monkeypatch.setenv(var, "1")
            else:
                monkeypatch.delenv(var, raising=False)
        dev = restirpt.Device(0)      # (the switches are read when the context is created)
        b = Backend("cuda", bench_scene, w, h, dev)
        drv = FrameDriver(bench_scene.camera(w, h))
        for _ in range(3):
            cur, prev = drv.begin_frame()
            b.set_camera(cur, prev)
            for name in passes:
                b.run(name, gs if name.startswith("gris_") else None)
            b.flip()
        out[mode] = (b.read("GRIS_PREV" if method == "gris" else "GI_PREV"), b.read("INDIRECT_OUTPUT"))
        b.close()
        dev.close()
    assert out["two"][1][..., :3].mean() > 0
    assert bitwise_mismatch(out["two"][0], out["one"][0]) == 0
    assert bitwise_mismatch(out["two"][1], out["one"][1]) == 0


def test_shading_from_the_shifts_reconnection_data_equals_replay_shading(bench_scene, monkeypatch):
    """The spatial pass shades a neighbour's winning sample from the reconnection data its shift left in the ShiftTask
    (spatialShadeFromTask, the reference's GRISReconnectionData idea, gris_retrace.glsl:238-320); RPT_NO_SHADE_FROM_TASK=1 replays
    every selected sample again as the shader does (gris_resample_spatial.glsl:88-131).  Same functions on the same operands:
    the film must be the same bits — on VeachAjar, whose glass makes thousands of samples reconnect beyond the first bounce."""
    w, h = 960, 540
    gs = GRISSettings(2, 1.0, 1, 1, 20)
    out = {}
    for mode in ("task", "replay"):
        if mode == "replay":
            monkeypatch.setenv("RPT_NO_SHADE_FROM_TASK", "1")
        else:
            monkeypatch.delenv("RPT_NO_SHADE_FROM_TASK", raising=False)
        dev = restirpt.Device(0)      # (the switches are read when the context is created)
        b = Backend("cuda", bench_scene, w, h, dev)
        drv = FrameDriver(bench_scene.camera(w, h))
        for i in range(3):
            cur, prev = drv.begin_frame(move=(0.003 * i, 0.0, 0.001))
            b.set_camera(cur, prev)
            for name in ("gbuffer", "gris_pathtrace", "gris_temporal", "gris_spatial"):
                b.run(name, None if name == "gbuffer" else gs)
            b.flip()
        out[mode] = (b.read("GRIS_PREV"), b.read("INDIRECT_OUTPUT"))
        b.close()
        dev.close()
    assert out["task"][1][..., :3].mean() > 0
    assert bitwise_mismatch(out["task"][0], out["replay"][0]) == 0
    assert bitwise_mismatch(out["task"][1], out["replay"][1]) == 0


def test_replay_wavefront_equals_inline_replay(bench_scene, monkeypatch):
    """The prefixes the reuse passes have to replay (samples that reconnect beyond the first bounce) run as a wavefront of their
    own — one queue-traversal launch and one rwStepKernel per bounce; RPT_NO_REPLAY_WAVEFRONT=1 replays them one thread per
    prefix with in-line traversal (the list kernels).  Both call replayVertex (gris_retrace.glsl:62-135) in the same order on
    the same operands, so reservoirs and film must be the same bits — on VeachAjar, where thousands of samples per frame take
    this way."""
    w, h = 960, 540
    gs = GRISSettings(2, 1.0, 1, 1, 20)
    out = {}
    for mode in ("wavefront", "inline"):
        if mode == "inline":
            monkeypatch.setenv("RPT_NO_REPLAY_WAVEFRONT", "1")
            monkeypatch.delenv("RPT_RW_MIN_LIST", raising=False)
        else:
            monkeypatch.delenv("RPT_NO_REPLAY_WAVEFRONT", raising=False)
            monkeypatch.setenv("RPT_RW_MIN_LIST", "0")      # (by default the wavefront is only used when the last frame's list was long)
        dev = restirpt.Device(0)      # (the switches are read when the context is created)
        b = Backend("cuda", bench_scene, w, h, dev)
        drv = FrameDriver(bench_scene.camera(w, h))
        for i in range(4):
            cur, prev = drv.begin_frame(move=(0.003 * i, 0.001, 0.0))
            b.set_camera(cur, prev)
            for name in ("gbuffer", "gris_pathtrace", "gris_temporal", "gris_spatial"):
                b.run(name, None if name == "gbuffer" else gs)
            b.flip()
        prevr = b.read("GRIS_PREV")
        rc = (C.c_uint32 * 16)()
        dev.lib.rpt_reuse_counters(b.frame, rc)
        out[mode] = (prevr, b.read("GRIS_TEMP"), b.read("INDIRECT_OUTPUT"), list(rc))
        b.close()
        dev.close()
    flags = out["wavefront"][0]["flags"]
    valid = out["wavefront"][0]["rcIsec"]["instanceIdx"] != 0xffffffff
    assert (valid & ((flags & 0xff) > 1)).sum() > 1000, "the scene must have samples that reconnect beyond the first bounce"
    for k in range(3):
        assert bitwise_mismatch(out["wavefront"][k], out["inline"][k]) == 0, k
    assert out["wavefront"][3][5] > 1000 and out["inline"][3][5] == 0, "rpt_reuse_counters: the replays must have taken the way under test"


def test_overlapped_frames_equal_serial_frames(bench_scene, monkeypatch):
    """Two frames in flight: the reuse passes of frame k (rpt_gris_temporal, rpt_gris_spatial, the post-process) run on the frame's
    late stream set while rpt_gbuffer and rpt_gris_pathtrace of frame k+1 are already running on the frame's stream (capi.cu,
    LateScope; three-slot rotation of G-buffer / motion / GRIS reservoirs, two sets of wavefront state); RPT_NO_FRAME_OVERLAP=1 keeps
    one frame at a time.  Six frames with a camera dolly and NO read or synchronisation in between (any read would join the
    streams), at a film size where the passes really overlap: reservoirs, accumulated film and every tone-mapped image must be
    the same bits either way."""
    w, h = 1920, 1080
    gs = GRISSettings(2, 1.0, 1, 1, 20)
    post = restirpt.PostSettings(1, 1)
    out = {}
    for mode in ("overlap", "serial"):
        if mode == "serial":
            monkeypatch.setenv("RPT_NO_FRAME_OVERLAP", "1")
        else:
            monkeypatch.delenv("RPT_NO_FRAME_OVERLAP", raising=False)
        dev = restirpt.Device(0)
        b = Backend("cuda", bench_scene, w, h, dev)
        drv = FrameDriver(bench_scene.camera(w, h))
        images = [np.zeros((h, w, 4), dtype=np.uint8) for _ in range(6)]
        tickets = []
        for i in range(6):
            cur, prev = drv.begin_frame(move=(0.004 * i, 0.002, 0.0))
            b.set_camera(cur, prev)
            for name in ("gbuffer", "gris_pathtrace", "gris_temporal", "gris_spatial"):
                b.run(name, None if name == "gbuffer" else gs)
            t = C.c_uint64()
            restirpt.check(dev.ctx, dev.lib.rpt_postprocess_async(b.frame, C.byref(post), images[i].ctypes.data_as(restirpt.P), C.byref(t)), "postprocess_async")
            tickets.append(t.value)
            b.flip()
        for t in tickets[-2:]:
            restirpt.check(dev.ctx, dev.lib.rpt_readback_wait(b.frame, t), "readback_wait")
        out[mode] = (b.read("GRIS_PREV"), b.read("GRIS_TEMP"), b.read("INDIRECT_OUTPUT"), images)
        b.close()
        dev.close()
    assert out["overlap"][2][..., :3].mean() > 0
    for k in range(3):
        assert bitwise_mismatch(out["overlap"][k], out["serial"][k]) == 0, k
    for i in range(6):
        assert np.array_equal(out["overlap"][3][i], out["serial"][3][i]), f"tone-mapped image of frame {i}"
        assert out["overlap"][3][i][..., :3].mean() > 1


def test_pipelined_readback_delivers_the_blocking_calls_images(built):
    """Renderer::drawFrameAsync (rpt_postprocess_async: three device images, a copy stream, tickets) against Renderer::drawFrame
    (rpt_postprocess with a host pointer) on the same seeds: the same RGBA8 image for every frame, also when the host collects
    a frame's image only after the next frame has been issued, and when blocking and pipelined calls are mixed"""
    import ctypes as C
    import torch
    from restirpt import P, GRISSettings
    host = restirpt.host_lib()
    sc = restirpt.HostScene.room(6000, 7)
    w, h = 320, 180
    gs = GRISSettings(2, 1.0, 1, 1, 20)

    def renderer():
        r = host.rh_renderer_create(sc.handle, w, h, 0, 0, h, 0)
        assert r, host.rh_last_error()
        host.rh_renderer_set_methods(r, 0, 3, 1, 1, 0)
        host.rh_renderer_set_gris(r, C.byref(gs))
        return r

    n = 7
    ra, rb = renderer(), renderer()
    want = []
    for i in range(n):
        img = np.zeros((h, w, 4), dtype=np.uint8)
        assert host.rh_renderer_draw_frame(ra, restirpt.hash2(50 + i), img.ctypes.data_as(P)) == 0, host.rh_last_error()
        want.append(img)
    pinned = [torch.empty(w * h * 4, dtype=torch.uint8).pin_memory() for _ in range(2)]
    got, ticket, prev = [], C.c_uint64(), None
    for i in range(n):
        if i == 4:   # a blocking frame in between must not disturb the images in flight
            collect = prev
            assert host.rh_renderer_wait_readback(rb, collect) == 0
            got.append(pinned[(i - 1) & 1].numpy().reshape(h, w, 4).copy())
            img = np.zeros((h, w, 4), dtype=np.uint8)
            assert host.rh_renderer_draw_frame(rb, restirpt.hash2(50 + i), img.ctypes.data_as(P)) == 0, host.rh_last_error()
            got.append(img)
            prev = None
            continue
        assert host.rh_renderer_draw_frame_async(rb, restirpt.hash2(50 + i), P(pinned[i & 1].data_ptr()), C.byref(ticket)) == 0, host.rh_last_error()
        if prev is not None:
            assert host.rh_renderer_wait_readback(rb, prev) == 0
            got.append(pinned[(i - 1) & 1].numpy().reshape(h, w, 4).copy())
        prev = ticket.value
    assert host.rh_renderer_wait_readback(rb, prev) == 0
    got.append(pinned[(n - 1) & 1].numpy().reshape(h, w, 4).copy())
    assert len(got) == n
    for i in range(n):
        assert np.array_equal(got[i], want[i]), i
    assert want[-1][..., :3].mean() > 1
    assert host.rh_renderer_wait_readback(rb, 99) != 0      # no such ticket
    host.rh_renderer_destroy(ra); host.rh_renderer_destroy(rb)
