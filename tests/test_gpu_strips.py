"""Multi-GPU strip logic on ONE GPU: two (or four) strip frames of the same film in one process, connected as peers,
must reproduce the full-film result bit-for-bit (RNG is a pure function of the global pixel) — with a static camera and,
because the boundary rows of the final reservoirs are mirrored into the neighbours' halo rows, also with a moving one as
long as a reprojection does not jump further than the halo.  The same kernels / peer stores / device-side flags run as in
the one-process-per-GPU deployment; only the transport under the peer pointers differs (same-device memory here, NVLink
there — that path is checked by bench.py on N GPUs: `strip_image_equal`)."""
import ctypes as C

import numpy as np
import pytest

import restirpt
from restirpt import GRISSettings, DISettings, PeerInfo, P
from restirpt.multigpu import partition
from common import FrameDriver, bitwise_mismatch

pytestmark = pytest.mark.gpu
HALO = 21


def _connect(lib, dev, frames):
    infos = []
    for f in frames:
        info = PeerInfo()
        assert lib.rpt_frame_export_peer(f, C.byref(info)) == 0
        infos.append(info)
    for i, f in enumerate(frames):
        up = C.byref(infos[i - 1]) if i > 0 else None
        down = C.byref(infos[i + 1]) if i + 1 < len(frames) else None
        st = lib.rpt_frame_connect_peers(f, up, down)
        assert st == 0, lib.rpt_last_error(dev.ctx)


IMG = {"gris": "INDIRECT_OUTPUT", "di": "DIRECT_OUTPUT", "gi": "INDIRECT_OUTPUT"}
RES = {"gris": "GRIS_PREV", "di": "DI_PREV", "gi": "GI_PREV"}


def _run(dev, scene_h, frames_spec, width, height, cam, nframes, method, moves=None, gather=False):
    """frames_spec: list of (row0, row1, halo).  Returns per-strip (INDIRECT/DIRECT output, final reservoirs) of the owned rows;
    with gather=True additionally the RGBA8 film gathered on the first strip, as the last element."""
    lib = dev.lib
    frames = [dev.frame(width, height, r0, r1, halo) for (r0, r1, halo) in frames_spec]
    if len(frames) > 1:
        _connect(lib, dev, frames)
    if gather:
        info = restirpt.GatherInfo()
        assert lib.rpt_frame_gather_create(frames[0], len(frames), C.byref(info)) == 0, lib.rpt_last_error(dev.ctx)
        for i, f in enumerate(frames):
            assert lib.rpt_frame_gather_connect(f, C.byref(info), i) == 0, lib.rpt_last_error(dev.ctx)
    gs, ds, ps = GRISSettings(2, 1.0, 1, 1, 20), DISettings(0, 0, 1, 1), restirpt.PostSettings(1, 1, 0, 0)
    drv = FrameDriver(cam)
    film = np.zeros((height, width, 4), dtype=np.uint8)
    for i in range(nframes):
        cur, prev = drv.begin_frame(move=None if moves is None else moves[i])
        # stage order matters in one host thread: every strip's temporal pass is enqueued before any spatial pass
        for f in frames:
            assert lib.rpt_set_camera(f, C.byref(cur), C.byref(prev)) == 0
            assert lib.rpt_gbuffer(f, scene_h) == 0
            if method == "gris":
                assert lib.rpt_gris_pathtrace(f, scene_h, C.byref(gs)) == 0
                assert lib.rpt_gris_temporal(f, scene_h, C.byref(gs)) == 0
            elif method == "di":
                assert lib.rpt_di_pathgen(f, scene_h, C.byref(ds)) == 0
                assert lib.rpt_di_temporal(f, scene_h, C.byref(ds)) == 0
            else:
                assert lib.rpt_gi_restir(f, scene_h) == 0
        for f in frames:
            if method == "gris":
                assert lib.rpt_gris_spatial(f, scene_h, C.byref(gs)) == 0
            elif method == "di":
                assert lib.rpt_di_spatial(f, scene_h, C.byref(ds)) == 0
        if gather:
            for f in frames:
                assert lib.rpt_postprocess(f, C.byref(ps), None) == 0
            assert lib.rpt_gather_output(frames[0], film.ctypes.data_as(P)) == 0, lib.rpt_last_error(dev.ctx)
        for f in frames:
            lib.rpt_sync(f)
            lib.rpt_frame_flip(f)
    outs = []
    for f, (r0, r1, halo) in zip(frames, frames_spec):
        assert lib.rpt_frame_peer_error(f) == 0
        b, e = C.c_uint32(), C.c_uint32()
        lib.rpt_frame_rows(f, C.byref(b), C.byref(e))
        img = restirpt.read_buffer(lib, f, restirpt.BUF[IMG[method]], width, e.value - b.value)
        res = restirpt.read_buffer(lib, f, restirpt.BUF[RES[method]], width, e.value - b.value)
        outs.append((img[r0 - b.value: r1 - b.value], res[r0 - b.value: r1 - b.value]))
    for f in frames:
        lib.rpt_frame_gather_disconnect(f)
        lib.rpt_frame_disconnect_peers(f)
    for f in frames:
        lib.rpt_frame_destroy(f)
    if gather:
        outs.append(film)
    return outs


@pytest.mark.parametrize("method", ["gris", "di"])
@pytest.mark.parametrize("strips", [2, 4])
def test_strips_equal_full_film(built, method, strips):
    sc = restirpt.HostScene.room(5000, 11)
    dev = restirpt.Device(0)
    w, h = 160, 120
    scene_h = dev.scene(sc.desc)
    cam = sc.camera(w, h)
    full = _run(dev, scene_h, [(0, h, 0)], w, h, cam, 3, method)[0]
    parts = _run(dev, scene_h, [(r0, r1, HALO) for r0, r1 in partition(h, strips)], w, h, cam, 3, method)
    img = np.concatenate([p[0] for p in parts], axis=0)
    res = np.concatenate([p[1] for p in parts], axis=0)
    assert bitwise_mismatch(img, full[0]) == 0
    assert bitwise_mismatch(res, full[1]) == 0
    assert np.isfinite(img).all() and img[..., :3].mean() > 0
    dev.lib.rpt_scene_destroy(scene_h)


# the camera moves between frames: temporal reuse follows the motion vectors across the cuts.  The final reservoirs of the boundary
# rows are mirrored into the neighbours' halo rows, so the strips still equal the uncut film (ReSTIR GI included: its one pass is
# ordered by its own epoch flags)
MOVES = [(0.0, 0.0, 0.0), (0.02, 0.03, 0.0), (0.02, -0.02, 0.01), (-0.03, 0.02, 0.0), (0.0, 0.0, 0.0)]


@pytest.mark.parametrize("method", ["gris", "di", "gi"])
@pytest.mark.parametrize("strips", [2, 4])
def test_strips_equal_full_film_with_a_moving_camera(built, method, strips):
    sc = restirpt.HostScene.room(5000, 11)
    dev = restirpt.Device(0)
    w, h = 160, 120
    scene_h = dev.scene(sc.desc)
    cam = sc.camera(w, h)
    full = _run(dev, scene_h, [(0, h, 0)], w, h, cam, len(MOVES), method, moves=MOVES)[0]
    parts = _run(dev, scene_h, [(r0, r1, HALO) for r0, r1 in partition(h, strips)], w, h, cam, len(MOVES), method, moves=MOVES)
    img = np.concatenate([p[0] for p in parts], axis=0)
    res = np.concatenate([p[1] for p in parts], axis=0)
    assert bitwise_mismatch(img, full[0]) == 0
    assert bitwise_mismatch(res, full[1]) == 0
    assert (res["sampleCount"] > 1).mean() > 0.3, "temporal history must have been found"


def test_gathered_film_equals_full_film(built):
    """rpt_frame_gather_*: every strip's post-process pass stores its rows into the film image on the first strip's device;
    rpt_gather_output returns the film — equal to the post-processed image of the uncut film"""
    sc = restirpt.HostScene.room(5000, 11)
    dev = restirpt.Device(0)
    w, h = 160, 120
    scene_h = dev.scene(sc.desc)
    cam = sc.camera(w, h)
    full = _run(dev, scene_h, [(0, h, 0)], w, h, cam, 3, "gris", gather=True)
    parts = _run(dev, scene_h, [(r0, r1, HALO) for r0, r1 in partition(h, 4)], w, h, cam, 3, "gris", gather=True)
    assert np.array_equal(full[-1], parts[-1])
    assert parts[-1][..., :3].mean() > 1 and (parts[-1][..., 3] == 255).all()
    dev.lib.rpt_scene_destroy(scene_h)


def test_renderer_draws_in_process_strips_stage_by_stage(built):
    """host Renderer: rh_renderer_draw_frame refuses a strip connected to a neighbour of the same process (its spatial pass would
    wait for a temporal pass that is not enqueued yet); rh_draw_strips drives all strips stage by stage and reproduces the film"""
    host, lib = restirpt.host_lib(), restirpt.device_lib()
    sc = restirpt.HostScene.room(5000, 11)
    w, h = 160, 120
    gs = GRISSettings(2, 1.0, 1, 1, 20)

    def make(r0, r1, halo):
        r = host.rh_renderer_create(sc.handle, w, h, 0, r0, r1, halo)
        assert r, host.rh_last_error()
        host.rh_renderer_set_methods(r, 0, 3, 1, 1, 0)
        host.rh_renderer_set_gris(r, C.byref(gs))
        return r

    full = make(0, h, 0)
    img_full = np.zeros((h, w, 4), dtype=np.uint8)
    for i in range(3):
        assert host.rh_renderer_draw_frame(full, restirpt.hash2(i + 1), img_full.ctypes.data_as(P)) == 0
    bounds = partition(h, 2)
    rs = [make(r0, r1, HALO) for r0, r1 in bounds]
    frames = [P(host.rh_renderer_frame(r)) for r in rs]

    class _Dev:   # (_connect only needs .ctx for error messages)
        ctx = None
    _connect(lib, _Dev, frames)
    assert lib.rpt_frame_peers_in_process(frames[0]) == 1
    assert host.rh_renderer_draw_frame(rs[0], 1, None) != 0 and b"rh_draw_strips" in host.rh_last_error()
    outs = [np.zeros((r1 - r0, w, 4), dtype=np.uint8) for r0, r1 in bounds]
    arr_r = (P * 2)(*rs)
    arr_o = (P * 2)(*[o.ctypes.data for o in outs])
    for i in range(3):
        assert host.rh_draw_strips(arr_r, 2, restirpt.hash2(i + 1), arr_o) == 0, host.rh_last_error()
    assert np.array_equal(np.concatenate(outs, axis=0), img_full)
    for f in frames:
        assert lib.rpt_frame_peer_error(f) == 0
        lib.rpt_frame_disconnect_peers(f)
    for r in rs + [full]:
        host.rh_renderer_destroy(r)


def test_a_missing_neighbour_becomes_an_error_and_reconnecting_recovers(built):
    """A strip whose neighbour is never driven: the device-side wait gives up after its time limit, the condition is sticky —
    every later pass of that frame fails with RPT_ERR_PEER — and disconnecting / reconnecting (epochs and flags restart at
    zero on both sides, whatever number of frames each has rendered) makes the pair work again."""
    sc = restirpt.HostScene.cornell()
    dev = restirpt.Device(0)
    lib = dev.lib
    w, h = 96, 64
    scene_h = dev.scene(sc.desc)
    cam = sc.camera(w, h)
    (a0, a1), (b0, b1) = partition(h, 2)
    fa, fb = dev.frame(w, h, a0, a1, HALO), dev.frame(w, h, b0, b1, HALO)
    gs = GRISSettings(2, 1.0, 1, 1, 20)
    drv = FrameDriver(cam)

    def frame_of(frames, spatial=True):
        cur, prev = drv.begin_frame()
        status = []
        for f in frames:
            lib.rpt_set_camera(f, C.byref(cur), C.byref(prev))
            status += [lib.rpt_gbuffer(f, scene_h), lib.rpt_gris_pathtrace(f, scene_h, C.byref(gs)), lib.rpt_gris_temporal(f, scene_h, C.byref(gs))]
        if spatial:
            for f in frames:
                status.append(lib.rpt_gris_spatial(f, scene_h, C.byref(gs)))
        for f in frames:
            lib.rpt_sync(f)
            lib.rpt_frame_flip(f)
        return status

    # strip A renders two frames alone (unconnected) first, so the two strips have different histories and ping-pong phases
    for _ in range(2):
        assert all(s == 0 for s in frame_of([fa]))
    _connect(lib, dev, [fa, fb])
    frame_of([fa])                       # B is not driven: A's spatial pass waits for B's temporal pass and times out (~4 s)
    assert lib.rpt_frame_peer_error(fa) != 0
    assert lib.rpt_gbuffer(fa, scene_h) == -6 and b"hand-over" in lib.rpt_last_error(dev.ctx)
    for f in (fa, fb):
        assert lib.rpt_frame_disconnect_peers(f) == 0
    _connect(lib, dev, [fa, fb])
    for _ in range(3):
        assert all(s == 0 for s in frame_of([fa, fb]))
    for f in (fa, fb):
        assert lib.rpt_frame_peer_error(f) == 0
        b, e = C.c_uint32(), C.c_uint32()
        lib.rpt_frame_rows(f, C.byref(b), C.byref(e))
        img = restirpt.read_buffer(lib, f, restirpt.BUF["INDIRECT_OUTPUT"], w, e.value - b.value)
        assert np.isfinite(img).all()
    for f in (fa, fb):
        lib.rpt_frame_disconnect_peers(f)
    for f in (fa, fb):
        lib.rpt_frame_destroy(f)
    dev.lib.rpt_scene_destroy(scene_h)
