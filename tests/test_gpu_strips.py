"""Multi-GPU strip logic on ONE GPU: two (or four) strip frames of the same film in one process, connected as peers,
must reproduce the full-film result bit-for-bit for a static camera (RNG is a pure function of the global pixel).
The same kernels / peer stores / device-side flags run as in the one-process-per-GPU deployment; only the transport
under the peer pointers differs (same-device memory here, NVLink there)."""
import ctypes as C

import numpy as np
import pytest

import restirpt
from restirpt import GRISSettings, DISettings, PeerInfo, P
from restirpt.multigpu import partition
from common import FrameDriver, bitwise_mismatch

pytestmark = pytest.mark.gpu
HALO = 21


def _run(dev, scene_h, frames_spec, width, height, cam, nframes, method):
    """frames_spec: list of (row0, row1, halo).  Returns per-strip (INDIRECT/DIRECT output, final reservoirs)."""
    lib = dev.lib
    frames = [dev.frame(width, height, r0, r1, halo) for (r0, r1, halo) in frames_spec]
    if len(frames) > 1:
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
    gs, ds = GRISSettings(2, 1.0, 1, 1, 20), DISettings(0, 0, 1, 1)
    drv = FrameDriver(cam)
    for _ in range(nframes):
        cur, prev = drv.begin_frame()
        # stage order matters in one host thread: every strip's temporal pass is enqueued before any spatial pass
        for f in frames:
            assert lib.rpt_set_camera(f, C.byref(cur), C.byref(prev)) == 0
            assert lib.rpt_gbuffer(f, scene_h) == 0
            if method == "gris":
                assert lib.rpt_gris_pathtrace(f, scene_h, C.byref(gs)) == 0
                assert lib.rpt_gris_temporal(f, scene_h, C.byref(gs)) == 0
            else:
                assert lib.rpt_di_pathgen(f, scene_h, C.byref(ds)) == 0
                assert lib.rpt_di_temporal(f, scene_h, C.byref(ds)) == 0
        for f in frames:
            if method == "gris":
                assert lib.rpt_gris_spatial(f, scene_h, C.byref(gs)) == 0
            else:
                assert lib.rpt_di_spatial(f, scene_h, C.byref(ds)) == 0
        for f in frames:
            lib.rpt_sync(f)
            lib.rpt_frame_flip(f)
    outs = []
    for f, (r0, r1, halo) in zip(frames, frames_spec):
        assert lib.rpt_frame_peer_error(f) == 0
        b, e = C.c_uint32(), C.c_uint32()
        lib.rpt_frame_rows(f, C.byref(b), C.byref(e))
        img_id = restirpt.BUF["INDIRECT_OUTPUT" if method == "gris" else "DIRECT_OUTPUT"]
        res_id = restirpt.BUF["GRIS_PREV" if method == "gris" else "DI_PREV"]
        img = restirpt.read_buffer(lib, f, img_id, width, e.value - b.value)
        res = restirpt.read_buffer(lib, f, res_id, width, e.value - b.value)
        outs.append((img[r0 - b.value: r1 - b.value], res[r0 - b.value: r1 - b.value]))
        lib.rpt_frame_destroy(f)
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
