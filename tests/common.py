"""Shared test harness: drives the CUDA library (rpt_*) and the CPU oracle (orc_*) through the same calls."""
import ctypes as C

import numpy as np

import restirpt
from restirpt import BUF, BUF_DTYPE, Camera, DISettings, GRISSettings, PostSettings, P
from oracle import binding


class Backend:
    """One scene + one full-film frame on either implementation."""

    def __init__(self, kind, host_scene, width, height, device=None):
        self.kind = kind
        self.w, self.h = width, height
        if kind == "cuda":
            self.lib = restirpt.device_lib()
            self.prefix = "rpt"
            self.dev = device or restirpt.Device(0)
            self.scene = self.dev.scene(host_scene.desc)
            self.frame = self.dev.frame(width, height)
        elif kind == "oracle":
            self.lib = binding.oracle_lib()
            self.prefix = "orc"
            self.scene = P(self.lib.orc_scene_create(C.byref(host_scene.desc)))
            self.frame = P(self.lib.orc_frame_create(width, height))
        else:
            raise ValueError(kind)

    def _fn(self, name):
        return getattr(self.lib, f"{self.prefix}_{name}")

    def _ck(self, status, what):
        if self.kind == "cuda":
            restirpt.check(self.dev.ctx, status, what)

    def set_camera(self, cur, prev):
        self._ck(self._fn("set_camera")(self.frame, C.byref(cur), C.byref(prev)), "set_camera")

    def run(self, name, settings=None):
        fn = self._fn(name)
        if settings is None:
            self._ck(fn(self.frame, self.scene), name)
        else:
            self._ck(fn(self.frame, self.scene, C.byref(settings)), name)

    def postprocess(self, settings):
        out = np.zeros((self.h, self.w, 4), dtype=np.uint8)
        self._ck(self._fn("postprocess")(self.frame, C.byref(settings), out.ctypes.data_as(P)), "postprocess")
        return out

    def flip(self):
        self._fn("frame_flip")(self.frame)

    def clear(self):
        self._fn("frame_clear")(self.frame)

    def read(self, buf):
        buf_id = BUF[buf] if isinstance(buf, str) else buf
        return restirpt.read_buffer(self.lib, self.frame, buf_id, self.w, self.h, self.prefix)

    def trace_closest(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        n = rays.shape[0]
        out = np.zeros(n, dtype=restirpt.ISEC_DTYPE)
        if self.kind == "cuda":
            self._ck(self.lib.rpt_trace_closest(self.dev.ctx, self.scene, rays.ctypes.data_as(P), n, out.ctypes.data_as(P)), "trace_closest")
        else:
            self.lib.orc_trace_closest(self.scene, rays.ctypes.data_as(P), n, out.ctypes.data_as(P))
        return out

    def trace_shadow(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        n = rays.shape[0]
        out = np.zeros(n, dtype=np.uint8)
        if self.kind == "cuda":
            self._ck(self.lib.rpt_trace_shadow(self.dev.ctx, self.scene, rays.ctypes.data_as(P), n, out.ctypes.data_as(P)), "trace_shadow")
        else:
            self.lib.orc_trace_shadow(self.scene, rays.ctypes.data_as(P), n, out.ctypes.data_as(P))
        return out

    def close(self):
        self._fn("frame_destroy")(self.frame)
        self._fn("scene_destroy")(self.scene)


METHOD_PASSES = {
    # (direct, indirect) method names -> pass list, as Renderer::drawFrame sequences them
    "naive": [("di_naive", None), ("gi_naive", None)],
    "naive_rt": [("di_naive_rt", None), ("gi_naive", None)],   # RayTracing-pipeline mode: di_naive.rgen
    "di": [("di_pathgen", "di"), ("di_temporal", "di"), ("di_spatial", "di")],
    "gi": [("gi_restir", None)],
    "gris": [("gris_pathtrace", "gris"), ("gris_temporal", "gris"), ("gris_spatial", "gris")],
}


class FrameDriver:
    """Python mirror of the headless Renderer's per-frame camera handling (reference src/Renderer.cpp:358-368,
    654-660), used to drive both backends with identical inputs."""

    def __init__(self, camera, accumulate=False):
        self.host = restirpt.host_lib()
        self.cam = camera.copy()
        self.prev = camera.copy()
        self.accumulate = accumulate
        self.frame_no = 0
        self.clear_next = False

    def begin_frame(self, seed=None, move=None):
        if move is not None:
            d = (C.c_float * 3)(*move)
            self.host.rh_camera_move(C.byref(self.cam), d)
        if not self.accumulate:
            self.host.rh_camera_update(C.byref(self.cam))
        if self.clear_next:
            self.cam.frameIndex = 0x80000000
            self.clear_next = False
        self.cam.seed = restirpt.hash2(self.frame_no + 1) if seed is None else seed
        cur, prev = self.cam.copy(), self.prev.copy()
        self.prev = self.cam.copy()
        self.host.rh_camera_next_frame(C.byref(self.cam), self.cam.seed)
        self.frame_no += 1
        return cur, prev


def run_frames(backend, camera, method, frames, di=None, gris=None, moves=None, accumulate=False, snapshot=None):
    """Run `frames` frames of gbuffer + method passes; `snapshot(frame_idx, pass_name, backend)` is called after
    every pass."""
    driver = FrameDriver(camera, accumulate)
    settings = {"di": di or DISettings(0, 0, 1, 1), "gris": gris or GRISSettings(2, 1.0, 1, 1, 20)}
    for i in range(frames):
        cur, prev = driver.begin_frame(move=None if moves is None else moves[i])
        backend.set_camera(cur, prev)
        backend.run("gbuffer")
        if snapshot:
            snapshot(i, "gbuffer", backend)
        for name, skey in METHOD_PASSES[method]:
            backend.run(name, settings[skey] if skey else None)
            if snapshot:
                snapshot(i, name, backend)
        backend.flip()


def bitwise_mismatch(a, b):
    """Number of array elements (pixels) whose bytes differ."""
    av = a.view(np.uint8).reshape(a.shape + (-1,)) if a.dtype.fields else np.ascontiguousarray(a).view(np.uint8).reshape(a.shape[:2] + (-1,))
    bv = b.view(np.uint8).reshape(b.shape + (-1,)) if b.dtype.fields else np.ascontiguousarray(b).view(np.uint8).reshape(b.shape[:2] + (-1,))
    return int(np.any(av != bv, axis=-1).sum())


def camera_rays(cam, width, height):
    """Pixel-centre primary rays of a camera, float64 numpy (independent of both implementations)."""
    ys, xs = np.mgrid[0:height, 0:width]
    u = (xs + 0.5) / width
    v = 1.0 - (ys + 0.5) / height
    ndc_x, ndc_y = u * 2 - 1, v * 2 - 1
    aspect = width / height
    t = np.tan(np.radians(cam.FOV * 0.5))
    d = np.stack([ndc_x * aspect * t, ndc_y * t, np.ones_like(ndc_x)], -1)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    right, up, front = (np.array(list(c)) for c in (cam.right, cam.up, cam.front))
    w = d[..., 0:1] * right + d[..., 1:2] * up + d[..., 2:3] * front
    w /= np.linalg.norm(w, axis=-1, keepdims=True)
    return np.array(list(cam.pos)), w


def random_rays(rng, n, lo, hi, tmin=1e-4, tmax=1e7):
    o = rng.uniform(lo, hi, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0:3] = o
    rays[:, 3] = tmin
    rays[:, 4:7] = d
    rays[:, 7] = tmax
    return rays
