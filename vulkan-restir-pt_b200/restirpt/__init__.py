"""ctypes view of the two product libraries (no compute happens in Python).

* ``librestirpt.so``       — CUDA kernels + C ABI, ``include/restirpt.h``
* ``librestirpt_host.so``  — C++ host (Scene / Camera / alias table / headless Renderer), ``include/restirpt_host.h``

Used by the tests, ``bench.py`` and ``__graft_entry__``.  Loading fails loudly when the libraries have not been
built (``make`` / ``__graft_entry__.build()``); there is no fallback of any kind.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RPT_LIB_DIR: an experiment build of the same two libraries (Makefile: LIBDIR=...), for A/B measurements
LIB_DIR = os.environ.get("RPT_LIB_DIR") or os.path.join(os.path.dirname(_HERE), "lib")
REPO_ROOT = os.path.dirname(os.path.dirname(_HERE))


# ---- struct layouts (include/restirpt.h) ---------------------------------------------------------------------
class Material(C.Structure):
    _fields_ = [("baseColor", C.c_float * 3), ("type", C.c_uint32), ("textureIdx", C.c_uint32),
                ("metallic", C.c_float), ("roughness", C.c_float), ("ior", C.c_float)]


class MeshVertex(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("uvx", C.c_float), ("norm", C.c_float * 3), ("uvy", C.c_float)]


class ObjectInstance(C.Structure):
    _fields_ = [("transform", C.c_float * 16), ("transformInv", C.c_float * 16), ("transformInvT", C.c_float * 16),
                ("radiance", C.c_float * 3), ("pad0", C.c_float), ("indexOffset", C.c_uint32),
                ("indexCount", C.c_uint32), ("matIndex", C.c_uint32), ("pad2", C.c_float)]


class TriangleLight(C.Structure):
    _fields_ = [("v0", C.c_float * 3), ("nx", C.c_float), ("v1", C.c_float * 3), ("ny", C.c_float),
                ("v2", C.c_float * 3), ("nz", C.c_float), ("radiance", C.c_float * 3), ("area", C.c_float)]


class LightSampleTableElement(C.Structure):
    _fields_ = [("prob", C.c_float), ("failId", C.c_uint32)]


class Camera(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("projView", C.c_float * 16),
                ("lastProjView", C.c_float * 16),
                ("pos", C.c_float * 3), ("FOV", C.c_float), ("angle", C.c_float * 3), ("nearZ", C.c_float),
                ("front", C.c_float * 3), ("farZ", C.c_float), ("right", C.c_float * 3), ("lensRadius", C.c_float),
                ("up", C.c_float * 3), ("focalDist", C.c_float),
                ("filmSize", C.c_uint32 * 2), ("frameIndex", C.c_uint32), ("seed", C.c_uint32)]

    def copy(self):
        c = Camera()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(Camera))
        return c


class Intersection(C.Structure):
    _fields_ = [("bary", C.c_float * 2), ("instanceIdx", C.c_uint32), ("triangleIdx", C.c_uint32)]


class DISettings(C.Structure):
    _fields_ = [("shiftType", C.c_uint32), ("sampleType", C.c_uint32), ("temporalReuse", C.c_uint32),
                ("spatialReuse", C.c_uint32)]


class GRISSettings(C.Structure):
    _fields_ = [("shiftType", C.c_uint32), ("rrScale", C.c_float), ("temporalReuse", C.c_uint32),
                ("spatialReuse", C.c_uint32), ("cap", C.c_uint32)]


class PostSettings(C.Structure):
    _fields_ = [("toneMapping", C.c_uint32), ("correctGamma", C.c_uint32), ("noDirect", C.c_uint32),
                ("noIndirect", C.c_uint32)]


class TextureDesc(C.Structure):
    _fields_ = [("rgba8", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("filter", C.c_uint32)]


class SceneDesc(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("numVertices", C.c_uint32),
                ("indices", C.c_void_p), ("numIndices", C.c_uint32),
                ("materials", C.c_void_p), ("numMaterials", C.c_uint32),
                ("materialIndices", C.c_void_p), ("numMaterialIndices", C.c_uint32),
                ("instances", C.c_void_p), ("numInstances", C.c_uint32),
                ("triangleLights", C.c_void_p), ("numTriangleLights", C.c_uint32),
                ("lightSampleTable", C.c_void_p),
                ("textures", C.c_void_p), ("numTextures", C.c_uint32), ("flags", C.c_uint32)]


SCENE_TWO_LEVEL = 1


class Counters(C.Structure):
    _fields_ = [("closestRays", C.c_uint64), ("shadowRays", C.c_uint64), ("nodeVisits", C.c_uint64),
                ("triTests", C.c_uint64), ("shadedHits", C.c_uint64), ("shadowNodeVisits", C.c_uint64),
                ("shadowTriTests", C.c_uint64), ("maxNodeVisits", C.c_uint64)]


class PeerInfo(C.Structure):
    _fields_ = [("grisTempHandle", C.c_uint8 * 64), ("diTempHandle", C.c_uint8 * 64), ("flagsHandle", C.c_uint8 * 64),
                ("grisHandle", C.c_uint8 * 64 * 3), ("diHandle", C.c_uint8 * 64 * 2), ("giHandle", C.c_uint8 * 64 * 2),
                ("grisTempPtr", C.c_uint64), ("diTempPtr", C.c_uint64), ("flagsPtr", C.c_uint64),
                ("grisPtr", C.c_uint64 * 3), ("diPtr", C.c_uint64 * 2), ("giPtr", C.c_uint64 * 2), ("pid", C.c_uint64),
                ("device", C.c_int32), ("rowBegin", C.c_uint32), ("rowEnd", C.c_uint32), ("storeBegin", C.c_uint32),
                ("storeEnd", C.c_uint32), ("cur", C.c_uint32), ("pad", C.c_uint32 * 2)]


class GatherInfo(C.Structure):
    _fields_ = [("imageHandle", C.c_uint8 * 64), ("flagsHandle", C.c_uint8 * 64), ("imagePtr", C.c_uint64),
                ("flagsPtr", C.c_uint64), ("pid", C.c_uint64), ("device", C.c_int32), ("width", C.c_uint32),
                ("height", C.c_uint32), ("numStrips", C.c_uint32)]


class PassStats(C.Structure):
    _fields_ = [("ms", C.c_double * 12), ("launches", C.c_uint64 * 12), ("kernelMs", C.c_double * 9),
                ("kernelLaunches", C.c_uint64 * 9)]


KERNEL_NAMES = ["trace_closest", "trace_any", "gris_begin", "gris_bounce", "gris_tail", "reuse_gen", "reuse_merge", "tail_wait",
                "trace_pair"]
PASS_NAMES = ["gbuffer", "di_naive", "gi_naive", "di_pathgen", "di_temporal", "di_spatial", "gi_restir",
              "gris_pathtrace", "gris_temporal", "gris_spatial", "visualize_as", "postprocess"]


class BvhStats(C.Structure):
    _fields_ = [("numTriangles", C.c_uint32), ("numNodes", C.c_uint32), ("nodeBytes", C.c_uint64),
                ("triBytes", C.c_uint64), ("buildMs", C.c_float), ("sahCost", C.c_float),
                ("twoLevel", C.c_uint32), ("numMeshes", C.c_uint32), ("numTlasNodes", C.c_uint32),
                ("numInstanceRecords", C.c_uint32), ("tlasBuildMs", C.c_float), ("pad", C.c_uint32)]


assert C.sizeof(Material) == 32 and C.sizeof(MeshVertex) == 32 and C.sizeof(ObjectInstance) == 224
assert C.sizeof(TriangleLight) == 64 and C.sizeof(Camera) == 352 and C.sizeof(Intersection) == 16

# buffer ids (RptBufferId) and numpy views of one pixel of each
BUF = dict(DIRECT_OUTPUT=0, INDIRECT_OUTPUT=1, DEPTH_NORMAL=2, DEPTH_NORMAL_PREV=3, ALBEDO_MATID=4,
           ALBEDO_MATID_PREV=5, MOTION=6, DI_THIS=7, DI_PREV=8, DI_TEMP=9, GI_THIS=10, GI_PREV=11,
           GRIS_THIS=12, GRIS_PREV=13, GRIS_TEMP=14, PRIMARY_ISEC=15)

ISEC_DTYPE = np.dtype([("bary", "<f4", 2), ("instanceIdx", "<u4"), ("triangleIdx", "<u4")])
DI_DTYPE = np.dtype([("isec", ISEC_DTYPE), ("Li", "<f4", 3), ("pad0", "<f4"), ("jacobian", "<f4"),
                     ("samplePdf", "<f4"), ("rng", "<u4"), ("isLightSample", "<u4"), ("sampleCount", "<u4"),
                     ("resampleWeight", "<f4"), ("contribWeight", "<f4"), ("weight", "<f4")])
GI_DTYPE = np.dtype([("rcIsec", ISEC_DTYPE), ("rcLo", "<f4", 3), ("rcPrevCoord", "<u4"), ("sampleCount", "<u4"),
                     ("resampleWeight", "<f4"), ("contribWeight", "<f4"), ("pad0", "<f4")])
GRIS_DTYPE = np.dtype([("rcIsec", ISEC_DTYPE), ("rcLi", "<f4", 3), ("rcRng", "<u4"), ("rcWi", "<f4", 3),
                       ("flags", "<u4"), ("pad", "<f4", 2), ("rcPrevSamplePdf", "<f4"), ("rcJacobian", "<f4"),
                       ("F", "<f4", 3), ("primaryRng", "<u4"), ("sampleCount", "<f4"), ("resampleWeight", "<f4"),
                       ("contribWeight", "<f4"), ("pad0", "<f4")])
BUF_DTYPE = {0: np.dtype(("<f4", 4)), 1: np.dtype(("<f4", 4)), 2: np.dtype(("<f4", 4)), 3: np.dtype(("<f4", 4)),
             4: np.dtype(("<u4", 2)), 5: np.dtype(("<u4", 2)), 6: np.dtype(("<f4", 2)),
             7: DI_DTYPE, 8: DI_DTYPE, 9: DI_DTYPE, 10: GI_DTYPE, 11: GI_DTYPE,
             12: GRIS_DTYPE, 13: GRIS_DTYPE, 14: GRIS_DTYPE, 15: ISEC_DTYPE}
assert DI_DTYPE.itemsize == 64 and GI_DTYPE.itemsize == 48 and GRIS_DTYPE.itemsize == 96


class RestirptError(RuntimeError):
    pass


def _load(name):
    path = os.path.join(LIB_DIR, name)
    if not os.path.exists(path):
        raise RestirptError(f"{path} is missing: build it with `make` (or __graft_entry__.build()); "
                            "there is no fallback implementation")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


_dev = None
_host = None

# every symbol include/restirpt.h declares: (restype, argtypes)
P = C.c_void_p
DEVICE_API = {
    "rpt_version": (C.c_int, []),
    "rpt_last_error": (C.c_char_p, [P]),
    "rpt_ctx_create": (C.c_int, [C.c_int, C.POINTER(P)]),
    "rpt_ctx_destroy": (None, [P]),
    "rpt_scene_create": (C.c_int, [P, C.POINTER(SceneDesc), C.POINTER(P)]),
    "rpt_scene_end_motion": (C.c_int, [P]),
    "rpt_scene_destroy": (None, [P]),
    "rpt_scene_update_instances": (C.c_int, [P, P, C.c_uint32]),
    "rpt_scene_bvh_stats": (C.c_int, [P, C.POINTER(BvhStats)]),
    "rpt_frame_create": (C.c_int, [P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(P)]),
    "rpt_frame_destroy": (None, [P]),
    "rpt_frame_clear": (C.c_int, [P]),
    "rpt_frame_flip": (C.c_int, [P]),
    "rpt_frame_stream": (P, [P]),
    "rpt_frame_join": (C.c_int, [P]),
    "rpt_set_camera": (C.c_int, [P, C.POINTER(Camera), C.POINTER(Camera)]),
    "rpt_gbuffer": (C.c_int, [P, P]),
    "rpt_di_naive": (C.c_int, [P, P]),
    "rpt_di_naive_rt": (C.c_int, [P, P]),
    "rpt_gi_naive": (C.c_int, [P, P]),
    "rpt_di_pathgen": (C.c_int, [P, P, C.POINTER(DISettings)]),
    "rpt_di_temporal": (C.c_int, [P, P, C.POINTER(DISettings)]),
    "rpt_di_spatial": (C.c_int, [P, P, C.POINTER(DISettings)]),
    "rpt_gi_restir": (C.c_int, [P, P]),
    "rpt_gris_pathtrace": (C.c_int, [P, P, C.POINTER(GRISSettings)]),
    "rpt_gris_temporal": (C.c_int, [P, P, C.POINTER(GRISSettings)]),
    "rpt_gris_spatial": (C.c_int, [P, P, C.POINTER(GRISSettings)]),
    "rpt_visualize_as": (C.c_int, [P, P]),
    "rpt_postprocess_async": (C.c_int, [P, C.POINTER(PostSettings), P, C.POINTER(C.c_uint64)]),
    "rpt_readback_wait": (C.c_int, [P, C.c_uint64]),
    "rpt_postprocess": (C.c_int, [P, C.POINTER(PostSettings), P]),
    "rpt_sync": (C.c_int, [P]),
    "rpt_frame_export_peer": (C.c_int, [P, C.POINTER(PeerInfo)]),
    "rpt_frame_connect_peers": (C.c_int, [P, C.POINTER(PeerInfo), C.POINTER(PeerInfo)]),
    "rpt_frame_disconnect_peers": (C.c_int, [P]),
    "rpt_frame_peer_error": (C.c_int, [P]),
    "rpt_frame_peers_in_process": (C.c_int, [P]),
    "rpt_frame_gather_create": (C.c_int, [P, C.c_uint32, C.POINTER(GatherInfo)]),
    "rpt_frame_gather_connect": (C.c_int, [P, C.POINTER(GatherInfo), C.c_uint32]),
    "rpt_frame_gather_disconnect": (C.c_int, [P]),
    "rpt_gather_output": (C.c_int, [P, P]),
    "rpt_frame_timing": (C.c_int, [P, C.c_int]),
    "rpt_frame_pass_stats": (C.c_int, [P, C.POINTER(PassStats)]),
    "rpt_buffer_stride": (C.c_size_t, [C.c_int]),
    "rpt_frame_rows": (C.c_int, [P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rpt_read": (C.c_int, [P, C.c_int, P, C.c_size_t]),
    "rpt_write": (C.c_int, [P, C.c_int, P, C.c_size_t]),
    "rpt_device_ptr": (P, [P, C.c_int]),
    "rpt_trace_closest": (C.c_int, [P, P, P, C.c_uint32, P]),
    "rpt_trace_shadow": (C.c_int, [P, P, P, C.c_uint32, P]),
    "rpt_trace_bench": (C.c_int, [P, P, P, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), P, P]),
    "rpt_wavefront_counters": (C.c_int, [P, C.POINTER(C.c_uint32)]),
    "rpt_reuse_counters": (C.c_int, [P, C.POINTER(C.c_uint32)]),
    "rpt_membench": (C.c_int, [P, C.c_size_t, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "rpt_counters_enable": (C.c_int, [P, C.c_int]),
    "rpt_counters_reset": (C.c_int, [P]),
    "rpt_counters_read": (C.c_int, [P, C.POINTER(Counters)]),
}

HALO_FN = C.CFUNCTYPE(None, P, P, C.c_int)
HOST_API = {
    "rh_last_error": (C.c_char_p, []),
    "rh_scene_load_xml": (P, [C.c_char_p]),
    "rh_scene_cornell": (P, []),
    "rh_scene_room": (P, [C.c_uint32, C.c_uint32]),
    "rh_scene_field": (P, [C.c_uint32, C.c_uint32, C.c_uint32]),
    "rh_scene_field_shared": (P, [C.c_uint32, C.c_uint32, C.c_uint32]),
    "rh_scene_set_two_level": (None, [P, C.c_int]),
    "rh_scene_destroy": (None, [P]),
    "rh_scene_desc": (None, [P, C.POINTER(SceneDesc)]),
    "rh_scene_camera": (None, [P, C.POINTER(Camera)]),
    "rh_scene_num_triangles": (C.c_uint32, [P]),
    "rh_camera_init": (None, [C.POINTER(Camera), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_uint32,
                              C.c_uint32, C.c_float, C.c_float]),
    "rh_camera_look_at": (None, [C.POINTER(Camera), C.POINTER(C.c_float)]),
    "rh_camera_set_film": (None, [C.POINTER(Camera), C.c_uint32, C.c_uint32]),
    "rh_camera_set_planes": (None, [C.POINTER(Camera), C.c_float, C.c_float]),
    "rh_camera_move": (None, [C.POINTER(Camera), C.POINTER(C.c_float)]),
    "rh_camera_update": (None, [C.POINTER(Camera)]),
    "rh_camera_next_frame": (None, [C.POINTER(Camera), C.c_uint32]),
    "rh_build_alias_table": (None, [C.POINTER(C.c_float), C.c_uint32, C.POINTER(LightSampleTableElement)]),
    "rh_renderer_create": (P, [P, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]),
    "rh_renderer_destroy": (None, [P]),
    "rh_renderer_set_methods": (None, [P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "rh_renderer_set_pipeline_mode": (None, [P, C.c_int]),
    "rh_renderer_update_instances": (C.c_int, [P, P]),
    "rh_scene_set_object_transform": (C.c_int, [P, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "rh_renderer_set_gris": (None, [P, C.POINTER(GRISSettings)]),
    "rh_renderer_set_di": (None, [P, C.POINTER(DISettings)]),
    "rh_renderer_clear_reservoirs": (None, [P]),
    "rh_renderer_camera_move": (None, [P, C.POINTER(C.c_float)]),
    "rh_renderer_camera": (None, [P, C.POINTER(Camera)]),
    "rh_renderer_set_halo_exchange": (None, [P, HALO_FN, P]),
    "rh_renderer_draw_frame": (C.c_int, [P, C.c_uint32, P]),
    "rh_renderer_draw_frame_async": (C.c_int, [P, C.c_uint32, P, C.POINTER(C.c_uint64)]),
    "rh_renderer_wait_readback": (C.c_int, [P, C.c_uint64]),
    "rh_draw_strips": (C.c_int, [C.POINTER(P), C.c_uint32, C.c_uint32, C.POINTER(P)]),
    "rh_renderer_frame": (P, [P]),
    "rh_renderer_scene": (P, [P]),
    "rh_renderer_ctx": (P, [P]),
    "rh_xml_dump": (C.c_size_t, [C.c_char_p, C.c_char_p, C.c_size_t]),
    "rh_write_png": (C.c_int, [C.c_char_p, P, C.c_uint32, C.c_uint32]),
    "rh_read_image": (P, [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rh_free_image": (None, [P]),
}


def _bind(lib, table):
    for name, (res, args) in table.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args


def device_lib():
    global _dev
    if _dev is None:
        _dev = _load("librestirpt.so")
        _bind(_dev, DEVICE_API)
    return _dev


def host_lib():
    global _host
    if _host is None:
        device_lib()
        _host = _load("librestirpt_host.so")
        _bind(_host, HOST_API)
    return _host


def hash2(seed):
    """reference math.glsl:227-234 (used for the per-frame seed sequence seed[f] = hash2(f + 1))"""
    seed &= 0xffffffff
    seed = (seed ^ 61) ^ (seed >> 16)
    seed = (seed * 9) & 0xffffffff
    seed = seed ^ (seed >> 4)
    seed = (seed * 0x27d4eb2d) & 0xffffffff
    seed = seed ^ (seed >> 15)
    return seed


def check(ctx, status, what):
    if status != 0:
        msg = device_lib().rpt_last_error(ctx)
        raise RestirptError(f"{what} failed ({status}): {msg.decode() if msg else ''}")


# ---- thin object wrappers --------------------------------------------------------------------------------------
class HostScene:
    """Scene of the C++ host library (XML / procedural), exposing the RptSceneDesc view."""

    def __init__(self, handle):
        if not handle:
            raise RestirptError("scene creation failed: " + host_lib().rh_last_error().decode())
        self.handle = handle
        self.desc = SceneDesc()
        host_lib().rh_scene_desc(handle, C.byref(self.desc))

    @staticmethod
    def cornell():
        return HostScene(host_lib().rh_scene_cornell())

    @staticmethod
    def room(tris=20000, seed=1):
        return HostScene(host_lib().rh_scene_room(tris, seed))

    @staticmethod
    def field(subdiv=2, grid=4, seed=42, shared=False, two_level=False):
        """shared: one copy of the mesh referenced by every instance; two_level: BLAS per unique mesh + TLAS"""
        sc = HostScene((host_lib().rh_scene_field_shared if shared else host_lib().rh_scene_field)(subdiv, grid, seed))
        if two_level:
            sc.set_two_level(True)
        return sc

    def set_two_level(self, on=True):
        host_lib().rh_scene_set_two_level(self.handle, 1 if on else 0)
        host_lib().rh_scene_desc(self.handle, C.byref(self.desc))

    @staticmethod
    def xml(path):
        return HostScene(host_lib().rh_scene_load_xml(path.encode()))

    def set_object_transform(self, object_idx, pos, scale=(1.0, 1.0, 1.0), rot_deg=(0.0, 0.0, 0.0)):
        """place object model `object_idx` anew (the XML <transform> attributes); self.desc stays valid (same arrays)"""
        f3 = C.c_float * 3
        if host_lib().rh_scene_set_object_transform(self.handle, object_idx, f3(*pos), f3(*scale), f3(*rot_deg)) != 0:
            raise RestirptError(host_lib().rh_last_error().decode())
        host_lib().rh_scene_desc(self.handle, C.byref(self.desc))

    def camera(self, width=None, height=None):
        cam = Camera()
        host_lib().rh_scene_camera(self.handle, C.byref(cam))
        if width:
            host_lib().rh_camera_set_film(C.byref(cam), width, height)
        # the reference Renderer overrides the planes (src/Renderer.cpp:124)
        host_lib().rh_camera_set_planes(C.byref(cam), 0.001, 200.0)
        C.memmove(cam.lastProjView, cam.projView, 64)
        return cam

    @property
    def num_triangles(self):
        return host_lib().rh_scene_num_triangles(self.handle)

    def close(self):
        if self.handle:
            host_lib().rh_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Device:
    """RptCtx + helpers.  Raises when no CUDA device is present (the library has no CPU path)."""

    def __init__(self, index=0):
        self.lib = device_lib()
        self.ctx = P()
        check(None, self.lib.rpt_ctx_create(index, C.byref(self.ctx)), "rpt_ctx_create")

    def scene(self, desc):
        s = P()
        check(self.ctx, self.lib.rpt_scene_create(self.ctx, C.byref(desc), C.byref(s)), "rpt_scene_create")
        return s

    def frame(self, w, h, row_begin=0, row_end=None, halo=0):
        f = P()
        check(self.ctx, self.lib.rpt_frame_create(self.ctx, w, h, row_begin, h if row_end is None else row_end, halo,
                                                  C.byref(f)), "rpt_frame_create")
        return f

    def close(self):
        if self.ctx:
            self.lib.rpt_ctx_destroy(self.ctx)
            self.ctx = None


def read_buffer(lib, frame, buf_id, width, rows, prefix="rpt"):
    """Read one frame buffer into a structured numpy array of shape (rows, width)."""
    dt = BUF_DTYPE[buf_id]
    out = np.zeros((rows, width), dtype=dt)
    fn = getattr(lib, prefix + "_read")
    status = fn(frame, buf_id, out.ctypes.data_as(P), out.nbytes)
    if status != 0:
        raise RestirptError(f"{prefix}_read({buf_id}) failed: {status}")
    return out


def read_image(path):
    """Decode a PNG / JPEG / binary PPM with the host library (rh_read_image): (height, width, 4) uint8."""
    host = host_lib()
    w, h = C.c_uint32(), C.c_uint32()
    ptr = host.rh_read_image(os.fsencode(path), C.byref(w), C.byref(h))
    if not ptr:
        raise RestirptError(host.rh_last_error().decode())
    try:
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(h.value, w.value, 4)).copy()
    finally:
        host.rh_free_image(ptr)
