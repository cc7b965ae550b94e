"""TEST INFRASTRUCTURE.  Runs the reference's own compute shaders on the CPU (oracle/_ref/libref.so: src/shader/*.comp compiled by
g++ from /root/reference through oracle/ref/glsl_to_cpp.py + glsl_compat.h) on the buffers of an oracle frame, so that a test can
put the reference's shader text and the oracle's restatement side by side on identical inputs.  The parts the Vulkan driver supplies
(ray / triangle intersection, texture and G-buffer filtering) are the oracle's definitions, passed in as call-backs."""
import ctypes as C
import os

from restirpt import BUF, P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref.so")

# oracle pass (tests/common.py METHOD_PASSES) -> reference compute shader (src/shader/<name>.comp)
SHADER_OF_PASS = {
    "di_naive": "di_naive", "gi_naive": "gi_naive",
    "di_pathgen": "di_path_gen", "di_temporal": "di_temporal", "di_spatial": "di_spatial",
    "gi_restir": "gi_resample_temporal",
    "gris_pathtrace": "gris_path_trace", "gris_temporal": "gris_resample_temporal", "gris_spatial": "gris_resample_spatial",
    "visualize_as": "as_visualize",
}


class RefDriver(C.Structure):
    _fields_ = [("user", P), ("traceClosest", P), ("traceAny", P), ("countCandidates", P), ("sampleTexture", P), ("sampleDepthNormal", P)]


class RefBindings(C.Structure):
    _fields_ = [("camera", P), ("prevCamera", P),
                ("materials", P), ("materialIndices", P), ("vertices", P), ("indices", P), ("instances", P), ("lights", P), ("lightTable", P),
                ("numMaterials", C.c_uint32), ("numMaterialIndices", C.c_uint32), ("numVertices", C.c_uint32), ("numIndices", C.c_uint32),
                ("numInstances", C.c_uint32), ("numLights", C.c_uint32), ("numTextures", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("directOutput", P), ("indirectOutput", P), ("depthNormal", P), ("depthNormalPrev", P),
                ("albedoMatId", P), ("albedoMatIdPrev", P), ("motion", P),
                ("di", P), ("diPrev", P), ("diTemp", P), ("gi", P), ("giPrev", P), ("gris", P), ("grisPrev", P), ("grisTemp", P), ("grisRc", P),
                ("push", P), ("pushBytes", C.c_uint32),
                ("driver", RefDriver)]


def available():
    return os.path.exists(REF_LIB)


class RefShaderBackend:
    """Looks like tests/common.py's Backend("oracle", ...) — same frame, camera and ping-pong handling, G-buffer by the oracle
    (the reference rasterises it: GBuffer.vert / .frag are not compute shaders) — but every ray pass is the reference's shader.
    contract=False: IEEE built-ins + libm; contract=True: the numeric contract's built-in library (glsl_compat.h)."""

    kind = "reference-shaders"

    def __init__(self, oracle_backend, host_scene, contract, threads=0):
        self.o = oracle_backend
        self.w, self.h = oracle_backend.w, oracle_backend.h
        self.host_scene = host_scene          # keeps the arrays of the scene description alive
        self.lib = C.CDLL(REF_LIB)
        self.fn = self.lib.refc_run_shader if contract else self.lib.ref_run_shader
        self.fn.restype, self.fn.argtypes = C.c_int, [C.c_char_p, C.POINTER(RefBindings), C.c_int]
        self.threads = threads
        ol = self.o.lib
        ol.orc_driver_create.restype, ol.orc_driver_create.argtypes = P, [P, P]
        ol.orc_driver_destroy.argtypes = [P]
        ol.orc_frame_ptr.restype, ol.orc_frame_ptr.argtypes = P, [P, C.c_int]
        self.ctx = P(ol.orc_driver_create(self.o.scene, self.o.frame))
        self.cur = self.prev = None
        self.rc = (C.c_uint8 * (48 * self.w * self.h))()     # uGRISReconnectionData: bound, never read by a live shader path

    def _cb(self, name):
        return C.cast(getattr(self.o.lib, name), P)

    def set_camera(self, cur, prev):
        self.cur, self.prev = cur.copy(), prev.copy()
        self.o.set_camera(cur, prev)

    def flip(self):
        self.o.flip()

    def clear(self):
        self.o.clear()

    def read(self, buf):
        return self.o.read(buf)

    def run(self, name, settings=None):
        if name == "gbuffer":
            return self.o.run(name)
        d = self.host_scene.desc
        ptr = lambda b: P(self.o.lib.orc_frame_ptr(self.o.frame, BUF[b]))
        b = RefBindings()
        b.camera, b.prevCamera = C.cast(C.byref(self.cur), P), C.cast(C.byref(self.prev), P)
        b.materials, b.materialIndices = C.cast(d.materials, P), C.cast(d.materialIndices, P)
        b.vertices, b.indices, b.instances = C.cast(d.vertices, P), C.cast(d.indices, P), C.cast(d.instances, P)
        b.lights, b.lightTable = C.cast(d.triangleLights, P), C.cast(d.lightSampleTable, P)
        b.numMaterials, b.numMaterialIndices, b.numVertices, b.numIndices = d.numMaterials, d.numMaterialIndices, d.numVertices, d.numIndices
        b.numInstances, b.numLights, b.numTextures, b.width, b.height = d.numInstances, d.numTriangleLights, d.numTextures, self.w, self.h
        b.directOutput, b.indirectOutput = ptr("DIRECT_OUTPUT"), ptr("INDIRECT_OUTPUT")
        b.depthNormal, b.depthNormalPrev = ptr("DEPTH_NORMAL"), ptr("DEPTH_NORMAL_PREV")
        b.albedoMatId, b.albedoMatIdPrev, b.motion = ptr("ALBEDO_MATID"), ptr("ALBEDO_MATID_PREV"), ptr("MOTION")
        b.di, b.diPrev, b.diTemp = ptr("DI_THIS"), ptr("DI_PREV"), ptr("DI_TEMP")
        b.gi, b.giPrev = ptr("GI_THIS"), ptr("GI_PREV")
        b.gris, b.grisPrev, b.grisTemp = ptr("GRIS_THIS"), ptr("GRIS_PREV"), ptr("GRIS_TEMP")
        b.grisRc = C.cast(self.rc, P)
        if settings is not None:
            b.push, b.pushBytes = C.cast(C.byref(settings), P), C.sizeof(settings)
        b.driver = RefDriver(self.ctx, self._cb("orc_cb_trace_closest"), self._cb("orc_cb_trace_any"), self._cb("orc_cb_count_candidates"),
                             self._cb("orc_cb_sample_texture"), self._cb("orc_cb_sample_depth_normal"))
        rc = self.fn(SHADER_OF_PASS[name].encode(), C.byref(b), self.threads)
        assert rc == 0, (name, rc)

    def close(self):
        self.o.lib.orc_driver_destroy(self.ctx)
        self.o.close()
