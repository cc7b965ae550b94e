"""The oracle pinned against the reference ITSELF, run here: the reference's ten ray-query compute shaders (src/shader/*.comp with
everything they #include) are compiled by g++ from where they lie under /root/reference — oracle/ref/Makefile: glsl_to_cpp.py does a
purely syntactic rewrite into a temporary directory, glsl_compat.h supplies GLSL's types and built-ins — and executed on the CPU, one
main() per pixel, on the buffers of an oracle frame (tests/ref_shaders.py).  What the Vulkan driver supplies to the reference
(ray / triangle intersection, texture + G-buffer filtering, the rasterised G-buffer) is supplied by the oracle's definitions.

Two builds of GLSL's built-in function library, two bars:
  * numeric-contract built-ins (dot / cross / normalize / mix / reflect / matrix * vector / vector / scalar / sin / cos / tan
    evaluated as DESIGN.md §2 prescribes — choices GLSL leaves to the implementation): the reference's text and the oracle's
    restatement must then agree BIT FOR BIT on every buffer of every pass over multi-frame sequences with camera motion — any
    difference would be a difference in the algorithm;
  * IEEE built-ins + libm ("a GPU whose built-ins are exact"): every pass, started from identical inputs, must agree within the
    tolerance written below on all but a few per cent of the pixels (a discrete choice — a resampling pick, a Russian-roulette
    kill — can flip on the last bit of a float), and the image means within 1 %.
oracle/_ref/libref.so is built in the container (it needs /root/reference) and travels to the GPU box git-ignored; without it these
tests skip loudly."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import prepare_assets
import ref_shaders
import restirpt
from common import Backend, FrameDriver, METHOD_PASSES, bitwise_mismatch
from restirpt import BUF, DISettings, GRISSettings

pytestmark = pytest.mark.skipif(not ref_shaders.available(), reason="oracle/_ref/libref.so is not built (make -C oracle/ref, needs "
                                "/root/reference): the reference's shaders are NOT checked in this run")

MOVES = [(0.0, 0.0, 0.0), (0.02, 0.01, 0.0), (0.0, -0.01, 0.01), (0.01, 0.0, -0.02)]
WRITTEN = {   # the buffers each pass writes (layouts.glsl:180-195)
    "di_naive": ["DIRECT_OUTPUT"], "gi_naive": ["INDIRECT_OUTPUT"], "visualize_as": ["DIRECT_OUTPUT"],
    "di_pathgen": ["DI_THIS"], "di_temporal": ["DI_TEMP"], "di_spatial": ["DI_THIS", "DIRECT_OUTPUT"],
    "gi_restir": ["GI_THIS", "INDIRECT_OUTPUT"],
    "gris_pathtrace": ["GRIS_THIS"], "gris_temporal": ["GRIS_TEMP"], "gris_spatial": ["GRIS_THIS", "INDIRECT_OUTPUT"],
}
ALL_BUFFERS = ["DIRECT_OUTPUT", "INDIRECT_OUTPUT", "DI_THIS", "DI_PREV", "DI_TEMP", "GI_THIS", "GI_PREV", "GRIS_THIS", "GRIS_PREV", "GRIS_TEMP"]


def _scene(name):
    if name == "cornell":
        return restirpt.HostScene.cornell()
    if name == "room":
        return restirpt.HostScene.room(6000, 7)
    if name == "field_tlas":
        return restirpt.HostScene.field(1, 3, shared=True, two_level=True)
    xml = prepare_assets.ajar_xml()
    if not xml:
        pytest.skip("assets/_ref/VeachAjar is not prepared (tools/prepare_assets.py needs /root/reference): VeachAjar NOT checked in this run")
    return restirpt.HostScene.xml(xml)


def _lockstep(scene_name, w, h, method, frames, contract, di=None, gris=None, resync=False):
    """oracle and reference shaders side by side; yields (frame, pass, buffer, oracle array, reference array)"""
    sc = _scene(scene_name)
    a = Backend("oracle", sc, w, h)
    b = ref_shaders.RefShaderBackend(Backend("oracle", sc, w, h), sc, contract)
    settings = {"di": di or DISettings(0, 0, 1, 1), "gris": gris or GRISSettings(2, 1.0, 1, 1, 20)}
    drv = FrameDriver(sc.camera(w, h))
    try:
        for i in range(frames):
            cur, prev = drv.begin_frame(move=MOVES[i % len(MOVES)])
            for be in (a, b):
                be.set_camera(cur, prev)
                be.run("gbuffer")
            for name, skey in METHOD_PASSES[method]:
                if resync:      # the pass starts from the oracle's state on both sides
                    for k in ALL_BUFFERS:
                        arr = np.ascontiguousarray(a.read(k))
                        assert b.o.lib.orc_write(b.o.frame, BUF[k], arr.ctypes.data_as(restirpt.P), arr.nbytes) == 0
                for be in (a, b):
                    be.run(name, settings[skey] if skey else None)
                for k in WRITTEN[name]:
                    yield i, name, k, a.read(k), b.read(k)
            a.flip()
            b.flip()
    finally:
        a.close()
        b.close()


# ---- numeric-contract built-ins: bit for bit ---------------------------------------------------------------------------------
CASES = [
    ("cornell", 64, 36, "naive", {}),
    ("cornell", 64, 36, "di", {"di": (0, 0, 1, 1)}), ("cornell", 64, 36, "di", {"di": (1, 2, 1, 1)}), ("cornell", 64, 36, "di", {"di": (0, 1, 0, 1)}),
    ("cornell", 64, 36, "gi", {}),
    ("cornell", 64, 36, "gris", {"gris": (2, 1.0, 1, 1, 20)}), ("cornell", 64, 36, "gris", {"gris": (0, 1.0, 1, 1, 20)}),
    ("cornell", 64, 36, "gris", {"gris": (1, 0.5, 1, 0, 8)}), ("cornell", 64, 36, "gris", {"gris": (2, 1.0, 0, 1, 20)}),
    ("room", 96, 54, "naive", {}), ("room", 96, 54, "di", {"di": (0, 2, 1, 1)}), ("room", 96, 54, "gi", {}),
    ("room", 96, 54, "gris", {"gris": (2, 1.0, 1, 1, 20)}),
    ("field_tlas", 64, 36, "gris", {"gris": (2, 1.0, 1, 1, 20)}),
    # the shipped scene: 4 textures, glass / metal / metallic-workflow materials, 22 instances with transforms
    ("ajar", 160, 90, "naive", {}), ("ajar", 160, 90, "di", {"di": (0, 0, 1, 1)}), ("ajar", 160, 90, "gi", {}),
    ("ajar", 160, 90, "gris", {"gris": (2, 1.0, 1, 1, 20)}), ("ajar", 160, 90, "gris", {"gris": (0, 1.0, 1, 1, 20)}),
]


@pytest.mark.parametrize("scene,w,h,method,kw", CASES, ids=lambda v: str(v).replace(" ", "") if not isinstance(v, (str, int)) else str(v))
def test_reference_shaders_equal_the_oracle_bit_for_bit(scene, w, h, method, kw):
    """src/shader/{di_naive, gi_naive, di_path_gen, di_temporal, di_spatial, gi_resample_temporal, gris_path_trace,
    gris_resample_temporal, gris_resample_spatial}.comp against oracle_passes.cpp: 3-4 frames with camera motion (temporal
    reprojection, ping-pong buffers and accumulation all in play, nothing re-synchronised between passes or frames)"""
    di = DISettings(*kw["di"]) if "di" in kw else None
    gris = GRISSettings(*kw["gris"]) if "gris" in kw else None
    frames = 4 if scene == "cornell" else 3
    checked = lit = 0
    for i, name, buf, want, got in _lockstep(scene, w, h, method, frames, True, di=di, gris=gris):
        assert bitwise_mismatch(want, got) == 0, f"frame {i} after {name}: {buf} differs in {bitwise_mismatch(want, got)} of {w * h} pixels"
        checked += 1
        if buf.endswith("OUTPUT"):
            lit += int(np.count_nonzero(np.ascontiguousarray(got).view(np.float32).reshape(h, w, -1)[..., :3].sum(-1) > 0))
    assert checked >= frames * len(METHOD_PASSES[method]) and lit > 0      # (the frames are not black)


def test_reference_as_visualize_equals_the_oracle():
    sc = _scene("room")
    w, h = 96, 54
    a = Backend("oracle", sc, w, h)
    b = ref_shaders.RefShaderBackend(Backend("oracle", sc, w, h), sc, True)
    cam = sc.camera(w, h)
    for be in (a, b):
        be.set_camera(cam, cam)
        be.run("gbuffer")
        be.run("visualize_as")
    assert bitwise_mismatch(a.read("DIRECT_OUTPUT"), b.read("DIRECT_OUTPUT")) == 0
    assert np.ascontiguousarray(a.read("DIRECT_OUTPUT")).view(np.float32).max() > 0
    a.close()
    b.close()


# ---- IEEE built-ins + libm: within tolerance ---------------------------------------------------------------------------------
def _agreeing_pixels(want, got, rtol=2e-3, atol=1e-6):
    """pixels all of whose 32-bit words are either the same bits or, read as floats, within rtol"""
    h, w = want.shape[:2]
    a = np.ascontiguousarray(want).view(np.uint32).reshape(h, w, -1)
    b = np.ascontiguousarray(got).view(np.uint32).reshape(h, w, -1)
    with np.errstate(invalid="ignore", over="ignore"):
        fa, fb = a.view(np.float32).astype(np.float64), b.view(np.float32).astype(np.float64)
        close = np.abs(fa - fb) <= rtol * np.maximum(np.abs(fa), np.abs(fb)) + atol
    return np.all((a == b) | close, axis=-1)


@pytest.mark.parametrize("scene,w,h,method", [("cornell", 64, 36, "naive"), ("cornell", 64, 36, "di"), ("cornell", 64, 36, "gi"),
                                               ("cornell", 64, 36, "gris"), ("ajar", 160, 90, "gris"), ("ajar", 160, 90, "di")])
def test_reference_shaders_with_ieee_builtins_agree_within_tolerance(scene, w, h, method):
    """the same shaders with plain IEEE dot / cross / normalize ... and libm's sin / cos / tan — nothing of the numeric contract —
    against the oracle, every pass from identical inputs: at least 96 % of the pixels agree to 2e-3 in every word of every buffer
    the pass writes (the rest are discrete choices that flipped on a last bit: a different light, neighbour or path survives),
    and the film means agree to 1 %"""
    worst = 1.0
    for i, name, buf, want, got in _lockstep(scene, w, h, method, 3, False, resync=True):
        ok = _agreeing_pixels(want, got)
        worst = min(worst, ok.mean())
        assert ok.mean() >= 0.96, f"frame {i} after {name}: {buf} agrees in {ok.mean():.4f} of the pixels"
        if buf.endswith("OUTPUT"):
            ma = np.ascontiguousarray(want).view(np.float32).reshape(h, w, -1)[..., :3].astype(np.float64).mean()
            mb = np.ascontiguousarray(got).view(np.float32).reshape(h, w, -1)[..., :3].astype(np.float64).mean()
            assert abs(ma - mb) <= 0.01 * abs(ma) + 1e-7, (i, name, buf, ma, mb)
    print(f"{scene} {method}: worst per-pass pixel agreement {worst:.4f}")


# ---- the BASELINE configurations themselves ----------------------------------------------------------------------------------
@pytest.mark.parametrize("scene,w,h,method,kw,frames", [
    ("cornell", 640, 360, "naive", {}, 3),                                   # config 1 (seeds hash2(1), hash2(2), hash2(3))
    ("ajar", 1280, 720, "di", {"di": (0, 0, 1, 1)}, 2),                      # config 2: ReSTIR DI {Reconnection, Light, 1, 1}
    ("ajar", 1920, 1080, "gris", {"gris": (2, 1.0, 1, 1, 20)}, 2),           # config 3: ReSTIR PT {Hybrid, 1, 1, 1, 20}
], ids=["config1-cornell-640x360-naive", "config2-ajar-1280x720-restir-di", "config3-ajar-1920x1080-restir-pt"])
def test_reference_shaders_equal_the_oracle_on_the_baseline_configurations(scene, w, h, method, kw, frames):
    """BASELINE.json's configurations at their own film sizes, camera dolly between the frames: the reference's shaders and the
    oracle write the same bits into every reservoir and every output pixel (tests/test_gpu_baseline_configs.py holds the CUDA
    library to the oracle on the same configurations)"""
    di = DISettings(*kw["di"]) if "di" in kw else None
    gris = GRISSettings(*kw["gris"]) if "gris" in kw else None
    n = 0
    for i, name, buf, want, got in _lockstep(scene, w, h, method, frames, True, di=di, gris=gris):
        assert bitwise_mismatch(want, got) == 0, f"frame {i} after {name}: {buf} differs in {bitwise_mismatch(want, got)} of {w * h} pixels"
        n += 1
    assert n >= frames * len(METHOD_PASSES[method])
