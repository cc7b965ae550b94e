"""Pins against code the REFERENCE itself ships, compiled from where it lies by the committed recipe oracle/ref/Makefile into
oracle/_ref/libref.so (git-ignored; it travels to the GPU box like the other built libraries):

  * src/util/AliasTable.h (DiscreteSampler1D)  vs  rh_build_alias_table   — bit-exact, on VeachAjar's light powers and on random vectors
  * ext/pugixml (src/Scene.cpp:107-190 parses scenes with it)  vs  host/XmlLite.h — identical element tree for ajar.xml
  * ext/stb/stb_image.h (zvk/core/HostImage.cpp:70-75)  vs  host/Image.cpp — PNG bit-exact; baseline JPEG within the stated bound
    (two conforming JPEG decoders differ in IDCT / chroma-upsampling rounding; the measured workloads do not depend on it: the
    texel side-cars the host loads are written by the reference's decoder, tools/prepare_assets.py)

The rest of the reference (Win32 host + Vulkan ray-query shaders + glm) cannot be built here, so the shader path stays pinned by
the oracle only."""
import ctypes as C
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import prepare_assets
import restirpt
from restirpt import LightSampleTableElement

REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref.so")


@pytest.fixture(scope="module")
def ref(built):
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libref.so is not built (make -C oracle/ref, needs /root/reference): reference pins NOT checked in this run")
    lib = C.CDLL(REF_LIB)
    lib.ref_build_alias_table.argtypes = [C.POINTER(C.c_float), C.c_uint32, C.c_void_p]
    lib.ref_stbi_load_rgba8.restype = C.c_void_p
    lib.ref_stbi_load_rgba8.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ref_stbi_free.argtypes = [C.c_void_p]
    lib.ref_xml_dump.restype = C.c_size_t
    lib.ref_xml_dump.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    return lib


def _tables(ref, power):
    power = np.ascontiguousarray(power, dtype=np.float32)
    n = power.shape[0]
    ours = (LightSampleTableElement * (n + 1))()
    theirs = (LightSampleTableElement * (n + 1))()
    p = power.ctypes.data_as(C.POINTER(C.c_float))
    restirpt.host_lib().rh_build_alias_table(p, n, ours)
    ref.ref_build_alias_table(p, n, theirs)
    return bytes(ours), bytes(theirs)


def test_alias_table_equals_the_references_builder_on_random_powers(ref):
    rng = np.random.default_rng(2024)
    for n in (1, 2, 3, 7, 64, 1000, 4099):
        for dist in ("uniform", "lognormal", "equal", "one-hot"):
            if dist == "uniform":
                power = rng.uniform(0.01, 10.0, size=n)
            elif dist == "lognormal":
                power = rng.lognormal(0.0, 3.0, size=n)
            elif dist == "equal":
                power = np.full(n, 2.5)
            else:
                power = np.full(n, 1e-6); power[n // 2] = 1e3
            ours, theirs = _tables(ref, power)
            assert ours == theirs, f"n={n} {dist}"


def test_alias_table_equals_the_references_builder_on_the_shipped_scenes_lights(ref):
    xml = prepare_assets.ajar_xml()
    scenes = [restirpt.HostScene.cornell(), restirpt.HostScene.room(6000, 7)] + ([restirpt.HostScene.xml(xml)] if xml else [])
    for sc in scenes:
        n = sc.desc.numTriangleLights
        lights = np.ctypeslib.as_array(C.cast(sc.desc.triangleLights, C.POINTER(C.c_float)), shape=(n, 16))
        # the power the reference feeds its sampler: luminance(radiance) * area (src/Scene.cpp:296-316)
        lum = lights[:, 12] * 0.299 + lights[:, 13] * 0.587 + lights[:, 14] * 0.114
        ours, theirs = _tables(ref, (lum * lights[:, 15]).astype(np.float32))
        assert ours == theirs
        # and the table the scene itself carries was built from a power vector the reference's builder maps to the same bytes
        table = bytes((LightSampleTableElement * (n + 1)).from_address(sc.desc.lightSampleTable))
        stored = np.frombuffer(table, dtype=[("prob", "<f4"), ("failId", "<u4")])
        assert stored["failId"][0] == n and np.isfinite(stored["prob"]).all()


def test_xml_reader_sees_the_tree_pugixml_sees(ref):
    xml = prepare_assets.ajar_xml()
    if xml is None:
        pytest.skip("assets/_ref/VeachAjar is not prepared: XML pin NOT checked in this run")
    host = restirpt.host_lib()
    need = ref.ref_xml_dump(xml.encode(), None, 0)
    assert need > 100
    a, b = C.create_string_buffer(need + 16), C.create_string_buffer(need + 16)
    assert ref.ref_xml_dump(xml.encode(), a, need + 16) == need
    assert host.rh_xml_dump(xml.encode(), b, need + 16) == need, host.rh_last_error()
    assert a.value == b.value
    assert a.value.count(b"\n") > 40 and b" model " in a.value or b"model" in a.value


def _stb(ref, path):
    w, h = C.c_int(), C.c_int()
    p = ref.ref_stbi_load_rgba8(os.fsencode(path), C.byref(w), C.byref(h))
    assert p, path
    try:
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(h.value, w.value, 4)).copy()
    finally:
        ref.ref_stbi_free(p)


def test_png_decoder_equals_stb_image_bit_for_bit(ref, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(5)
    cases = {"rgb": rng.integers(0, 256, (37, 53, 3), dtype=np.uint8), "rgba": rng.integers(0, 256, (16, 21, 4), dtype=np.uint8),
             "gray": rng.integers(0, 256, (19, 8), dtype=np.uint8)}
    for name, arr in cases.items():
        path = str(tmp_path / f"{name}.png")
        Image.fromarray(arr).save(path)
        assert np.array_equal(restirpt.read_image(path), _stb(ref, path)), name
    Image.fromarray(cases["rgb"]).convert("P", palette=Image.ADAPTIVE, colors=64).save(str(tmp_path / "pal.png"))
    assert np.array_equal(restirpt.read_image(str(tmp_path / "pal.png")), _stb(ref, str(tmp_path / "pal.png")))
    xml = prepare_assets.ajar_xml()
    if xml:
        png = os.path.join(os.path.dirname(xml), "textures", "checkerboxsmall.png")
        assert np.array_equal(restirpt.read_image(png), _stb(ref, png))


def test_jpeg_decoder_stays_within_two_levels_of_stb_image(ref):
    """stb_image's JPEG path uses its own integer IDCT and a filtered chroma upsampler; host/Image.cpp follows the specification's
    (floating-point IDCT, centred-sample upsampling).  Both are conforming; the bound below is what the test pins."""
    xml = prepare_assets.ajar_xml()
    if xml is None:
        pytest.skip("assets/_ref/VeachAjar is not prepared: JPEG pin NOT checked in this run")
    tex = os.path.join(os.path.dirname(xml), "textures")
    for name in sorted(os.listdir(tex)):
        if not name.lower().endswith((".jpg", ".jpeg")):
            continue
        a = restirpt.read_image(os.path.join(tex, name)).astype(np.int32)
        b = _stb(ref, os.path.join(tex, name)).astype(np.int32)
        assert a.shape == b.shape
        d = np.abs(a - b)
        assert d.max() <= 4 and d.mean() < 0.6, (name, int(d.max()), float(d.mean()))


def test_the_texel_sidecars_the_host_loads_are_the_reference_decoders(ref):
    xml = prepare_assets.ajar_xml()
    if xml is None:
        pytest.skip("assets/_ref/VeachAjar is not prepared: side-car pin NOT checked in this run")
    tex = os.path.join(os.path.dirname(xml), "textures")
    n = 0
    for name in sorted(os.listdir(tex)):
        if name.lower().endswith((".png", ".jpg", ".jpeg")):
            side = os.path.join(tex, name + ".ppm")
            assert os.path.exists(side), side
            assert np.array_equal(restirpt.read_image(side)[..., :3], _stb(ref, os.path.join(tex, name))[..., :3]), name
            n += 1
    assert n == 4


# ---- the shader library itself: math.glsl / material.glsl / light_sampling.glsl compiled by g++ from where they lie ----------
# (oracle/ref: glsl_to_cpp.py rewrites GLSL-only syntax, glsl_compat.h supplies vecN / the built-ins).  Integer work is compared
# bit for bit; floating point within the tolerances written below — the oracle evaluates sin / cos / pow through its own
# deterministic routines and fuses a few multiply-adds explicitly (the numeric contract it shares with the CUDA kernels), the
# reference's text compiled here uses libm and plain IEEE operations.
@pytest.fixture(scope="module")
def glsl(ref):
    from oracle import binding
    f32p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    sig = {
        "ref_hash2": (C.c_uint32, [C.c_uint32]), "ref_make_seed": (C.c_uint32, [C.c_uint32] * 3), "ref_sample1f": (C.c_float, [u32p]),
        "ref_sample4f": (None, [u32p, f32p]), "ref_sample3f": (None, [u32p, f32p]), "ref_concentric_disk": (None, [C.c_float, C.c_float, f32p]),
        "ref_cosine_hemisphere": (None, [f32p, C.c_float, C.c_float, f32p]), "ref_uv_to_bary": (None, [C.c_float, C.c_float, f32p]),
        "ref_luminance": (C.c_float, [f32p]),
        "ref_eval_bsdf": (None, [C.c_void_p] + [f32p] * 6), "ref_sample_bsdf": (C.c_int, [C.c_void_p] + [f32p] * 7 + [u32p]),
        "ref_is_bsdf_delta": (C.c_int, [C.c_void_p]), "ref_is_bsdf_connectible": (C.c_int, [C.c_void_p]),
        "ref_sample_light": (None, [C.c_void_p, C.c_void_p] + [f32p] * 8 + [u32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(ref, name)
        fn.restype, fn.argtypes = res, args
    return ref, binding.oracle_lib()


def _f(*v):
    return (C.c_float * len(v))(*v)


def test_rng_equals_the_references_math_glsl_bit_for_bit(glsl):
    ref, orc = glsl
    rng = np.random.default_rng(1)
    for x in [0, 1, 61, 0x7fffffff, 0x80000000, 0xffffffff] + [int(v) for v in rng.integers(0, 2**32, 2000)]:
        assert ref.ref_hash2(x) == orc.orc_hash2(x)
    for seed, x, y in rng.integers(0, 2**32, (2000, 3)):
        seed, x, y = int(seed), int(x) % 4096, int(y) % 4096
        assert ref.ref_make_seed(seed, x, y) == orc.orc_make_seed(seed, x, y)
    a, b = C.c_uint32(12345), C.c_uint32(12345)
    for _ in range(5000):     # the stream, including the value that rounds to exactly 1.0
        assert np.float32(ref.ref_sample1f(C.byref(a))).view(np.uint32) == np.float32(orc.orc_sample1f(C.byref(b))).view(np.uint32)
        assert a.value == b.value
    # sample4f / sample3f consume the stream x, y, z, w (GLSL evaluates constructor arguments left to right)
    a, b = C.c_uint32(777), C.c_uint32(777)
    out = _f(0, 0, 0, 0)
    ref.ref_sample4f(C.byref(a), out)
    want = [orc.orc_sample1f(C.byref(b)) for _ in range(4)]
    assert list(out) == want and a.value == b.value
    ref.ref_sample3f(C.byref(a), out)
    assert list(out)[:3] == [orc.orc_sample1f(C.byref(b)) for _ in range(3)]


def test_sampling_helpers_equal_the_references_math_glsl(glsl):
    ref, orc = glsl
    rng = np.random.default_rng(2)
    a, b = _f(0, 0), _f(0, 0)
    for u, v in np.concatenate([rng.random((3000, 2)), [[0.0, 0.0], [1.0, 1.0], [0.5, 0.5], [0.25, 0.75]]]).astype(np.float32):
        ref.ref_concentric_disk(u, v, a)
        orc.orc_concentric_disk(u, v, b)
        if u == 0.5 and v == 0.5:      # the remapped centre is 0 / 0 in the shader (only the raw (0, 0) input is guarded): NaN on both sides
            assert all(math.isnan(x) for x in (a[0], a[1], b[0], b[1]))
            continue
        assert abs(a[0] - b[0]) <= 2e-6 and abs(a[1] - b[1]) <= 2e-6, (u, v)      # cos / sin: libm against the oracle's polynomial
    n = _f(0.3, -0.5, 0.81)
    a3, b3 = _f(0, 0, 0), _f(0, 0, 0)
    for u, v in rng.random((200, 2)).astype(np.float32):
        ref.ref_uv_to_bary(u, v, a)
        r = np.sqrt(np.float32(v))
        assert a[0] == np.float32(1.0) - r and a[1] == np.float32(u) * r             # uvToBary: IEEE sqrt, exact
        ref.ref_cosine_hemisphere(n, u, v, a3)
        assert abs(np.linalg.norm(list(a3)) - 1.0) < 1e-5 and np.dot(list(a3), list(n)) > -1e-5


MATS = [(1, 0.0, 0.5, 1.5), (2, 0.0, 0.5, 1.5), (2, 1.0, 0.17, 1.5), (2, 0.3, 0.05, 1.5), (2, 0.95, 0.005, 1.5), (3, 0.0, 0.3, 1.4),
        (3, 0.0, 0.005, 2.0), (4, 0.0, 0.0, 1.5), (4, 0.0, 0.0, 1.0 / 1.5), (6, 0.0, 0.0, 1.0)]


def _dirs(rng, n):
    v = rng.normal(size=(n, 3)).astype(np.float32)
    return v / np.linalg.norm(v, axis=1, keepdims=True)


@pytest.mark.parametrize("mtype,metallic,roughness,ior", MATS)
def test_bsdf_library_equals_the_references_material_glsl(glsl, mtype, metallic, roughness, ior):
    """evalBSDF / evalPdf / sampleBSDF / isBSDFDelta / isBSDFConnectible of the oracle against the reference's own text for every
    material type the reference renders (material.glsl:286-360), random directions on both sides of the surface"""
    ref, orc = glsl
    mat = restirpt.Material()
    mat.baseColor[:] = [0.8, 0.6, 0.3]
    mat.type, mat.textureIdx, mat.metallic, mat.roughness, mat.ior = mtype, 0xffffffff, metallic, roughness, ior
    mp = C.cast(C.byref(mat), C.c_void_p)
    assert ref.ref_is_bsdf_delta(mp) == orc.orc_is_bsdf_delta(mp) and ref.ref_is_bsdf_connectible(mp) == orc.orc_is_bsdf_connectible(mp)
    rng = np.random.default_rng(mtype * 100 + int(roughness * 1000))
    albedo = _f(0.7, 0.5, 0.4)
    n_eval = n_sample = 0
    for nrm, wo, wi, r in zip(_dirs(rng, 600), _dirs(rng, 600), _dirs(rng, 600), rng.random((600, 3)).astype(np.float32)):
        if np.dot(nrm, wo) < 0:
            wo = -wo if mtype != 4 else wo           # (the dielectric is entered from both sides)
        fa, fb, pa, pb = _f(0, 0, 0), _f(0, 0, 0), C.c_float(), C.c_float()
        ref.ref_eval_bsdf(mp, albedo, _f(*nrm), _f(*wo), _f(*wi), fa, C.byref(pa))
        orc.orc_eval_bsdf(mp, albedo, _f(*nrm), _f(*wo), _f(*wi), fb, C.byref(pb))
        # off the peak the same 1 - cos^2 cancellation is compared with the lobe's width alpha = roughness^2 (only visible below 0.01)
        rt = 3e-4 + (4e-7 / roughness ** 2 if mtype in (2, 3) else 0.0)
        assert np.allclose(list(fa), list(fb), rtol=rt, atol=1e-6), (list(fa), list(fb))
        assert abs(pa.value - pb.value) <= rt * abs(pa.value) + 1e-6
        n_eval += any(v > 0 for v in fa)
        wa, wb, ba, bb, ta, tb = _f(0, 0, 0), _f(0, 0, 0), _f(0, 0, 0), _f(0, 0, 0), C.c_uint32(), C.c_uint32()
        oka = ref.ref_sample_bsdf(mp, albedo, _f(*nrm), _f(*wo), _f(*r), wa, ba, C.byref(pa), C.byref(ta))
        okb = orc.orc_sample_bsdf(mp, albedo, _f(*nrm), _f(*wo), _f(*r), wb, bb, C.byref(pb), C.byref(tb))
        assert oka == okb and ta.value == tb.value, (oka, okb, ta.value, tb.value)
        if oka:
            n_sample += 1
            # the sampled direction: the visible-normal sampler takes sqrt(1 - |p|^2) of a disk point, which amplifies the last bits of
            # cos / sin near the rim (2e-4); everything else is within 2e-5
            assert np.allclose(list(wa), list(wb), atol=2e-4 if mtype in (2, 3) else 2e-5), (list(wa), list(wb))
            # at the sampled direction a glossy lobe sits on its peak, where GTR2's denominator is 1 - cos^2 + alpha^2 cos^2: the float32
            # cancellation in 1 - cos^2 (a few 1e-7) is compared with alpha^2 = roughness^4, so the value is only defined to
            # ~1e-6 / alpha^2 in ANY float32 evaluation order (the random directions above are off the peak and agree to 3e-4).
            # Below roughness 0.02 the sampled value carries no comparable digits (the metallic workflow at roughness 0.005 returns
            # 3.6e4 from the reference's libm direction and inf from the oracle's: cos rounds to exactly 1) — only the direction,
            # the lobe type and the validity flag are held there.
            rtol = 1e-3 + (1e-6 / roughness ** 4 if mtype in (2, 3) else 0.0)
            if rtol < 0.5:
                if ta.value & 4:      # Specular: delta lobes carry their weight in bsdf / pdf directly
                    assert np.allclose(list(ba), list(bb), rtol=rtol, atol=1e-6) and abs(pa.value - pb.value) <= rtol * abs(pa.value) + 1e-6
                else:                 # value and density the reference returns = the oracle's evaluation AT THE REFERENCE'S direction
                    orc.orc_eval_bsdf(mp, albedo, _f(*nrm), _f(*wo), wa, fb, C.byref(pb))
                    assert np.allclose(list(ba), list(fb), rtol=rtol, atol=1e-6), (list(ba), list(fb))
                    assert abs(pa.value - pb.value) <= rtol * abs(pa.value) + 1e-6
    assert n_sample > 100 and (n_eval > 50 or mtype in (4, 6) or (mtype == 3 and roughness < 0.01))


def test_light_sampling_equals_the_references_light_sampling_glsl(glsl):
    """sampleLightByPower (alias-table pick + uniform point on the triangle + solid-angle pdf, light_sampling.glsl:6-37) on the
    shipped scene's light table and on a many-light table"""
    ref, orc = glsl
    from common import Backend
    scenes = [restirpt.HostScene.cornell(), restirpt.HostScene.room(2000, 3)]
    xml = prepare_assets.ajar_xml()
    if xml:
        scenes.append(restirpt.HostScene.xml(xml))
    rng = np.random.default_rng(9)
    for sc in scenes:
        b = Backend("oracle", sc, 8, 8)
        d = sc.desc
        for _ in range(400):
            p = _f(*rng.uniform(-1.5, 1.5, 3))
            r4 = _f(*rng.random(4).astype(np.float32))
            A = [_f(0, 0, 0), _f(0, 0, 0), C.c_float(), C.c_float(), C.c_float(), _f(0, 0), C.c_uint32()]
            Bv = [_f(0, 0, 0), _f(0, 0, 0), C.c_float(), C.c_float(), C.c_float(), _f(0, 0), C.c_uint32()]
            ref.ref_sample_light(d.lightSampleTable, d.triangleLights, p, r4, A[0], A[1], C.byref(A[2]), C.byref(A[3]), C.byref(A[4]), A[5], C.byref(A[6]))
            orc.orc_sample_light(b.scene, p, r4, Bv[0], Bv[1], C.byref(Bv[2]), C.byref(Bv[3]), C.byref(Bv[4]), Bv[5], C.byref(Bv[6]))
            assert A[6].value == Bv[6].value
            assert np.allclose(list(A[0]), list(Bv[0]), rtol=1e-5) and np.allclose(list(A[1]), list(Bv[1]), atol=2e-6)
            for k in (2, 3, 4):
                assert abs(A[k].value - Bv[k].value) <= 2e-5 * abs(A[k].value) + 1e-7, k
            assert list(A[5]) == list(Bv[5])                     # barycentrics: one IEEE sqrt, two multiplies — the same bits
        b.close()


# ---- the same library with the numeric contract's built-ins (refc_*): bit for bit -------------------------------------------------
def test_bsdf_library_with_contract_builtins_equals_the_oracle_bit_for_bit(glsl):
    """material.glsl compiled with the built-in library of the numeric contract (glsl_compat.h -DGLSL_BUILTINS_CONTRACT: dot / cross /
    normalize / mix / reflect / mat3 * vec3 / sin / cos as oracle_math.h evaluates them): evalBSDF, evalPdf and sampleBSDF of every
    material type return the oracle's bits on 3000 random configurations each — including the dielectric, whose Fresnel term is
    the #else branch of material.glsl:41-57 (`#if MATERIAL_DIELECTRIC_USE_SCHLICK_APPROX` is false: the macro is `true`, no macro)"""
    ref, orc = glsl
    f32p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    ref.refc_eval_bsdf.restype, ref.refc_eval_bsdf.argtypes = None, [C.c_void_p] + [f32p] * 6
    ref.refc_sample_bsdf.restype, ref.refc_sample_bsdf.argtypes = C.c_int, [C.c_void_p] + [f32p] * 7 + [u32p]
    bits = lambda a: np.array(list(a), dtype=np.float32).view(np.uint32).tolist()
    for mtype, metallic, roughness, ior in MATS + [(0, 0.0, 0.5, 1.5), (5, 0.0, 0.2, 1.5)]:
        mat = restirpt.Material()
        mat.baseColor[:] = [0.8, 0.6, 0.3]
        mat.type, mat.textureIdx, mat.metallic, mat.roughness, mat.ior = mtype, 0xffffffff, metallic, roughness, ior
        mp = C.cast(C.byref(mat), C.c_void_p)
        rng = np.random.default_rng(mtype * 7 + 1)
        albedo = _f(0.7, 0.5, 0.4)
        for nrm, wo, wi, r in zip(_dirs(rng, 3000), _dirs(rng, 3000), _dirs(rng, 3000), rng.random((3000, 3)).astype(np.float32)):
            if np.dot(nrm, wo) < 0 and mtype != 4:
                wo = -wo
            fa, fb, pa, pb = _f(0, 0, 0), _f(0, 0, 0), C.c_float(), C.c_float()
            ref.refc_eval_bsdf(mp, albedo, _f(*nrm), _f(*wo), _f(*wi), fa, C.byref(pa))
            orc.orc_eval_bsdf(mp, albedo, _f(*nrm), _f(*wo), _f(*wi), fb, C.byref(pb))
            assert bits(fa) == bits(fb) and bits([pa.value]) == bits([pb.value]), (mtype, list(fa), list(fb))
            wa, wb, ba, bb, ta, tb = _f(0, 0, 0), _f(0, 0, 0), _f(0, 0, 0), _f(0, 0, 0), C.c_uint32(), C.c_uint32()
            oka = ref.refc_sample_bsdf(mp, albedo, _f(*nrm), _f(*wo), _f(*r), wa, ba, C.byref(pa), C.byref(ta))
            okb = orc.orc_sample_bsdf(mp, albedo, _f(*nrm), _f(*wo), _f(*r), wb, bb, C.byref(pb), C.byref(tb))
            assert oka == okb and ta.value == tb.value, (mtype, oka, okb, ta.value, tb.value)
            if oka:
                assert bits(wa) == bits(wb) and bits(ba) == bits(bb) and bits([pa.value]) == bits([pb.value]), (mtype, list(wa), list(wb))


def test_light_sampling_with_contract_builtins_equals_the_oracle_bit_for_bit(glsl):
    ref, orc = glsl
    from common import Backend
    f32p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    ref.refc_sample_light.restype, ref.refc_sample_light.argtypes = None, [C.c_void_p, C.c_void_p] + [f32p] * 8 + [u32p]
    scenes = [restirpt.HostScene.cornell(), restirpt.HostScene.room(2000, 3)]
    xml = prepare_assets.ajar_xml()
    if xml:
        scenes.append(restirpt.HostScene.xml(xml))
    rng = np.random.default_rng(19)
    for sc in scenes:
        b = Backend("oracle", sc, 8, 8)
        d = sc.desc
        for _ in range(3000):
            p = _f(*rng.uniform(-2.0, 2.0, 3))
            r4 = _f(*rng.random(4).astype(np.float32))
            A = [_f(0, 0, 0), _f(0, 0, 0), C.c_float(), C.c_float(), C.c_float(), _f(0, 0), C.c_uint32()]
            Bv = [_f(0, 0, 0), _f(0, 0, 0), C.c_float(), C.c_float(), C.c_float(), _f(0, 0), C.c_uint32()]
            ref.refc_sample_light(d.lightSampleTable, d.triangleLights, p, r4, A[0], A[1], C.byref(A[2]), C.byref(A[3]), C.byref(A[4]), A[5], C.byref(A[6]))
            orc.orc_sample_light(b.scene, p, r4, Bv[0], Bv[1], C.byref(Bv[2]), C.byref(Bv[3]), C.byref(Bv[4]), Bv[5], C.byref(Bv[6]))
            va = np.array(list(A[0]) + list(A[1]) + [A[2].value, A[3].value, A[4].value] + list(A[5]), dtype=np.float32).view(np.uint32)
            vb = np.array(list(Bv[0]) + list(Bv[1]) + [Bv[2].value, Bv[3].value, Bv[4].value] + list(Bv[5]), dtype=np.float32).view(np.uint32)
            assert A[6].value == Bv[6].value and np.array_equal(va, vb)
        b.close()
