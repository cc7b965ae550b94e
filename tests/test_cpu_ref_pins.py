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
