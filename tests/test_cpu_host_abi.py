"""CPU-side tests (no GPU): host library (Scene / Camera / alias table / OBJ + XML readers / PNG writer) and the
shape of the C ABI (every symbol include/*.h declares is exported; no compute without a device)."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import restirpt
from restirpt import Camera, LightSampleTableElement, P

ROOT = restirpt.REPO_ROOT


def test_every_declared_symbol_is_exported(built):
    dev, host = restirpt.device_lib(), restirpt.host_lib()
    for header, lib, prefix in (("restirpt.h", dev, "rpt_"), ("restirpt_host.h", host, "rh_")):
        text = open(os.path.join(ROOT, "include", header)).read()
        names = set(re.findall(r"\b(%s\w+)\s*\(" % prefix, text))
        assert len(names) > 20
        for n in sorted(names):
            assert hasattr(lib, n), f"{n} declared in include/{header} but not exported"
    table = set(restirpt.DEVICE_API) | set(restirpt.HOST_API)
    declared = set(re.findall(r"\b(rpt_\w+|rh_\w+)\s*\(", open(os.path.join(ROOT, "include", "restirpt.h")).read() +
                              open(os.path.join(ROOT, "include", "restirpt_host.h")).read()))
    assert declared <= table, f"python binding misses {declared - table}"


def test_struct_sizes_match_reference_layouts(built):
    dev = restirpt.device_lib()
    want = {0: 16, 1: 16, 2: 16, 3: 16, 4: 8, 5: 8, 6: 8, 7: 64, 8: 64, 9: 64, 10: 48, 11: 48, 12: 96, 13: 96, 14: 96, 15: 16}
    for k, v in want.items():
        assert dev.rpt_buffer_stride(k) == v
        assert restirpt.BUF_DTYPE[k].itemsize == v
    assert C.sizeof(restirpt.GRISSettings) == 20 and C.sizeof(restirpt.DISettings) == 16 and C.sizeof(restirpt.PostSettings) == 16


def test_no_device_means_error_not_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    ctx = P()
    st = restirpt.device_lib().rpt_ctx_create(0, C.byref(ctx))
    assert st == -2 and not ctx.value
    assert b"no CPU path" in restirpt.device_lib().rpt_last_error(None)
    with pytest.raises(restirpt.RestirptError):
        restirpt.Device(0)


def test_alias_table_matches_distribution(built):
    host = restirpt.host_lib()
    rng = np.random.default_rng(1)
    w = rng.random(37).astype(np.float32) ** 3 + 1e-3
    table = (LightSampleTableElement * (len(w) + 1))()
    host.rh_build_alias_table(w.ctypes.data_as(C.POINTER(C.c_float)), len(w), table)
    assert table[0].failId == len(w) and abs(table[0].prob - w.sum()) < 1e-4
    # exact reconstruction of the probabilities the shader-side lookup realises (light_sampling.glsl:25-30)
    p = np.zeros(len(w))
    for i in range(len(w)):
        prob, fail = table[i + 1].prob, table[i + 1].failId
        p[i] += min(prob, 1.0) / len(w)
        if prob < 1.0:
            assert 1 <= fail <= len(w)
            p[fail - 1] += (1.0 - prob) / len(w)
    assert np.allclose(p, w / w.sum(), atol=2e-6)


def test_camera_matches_pinhole_projection(built):
    host = restirpt.host_lib()
    cam = Camera()
    host.rh_camera_init(C.byref(cam), (C.c_float * 3)(1.0, -2.0, 0.5), (C.c_float * 3)(30.0, 10.0, 0.0), 45.0, 640, 360, 0.001, 200.0)
    assert C.sizeof(cam) == 352 and cam.filmSize[0] == 640 and cam.frameIndex == 0
    front, right, up = (np.array(v[:]) for v in (cam.front, cam.right, cam.up))
    assert abs(front @ right) < 1e-6 and abs(front @ up) < 1e-6 and abs(np.linalg.norm(front) - 1) < 1e-6
    # a point seen through pixel uv must project (projView, y flipped like proj[1][1] *= -1) back to uv
    pv = np.array(cam.projView[:]).reshape(4, 4).T
    for u, v in [(0.5, 0.5), (0.1, 0.8), (0.9, 0.2)]:
        ndc = np.array([2 * u - 1, 2 * (1 - v) - 1])
        t = math.tan(math.radians(22.5))
        d = right * ndc[0] * (640 / 360) * t + up * ndc[1] * t + front
        Pw = np.array(cam.pos[:]) + 3.0 * d / np.linalg.norm(d)
        clip = pv @ np.append(Pw, 1.0)
        uv = clip[:2] / clip[3] * 0.5 + 0.5
        assert abs(uv[0] - u) < 1e-4 and abs(uv[1] - v) < 1e-4
    host.rh_camera_next_frame(C.byref(cam), 99)
    assert cam.seed == 99 and cam.frameIndex == 1 and cam.lastProjView[:] == cam.projView[:]
    host.rh_camera_update(C.byref(cam))
    assert cam.frameIndex == 0


def test_procedural_scenes(built):
    sc = restirpt.HostScene.cornell()
    assert sc.num_triangles == 36 and sc.desc.numTriangleLights == 2 and sc.desc.numInstances == 7
    room = restirpt.HostScene.room(3000, 1)
    assert room.num_triangles > 2500 and room.desc.numTextures == 1
    field = restirpt.HostScene.field(1, 3, 42)
    assert field.desc.numInstances == 10


def test_xml_and_obj_loader(tmp_path, built):
    (tmp_path / "models").mkdir()
    (tmp_path / "models" / "quad.obj").write_text(
        "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\nf 1/1/1 2/2/1 3/3/1 4/4/1\n")
    (tmp_path / "models" / "light.obj").write_text("v 0 0 2\nv 1 0 2\nv 0 1 2\nvn 0 0 -1\nf 1//1 3//1 2//1\n")
    (tmp_path / "scene.xml").write_text("""<?xml version="1.0"?>
<scene name="t"><integrator type="path"><size width="64" height="48" /></integrator>
<camera type="thinLens"><position value="0.5 -3 0.5" /><lookAt value="0.5 0 0.5" /><fov value="40" /></camera>
<modelInstances>
 <modelInstance path="models/light.obj" name="l" type="light"><transform translate="0 0 0" scale="1 1 1" rotate="0 0 0" /><radiance value="10 10 10" /></modelInstance>
 <modelInstance path="models/quad.obj" name="q" type="object"><transform translate="0 0 1" scale="2 2 2" rotate="0 0 0" />
  <material type="metalWorkflow"><baseColor value="0.5 0.25 0.125" /><metallic value="1.0" /><roughness value="0.3" /></material></modelInstance>
 <modelInstance path="models/quad.obj" name="plain" type="object"><transform translate="0 0 0" scale="1 1 1" rotate="0 0 0" /></modelInstance>
</modelInstances></scene>""")
    sc = restirpt.HostScene.xml(str(tmp_path / "scene.xml"))
    d = sc.desc
    assert d.numInstances == 2 and d.numTriangleLights == 1 and d.numIndices == 12 and d.numVertices == 8
    mats = np.ctypeslib.as_array(C.cast(d.materials, C.POINTER(C.c_float)), (d.numMaterials, 8))
    assert d.numMaterials == 3
    assert np.allclose(mats[0, :3], [1, 0, 1])                       # magenta placeholder (Resource.cpp:36-40)
    assert np.allclose(mats[1, :3], [0.5, 0.25, 0.125]) and mats[1].view(np.uint32)[3] == 2
    assert np.allclose(mats[2, :3], [0.6, 0.6, 0.6]) and mats[2].view(np.uint32)[3] == 1   # OBJ default material
    verts = np.ctypeslib.as_array(C.cast(d.vertices, C.POINTER(C.c_float)), (d.numVertices, 8))
    assert np.allclose(sorted(verts[:4, 7]), [0, 0, 1, 1])           # FlipUVs: v -> 1 - v
    light = np.ctypeslib.as_array(C.cast(d.triangleLights, C.POINTER(C.c_float)), (1, 16))[0]
    assert abs(light[15] - 0.5) < 1e-6 and np.allclose(light[12:15], 20.0)   # radiance = power / area
    inst = np.ctypeslib.as_array(C.cast(d.instances, C.POINTER(C.c_float)), (2, 56))
    M = inst[0, :16].reshape(4, 4).T
    # T * Rx(90 deg) * S(x, z, y): model +y becomes world +z (Model.cpp:11-21)
    assert np.allclose((M @ [0, 1, 0, 1])[:3], [0, 0, 3], atol=1e-5)
    cam = sc.camera()
    assert np.allclose(cam.front[:], [0, 1, 0], atol=1e-5) and cam.filmSize[0] == 64


def test_obj_objects_groups_and_mtl_materials(tmp_path, built):
    """OBJ import with assimp's structure (ObjFileParser / ObjFileImporter) as Resource::createNewModelInstance sees it:
    meshes per object and material, material list = DefaultMaterial + .mtl materials (+ undefined ones), the objects
    visited in reverse order (the reference's node stack), textures from map_Kd, the XML material applied to every mesh."""
    (tmp_path / "m").mkdir()
    tex = bytes([255, 0, 0, 0, 255, 0, 0, 0, 255, 255, 255, 255])
    (tmp_path / "m" / "tex.png.ppm").write_bytes(b"P6\n2 2\n255\n" + tex)
    (tmp_path / "m" / "two.mtl").write_text(
        "newmtl red\nKd 0.8 0.1 0.1\nKa 0 0 0\n\nnewmtl textured\nKd 0.5 0.5 0.5\nmap_Kd -s 1 1 1 tex.png\n")
    (tmp_path / "m" / "two.obj").write_text(
        "mtllib two.mtl\n"
        "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 1\nv 1 0 1\nv 1 1 1\nv 0 1 1\nvn 0 0 1\nvt 0 0.25\n"
        "f 1//1 2//1 3//1\n"                      # before any o / usemtl: object 'defaultobject', default material
        "o first\nusemtl red\nf 1//1 3//1 4//1\n"
        "usemtl textured\nf 5/1/1 6/1/1 7/1/1 8/1/1\n"   # second mesh of 'first' (a quad: 2 triangles)
        "g second\nusemtl ghost\nf 5//1 7//1 8//1\n")   # a group is an object too; 'ghost' is not in the .mtl
    xml_head = """<?xml version="1.0"?><scene name="t"><integrator type="path"><size width="32" height="32" /></integrator>
<camera type="thinLens"><position value="0 -3 0" /><lookAt value="0 0 0" /><fov value="40" /></camera><modelInstances>"""
    (tmp_path / "plain.xml").write_text(xml_head + """
 <modelInstance path="m/two.obj" name="a" type="object"><transform translate="0 0 0" scale="1 1 1" rotate="0 0 0" /></modelInstance>
</modelInstances></scene>""")
    sc = restirpt.HostScene.xml(str(tmp_path / "plain.xml"))
    d = sc.desc
    assert d.numIndices == 15 and d.numVertices == 3 + 3 + 4 + 3 and d.numInstances == 1 and d.numTextures == 1
    mats = np.ctypeslib.as_array(C.cast(d.materials, C.POINTER(C.c_float)), (d.numMaterials, 8))
    # pool: [0] magenta placeholder, then DefaultMaterial, red, textured, ghost
    assert d.numMaterials == 5
    assert np.allclose(mats[1, :3], 0.6) and np.allclose(mats[2, :3], [0.8, 0.1, 0.1]) and np.allclose(mats[4, :3], 0.6)
    tex_idx = mats.view(np.uint32)[:, 4]
    assert tex_idx[3] == 0 and tex_idx[1] == 0xffffffff and tex_idx[2] == 0xffffffff
    # objects in reverse order (second, first, defaultobject), meshes of an object in order: ghost | red, textured x2 | default
    mi = np.ctypeslib.as_array(C.cast(d.materialIndices, C.POINTER(C.c_int32)), (d.numMaterialIndices,))
    assert mi.tolist() == [4, 2, 3, 3, 1]
    verts = np.ctypeslib.as_array(C.cast(d.vertices, C.POINTER(C.c_float)), (d.numVertices, 8))
    assert np.allclose(verts[0, :3], [0, 0, 1])            # first vertex of the 'second' group's triangle (v5)
    assert np.allclose(verts[6:10, 7], 0.75)               # the textured quad's vt 0.25, V flipped
    # an XML <material> replaces the type of every mesh's material but keeps their colours / textures when it has no baseColor
    (tmp_path / "override.xml").write_text(xml_head + """
 <modelInstance path="m/two.obj" name="a" type="object"><transform translate="0 0 0" scale="1 1 1" rotate="0 0 0" />
  <material type="metalWorkflow"><metallic value="1.0" /><roughness value="0.3" /></material></modelInstance>
</modelInstances></scene>""")
    sc2 = restirpt.HostScene.xml(str(tmp_path / "override.xml"))
    m2 = np.ctypeslib.as_array(C.cast(sc2.desc.materials, C.POINTER(C.c_float)), (sc2.desc.numMaterials, 8))
    assert (m2.view(np.uint32)[1:5, 3] == 2).all() and np.allclose(m2[2, :3], [0.8, 0.1, 0.1]) and m2.view(np.uint32)[3, 4] == 0


def test_obj_face_with_a_missing_vertex_is_an_error(tmp_path, built):
    """found by fuzzing the loader under AddressSanitizer (profiles/r1_21_host_sanitizers.txt): an index past the vertex
    list (or 0) used to read outside the array; assimp rejects such files too"""
    (tmp_path / "m").mkdir()
    head = """<?xml version="1.0"?><scene name="t"><integrator type="path"><size width="32" height="32" /></integrator>
<camera type="thinLens"><position value="0 -3 0" /><lookAt value="0 0 0" /><fov value="40" /></camera><modelInstances>
 <modelInstance path="m/bad.obj" name="a" type="object"><transform translate="0 0 0" scale="1 1 1" rotate="0 0 0" /></modelInstance>
</modelInstances></scene>"""
    (tmp_path / "s.xml").write_text(head)
    for faces in ("f 1 2 9\n", "f 0 1 2\n", "f 1 2 -7\n"):
        (tmp_path / "m" / "bad.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\n" + faces)
        with pytest.raises(restirpt.RestirptError, match="does not exist"):
            restirpt.HostScene.xml(str(tmp_path / "s.xml"))
    (tmp_path / "m" / "bad.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 -1\n")   # negative = relative: fine
    assert restirpt.HostScene.xml(str(tmp_path / "s.xml")).desc.numIndices == 3


def test_set_object_transform_rewrites_the_instance(built):
    """dynamic scenes: rh_scene_set_object_transform rebuilds transform / inverse / inverse-transpose of one ObjectInstance
    (reference src/Model.cpp:11-21: T * Rz * Rx(+90) * Ry * S(x, z, y)) and leaves the others alone"""
    sc = restirpt.HostScene.cornell()
    n = sc.desc.numInstances
    before = np.ctypeslib.as_array(C.cast(sc.desc.instances, C.POINTER(C.c_float)), (n, 56)).copy()
    sc.set_object_transform(6, (0.25, -0.5, 0.125), (1.0, 1.0, 1.0), (0.0, 0.0, 0.0))
    after = np.ctypeslib.as_array(C.cast(sc.desc.instances, C.POINTER(C.c_float)), (n, 56)).copy()
    assert np.array_equal(before[:6], after[:6])
    M = after[6, :16].reshape(4, 4).T
    Minv = after[6, 16:32].reshape(4, 4).T
    MinvT = after[6, 32:48].reshape(4, 4).T
    assert np.allclose(M[:3, 3], [0.25, -0.5, 0.125]) and np.allclose(M @ Minv, np.eye(4), atol=1e-5) and np.allclose(MinvT, Minv.T)
    assert np.array_equal(before[6, 48:].view(np.uint32), after[6, 48:].view(np.uint32))   # radiance, index range untouched
    with pytest.raises(restirpt.RestirptError):
        sc.set_object_transform(99, (0, 0, 0))


def test_png_writer_roundtrip(tmp_path, built):
    from PIL import Image
    img = (np.random.default_rng(0).random((17, 31, 4)) * 255).astype(np.uint8)
    path = str(tmp_path / "x.png")
    assert restirpt.host_lib().rh_write_png(path.encode(), img.ctypes.data_as(P), 31, 17) == 0
    assert np.array_equal(np.asarray(Image.open(path)), img)


def test_film_layout_and_partition():
    import bench
    from restirpt.multigpu import partition, storage_rows
    # weak scaling: the same 16:9 view with N x the pixels of 1920x1080 (N = 4 is the 4K film of BASELINE config 4)
    assert bench.film_for(1) == (1920, 1080) and bench.film_for(2) == (2712, 1526)
    assert bench.film_for(4) == (3840, 2160) and bench.film_for(8) == (5432, 3056)
    for n in (1, 2, 4, 8):
        w, h = bench.film_for(n)
        assert abs(w * h / (1920.0 * 1080.0) - n) < 0.01 * n and abs(w / h - 16.0 / 9.0) < 2e-3
    for h, n in [(1080, 1), (2160, 4), (3056, 8), (1081, 4)]:
        parts = partition(h, n)
        assert parts[0][0] == 0 and parts[-1][1] == h and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert storage_rows(0, 270, 2160, 21) == (0, 291) and storage_rows(270, 540, 2160, 21) == (249, 561)


def _tri_soup(sc, with_normals=True):
    """world-independent triangle soup of a scene's object geometry: sorted (positions[, normals]) of every triangle corner"""
    d = sc.desc
    verts = np.ctypeslib.as_array(C.cast(d.vertices, C.POINTER(C.c_float)), (d.numVertices, 8))
    idx = np.ctypeslib.as_array(C.cast(d.indices, C.POINTER(C.c_uint32)), (d.numIndices,))
    tris = verts[idx].reshape(-1, 3, 8)
    cols = [0, 1, 2, 4, 5, 6] if with_normals else [0, 1, 2]
    rows = [tuple(np.round(t[:, cols].reshape(-1), 5)) for t in tris]
    # a triangle may start at any of its corners: rotate so that the smallest corner comes first (winding kept)
    out = []
    for t in tris:
        c = [tuple(np.round(t[k, cols], 5)) for k in range(3)]
        k = min(range(3), key=lambda j: c[j])
        out.append(c[k] + c[(k + 1) % 3] + c[(k + 2) % 3])
    return sorted(out), len(rows)


def test_ply_and_stl_models_equal_the_obj_model(tmp_path, built):
    """The reference hands every <modelInstance path> to assimp (src/Resource.cpp:100-181).  Besides OBJ the host library reads
    PLY (ascii, binary little / big endian, polygons, optional normals and texture coordinates) and STL (binary, ascii): the same
    cube through each must give the same triangles as through the OBJ reader — positions and winding always, normals where the
    file carries them — and generated smooth normals where it does not (aiProcess_GenSmoothNormals)."""
    import struct
    P = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    quads = [((0, 3, 2, 1), (0, 0, -1)), ((4, 5, 6, 7), (0, 0, 1)), ((0, 1, 5, 4), (0, -1, 0)),
             ((2, 3, 7, 6), (0, 1, 0)), ((1, 2, 6, 5), (1, 0, 0)), ((0, 4, 7, 3), (-1, 0, 0))]
    m = tmp_path / "m"
    m.mkdir()
    # OBJ with per-face normals (what the other files with normals must reproduce)
    obj = "".join(f"v {x} {y} {z}\n" for x, y, z in P) + "".join(f"vn {x} {y} {z}\n" for _, (x, y, z) in quads)
    obj += "".join("f " + " ".join(f"{v + 1}//{k + 1}" for v in q) + "\n" for k, (q, _) in enumerate(quads))
    (m / "cube.obj").write_text(obj)
    # PLY with normals needs one vertex per (position, normal): 24 vertices
    pv, pf = [], []
    for q, n in quads:
        base = len(pv)
        pv += [P[v] + n + (0.25 * i, 0.5) for i, v in enumerate(q)]
        pf.append([base, base + 1, base + 2, base + 3])
    hdr = ("ply\nformat {fmt} 1.0\ncomment made by the test\nelement vertex 24\nproperty float x\nproperty float y\nproperty float z\n"
           "property float nx\nproperty float ny\nproperty float nz\nproperty float s\nproperty float t\n"
           "element face 6\nproperty list uchar int vertex_indices\nend_header\n")
    (m / "cube_ascii.ply").write_text(hdr.format(fmt="ascii") + "".join(" ".join(str(c) for c in v) + "\n" for v in pv)
                                      + "".join("4 " + " ".join(map(str, f)) + "\n" for f in pf))
    for name, e in (("cube_le.ply", "<"), ("cube_be.ply", ">")):
        body = b"".join(struct.pack(e + "8f", *v) for v in pv) + b"".join(struct.pack(e + "B4i", 4, *f) for f in pf)
        (m / name).write_bytes(hdr.format(fmt="binary_little_endian" if e == "<" else "binary_big_endian").encode() + body)
    # PLY without normals, shared vertices, double coordinates, an extra per-vertex property and an extra element
    (m / "cube_shared.ply").write_text(
        "ply\nformat ascii 1.0\nelement vertex 8\nproperty double x\nproperty double y\nproperty double z\nproperty uchar red\n"
        "element face 6\nproperty list uchar uint vertex_index\nelement edge 1\nproperty int a\nproperty int b\nend_header\n"
        + "".join(f"{x} {y} {z} 200\n" for x, y, z in P) + "".join("4 " + " ".join(map(str, q)) + "\n" for q, _ in quads) + "0 1\n")
    # STL: triangles with facet normals
    tris = []
    for q, n in quads:
        tris += [(n, (P[q[0]], P[q[1]], P[q[2]])), (n, (P[q[0]], P[q[2]], P[q[3]]))]
    (m / "cube_bin.stl").write_bytes(b"solid made by the test".ljust(80, b" ") + struct.pack("<I", len(tris))
                                     + b"".join(struct.pack("<12fH", *n, *t[0], *t[1], *t[2], 0) for n, t in tris))
    (m / "cube_ascii.stl").write_text("solid cube\n" + "".join(
        f"facet normal {n[0]} {n[1]} {n[2]}\n outer loop\n" + "".join(f"  vertex {v[0]} {v[1]} {v[2]}\n" for v in t) + " endloop\nendfacet\n"
        for n, t in tris) + "endsolid cube\n")
    (m / "light.obj").write_text("v 0 0 3\nv 1 0 3\nv 0 1 3\nvn 0 0 -1\nf 1//1 3//1 2//1\n")

    def scene(model, kind="object"):
        xml = tmp_path / (model.replace(".", "_") + ".xml")
        xml.write_text(f"""<?xml version="1.0"?><scene name="t"><integrator type="path"><size width="32" height="32" /></integrator>
<camera type="thinLens"><position value="0.5 -3 0.5" /><lookAt value="0.5 0 0.5" /><fov value="40" /></camera><modelInstances>
 <modelInstance path="m/light.obj" name="l" type="light"><transform translate="0 0 0" scale="1 1 1" rotate="0 0 0" /><radiance value="5 5 5" /></modelInstance>
 <modelInstance path="m/{model}" name="a" type="{kind}"><transform translate="0 0 0" scale="1 1 1" rotate="0 0 0" />{'<radiance value="1 1 1" />' if kind == 'light' else ''}</modelInstance>
</modelInstances></scene>""")
        return restirpt.HostScene.xml(str(xml))

    want, n = _tri_soup(scene("cube.obj"))
    assert n == 12
    for model in ("cube_ascii.ply", "cube_le.ply", "cube_be.ply", "cube_bin.stl", "cube_ascii.stl"):
        got, k = _tri_soup(scene(model))
        assert k == 12 and got == want, model
    # texture coordinates of the PLY: s kept, t flipped (aiProcess_FlipUVs)
    d = scene("cube_le.ply").desc
    verts = np.ctypeslib.as_array(C.cast(d.vertices, C.POINTER(C.c_float)), (d.numVertices, 8))
    assert np.allclose(sorted(set(np.round(verts[:, 3], 5))), [0, 0.25, 0.5, 0.75]) and np.allclose(verts[:, 7], 0.5)
    mats = np.ctypeslib.as_array(C.cast(d.materials, C.POINTER(C.c_float)), (d.numMaterials, 8))
    assert np.allclose(mats[1, :3], 0.6)      # assimp's default material
    # shared vertices without normals: same positions / winding, smooth normals = normalised sum of the three faces at a corner
    want_pos, _ = _tri_soup(scene("cube.obj"), with_normals=False)
    sc = scene("cube_shared.ply")
    got_pos, k = _tri_soup(sc, with_normals=False)
    assert k == 12 and got_pos == want_pos
    d = sc.desc
    verts = np.ctypeslib.as_array(C.cast(d.vertices, C.POINTER(C.c_float)), (d.numVertices, 8))
    assert d.numVertices == 8
    for v in verts:
        outward = (v[0:3] * 2 - 1) / np.sqrt(3.0)
        assert np.allclose(v[4:7], outward, atol=1e-5)
    # as a light without normals: flat normals (aiProcess_GenNormals), 12 triangle lights
    assert scene("cube_shared.ply", "light").desc.numTriangleLights == 1 + 12
    # errors are reported, not crashes
    (m / "bad.ply").write_text("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                               "element face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n0 1 0\n3 0 1 7\n")
    with pytest.raises(Exception, match="vertex that does not exist"):
        scene("bad.ply")
    (m / "bad.stl").write_bytes(b"x" * 100)
    with pytest.raises(Exception):
        scene("bad.stl")
    (m / "cube.fbx").write_bytes(b"whatever")
    with pytest.raises(Exception, match="unsupported model format"):
        scene("cube.fbx")
