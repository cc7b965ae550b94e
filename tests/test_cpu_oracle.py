"""CPU-side tests (no GPU): the oracle's building blocks against independent restatements written here in
Python / numpy float64 straight from the reference GLSL, plus known-answer vectors for the integer RNG.

The reference ships no tests or golden vectors; the pin against the reference itself is its shader source
compiled for the CPU (tests/test_cpu_ref_pins.py, tests/test_cpu_ref_shaders.py) — the checks here are the independent
second opinion on the building blocks."""
import ctypes as C
import math

import numpy as np
import pytest

import restirpt
from restirpt import Material, P
from oracle import binding


@pytest.fixture(scope="module")
def orc(built):
    return binding.oracle_lib()


# ---- integer RNG: reference math.glsl:227-266 restated with Python ints ------------------------------------------
def py_hash2(seed):
    seed &= 0xffffffff
    seed = (seed ^ 61) ^ (seed >> 16)
    seed = (seed * 9) & 0xffffffff
    seed ^= seed >> 4
    seed = (seed * 0x27d4eb2d) & 0xffffffff
    seed ^= seed >> 15
    return seed


def py_make_seed(seed, x, y):
    a = ((seed + x) & 0xffffffff) ^ ((y - 1) & 0xffffffff)
    b = (y * ((x - 2) & 0xffffffff)) & 0xffffffff
    return (py_hash2(a) + py_hash2(b)) & 0xffffffff


# known answers of Wang's hash: frozen outputs of the Python-int restatement above (the reference has no vectors)
HASH2_KAT = {0: 3232319850, 1: 663891101, 2: 3329832309, 61: 0, 0xffffffff: 1895078355, 123456789: 2974685137,
             0x80000000: 2903943700, 42: 1462734105}


def test_hash2_known_answers(orc):
    for k, v in HASH2_KAT.items():
        assert py_hash2(k) == v
        assert orc.orc_hash2(k) == v
        assert restirpt.hash2(k) == v


def test_make_seed_and_sample_stream(orc):
    rng = np.random.default_rng(0)
    for _ in range(200):
        seed, x, y = (int(v) for v in rng.integers(0, 2**32, 3))
        x %= 4096
        y %= 4096
        assert orc.orc_make_seed(seed, x, y) == py_make_seed(seed, x, y)
    state = C.c_uint32(py_make_seed(7, 3, 5))
    ref = state.value
    for _ in range(16):
        ref = py_hash2(ref)
        got = orc.orc_sample1f(C.byref(state))
        assert state.value == ref
        assert got == np.float32(np.float32(ref) / np.float32(4294967295.0))   # float(u) / 4294967295.0 in fp32
    # the divisor rounds to 2^32, so u = 0xffffffff gives exactly 1.0 (documented reference behaviour)
    assert np.float32(np.float32(0xffffffff) / np.float32(4294967295.0)) == np.float32(1.0)


def test_sincos_polynomial_accuracy(orc):
    s, c = C.c_float(), C.c_float()
    worst = 0.0
    for x in np.linspace(-3.5, 3.5, 4001):
        orc.orc_sincos(float(np.float32(x)), C.byref(s), C.byref(c))
        worst = max(worst, abs(s.value - math.sin(np.float32(x))), abs(c.value - math.cos(np.float32(x))))
    assert worst < 2.5e-7   # a couple of fp32 ulps, far inside Vulkan's 2^-11 allowance for sin/cos


def test_round_through_half_matches_numpy(orc):
    rng = np.random.default_rng(3)
    vals = np.concatenate([rng.normal(0, 0.05, 4000), rng.normal(0, 1e-5, 2000), [0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e-8, 5.96e-8, 2.98e-8, 2.9802322e-8]])
    for v in vals.astype(np.float32):
        want = np.float32(np.float16(v))
        got = np.float32(orc.orc_round_through_half(float(v)))
        assert want == got or (np.isinf(want) and np.isinf(got)), (v, want, got)


def test_concentric_disk(orc):
    out = (C.c_float * 2)()
    rng = np.random.default_rng(4)
    for u, v in rng.random((500, 2)):
        orc.orc_concentric_disk(float(np.float32(u)), float(np.float32(v)), out)
        a, b = 2 * np.float32(u) - 1, 2 * np.float32(v) - 1   # math.glsl:23-39 in float64
        if a * a > b * b:
            r, phi = a, math.pi * b / a * 0.25
        else:
            r, phi = b, math.pi * 0.5 - math.pi * a / b * 0.25
        assert abs(out[0] - r * math.cos(phi)) < 1e-6 and abs(out[1] - r * math.sin(phi)) < 1e-6
        assert out[0] ** 2 + out[1] ** 2 <= 1.0 + 1e-6


# ---- BSDFs: reference material.glsl restated in float64 ------------------------------------------------------------
def ref_eval(mat, albedo, n, wo, wi):
    def schlick_g(c, alpha):
        a = alpha * 0.5
        return c / (c * (1 - a) + a)

    def gtr2(c, alpha):
        if c < 1e-6:
            return 0.0
        aa = alpha * alpha
        d = c * c * (aa - 1) + 1
        return aa / (d * d * math.pi)

    def gtr2_pdf(n, m, wo, alpha):
        return gtr2(n @ m, alpha) * schlick_g(n @ wo, alpha) * abs(m @ wo) / abs(n @ wo)

    t = mat.type
    if t == 1:
        return albedo / math.pi, abs(n @ wi) / math.pi
    wh = (wo + wi) / np.linalg.norm(wo + wi)
    cos_o, cos_i = n @ wo, n @ wi
    alpha = mat.roughness ** 2
    if t == 2:
        pdf = (max(n @ wi, 0) / math.pi) * (1 - 1 / (2 - mat.metallic)) + gtr2_pdf(n, wh, wo, alpha) / (4 * abs(wh @ wo)) * (1 / (2 - mat.metallic))
        if cos_i * cos_o < 1e-7:
            return np.zeros(3), pdf
        f0 = 0.08 * (1 - mat.metallic) + albedo * mat.metallic
        f = f0 + (1 - f0) * (1 - wh @ wo) ** 5
        g = schlick_g(abs(cos_o), alpha) * schlick_g(abs(cos_i), alpha)
        d = gtr2(n @ wh, alpha)
        spec = g * d / (4 * cos_i * cos_o)
        return albedo / math.pi * (1 - mat.metallic) * (1 - f) + spec * f, pdf
    if t == 3:
        if mat.roughness < 0.01:
            return np.zeros(3), 0.0
        pdf = gtr2_pdf(n, wh, wo, alpha) / (4 * abs(wh @ wo))
        if cos_i * cos_o < 1e-7:
            return np.zeros(3), pdf
        f0 = abs(1 - mat.ior) / (1 + mat.ior)
        f = f0 + (1 - f0) * (1 - abs(wh @ wo)) ** 5
        g = schlick_g(abs(cos_o), alpha) * schlick_g(abs(cos_i), alpha)
        return albedo * f * g * gtr2(n @ wh, alpha) / (4 * cos_i * cos_o), pdf
    return np.zeros(3), 0.0


def _vec(a):
    return (C.c_float * 3)(*[float(x) for x in a])


def _unit(rng):
    v = rng.normal(size=3)
    return v / np.linalg.norm(v)


@pytest.mark.parametrize("mtype,metallic,roughness", [(1, 0, 1), (2, 0.0, 0.3), (2, 1.0, 0.17), (2, 0.5, 0.6), (3, 0, 0.25), (4, 0, 0), (6, 0, 0)])
def test_bsdf_eval_against_float64_restatement(orc, mtype, metallic, roughness):
    rng = np.random.default_rng(10 + mtype)
    mat = Material((C.c_float * 3)(0.8, 0.7, 0.6), mtype, 0xffffffff, metallic, roughness, 1.5)
    out, pdf = (C.c_float * 3)(), C.c_float()
    for _ in range(300):
        n = _unit(rng)
        wo, wi = _unit(rng), _unit(rng)
        if n @ wo < 0:
            wo = -wo
        if n @ wi < 0:
            wi = -wi
        n32, wo32, wi32 = (np.float32(v).astype(np.float64) for v in (n, wo, wi))
        albedo = np.array([0.8, 0.7, 0.6], dtype=np.float32).astype(np.float64)
        orc.orc_eval_bsdf(C.byref(mat), _vec(albedo), _vec(n32), _vec(wo32), _vec(wi32), out, C.byref(pdf))
        f_ref, pdf_ref = ref_eval(mat, albedo, n32, wo32, wi32)
        assert np.allclose(np.array(out[:]), f_ref, rtol=2e-4, atol=1e-6)
        assert abs(pdf.value - pdf_ref) <= 2e-4 * abs(pdf_ref) + 1e-6


@pytest.mark.parametrize("mtype,metallic,roughness", [(1, 0, 1), (2, 0.0, 0.3), (2, 1.0, 0.17), (3, 0, 0.25)])
def test_bsdf_sampling_is_consistent_with_eval(orc, mtype, metallic, roughness):
    """sampleBSDF must return exactly evalBSDF / evalPdf for the direction it picked, inside the upper hemisphere,
    and E[f cos / pdf] must stay below 1 (energy conservation of the lobes)"""
    rng = np.random.default_rng(20 + mtype)
    mat = Material((C.c_float * 3)(0.9, 0.9, 0.9), mtype, 0xffffffff, metallic, roughness, 1.5)
    n = np.array([0.0, 0.0, 1.0])
    wo = np.array([0.3, -0.2, 0.93])
    wo /= np.linalg.norm(wo)
    wi, bsdf, f2 = (C.c_float * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
    pdf, pdf2, typ = C.c_float(), C.c_float(), C.c_uint32()
    acc, cnt = 0.0, 0
    for _ in range(4000):
        r3 = rng.random(3)
        ok = orc.orc_sample_bsdf(C.byref(mat), _vec([0.9] * 3), _vec(n), _vec(wo), _vec(r3), wi, bsdf, C.byref(pdf), C.byref(typ))
        cnt += 1
        if not ok or pdf.value < 1e-6:
            continue
        w = np.array(wi[:])
        assert abs(np.linalg.norm(w) - 1) < 1e-4 and w[2] >= -1e-6
        orc.orc_eval_bsdf(C.byref(mat), _vec([0.9] * 3), _vec(n), _vec(wo), wi, f2, C.byref(pdf2))
        assert np.array_equal(np.array(bsdf[:]), np.array(f2[:])) and pdf.value == pdf2.value
        acc += bsdf[0] * abs(w[2]) / pdf.value
    assert acc / cnt < 1.02


# ---- ray casting: oracle BVH vs its brute-force definition vs numpy float64 Möller–Trumbore --------------------------
def test_oracle_traversal_against_numpy(orc):
    from common import Backend, random_rays
    sc = restirpt.HostScene.cornell()
    b = Backend("oracle", sc, 8, 8)
    rng = np.random.default_rng(7)
    rays = random_rays(rng, 3000, -0.9, 0.9)
    rays[:, 2] = np.abs(rays[:, 2]) + 0.05
    got = b.trace_closest(rays)
    # flattened world triangles, float64
    d = sc.desc
    verts = np.ctypeslib.as_array(C.cast(d.vertices, C.POINTER(C.c_float)), (d.numVertices, 8))
    idx = np.ctypeslib.as_array(C.cast(d.indices, C.POINTER(C.c_uint32)), (d.numIndices,))
    inst = np.ctypeslib.as_array(C.cast(d.instances, C.POINTER(C.c_float)), (d.numInstances, 56))
    inst_u = inst.view(np.uint32)
    lights = np.ctypeslib.as_array(C.cast(d.triangleLights, C.POINTER(C.c_float)), (d.numTriangleLights, 16))
    tris, ids = [], []
    for i, L in enumerate(lights):
        tris.append([L[0:3], L[4:7], L[8:11]]); ids.append((0, i))
    for k in range(d.numInstances):
        M = inst[k, 0:16].reshape(4, 4).T.astype(np.float64)
        off, cnt = int(inst_u[k, 52]), int(inst_u[k, 53])
        for t in range(cnt // 3):
            p = [M @ np.append(verts[idx[off + 3 * t + c], 0:3].astype(np.float64), 1.0) for c in range(3)]
            tris.append([q[:3] for q in p]); ids.append((k + 1, t))
    tris = np.array(tris, dtype=np.float64)
    v0, e1, e2 = tris[:, 0], tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]
    agree = 0
    id_index = {v: i for i, v in enumerate(ids)}
    for r, g in zip(rays, got):
        o, dd = r[0:3].astype(np.float64), r[4:7].astype(np.float64)
        pv = np.cross(dd, e2)
        det = np.einsum("ij,ij->i", e1, pv)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            s = o - v0
            u = np.einsum("ij,ij->i", s, pv) * inv
            q = np.cross(s, e1)
            v = (q @ dd) * inv
            t = np.einsum("ij,ij->i", e2, q) * inv
        ok = (u >= 0) & (v >= 0) & (u + v <= 1) & (t > 1e-4) & (t < 1e7)
        if ok.any():
            tbest = np.where(ok, t, np.inf).min()
            pick = (int(g["instanceIdx"]), int(g["triangleIdx"]))
            if pick in id_index:
                # coplanar faces (box bottoms on the floor) are exact ties: accept any triangle at the minimum distance
                agree += abs(t[id_index[pick]] - tbest) <= 1e-5 * max(1.0, tbest)
        else:
            agree += int(g["instanceIdx"]) == 0xffffffff
    # fp32 vs fp64 can only disagree for rays grazing an edge; with 3000 random rays that is at most a handful
    assert agree >= len(rays) - 3
    b.lib.orc_scene_set_brute_force(b.scene, 1)
    assert np.array_equal(b.trace_closest(rays), got)
    b.close()


def test_oracle_furnace_lambert_energy(orc):
    """white-furnace style sanity: naive PT in the Cornell box produces finite, non-negative, bounded radiance and
    accumulates (running mean over frameIndex, gi_naive.comp:23-27)"""
    from common import Backend, run_frames
    sc = restirpt.HostScene.cornell()
    b = Backend("oracle", sc, 64, 36)
    run_frames(b, sc.camera(64, 36), "naive", 4, accumulate=True)
    out = b.read("INDIRECT_OUTPUT")
    assert np.isfinite(out).all() and (out[..., :3] >= 0).all() and out[..., :3].max() <= 1e4
    assert 0.01 < out[..., :3].mean() < 5.0
    b.close()


# ---- two-level scenes: the oracle's instanced definition (object-space triangle test under every instance) ----------------
def test_oracle_two_level_equals_its_brute_force_and_agrees_with_the_flattened_scene(orc):
    """RPT_SCENE_TWO_LEVEL (reference src/Scene.cpp:448-547: BLAS per mesh, TLAS of instances, ray transformed at the instance
    boundary): (1) the per-mesh BVHs are only accelerators of the instanced brute force — same bits; (2) against the flattened
    world-space definition of the same scene the triangle hit is the same except for rays within rounding of an edge, and t is
    shared because the direction is not renormalised (barycentrics agree to ~1e-4); (3) one copy of the shared mesh."""
    from common import Backend, random_rays, camera_rays
    sc = restirpt.HostScene.field(1, 3, 42, shared=True)
    w, h = 64, 36
    flat = Backend("oracle", sc, w, h)
    sc.set_two_level(True)
    two = Backend("oracle", sc, w, h)
    assert flat.lib.orc_scene_num_triangles(flat.scene) == two.lib.orc_scene_num_triangles(two.scene)   # flattened numbering is the tie order
    cam = sc.camera(w, h)
    o, d = camera_rays(cam, w, h)
    rays = np.zeros((w * h, 8), dtype=np.float32)
    rays[:, 0:3] = o; rays[:, 3] = 1e-4; rays[:, 4:7] = d.reshape(-1, 3); rays[:, 7] = 1e7
    rng = np.random.default_rng(5)
    rays = np.concatenate([rays, random_rays(rng, 3000, -2.0, 2.0)])
    a, b = flat.trace_closest(rays), two.trace_closest(rays)
    same = (a["instanceIdx"] == b["instanceIdx"]) & (a["triangleIdx"] == b["triangleIdx"])
    assert same.mean() > 0.998, same.mean()
    hit = same & (a["instanceIdx"] != 0xffffffff)
    assert hit.mean() > 0.5 and (a["instanceIdx"][hit] > 1).mean() > 0.03     # the instanced blobs are hit, not only the room
    assert np.abs(a["bary"][hit] - b["bary"][hit]).max() < 1e-3
    sa, sb = flat.trace_shadow(rays), two.trace_shadow(rays)
    assert (sa == sb).mean() > 0.998
    two.lib.orc_scene_set_brute_force(two.scene, 1)
    assert np.array_equal(two.trace_closest(rays), b)
    assert np.array_equal(two.trace_shadow(rays), sb)
    flat.close(); two.close()


def test_shared_field_scene_references_one_mesh(orc):
    sc = restirpt.HostScene.field(1, 3, 42, shared=True)
    dup = restirpt.HostScene.field(1, 3, 42)
    d, e = sc.desc, dup.desc
    assert d.numInstances == e.numInstances == 10 and d.flags == 0
    inst = (restirpt.ObjectInstance * d.numInstances).from_address(d.instances)
    ranges = {(i.indexOffset, i.indexCount) for i in inst}
    assert len(ranges) == 2                                  # the room + ONE blob mesh
    assert d.numIndices < e.numIndices / 4
    other = (restirpt.ObjectInstance * e.numInstances).from_address(e.instances)
    for a, b in zip(inst, other):                            # same placements as the duplicated-geometry scene
        assert list(a.transform) == list(b.transform)
    sc.set_two_level(True)
    assert sc.desc.flags == restirpt.SCENE_TWO_LEVEL


def test_oracle_per_instance_motion_vectors(orc):
    """orc_scene_set_prev_instances (the oracle's twin of rpt_scene_update_instances .. rpt_scene_end_motion): the motion image
    carries an object's own movement — for a translation across the view it is the projected shift, to first order"""
    from common import Backend
    sc = restirpt.HostScene.cornell()
    n = sc.desc.numInstances
    before = (restirpt.ObjectInstance * n)()
    C.memmove(before, sc.desc.instances, C.sizeof(before))
    sc.set_object_transform(6, (0.1, 0.0, 0.0), (1.0, 1.0, 1.0), (0.0, 0.0, 0.0))   # the short box, +0.1 along x
    w, h = 160, 90
    b = Backend("oracle", sc, w, h)
    cam = sc.camera(w, h)
    b.clear(); b.set_camera(cam, cam); b.run("gbuffer")
    assert np.abs(b.read("MOTION")).max() < 1e-3                       # camera-only motion of a static camera
    b.lib.orc_scene_set_prev_instances(b.scene, C.cast(before, C.c_void_p), n)
    b.run("gbuffer")
    mv = b.read("MOTION")
    ids = b.read("ALBEDO_MATID")[..., 1] & 0xffff
    depth = b.read("DEPTH_NORMAL")[..., 0]
    box = (ids == 6) & (depth > 0)
    assert box.sum() > 200 and np.abs(mv[~box]).max() < 1e-3
    # the camera looks down +y from (0, -3.4, 1): x is screen-right, so last frame the point was further LEFT (negative u motion);
    # at distance d the shift is 0.1 / (2 d tan(fov/2) aspect) of the film width
    d = depth[box]
    expect = -0.1 / (2.0 * d * math.tan(math.radians(22.5)) * (w / h))
    assert np.abs(mv[box][:, 0] - expect).max() < 0.25 * np.abs(expect).max()
    assert np.abs(mv[box][:, 1]).max() < 0.2 * np.abs(expect).max()
    b.lib.orc_scene_set_prev_instances(b.scene, None, 0)
    b.run("gbuffer")
    assert np.abs(b.read("MOTION")).max() < 1e-3
    b.close()
