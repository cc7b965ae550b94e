"""GPU parity tests proper: the CUDA library (through its C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): closest-hit primitive IDs bit-exact; every G-buffer / reservoir / output buffer
is compared BIT-EXACTLY as well, because both sides implement the same fp32 numeric contract (DESIGN.md
§numerics).  The only toleranced comparison is the 8-bit post-process output (pow() comes from two different
libms): at most 1 LSB.
"""
import numpy as np
import pytest

import restirpt
from restirpt import DISettings, GRISSettings, PostSettings
from common import Backend, run_frames, bitwise_mismatch, camera_rays, random_rays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def device(built):
    dev = restirpt.Device(0)
    yield dev


SCENES = {
    "cornell": lambda: restirpt.HostScene.cornell(),
    "room": lambda: restirpt.HostScene.room(6000, 7),
    "field": lambda: restirpt.HostScene.field(1, 3, 42),
    # two-level scenes (RPT_SCENE_TWO_LEVEL: BLAS per unique mesh in object space + TLAS, reference src/Scene.cpp:448-547): nine
    # instances of ONE mesh, and the room with a BLAS per object — against the oracle's instanced definition
    "field_tlas": lambda: restirpt.HostScene.field(1, 3, 42, shared=True, two_level=True),
    "room_tlas": lambda: _two_level(restirpt.HostScene.room(6000, 7)),
}


def _two_level(sc):
    sc.set_two_level(True)
    return sc


@pytest.fixture(scope="module", params=list(SCENES))
def pair(request, device):
    sc = SCENES[request.param]()
    w, h = 96, 54
    gpu = Backend("cuda", sc, w, h, device)
    cpu = Backend("oracle", sc, w, h)
    yield request.param, sc, gpu, cpu
    gpu.close()
    cpu.close()


def test_closest_hit_ids_random_rays(pair):
    name, sc, gpu, cpu = pair
    rng = np.random.default_rng(1234)
    rays = random_rays(rng, 20000, -3.0, 3.0)
    a, b = gpu.trace_closest(rays), cpu.trace_closest(rays)
    assert np.array_equal(a["instanceIdx"], b["instanceIdx"])
    assert np.array_equal(a["triangleIdx"], b["triangleIdx"])
    hit = a["instanceIdx"] != 0xffffffff
    assert hit.mean() > 0.05
    assert np.array_equal(a["bary"][hit].view(np.uint32), b["bary"][hit].view(np.uint32))


def test_queue_traversal_kernel_matches_oracle(pair):
    """the persistent dynamic-fetch traversal kernel of the wavefront passes (trace_queue.cu), closest and any hit"""
    import ctypes as C
    from restirpt import P
    name, sc, gpu, cpu = pair
    rng = np.random.default_rng(4321)
    rays = random_rays(rng, 30011, -3.0, 3.0)     # not a multiple of the warp size
    rays[::97, 4:7] = np.nan                      # degenerate rays are misses
    want = cpu.trace_closest(rays)
    ms = C.c_float()
    got = np.zeros(rays.shape[0], dtype=restirpt.ISEC_DTYPE)
    restirpt.check(gpu.dev.ctx, gpu.lib.rpt_trace_bench(gpu.dev.ctx, gpu.scene, rays.ctypes.data_as(P), rays.shape[0], 0, 1, 1, C.byref(ms),
                                                        got.ctypes.data_as(P), None), "rpt_trace_bench")
    assert np.array_equal(got["instanceIdx"], want["instanceIdx"]) and np.array_equal(got["triangleIdx"], want["triangleIdx"])
    hit = want["instanceIdx"] != 0xffffffff
    assert np.array_equal(got["bary"][hit].view(np.uint32), want["bary"][hit].view(np.uint32))
    rays[:, 7] = 1.5
    want_occ = cpu.trace_shadow(rays)
    occ = np.zeros(rays.shape[0], dtype=np.uint8)
    restirpt.check(gpu.dev.ctx, gpu.lib.rpt_trace_bench(gpu.dev.ctx, gpu.scene, rays.ctypes.data_as(P), rays.shape[0], 1, 1, 1, C.byref(ms),
                                                        None, occ.ctypes.data_as(P)), "rpt_trace_bench")
    assert np.array_equal(occ, want_occ)


def test_shadow_rays(pair):
    name, sc, gpu, cpu = pair
    rng = np.random.default_rng(99)
    rays = random_rays(rng, 20000, -2.5, 2.5, tmax=1.5)
    a, b = gpu.trace_shadow(rays), cpu.trace_shadow(rays)
    assert np.array_equal(a, b)
    assert a.any() and not a.all()


def test_oracle_bvh_equals_brute_force(pair):
    """pins the oracle's own accelerator against its brute-force definition"""
    name, sc, gpu, cpu = pair
    rng = np.random.default_rng(5)
    rays = random_rays(rng, 1500, -3.0, 3.0)
    fast = cpu.trace_closest(rays)
    cpu.lib.orc_scene_set_brute_force(cpu.scene, 1)
    slow = cpu.trace_closest(rays)
    cpu.lib.orc_scene_set_brute_force(cpu.scene, 0)
    assert np.array_equal(fast, slow)


PASS_BUFFERS = {
    "gbuffer": ["DEPTH_NORMAL", "ALBEDO_MATID", "MOTION", "PRIMARY_ISEC"],
    "di_naive": ["DIRECT_OUTPUT"], "di_naive_rt": ["DIRECT_OUTPUT"], "gi_naive": ["INDIRECT_OUTPUT"],
    "di_pathgen": ["DI_THIS"], "di_temporal": ["DI_TEMP"], "di_spatial": ["DI_THIS", "DIRECT_OUTPUT"],
    "gi_restir": ["GI_THIS", "INDIRECT_OUTPUT"],
    "gris_pathtrace": ["GRIS_THIS"], "gris_temporal": ["GRIS_TEMP"], "gris_spatial": ["GRIS_THIS", "INDIRECT_OUTPUT"],
}


def _compare_method(pair, method, frames, moves=None, **kw):
    name, sc, gpu, cpu = pair
    cam = sc.camera(gpu.w, gpu.h)
    shots = {"cuda": {}, "oracle": {}}

    def snap(i, pass_name, backend):
        for buf in PASS_BUFFERS[pass_name]:
            shots[backend.kind][(i, pass_name, buf)] = backend.read(buf)

    for b in (gpu, cpu):
        b.clear()
        run_frames(b, cam, method, frames, moves=moves, snapshot=snap, **kw)
    assert shots["cuda"].keys() == shots["oracle"].keys()
    bad = {k: bitwise_mismatch(shots["cuda"][k], shots["oracle"][k]) for k in shots["cuda"]}
    bad = {k: v for k, v in bad.items() if v}
    assert not bad, f"{name}/{method}: pixels differing per (frame, pass, buffer): {bad}"
    return shots["cuda"]


DOLLY = [(0.0, 0.0, 0.0), (0.01, 0.02, 0.0), (0.01, 0.02, 0.005), (0.0, 0.0, 0.0)]


def test_gbuffer_and_naive_passes_bit_exact(pair):
    shots = _compare_method(pair, "naive", 2)
    depth = shots[(0, "gbuffer", "DEPTH_NORMAL")][..., 0]
    assert (depth > 0).mean() > 0.3
    out = shots[(1, "gi_naive", "INDIRECT_OUTPUT")]
    assert np.isfinite(out).all() and out[..., :3].mean() > 0


def test_naive_direct_rt_pipeline_mode_bit_exact(pair):
    """di_naive.rgen (RayTracing-pipeline mode, reference src/RayTracing.h:28-30): one light sample with weight 1 —
    a different estimator from di_naive.comp, so the images must differ from the ray-query mode's but agree in the mean"""
    name, sc, gpu, cpu = pair
    rt = _compare_method(pair, "naive_rt", 2)
    rq = _compare_method(pair, "naive", 2)
    a, b = rt[(1, "di_naive_rt", "DIRECT_OUTPUT")][..., :3], rq[(1, "di_naive", "DIRECT_OUTPUT")][..., :3]
    assert np.isfinite(a).all()
    if b.mean() > 0:   # (a scene whose light is not directly visible from anywhere has both images black)
        assert a.mean() > 0 and not np.array_equal(a, b)


@pytest.mark.parametrize("shift,sample", [(0, 0), (0, 2), (1, 2)])
def test_restir_di_bit_exact(pair, shift, sample):
    _compare_method(pair, "di", 4, moves=DOLLY, di=DISettings(shift, sample, 1, 1))


def test_restir_gi_bit_exact(pair):
    _compare_method(pair, "gi", 4, moves=DOLLY)


def test_parity_scenes_reach_the_path_tracing_tail(pair):
    """the bit-exact GRIS comparisons below only cover grisTailKernel if some paths survive bounce 6 on these small films"""
    name, sc, gpu, cpu = pair
    import ctypes as C
    run_frames(gpu, sc.camera(gpu.w, gpu.h), "gris", 1)
    wc = (C.c_uint32 * 64)()
    gpu.lib.rpt_wavefront_counters(gpu.frame, wc)
    assert wc[4 * 7] > 0, f"{name}: no path alive at bounce 7"


@pytest.mark.parametrize("shift,temporal,spatial", [(2, 1, 1), (0, 1, 1), (2, 0, 1), (2, 1, 0)])
def test_restir_pt_gris_bit_exact(pair, shift, temporal, spatial):
    shots = _compare_method(pair, "gris", 4, moves=DOLLY, gris=GRISSettings(shift, 1.0, temporal, spatial, 20))
    r = shots[(3, "gris_spatial", "GRIS_THIS")]
    assert (r["rcIsec"]["instanceIdx"] != 0xffffffff).mean() > 0.05


def test_background_pixels_keep_the_reservoir_of_two_frames_ago(pair):
    """gris_path_trace.glsl:54-56 returns early on a background pixel, so in the reference's ping-pong pair that pixel's "this"
    reservoir still holds what the frame before last left there.  The CUDA side rotates its final reservoirs through THREE buffers
    (frame pipeline) and has to carry that reservoir over — checked here with a camera that moves far enough for
    surface pixels to turn into background and back, six frames, every buffer of every pass."""
    name, sc, gpu, cpu = pair
    moves = [(0.0, 0.0, 0.0), (0.6, 0.25, 0.0), (0.6, 0.25, 0.0), (-0.9, -0.3, 0.0), (-0.9, -0.3, 0.0), (0.6, 0.1, 0.0)]
    shots = _compare_method(pair, "gris", 6, moves=moves, gris=GRISSettings(2, 1.0, 1, 1, 20))
    valid = [shots[(i, "gbuffer", "DEPTH_NORMAL")][..., 0] > 0 for i in range(6)]
    turned_background = sum(int((valid[i - 2] & ~valid[i]).sum()) for i in range(2, 6))
    if not any((~v).any() for v in valid):
        pytest.skip(f"{name}: a closed scene, no background pixels")
    assert turned_background > 0, f"{name}: the camera move never turned a surface pixel into background"
    # and the stale reservoirs are really there (not zeros): some background pixel holds a reservoir with history after the pass
    assert any((shots[(i, "gris_pathtrace", "GRIS_THIS")]["sampleCount"][~valid[i]] > 0).any() for i in range(2, 6))


def test_postprocess_within_one_lsb(pair):
    name, sc, gpu, cpu = pair
    cam = sc.camera(gpu.w, gpu.h)
    for b in (gpu, cpu):
        b.clear()
        run_frames(b, cam, "naive", 1)
    for tm in (0, 1, 2):
        st = PostSettings(tm, 1, 0, 0)
        a, b = gpu.postprocess(st).astype(np.int32), cpu.postprocess(st).astype(np.int32)
        assert np.abs(a - b).max() <= 1
        assert a[..., :3].mean() > 1


def test_visualize_as(pair):
    name, sc, gpu, cpu = pair
    cam = sc.camera(gpu.w, gpu.h)
    for b in (gpu, cpu):
        b.clear()
        b.set_camera(cam, cam)
        b.run("visualize_as")
    assert bitwise_mismatch(gpu.read("DIRECT_OUTPUT"), cpu.read("DIRECT_OUTPUT")) == 0


def test_degenerate_rays_are_misses(pair):
    """NaN / infinite / zero-length / empty-interval rays: both sides answer 'no hit' (the oracle's brute-force
    path, which has no early-out, proves that this is what the triangle test itself yields)"""
    name, sc, gpu, cpu = pair
    rng = np.random.default_rng(17)
    rays = random_rays(rng, 64, -1.0, 1.0)
    nan, inf = np.float32("nan"), np.float32("inf")
    rays[0, 0] = nan; rays[1, 5] = nan; rays[2, 4:7] = 0.0; rays[3, 3], rays[3, 7] = 5.0, 1.0
    rays[4, 1] = inf; rays[5, 6] = -inf; rays[6, 7] = nan; rays[7, 0:3] = nan; rays[7, 4:7] = nan
    a, b = gpu.trace_closest(rays), cpu.trace_closest(rays)
    cpu.lib.orc_scene_set_brute_force(cpu.scene, 1)
    c = cpu.trace_closest(rays)
    cpu.lib.orc_scene_set_brute_force(cpu.scene, 0)
    assert np.array_equal(a["instanceIdx"], b["instanceIdx"]) and np.array_equal(b["instanceIdx"], c["instanceIdx"])
    assert (a["instanceIdx"][[0, 1, 3, 4, 5, 6, 7]] == 0xffffffff).all()
    assert np.array_equal(gpu.trace_shadow(rays), cpu.trace_shadow(rays))


def test_dynamic_scene_update_equals_a_fresh_scene(device):
    """rpt_scene_update_instances (new transforms, BVH rebuilt on the GPU) must leave the device scene in the state a scene
    created from scratch with those transforms has — closest-hit ids and a ReSTIR PT frame bit for bit — and that state
    must still match the oracle"""
    import ctypes as C
    sc = restirpt.HostScene.cornell()
    w, h = 96, 54
    moved = Backend("cuda", sc, w, h, device)
    sc.set_object_transform(6, (0.2, 0.15, 0.3), (1.0, 1.0, 1.0), (25.0, 0.0, 0.0))   # lift and turn the short box
    restirpt.check(device.ctx, device.lib.rpt_scene_update_instances(moved.scene, sc.desc.instances, sc.desc.numInstances),
                   "rpt_scene_update_instances")
    restirpt.check(device.ctx, device.lib.rpt_scene_end_motion(moved.scene), "rpt_scene_end_motion")   # (a fresh scene has no "previous placement")
    fresh = Backend("cuda", sc, w, h, device)
    cpu = Backend("oracle", sc, w, h)
    cam = sc.camera(w, h)
    o, d = camera_rays(cam, w, h)
    rays = np.zeros((w * h, 8), dtype=np.float32)
    rays[:, 0:3] = o; rays[:, 3] = 1e-4; rays[:, 4:7] = d.reshape(-1, 3); rays[:, 7] = 1e7
    a, b, c = moved.trace_closest(rays), fresh.trace_closest(rays), cpu.trace_closest(rays)
    assert np.array_equal(a, b) and np.array_equal(a["instanceIdx"], c["instanceIdx"]) and np.array_equal(a["triangleIdx"], c["triangleIdx"])
    for bk in (moved, fresh, cpu):
        run_frames(bk, cam, "gris", 2)
    for buf in ("GRIS_PREV", "INDIRECT_OUTPUT", "DEPTH_NORMAL_PREV"):
        assert bitwise_mismatch(moved.read(buf), fresh.read(buf)) == 0, buf
        assert bitwise_mismatch(moved.read(buf), cpu.read(buf)) == 0, buf
    # the geometry range of an instance cannot change
    bad = (restirpt.ObjectInstance * sc.desc.numInstances).from_address(sc.desc.instances)
    tampered = (restirpt.ObjectInstance * sc.desc.numInstances)()
    C.memmove(tampered, bad, C.sizeof(tampered))
    tampered[0].indexCount += 3
    assert device.lib.rpt_scene_update_instances(moved.scene, tampered, sc.desc.numInstances) < 0
    for bk in (moved, fresh, cpu):
        bk.close()


def test_two_level_scene_finds_the_flattened_scenes_triangles(device):
    """the same scene description built both ways: the BLAS / TLAS traversal intersects in object space, the flattened one in
    world space, so t and the barycentrics differ in their last bits — but the triangle a ray hits is the same except where two
    triangles are within rounding of each other (edges, grazing rays)"""
    sc = restirpt.HostScene.field(1, 3, 42, shared=True)
    w, h = 192, 108
    flat = Backend("cuda", sc, w, h, device)
    sc.set_two_level(True)
    two = Backend("cuda", sc, w, h, device)
    st = restirpt.BvhStats()
    device.lib.rpt_scene_bvh_stats(two.scene, st)
    assert st.twoLevel == 1 and st.numMeshes == 3 and st.numInstanceRecords == sc.desc.numInstances + 1   # room, blob, lights
    sf = restirpt.BvhStats()
    device.lib.rpt_scene_bvh_stats(flat.scene, sf)
    assert sf.twoLevel == 0 and st.numTriangles < sf.numTriangles / 4   # one copy of the blob instead of nine
    cam = sc.camera(w, h)
    o, d = camera_rays(cam, w, h)
    rays = np.zeros((w * h, 8), dtype=np.float32)
    rays[:, 0:3] = o; rays[:, 3] = 1e-4; rays[:, 4:7] = d.reshape(-1, 3); rays[:, 7] = 1e7
    rng = np.random.default_rng(99)
    rays = np.concatenate([rays, random_rays(rng, 20000, -3.0, 3.0)])
    a, b = flat.trace_closest(rays), two.trace_closest(rays)
    same = (a["instanceIdx"] == b["instanceIdx"]) & (a["triangleIdx"] == b["triangleIdx"])
    assert same.mean() > 0.999, same.mean()
    hit = same & (a["instanceIdx"] != 0xffffffff)
    assert hit.mean() > 0.5
    assert np.abs(a["bary"][hit] - b["bary"][hit]).max() < 1e-3
    assert (flat.trace_shadow(rays) == two.trace_shadow(rays)).mean() > 0.999
    flat.close(); two.close()


def test_two_level_update_instances_rebuilds_the_tlas_only(device):
    """on a two-level scene rpt_scene_update_instances leaves the BLASes alone (object space) and makes new instance records and
    a new TLAS; the result is the state of a scene created from scratch with those transforms, and matches the oracle"""
    sc = restirpt.HostScene.field(1, 3, 42, shared=True, two_level=True)
    w, h = 96, 54
    moved = Backend("cuda", sc, w, h, device)
    before = restirpt.BvhStats()
    device.lib.rpt_scene_bvh_stats(moved.scene, before)
    sc.set_object_transform(3, (0.4, -0.3, 1.4), (1.3, 1.3, 1.3), (0.0, 40.0, 10.0))
    sc.set_object_transform(7, (-0.9, 0.6, 0.8), (0.7, 0.7, 0.7), (75.0, 0.0, 0.0))
    restirpt.check(device.ctx, device.lib.rpt_scene_update_instances(moved.scene, sc.desc.instances, sc.desc.numInstances),
                   "rpt_scene_update_instances")
    restirpt.check(device.ctx, device.lib.rpt_scene_end_motion(moved.scene), "rpt_scene_end_motion")   # (a fresh scene has no "previous placement")
    after = restirpt.BvhStats()
    device.lib.rpt_scene_bvh_stats(moved.scene, after)
    assert after.numNodes == before.numNodes and after.numTriangles == before.numTriangles and after.tlasBuildMs > 0
    fresh = Backend("cuda", sc, w, h, device)
    cpu = Backend("oracle", sc, w, h)
    cam = sc.camera(w, h)
    o, d = camera_rays(cam, w, h)
    rays = np.zeros((w * h, 8), dtype=np.float32)
    rays[:, 0:3] = o; rays[:, 3] = 1e-4; rays[:, 4:7] = d.reshape(-1, 3); rays[:, 7] = 1e7
    a, b, c = moved.trace_closest(rays), fresh.trace_closest(rays), cpu.trace_closest(rays)
    assert np.array_equal(a, b) and np.array_equal(a["instanceIdx"], c["instanceIdx"]) and np.array_equal(a["triangleIdx"], c["triangleIdx"])
    for bk in (moved, fresh, cpu):
        run_frames(bk, cam, "gris", 2)
    for buf in ("GRIS_PREV", "INDIRECT_OUTPUT", "DEPTH_NORMAL_PREV"):
        assert bitwise_mismatch(moved.read(buf), fresh.read(buf)) == 0, buf
        assert bitwise_mismatch(moved.read(buf), cpu.read(buf)) == 0, buf
    for bk in (moved, fresh, cpu):
        bk.close()


def test_scene_arrays_are_validated_at_the_boundary(device):
    """rpt_scene_create checks index / material / instance / alias-table ranges on the host: RPT_ERR_INVALID with a message,
    not an out-of-bounds device read"""
    import ctypes as C
    from restirpt import P
    sc = restirpt.HostScene.cornell()

    def attempt(mutate):
        d = restirpt.SceneDesc()
        C.memmove(C.byref(d), C.byref(sc.desc), C.sizeof(d))
        keep = mutate(d)
        out = P()
        rc = device.lib.rpt_scene_create(device.ctx, C.byref(d), C.byref(out))
        assert rc < 0 and not out.value, rc
        return device.lib.rpt_last_error(device.ctx).decode(), keep

    def bad_index(d):
        idx = (C.c_uint32 * d.numIndices).from_address(d.indices)
        copy = (C.c_uint32 * d.numIndices)(*idx)
        copy[5] = d.numVertices
        d.indices = C.cast(copy, C.c_void_p).value
        return copy

    def bad_material(d):
        mi = (C.c_int32 * d.numMaterialIndices).from_address(d.materialIndices)
        copy = (C.c_int32 * d.numMaterialIndices)(*mi)
        copy[0] = d.numMaterials
        d.materialIndices = C.cast(copy, C.c_void_p).value
        return copy

    def bad_instance(d):
        n = d.numInstances
        src = (restirpt.ObjectInstance * n).from_address(d.instances)
        copy = (restirpt.ObjectInstance * n)()
        C.memmove(copy, src, C.sizeof(copy))
        copy[1].indexCount = d.numIndices + 3
        d.instances = C.cast(copy, C.c_void_p).value
        return copy

    def bad_table(d):
        n = d.numTriangleLights + 1
        src = (restirpt.LightSampleTableElement * n).from_address(d.lightSampleTable)
        copy = (restirpt.LightSampleTableElement * n)()
        C.memmove(copy, src, C.sizeof(copy))
        copy[1].failId = 0
        d.lightSampleTable = C.cast(copy, C.c_void_p).value
        return copy

    def bad_count(d):
        d.numMaterialIndices -= 1

    for mutate, word in ((bad_index, "index"), (bad_material, "material"), (bad_instance, "instance"), (bad_table, "failId"), (bad_count, "numMaterialIndices")):
        msg, _ = attempt(mutate)
        assert word in msg, (word, msg)


@pytest.mark.parametrize("two_level", [False, True])
def test_per_instance_motion_vectors(device, two_level):
    """After rpt_scene_update_instances the G-buffer's motion image follows every surface point back through its instance's
    previous placement (until rpt_scene_end_motion): bit-exact against the oracle's twin, different from the camera-only motion
    exactly on the moved object, and equal to it again once the motion has ended"""
    import ctypes as C
    sc = restirpt.HostScene.cornell()
    if two_level:
        sc.set_two_level(True)
    w, h = 160, 90
    gpu = Backend("cuda", sc, w, h, device)
    n = sc.desc.numInstances
    before = (restirpt.ObjectInstance * n)()
    C.memmove(before, sc.desc.instances, C.sizeof(before))
    sc.set_object_transform(6, (0.12, -0.1, 0.05), (1.0, 1.0, 1.0), (10.0, 0.0, 0.0))      # the short box slides and turns
    restirpt.check(device.ctx, device.lib.rpt_scene_update_instances(gpu.scene, sc.desc.instances, n), "rpt_scene_update_instances")
    cpu = Backend("oracle", sc, w, h)                                                       # the new placements ...
    cpu.lib.orc_scene_set_prev_instances(cpu.scene, C.cast(before, C.c_void_p), n)          # ... and the old ones
    cam = sc.camera(w, h)
    for b in (gpu, cpu):
        b.clear(); b.set_camera(cam, cam); b.run("gbuffer")
    moving, want = gpu.read("MOTION"), cpu.read("MOTION")
    assert bitwise_mismatch(moving, want) == 0
    ids = gpu.read("ALBEDO_MATID").reshape(h, w, 2)[..., 1] & 0xffff
    depth = gpu.read("DEPTH_NORMAL").reshape(h, w, 4)[..., 0]
    moved_px = (ids == 6) & (depth > 0)
    assert moved_px.sum() > 200
    mv = moving.reshape(h, w, 2)
    assert np.abs(mv[moved_px]).max() > 1.0 / w                     # the box moved by more than a pixel
    assert np.abs(mv[~moved_px]).max() < 1e-3                       # static camera: nothing else moves
    restirpt.check(device.ctx, device.lib.rpt_scene_end_motion(gpu.scene), "rpt_scene_end_motion")
    cpu.lib.orc_scene_set_prev_instances(cpu.scene, None, 0)
    for b in (gpu, cpu):
        b.run("gbuffer")
    still = gpu.read("MOTION")
    assert bitwise_mismatch(still, cpu.read("MOTION")) == 0 and np.abs(still).max() < 1e-3
    gpu.close(); cpu.close()


@pytest.mark.parametrize("scene_name", ["cornell", "room"])
def test_restir_pt_bit_exact_with_the_wavefront_forms_forced(built, monkeypatch, scene_name):
    """The library picks the form of two latency chains by size — the path tracer's tail (in-line kernel below 1.5 M pixels,
    wavefront rounds above) and the spatial pass's replays (in-line list kernel for short lists, replay wavefront from 80 k pairs) —
    so on the small parity films the wavefront forms would never run.  Here they are forced (RPT_WAVEFRONT_TAIL=1,
    RPT_RW_MIN_LIST=0; the switches are read when the context is created) and held to the same bar: every buffer of every pass
    equals the oracle bit for bit over four frames with a dolly."""
    import ctypes as C
    monkeypatch.setenv("RPT_WAVEFRONT_TAIL", "1")
    monkeypatch.setenv("RPT_RW_MIN_LIST", "0")
    dev = restirpt.Device(0)
    sc = SCENES[scene_name]()
    gpu, cpu = Backend("cuda", sc, 96, 54, dev), Backend("oracle", sc, 96, 54)
    try:
        _compare_method((scene_name, sc, gpu, cpu), "gris", 4, moves=DOLLY, gris=GRISSettings(2, 1.0, 1, 1, 20))
        rc = (C.c_uint32 * 16)()
        gpu.lib.rpt_reuse_counters(gpu.frame, rc)
        assert rc[3] + rc[5] == 0 or rc[5] > 0, "replays took the in-line list although the wavefront was forced"
    finally:
        gpu.close(); cpu.close()
        dev.close()
