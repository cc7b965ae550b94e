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
}


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
