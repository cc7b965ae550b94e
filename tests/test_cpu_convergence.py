"""Convergence of the estimators (BASELINE.json north_star: "converged means must match ... within stated noise bounds").

The reference's shaders cannot run here, so the converged-mean check is made between independent code
paths of the restated pipeline: the accumulated image of the ReSTIR PT passes (gris_path_trace / temporal / spatial)
against the accumulated image of the reference's plain unbiased path tracer (gi_naive.comp, NEE + MIS) on the same
scene and camera.  It runs on the CPU oracle; the CUDA library produces the same bits (tests/test_gpu_parity.py), so
every number below holds for it as well.

Measured at 1536 frames of 64x36 (tools: the same loop as below), ratio of the image means to the plain path tracer:
    ReSTIR PT without reuse   0.997 (cornell)   1.006 (room)   0.992 (field)     -> the same expectation
    hybrid shift, temporal + spatial reuse, M cap 20
                              0.977             0.977          0.971             -> 2-3 % darker
with a half-to-half spread of the accumulations of 0.5-3 %.  The small deficit of the reuse passes is the reference's
own estimator — it merges reservoirs with 1/M weights, the pairwise-MIS variant being an unfinished item of its README
 — and is pinned here as a band, not corrected.
Bounds at the 640 frames used below (fixed seeds, so the runs are reproducible): mean within +-4 % without reuse and
within [-7 %, +2 %] with reuse; 8x6-block means within 8 % / 20 % RMS of the mean (measured: 1-3 % / 4-11 %)."""
import numpy as np
import pytest

import restirpt
from restirpt import GRISSettings
from common import Backend, FrameDriver, METHOD_PASSES

W, H, FRAMES = 64, 36, 640

SCENES = {
    "cornell": lambda: restirpt.HostScene.cornell(),
    "field": lambda: restirpt.HostScene.field(1, 3, 42),
}


def _accumulate(sc, method, gris=None):
    b = Backend("oracle", sc, W, H)
    drv = FrameDriver(sc.camera(W, H), accumulate=True)
    for _ in range(FRAMES):
        cur, prev = drv.begin_frame()
        b.set_camera(cur, prev)
        b.run("gbuffer")
        for name, skey in METHOD_PASSES[method]:
            b.run(name, gris if skey else None)
        b.flip()
    img = b.read("INDIRECT_OUTPUT")[..., :3].astype(np.float64)
    b.close()
    return img


def _blocks(img):
    return img.reshape(H // 6, 6, W // 8, 8, 3).mean(axis=(1, 3))


@pytest.fixture(scope="module", params=list(SCENES))
def converged(request, built):
    sc = SCENES[request.param]()
    return (request.param, _accumulate(sc, "naive"), _accumulate(sc, "gris", GRISSettings(2, 1.0, 0, 0, 20)),
            _accumulate(sc, "gris", GRISSettings(2, 1.0, 1, 1, 20)))


def test_restir_pt_without_reuse_converges_to_the_plain_path_tracer(converged):
    name, ref, plain, _ = converged
    assert np.isfinite(plain).all() and ref.mean() > 1e-3
    ratio = plain.mean() / ref.mean()
    assert abs(ratio - 1.0) < 0.04, f"{name}: mean ratio {ratio:.4f}"
    rb, pb = _blocks(ref), _blocks(plain)
    lit = rb.sum(axis=-1) > 0.05 * rb.sum(axis=-1).mean()
    rel = np.sqrt(((pb[lit] - rb[lit]) ** 2).mean()) / rb[lit].mean()
    assert rel < 0.08, f"{name}: block RMS difference {rel:.3f} of the mean"


def test_restir_pt_with_reuse_stays_in_the_documented_band(converged):
    name, ref, plain, reuse = converged
    assert np.isfinite(reuse).all()
    ratio = reuse.mean() / ref.mean()
    assert 0.93 < ratio < 1.02, f"{name}: mean ratio {ratio:.4f}"
    rb, ub = _blocks(ref), _blocks(reuse)
    lit = rb.sum(axis=-1) > 0.05 * rb.sum(axis=-1).mean()
    rel = np.sqrt(((ub[lit] - rb[lit]) ** 2).mean()) / rb[lit].mean()
    assert rel < 0.20, f"{name}: block RMS difference {rel:.3f} of the mean"
    # reuse is not a no-op: the accumulated images differ from the no-reuse run pixel by pixel
    assert np.abs(reuse - plain).max() > 0
