"""restirpt_render on a GPU: a few ReSTIR PT frames of the built-in Cornell box through the C++ Renderer, screenshot
decoded back with the host library's own PNG reader.  (Named to run after the parity tests.)"""
import os
import subprocess

import numpy as np
import pytest

import restirpt

pytestmark = pytest.mark.gpu
BIN = os.path.join(restirpt.REPO_ROOT, "vulkan-restir-pt_b200", "bin", "restirpt_render")


def test_cli_renders_a_screenshot(built, tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("restirpt_render is not built on this box (tests/test_cpu_cli.py checks the build)")
    out = str(tmp_path / "shot.png")
    r = subprocess.run([BIN, "cornell", "--size", "96x54", "--frames", "6", "--seeds", "hash2", "--direct", "naive", "--out", out],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "frames/s" in r.stdout and "wrote" in r.stdout
    img = restirpt.read_image(out)
    assert img.shape == (54, 96, 4) and (img[..., 3] == 255).all()
    assert img[..., :3].mean() > 2 and np.unique(img[..., :3]).size > 16   # lit, not flat
