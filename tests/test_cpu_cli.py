"""restirpt_render, the headless stand-in of the reference executable (host/main.cpp): argument handling and the loud
failure without a CUDA device — the product has no CPU path."""
import os
import subprocess

import restirpt

BIN = os.path.join(restirpt.REPO_ROOT, "vulkan-restir-pt_b200", "bin", "restirpt_render")


def _run(*args):
    return subprocess.run([BIN, *args], capture_output=True, text=True, timeout=300)


def test_cli_is_built_and_documents_itself(built):
    assert os.path.exists(BIN), "make cli (or __graft_entry__.build()) builds it"
    r = _run("--help")
    assert r.returncode == 0 and "usage: restirpt_render" in r.stderr
    for word in ("--indirect", "restir-pt", "--shift", "hybrid", "--accumulate", "--out"):
        assert word in r.stderr


def test_cli_rejects_bad_arguments(built):
    for args, needle in ((("--direct", "bogus"), "is not one of"), (("--size", "12"), "expected WxH"), (("--frames",), "needs a value"),
                         (("--wat",), "unknown option"), (("a.xml", "b.xml"), "more than one scene")):
        r = _run(*args)
        assert r.returncode == 2 and needle in r.stderr, (args, r.stderr)


def test_cli_fails_loudly_on_a_missing_scene_or_device(built, tmp_path):
    r = _run(str(tmp_path / "nope.xml"))
    assert r.returncode == 1 and "cannot open" in r.stderr
    r = _run("cornell", "--size", "64x36", "--frames", "2")
    if r.returncode == 0:   # a CUDA device is present: it rendered
        assert "frames/s" in r.stdout
    else:
        assert r.returncode == 1 and "no CUDA device" in r.stderr and "no CPU path" in r.stderr
