"""The committed golden fixture (tests/golden/golden_v1.json, minted by tests/golden/make_golden.py) against the CPU
oracle: pins the oracle against drift.  The same fixture is checked against the CUDA path in test_gpu_golden.py."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden
from oracle import binding

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.json")))


def test_rng_known_answers_match_numpy_and_oracle(built):
    assert make_golden.rng_kat() == GOLDEN["rng"]
    orc = binding.oracle_lib()
    for k, v in GOLDEN["rng"]["hash2"].items():
        assert orc.orc_hash2(int(k)) == v
    for k, v in GOLDEN["rng"]["makeSeed"].items():
        seed, x, y = (int(t) for t in k.split(","))
        assert orc.orc_make_seed(seed, x, y) == v
    for k, bits in GOLDEN["rng"]["stream"].items():
        state = C.c_uint32(int(k))
        got = [int(np.float32(orc.orc_sample1f(C.byref(state))).view(np.uint32)) for _ in bits]
        assert got == bits


@pytest.mark.parametrize("case", sorted(GOLDEN["cases"]))
def test_oracle_reproduces_golden_buffers(built, case):
    g = GOLDEN["cases"][case]
    kw = {k: tuple(v) for k, v in g["settings"].items()}
    digests, means = make_golden.run_case(make_golden.oracle_factory, g["method"], kw)
    assert digests == g["sha256"]
    for k, v in g["mean"].items():
        assert means[k] == pytest.approx(v, rel=1e-12)
        assert v > 0.0


def test_bench_reference_arm_prints_the_contract_line(built):
    """`bench.py --impl reference` (the CPU oracle on a bounded sample of the bench workload) runs without a GPU and prints
    one JSON line with the keys the driver reads"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_bench_cpu_baseline_leg_and_its_fallback(built, monkeypatch):
    """the main arm's `cpu_baseline` comes from the reference arm run as a child process; if the child cannot run, from the
    same loop in-process — either way a positive frames/s figure with its sample description"""
    import subprocess
    import bench
    scene, _ = bench.load_scene()
    a = bench.cpu_baseline_leg(scene)
    assert a["value"] > 0 and a["kind"] == "port" and a["cores"] >= 1 and "process of its own" in a["sample"]

    def boom(*args, **kwargs):
        raise OSError("no child processes today")
    monkeypatch.setattr(subprocess, "run", boom)
    b = bench.cpu_baseline_leg(scene)
    assert b["value"] > 0 and b["kind"] == "port" and "process of its own" not in b["sample"]
    assert 0.2 < a["value"] / b["value"] < 5.0
