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
