"""The CUDA path (through the C ABI) against the committed golden fixture: every buffer of every method must hash to
the committed SHA-256 — bit-exact, no oracle involved at run time."""
import json
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden
import restirpt
from common import Backend

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.json")))


@pytest.mark.parametrize("case", sorted(GOLDEN["cases"]))
def test_cuda_reproduces_golden_buffers(built, case):
    g = GOLDEN["cases"][case]
    kw = {k: tuple(v) for k, v in g["settings"].items()}
    dev = restirpt.Device(0)

    def factory():
        sc = restirpt.HostScene.cornell()
        return Backend("cuda", sc, make_golden.W, make_golden.H, dev), sc.camera(make_golden.W, make_golden.H)

    digests, means = make_golden.run_case(factory, g["method"], kw)
    bad = sorted(k for k in g["sha256"] if digests.get(k) != g["sha256"][k])
    assert not bad, f"{case}: buffers differ from the golden fixture: {bad}"
