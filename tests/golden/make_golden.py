#!/usr/bin/env python3
"""Mint the golden fixture tests/golden/golden_v1.json.

The reference ships no test vectors and cannot run here, so these goldens are produced by the CPU
oracle (oracle/, the restatement of the reference shaders) and by independent numpy restatements of the integer RNG.
They pin (a) the oracle against drift and (b) the CUDA path against a fixed, committed answer:
  * RNG known answers (hash2 / makeSeed / sample1f bit patterns), computed in numpy only;
  * per-pass buffer digests (SHA-256 of the raw bytes) of 3-frame sequences with camera motion on the procedural
    Cornell box (BASELINE.json config 1, 64x36) for every method, plus mean radiance values.
Run:  python tests/golden/make_golden.py   (rewrites the JSON; commit the result)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vulkan-restir-pt_b200"), os.path.join(ROOT, "tests")]
import numpy as np

W, H, FRAMES = 64, 36, 3
MOVES = [(0.0, 0.0, 0.0), (0.02, 0.01, 0.0), (0.0, -0.01, 0.01)]
CASES = {
    # name: (method, settings kwargs)
    "naive": ("naive", {}),
    "di_reconnection_light": ("di", {"di": (0, 0, 1, 1)}),
    "di_replay_both": ("di", {"di": (1, 2, 1, 1)}),
    "gi": ("gi", {}),
    "gris_hybrid": ("gris", {"gris": (2, 1.0, 1, 1, 20)}),
    "gris_reconnection": ("gris", {"gris": (0, 1.0, 1, 1, 20)}),
}
BUFFERS = {
    "naive": ["DEPTH_NORMAL", "ALBEDO_MATID", "MOTION", "PRIMARY_ISEC", "DIRECT_OUTPUT", "INDIRECT_OUTPUT"],
    "di": ["DI_THIS", "DI_TEMP", "DIRECT_OUTPUT"],
    "gi": ["GI_THIS", "INDIRECT_OUTPUT"],
    "gris": ["GRIS_THIS", "GRIS_TEMP", "INDIRECT_OUTPUT"],
}


def np_hash2(seed):
    s = np.uint32(seed)
    with np.errstate(over="ignore"):
        s = (s ^ np.uint32(61)) ^ (s >> np.uint32(16))
        s = s * np.uint32(9)
        s = s ^ (s >> np.uint32(4))
        s = s * np.uint32(0x27d4eb2d)
        s = s ^ (s >> np.uint32(15))
    return np.uint32(s)


def np_make_seed(seed, x, y):
    with np.errstate(over="ignore"):
        a = np_hash2((np.uint32(seed) + np.uint32(x)) ^ (np.uint32(y) - np.uint32(1)))
        b = np_hash2(np.uint32(y) * (np.uint32(x) - np.uint32(2)))
        return np.uint32(a + b)


def rng_kat():
    out = {"hash2": {}, "makeSeed": {}, "stream": {}}
    for s in (0, 1, 2, 12345, 0xdeadbeef, 0xffffffff):
        out["hash2"][str(s)] = int(np_hash2(s))
    for seed, x, y in ((1, 0, 0), (2, 17, 5), (0x9e3779b9, 1919, 1079), (7, 2, 0)):
        out["makeSeed"][f"{seed},{x},{y}"] = int(np_make_seed(seed, x, y))
    for seed in (1, 0xabcdef01):
        st, bits = np.uint32(seed), []
        for _ in range(16):
            st = np_hash2(st)
            bits.append(int(np.float32(np.float32(st) / np.float32(4294967295.0)).view(np.uint32)))
        out["stream"][str(seed)] = bits
    return out


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(backend_factory, method, kw):
    """-> {"frame/pass/buffer": sha256}, {"buffer": mean} using tests/common.run_frames"""
    import restirpt
    from restirpt import DISettings, GRISSettings
    from common import run_frames
    b, cam = backend_factory()
    di = DISettings(*kw["di"]) if "di" in kw else None
    gris = GRISSettings(*kw["gris"]) if "gris" in kw else None
    digests, means = {}, {}

    def snap(i, pass_name, be):
        last = pass_name == {"naive": "gi_naive", "di": "di_spatial", "gi": "gi_restir", "gris": "gris_spatial"}[method]
        if not last:
            return
        for buf in BUFFERS[method]:
            arr = be.read(buf)
            digests[f"{i}/{buf}"] = digest(arr)
            if buf.endswith("OUTPUT"):
                means[f"{i}/{buf}"] = float(np.asarray(arr, dtype=np.float64)[..., :3].mean())

    run_frames(b, cam, method, FRAMES, di=di, gris=gris, moves=MOVES, snapshot=snap)
    b.close()
    return digests, means


def oracle_factory():
    import restirpt
    from common import Backend
    sc = restirpt.HostScene.cornell()
    return Backend("oracle", sc, W, H), sc.camera(W, H)


def main():
    golden = {"version": 1, "film": [W, H], "frames": FRAMES, "moves": MOVES, "rng": rng_kat(), "cases": {}}
    for name, (method, kw) in CASES.items():
        d, m = run_case(oracle_factory, method, kw)
        golden["cases"][name] = {"method": method, "settings": {k: list(v) for k, v in kw.items()}, "sha256": d, "mean": m}
        print(name, len(d), "digests", m)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.json")
    with open(path, "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
