"""Texture decoding of the host library (host/Image.cpp: PNG and JPEG written from the format specifications) against
PIL's decoders — the scene front-end's stand-in for the reference's stb_image call (src/Resource.cpp:26).

PNG is lossless: exact.  JPEG decoders are allowed to differ in the last bit of the inverse DCT and in how they
interpolate subsampled chroma, so the comparison is toleranced: mean absolute difference and the fraction of samples
that differ by more than a few levels (bounds written in the tests)."""
import io
import os
import struct
import sys
import zlib

import numpy as np
import pytest

import restirpt

PIL = pytest.importorskip("PIL.Image")
sys.path.insert(0, os.path.join(restirpt.REPO_ROOT, "tools"))
import prepare_assets


def _test_image(w, h, seed=0):
    """smooth gradients + a few hard edges + mild noise: exercises DC, low and high AC frequencies"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127 + 120 * np.sin(x / 9.0) * np.cos(y / 13.0), 255.0 * x / max(w - 1, 1), 255.0 * ((x // 8 + y // 5) % 2)], axis=-1)
    img += rng.normal(scale=6.0, size=img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def _decode(tmp_path, name, data):
    p = tmp_path / name
    p.write_bytes(data)
    return restirpt.read_image(str(p))


def _jpeg_close(mine, ref, mean_tol, outlier_tol):
    assert mine.shape[:2] == ref.shape[:2]
    assert (mine[..., 3] == 255).all()
    d = np.abs(mine[..., :3].astype(np.int32) - ref.astype(np.int32))
    assert d.mean() < mean_tol, f"mean |diff| {d.mean():.3f}"
    assert (d > 4).mean() < outlier_tol, f"{(d > 4).mean():.5f} of the samples differ by more than 4 levels (max {d.max()})"


@pytest.mark.parametrize("mode", ["L", "LA", "RGB", "RGBA", "P", "1", "I;16"])
def test_png_modes_exact(tmp_path, built, mode):
    rgb = _test_image(53, 37)
    if mode == "P":
        im = PIL.fromarray(rgb).quantize(37)
    elif mode == "I;16":
        im = PIL.fromarray((rgb[..., 0].astype(np.uint16) * 257 + 13).astype(np.uint16))
    elif mode in ("LA", "RGBA"):
        im = PIL.fromarray(np.dstack([rgb, 255 - rgb[..., :1]])).convert(mode)
    else:
        im = PIL.fromarray(rgb).convert(mode)
    buf = io.BytesIO()
    im.save(buf, format="PNG")
    mine = _decode(tmp_path, f"m_{mode.replace(';', '')}.png", buf.getvalue())
    if mode == "I;16":
        want = np.asarray(im) >> 8
        assert np.array_equal(mine[..., 0], want) and np.array_equal(mine[..., 1], want) and (mine[..., 3] == 255).all()
    else:
        assert np.array_equal(mine, np.asarray(im.convert("RGBA")))


def test_png_low_bit_depth_palette_and_transparency(tmp_path, built):
    idx = (np.arange(19 * 11).reshape(11, 19) % 4).astype(np.uint8)
    im = PIL.fromarray(idx, mode="P")
    im.putpalette([10, 20, 30, 200, 100, 0, 0, 255, 7, 90, 90, 90])
    buf = io.BytesIO()
    im.save(buf, format="PNG", bits=2, transparency=bytes([255, 0, 128, 255]))
    mine = _decode(tmp_path, "p2.png", buf.getvalue())
    assert np.array_equal(mine, np.asarray(PIL.open(io.BytesIO(buf.getvalue())).convert("RGBA")))


def test_png_adam7_interlace(tmp_path, built):
    """PIL cannot write interlaced files: the seven passes are assembled here (filter type 0 rows, one zlib stream)"""
    h, w = 21, 29
    img = np.dstack([_test_image(w, h, 3), np.full((h, w, 1), 200, np.uint8)])
    raw = b""
    for x0, y0, dx, dy in [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]:
        sub = img[y0::dy, x0::dx]
        if sub.size:
            raw += b"".join(b"\0" + row.tobytes() for row in sub)

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 1)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")
    assert np.array_equal(np.asarray(PIL.open(io.BytesIO(png)).convert("RGBA")), img)   # the file is well-formed
    assert np.array_equal(_decode(tmp_path, "adam7.png", png), img)


@pytest.mark.parametrize("kind,options", [
    ("baseline 4:4:4", dict(subsampling=0)),
    ("baseline 4:2:2", dict(subsampling=1)),
    ("baseline 4:2:0", dict(subsampling=2)),
    ("progressive 4:2:0", dict(subsampling=2, progressive=True)),
    ("progressive 4:4:4", dict(subsampling=0, progressive=True)),
    ("optimised tables", dict(subsampling=2, optimize=True)),
    ("grey", dict(grey=True)),
    ("grey progressive", dict(grey=True, progressive=True)),
])
def test_jpeg_against_pil(tmp_path, built, kind, options):
    options = dict(options)
    rgb = _test_image(203, 117, 1)   # not a multiple of the MCU size in either direction
    im = PIL.fromarray(rgb).convert("L") if options.pop("grey", False) else PIL.fromarray(rgb)
    buf = io.BytesIO()
    im.save(buf, format="JPEG", quality=88, **options)
    mine = _decode(tmp_path, "t.jpg", buf.getvalue())
    ref = np.asarray(PIL.open(io.BytesIO(buf.getvalue())).convert("RGB"))
    # inverse DCT rounding: <= 1 level almost everywhere; the image edges of subsampled chroma may differ more
    _jpeg_close(mine, ref, mean_tol=0.35, outlier_tol=0.004)


def test_jpeg_restart_intervals(tmp_path, built):
    rgb = _test_image(160, 96, 2)
    buf = io.BytesIO()
    try:
        PIL.fromarray(rgb).save(buf, format="JPEG", quality=90, subsampling=2, restart_marker_blocks=3)
    except TypeError:
        pytest.skip("this PIL cannot write restart markers")
    data = buf.getvalue()
    if b"\xff\xdd" not in data:
        pytest.skip("this PIL ignored restart_marker_blocks")
    mine = _decode(tmp_path, "r.jpg", data)
    _jpeg_close(mine, np.asarray(PIL.open(io.BytesIO(data)).convert("RGB")), mean_tol=0.35, outlier_tol=0.004)


def test_scene_textures_decode_like_pil(built):
    """the textures of the bench scene (two baseline files, one progressive 3008x2000 file)"""
    xml = prepare_assets.ajar_xml()
    if not xml:
        pytest.skip("VeachAjar asset not prepared")
    tex = os.path.join(os.path.dirname(xml), "textures")
    names = [n for n in sorted(os.listdir(tex)) if n.lower().endswith((".jpg", ".jpeg", ".png"))]
    assert names
    for n in names:
        mine = restirpt.read_image(os.path.join(tex, n))
        ref = np.asarray(PIL.open(os.path.join(tex, n)).convert("RGB"))
        _jpeg_close(mine, ref, mean_tol=0.1, outlier_tol=1e-4)   # measured: mean 0.02-0.03, at most 3 levels


def test_bad_files_fail_cleanly(tmp_path, built):
    for name, data in [("empty.png", b""), ("trunc.png", b"\x89PNG\r\n\x1a\n" + b"\0" * 9), ("trunc.jpg", b"\xff\xd8\xff\xc0\x00\x11\x08"),
                       ("noise.jpg", b"\xff\xd8" + bytes(range(256)) * 4), ("text.txt", b"hello")]:
        p = tmp_path / name
        p.write_bytes(data)
        with pytest.raises(restirpt.RestirptError):
            restirpt.read_image(str(p))
    with pytest.raises(restirpt.RestirptError):
        restirpt.read_image(str(tmp_path / "missing.png"))


def test_scene_loads_png_and_jpeg_textures_without_sidecars(tmp_path, built):
    """map_Kd textures of an OBJ model come straight from the image files (Scene::loadTextureFile -> readImage); a PPM
    side-car next to a file still wins, which keeps the texels of the measured workloads as they were"""
    import ctypes as C
    (tmp_path / "m").mkdir()
    rgb = _test_image(24, 16, 5)
    PIL.fromarray(rgb).save(str(tmp_path / "m" / "a.png"))
    PIL.fromarray(rgb).save(str(tmp_path / "m" / "b.jpg"), quality=95, subsampling=0)
    PIL.fromarray(rgb).save(str(tmp_path / "m" / "c.jpg"), quality=95, subsampling=0)
    (tmp_path / "m" / "c.jpg.ppm").write_bytes(b"P6\n1 1\n255\n" + bytes([9, 8, 7]))
    (tmp_path / "m" / "t.mtl").write_text("newmtl a\nKd 1 1 1\nmap_Kd a.png\nnewmtl b\nKd 1 1 1\nmap_Kd b.jpg\nnewmtl c\nKd 1 1 1\nmap_Kd c.jpg\n")
    (tmp_path / "m" / "t.obj").write_text("mtllib t.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\n"
                                          "usemtl a\nf 1/1/1 2/1/1 3/1/1\nusemtl b\nf 1/1/1 3/1/1 4/1/1\nusemtl c\nf 1/1/1 2/1/1 4/1/1\n")
    (tmp_path / "s.xml").write_text("""<?xml version="1.0"?><scene name="t"><integrator type="path"><size width="32" height="32" /></integrator>
<camera type="thinLens"><position value="0 -3 0" /><lookAt value="0 0 0" /><fov value="40" /></camera><modelInstances>
 <modelInstance path="m/t.obj" name="a" type="object"><transform translate="0 0 0" scale="1 1 1" rotate="0 0 0" /></modelInstance>
</modelInstances></scene>""")
    sc = restirpt.HostScene.xml(str(tmp_path / "s.xml"))
    d = sc.desc
    assert d.numTextures == 3
    tex = C.cast(d.textures, C.POINTER(restirpt.TextureDesc))
    got = {}
    for i in range(3):
        t = tex[i]
        got[(t.width, t.height, i)] = np.ctypeslib.as_array(C.cast(t.rgba8, C.POINTER(C.c_uint8)), (t.height, t.width, 4)).copy()
    sizes = sorted(k[:2] for k in got)
    assert sizes == [(1, 1), (24, 16), (24, 16)]
    full = [v for k, v in got.items() if k[:2] == (24, 16)]
    assert any(np.array_equal(v[..., :3], rgb) for v in full)                                        # the PNG, exact
    assert all(np.abs(v[..., :3].astype(int) - rgb.astype(int)).mean() < 3.0 for v in full)          # the JPEG, lossy but close
    assert [v for k, v in got.items() if k[:2] == (1, 1)][0].tolist() == [[[9, 8, 7, 255]]]           # the side-car


def test_random_files_against_pil(tmp_path, built):
    """seeded sweep over sizes (1..90, so every partial-MCU / one-column case), contents (noise, ramps, flat), JPEG encoder
    settings (quality, 4:4:4 / 4:2:2 / 4:2:0, progressive, optimised tables, restart intervals, grey) and PNG modes:
    PNG exact, JPEG within 4 levels everywhere (measured over 400 files: at most 3)"""
    rng = np.random.default_rng(123)
    for it in range(120):
        w, h = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        kind = rng.integers(0, 3)
        if kind == 0:
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        elif kind == 1:
            y, x = np.mgrid[0:h, 0:w]
            img = np.stack([(x * 3 + y) % 256, (y * 5) % 256, (x * y) % 256], -1).astype(np.uint8)
        else:
            img = np.full((h, w, 3), rng.integers(0, 256, 3), dtype=np.uint8)
        buf = io.BytesIO()
        if rng.integers(0, 2) == 0:
            opts = dict(quality=int(rng.integers(20, 101)), subsampling=int(rng.integers(0, 3)), progressive=bool(rng.integers(0, 2)),
                        optimize=bool(rng.integers(0, 2)))
            if rng.integers(0, 3) == 0:
                opts["restart_marker_blocks"] = int(rng.integers(1, 5))
            im = PIL.fromarray(img)
            if rng.integers(0, 4) == 0:
                im = im.convert("L")
                opts.pop("subsampling")
            try:
                im.save(buf, format="JPEG", **opts)
            except TypeError:
                opts.pop("restart_marker_blocks", None)
                im.save(buf, format="JPEG", **opts)
            mine = _decode(tmp_path, "f.jpg", buf.getvalue())
            ref = np.asarray(PIL.open(io.BytesIO(buf.getvalue())).convert("RGBA"))
            d = np.abs(mine.astype(np.int32) - ref.astype(np.int32))
            assert d.max() <= 4, f"file {it} ({w}x{h}, {opts}): max difference {d.max()}"
        else:
            mode = ["RGB", "RGBA", "L", "LA", "P", "1"][rng.integers(0, 6)]
            im = PIL.fromarray(img).quantize(int(rng.integers(2, 257))) if mode == "P" else PIL.fromarray(img).convert(mode)
            im.save(buf, format="PNG", compress_level=int(rng.integers(0, 10)))
            mine = _decode(tmp_path, "f.png", buf.getvalue())
            assert np.array_equal(mine, np.asarray(PIL.open(io.BytesIO(buf.getvalue())).convert("RGBA"))), f"file {it} ({w}x{h}, {mode})"


def test_ppm_header_larger_than_the_file_is_rejected_quickly(tmp_path, built):
    """a damaged size field must not turn into a multi-gigabyte allocation (found by fuzzing: profiles/r1_21_host_sanitizers.txt)"""
    import time
    p = tmp_path / "big.ppm"
    p.write_bytes(b"P6\n60000 60000\n255\n" + bytes(300))
    t0 = time.perf_counter()
    with pytest.raises(restirpt.RestirptError):
        restirpt.read_image(str(p))
    assert time.perf_counter() - t0 < 1.0
    ok = tmp_path / "ok.ppm"
    ok.write_bytes(b"P6\n# comment\n2 1\n255\n" + bytes([1, 2, 3, 4, 5, 6]))
    assert restirpt.read_image(str(ok)).tolist() == [[[1, 2, 3, 255], [4, 5, 6, 255]]]
