#!/usr/bin/env python3
"""Extract the reference's one shipped scene (res/model/VeachAjar.zip) into assets/_ref/ (git-ignored, but it
travels to the GPU box with gpurun) and write binary-PPM sidecars for its JPEG/PNG textures, decoded by the REFERENCE'S OWN
decoder: stb_image.h as vendored in the reference tree, compiled where it lies into oracle/_ref/libref.so (oracle/ref/Makefile;
the reference loads every texture with stbi_load(..., 4), zvk/core/HostImage.cpp:70-75).  The C++ host decodes PNG and JPEG
itself (host/Image.cpp, checked against stb in tests/test_cpu_ref_pins.py) but prefers a sidecar when there is one, so the
measured workloads shade the reference decoder's texels.  Data only — no reference source is copied.

The archive uses zip method 95 (XZ), which python's zipfile refuses, so members are decoded by hand.
Runs only where /root/reference exists (the build container); on the GPU box the prepared files are used."""
import lzma
import os
import struct
import sys
import zipfile
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_ZIP = "/root/reference/res/model/VeachAjar.zip"
OUT = os.path.join(ROOT, "assets", "_ref")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref.so")
USED_SUFFIXES = (".xml", ".obj", ".png", ".jpg", ".jpeg")


def extract(zip_path=DEFAULT_ZIP, out_dir=OUT):
    if not os.path.exists(zip_path):
        return None
    scene_xml = os.path.join(out_dir, "VeachAjar", "ajar.xml")
    marker = os.path.join(out_dir, "VeachAjar", ".prepared")
    if os.path.exists(marker):
        if os.path.exists(REF_LIB) and not open(marker).read().startswith("stb_image"):   # side-cars from an earlier PIL run
            open(marker, "w").write(write_sidecars(os.path.join(out_dir, "VeachAjar", "textures")) + "\n")
        return scene_xml
    zf = zipfile.ZipFile(zip_path)
    with open(zip_path, "rb") as f:
        for zi in zf.infolist():
            if zi.is_dir() or not zi.filename.lower().endswith(USED_SUFFIXES):
                continue
            if "teapot" in zi.filename:   # not referenced by ajar.xml
                continue
            f.seek(zi.header_offset)
            hdr = f.read(30)
            nlen, elen = struct.unpack("<HH", hdr[26:30])
            f.seek(zi.header_offset + 30 + nlen + elen)
            data = f.read(zi.compress_size)
            if zi.compress_type == 0:
                raw = data
            elif zi.compress_type == 8:
                raw = zlib.decompress(data, -15)
            elif zi.compress_type == 95:
                raw = lzma.decompress(data)
            else:
                raise RuntimeError(f"{zi.filename}: unsupported zip method {zi.compress_type}")
            dst = os.path.join(out_dir, zi.filename)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            with open(dst, "wb") as o:
                o.write(raw)
    decoder = write_sidecars(os.path.join(out_dir, "VeachAjar", "textures"))
    open(marker, "w").write(decoder + "\n")
    return scene_xml


def _stb_decode(lib, path):
    import ctypes as C
    w, h = C.c_int(), C.c_int()
    lib.ref_stbi_load_rgba8.restype = C.c_void_p
    lib.ref_stbi_load_rgba8.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ref_stbi_free.argtypes = [C.c_void_p]
    p = lib.ref_stbi_load_rgba8(os.fsencode(path), C.byref(w), C.byref(h))
    if not p:
        raise RuntimeError(f"stb_image could not decode {path}")
    try:
        rgba = C.string_at(p, w.value * h.value * 4)
    finally:
        lib.ref_stbi_free(p)
    rgb = bytearray(w.value * h.value * 3)
    rgb[0::3], rgb[1::3], rgb[2::3] = rgba[0::4], rgba[1::4], rgba[2::4]
    return w.value, h.value, bytes(rgb)


def write_sidecars(tex_dir):
    """<texture>.ppm next to every PNG / JPEG, decoded by stb_image through oracle/_ref/libref.so; PIL only when that library
    has not been built (then the marker says so and tests/test_cpu_ref_pins.py fails the side-car pin)."""
    lib = None
    if os.path.exists(REF_LIB):
        import ctypes as C
        lib = C.CDLL(REF_LIB)
    for name in sorted(os.listdir(tex_dir)):
        if not name.lower().endswith((".png", ".jpg", ".jpeg")):
            continue
        src = os.path.join(tex_dir, name)
        if lib is not None:
            w, h, rgb = _stb_decode(lib, src)
        else:
            from PIL import Image
            img = Image.open(src).convert("RGB")
            (w, h), rgb = img.size, img.tobytes()
        with open(src + ".ppm", "wb") as o:
            o.write(b"P6\n%d %d\n255\n" % (w, h))
            o.write(rgb)
    return "stb_image (oracle/_ref/libref.so)" if lib is not None else "PIL"


def ajar_xml():
    """Path of the prepared scene, or None when the asset is unavailable."""
    p = os.path.join(OUT, "VeachAjar", "ajar.xml")
    return p if os.path.exists(os.path.join(OUT, "VeachAjar", ".prepared")) else None


if __name__ == "__main__":
    print(extract(*(sys.argv[1:2] or [DEFAULT_ZIP])))
