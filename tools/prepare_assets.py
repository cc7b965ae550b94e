#!/usr/bin/env python3
"""Extract the reference's one shipped scene (res/model/VeachAjar.zip) into assets/_ref/ (git-ignored, but it
travels to the GPU box with gpurun) and write binary-PPM sidecars for its JPEG/PNG textures.  The C++ host decodes
PNG and JPEG itself (host/Image.cpp) but prefers a sidecar when there is one: the measured workloads and the golden
fixtures were made with these PIL-decoded texels.  Data only — no reference source is copied.

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
USED_SUFFIXES = (".xml", ".obj", ".png", ".jpg", ".jpeg")


def extract(zip_path=DEFAULT_ZIP, out_dir=OUT):
    if not os.path.exists(zip_path):
        return None
    scene_xml = os.path.join(out_dir, "VeachAjar", "ajar.xml")
    marker = os.path.join(out_dir, "VeachAjar", ".prepared")
    if os.path.exists(marker):
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
    from PIL import Image
    tex_dir = os.path.join(out_dir, "VeachAjar", "textures")
    for name in sorted(os.listdir(tex_dir)):
        if name.lower().endswith((".png", ".jpg", ".jpeg")):
            img = Image.open(os.path.join(tex_dir, name)).convert("RGB")
            with open(os.path.join(tex_dir, name + ".ppm"), "wb") as o:
                o.write(b"P6\n%d %d\n255\n" % img.size)
                o.write(img.tobytes())
    open(marker, "w").write("ok\n")
    return scene_xml


def ajar_xml():
    """Path of the prepared scene, or None when the asset is unavailable."""
    p = os.path.join(OUT, "VeachAjar", "ajar.xml")
    return p if os.path.exists(os.path.join(OUT, "VeachAjar", ".prepared")) else None


if __name__ == "__main__":
    print(extract(*(sys.argv[1:2] or [DEFAULT_ZIP])))
