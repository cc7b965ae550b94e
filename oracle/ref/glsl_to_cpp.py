#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (oracle/ref/Makefile).  Purely syntactic rewrite of a reference shader into something g++ accepts next to
glsl_compat.h — the arithmetic and the control flow are the reference's own text, read from where it lies:

  library mode   glsl_to_cpp.py <in.glsl> <out.inc>            one file, #include lines dropped (the shim includes in order)
  shader mode    glsl_to_cpp.py --shader <in.comp> <out.inc>   a whole compute shader with everything it #includes
in both modes `#version` / `#extension` lines are dropped and the C preprocessor (gcc -E in C mode, see preprocess()) resolves
#include / #if / #define

and then, in both modes:
  * `out T x` / `inout T x` parameters -> `T& x`, `in T x` -> `T x`;
  * swizzles `.xy .zw .yz .xyz .rgb .yzw` read as values -> member calls `.xy()` ...;
  * a vecN(...) constructor whose arguments draw random numbers -> vecN{...}: C++ evaluates braces left to right, as GLSL
    evaluates every argument list, so sample2f / sample3f / sample4f consume the stream in the reference's order;
  * (shader mode) `layout(...) uniform / buffer` interface blocks -> plain namespace-scope variables (unsized arrays -> bounds-checked glsl_buffer<T>),
    `layout(local_size...) in;` dropped;
  * (shader mode) every `struct` gets a value-initialising default constructor and a member-wise constructor (GLSL's
    `T(a, b, c)`), and its `bool` members become 4-byte `gbool` (the std430 / push-constant size);
  * (shader mode) scalar locals declared without an initialiser start at zero (`float depth;` -> `float depth{};`).
The output only ever exists in a temporary directory that the Makefile removes after compiling."""
import os
import re
import subprocess
import sys

SWIZZLES = "xy|zw|yz|xyz|rgb|yzw"
SCALARS = "float|int|uint|bool"


def braces_for_rng_constructors(text):
    out, i = [], 0
    pat = re.compile(r"\bvec[234]\(")
    while True:
        m = pat.search(text, i)
        if not m:
            out.append(text[i:])
            return "".join(out)
        depth, j = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(text[j], 0)
            j += 1
        body = text[m.end():j - 1]
        if "(rng)" in body:
            out.append(text[i:m.end() - 1] + "{" + braces_for_rng_constructors(body) + "}")
        else:
            out.append(text[i:m.end()] + braces_for_rng_constructors(body) + ")")
        i = j


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def preprocess(path, keep_includes):
    """#version / #extension dropped, then gcc -E in C mode with the shader directory on the include path.  C mode on purpose: no
    __cplusplus (HostDevice.h takes its GLSL branch), and an identifier that is not a macro evaluates to 0 inside #if — the rule of
    the GLSL preprocessor the reference is compiled with (glslang expands undefined identifiers to 0), which decides
    `#define MATERIAL_DIELECTRIC_USE_SCHLICK_APPROX true` / `#if MATERIAL_DIELECTRIC_USE_SCHLICK_APPROX` (material.glsl:22, :41):
    `true` is no macro, the #else branch — the exact Fresnel equations — is what the reference runs."""
    src = open(path).read()
    src = re.sub(r"^\s*#(version|extension)[^\n]*$", "", src, flags=re.M)
    if not keep_includes:
        src = re.sub(r'^\s*#include\s+"[^"]+"\s*$', "", src, flags=re.M)
    res = subprocess.run(["gcc", "-E", "-P", "-x", "c", "-undef", "-nostdinc", "-I", os.path.dirname(os.path.abspath(path)), "-"],
                         input=src, capture_output=True, text=True, check=True)
    return res.stdout


def interface_blocks(text):
    def members(body):
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.fullmatch(r"(\w+)\s+(\w+)\s*(\[\s*\])?", decl)
            assert m, decl
            out.append(("glsl_buffer<%s> %s;" if m.group(3) else "%s %s;") % (m.group(1), m.group(2)))
        return "\n".join(out)
    text = re.sub(r"layout\s*\(\s*local_size[^)]*\)\s*in\s*;", "", text)
    text = re.sub(r"layout\s*\([^)]*\)\s*(?:readonly\s+|writeonly\s+)*(?:uniform|buffer)\s+\w+\s*\{([^}]*)\}\s*;",
                  lambda m: members(m.group(1)), text)
    text = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+(\w+)\s+(\w+)\s*(\[\s*\])?\s*;",
                  lambda m: ("glsl_buffer<%s> %s;" if m.group(3) else "%s %s;") % (m.group(1), m.group(2)), text)
    return text


def struct_constructors(text):
    def one(m):
        name, body = m.group(1), m.group(2)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            mm = re.fullmatch(r"(\w+)\s+(\w+(?:\s*,\s*\w+)*)", decl)
            assert mm, (name, decl)
            ty = "gbool" if mm.group(1) == "bool" else mm.group(1)
            for f in re.split(r"\s*,\s*", mm.group(2)):
                fields.append((ty, f))
        decls = "".join("\t%s %s;\n" % f for f in fields)
        zero = ", ".join("%s()" % f for _, f in fields)
        params = ", ".join("%s %s_" % f for f in fields)
        inits = ", ".join("%s(%s_)" % (f, f) for _, f in fields)
        return "struct %s {\n%s\t%s() : %s {}\n\t%s(%s) : %s {}\n};" % (name, decls, name, zero, name, params, inits)
    return re.sub(r"\bstruct\s+(\w+)\s*\{([^{}]*)\}\s*;", one, text)


def zero_scalar_locals(text):
    def one(m):
        names = re.split(r"\s*,\s*", m.group(3))
        return "%s%s %s;" % (m.group(1), m.group(2), ", ".join(n + "{}" for n in names))
    # only inside function bodies (indented lines): globals and struct members keep their form
    return re.sub(r"^([ \t]+)(%s)\s+(\w+(?:\s*,\s*\w+)*)\s*;" % SCALARS, one, text, flags=re.M)


def main():
    args = sys.argv[1:]
    shader = args[0] == "--shader"
    if shader:
        args = args[1:]
        src = preprocess(args[0], True)
    else:
        src = preprocess(args[0], False)
    src = strip_comments(src)
    if shader:
        src = interface_blocks(src)
        src = struct_constructors(src)
        src = zero_scalar_locals(src)       # (struct members get `{}` too: a default member initialiser, harmless)
    src = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", src)
    src = re.sub(r"\bin\s+(\w+)\s+(\w+)", r"\1 \2", src)
    src = re.sub(r"\.(%s)\b(?!\s*\()" % SWIZZLES, r".\1()", src)
    src = braces_for_rng_constructors(src)
    open(args[1], "w").write(src)


if __name__ == "__main__":
    main()
