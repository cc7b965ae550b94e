#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (oracle/ref/Makefile).  Purely syntactic rewrite of a reference .glsl file into something g++ accepts next to
glsl_compat.h — the arithmetic is the reference's own text:
  * `#include "..."` lines dropped (the shim includes the files in order);
  * `out T x` / `inout T x` parameters -> `T& x`, `in T x` -> `T x`;
  * swizzles `.xy` / `.zw` read as values -> `.xy()` / `.zw()`;
  * a vecN(...) constructor whose arguments draw random numbers -> vecN{...}: C++ evaluates braces left to right, as GLSL evaluates
    every argument list, so sample2f / sample3f / sample4f consume the stream in the reference's order;
  * functions that pass a swizzle as an `out` argument (`dummy.yz`: two convenience overloads no live shader path calls) are dropped.
Usage: glsl_to_cpp.py <in.glsl> <out.inc>   — the output goes to a temporary directory that the Makefile removes after compiling."""
import re
import sys


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


def drop_functions_with(text, needle):
    """removes every top-level function definition whose body contains `needle`"""
    out, i = [], 0
    pat = re.compile(r"^[A-Za-z_][\w ]*\s+\w+\s*\([^;{}]*\)\s*\{", re.M)
    while True:
        m = pat.search(text, i)
        if not m:
            out.append(text[i:])
            return "".join(out)
        depth, j = 1, m.end()
        while depth:
            depth += {"{": 1, "}": -1}.get(text[j], 0)
            j += 1
        out.append(text[i:m.start()])
        if needle not in text[m.start():j]:
            out.append(text[m.start():j])
        i = j


def main():
    src = open(sys.argv[1]).read()
    src = re.sub(r'^\s*#include\s+"[^"]+"\s*$', "", src, flags=re.M)
    src = drop_functions_with(src, "dummy.")
    src = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", src)
    src = re.sub(r"\bin\s+(\w+)\s+(\w+)", r"\1 \2", src)
    src = re.sub(r"\.(xy|zw)\b(?!\s*\()", r".\1()", src)
    src = braces_for_rng_constructors(src)
    open(sys.argv[2], "w").write(src)


if __name__ == "__main__":
    main()
