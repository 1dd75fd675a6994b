#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY.  Build-time syntax bridge for the parity oracle.

Reads a reference kernel file where it lies (``/root/reference/AssignNN-*/code.cl``) and
writes a g++-compilable copy into ``oracle/_ref/`` (git-ignored; reference text is never
committed to this repo).  Exactly ONE mechanical, arithmetic-neutral rewrite is applied,
because C++ parses OpenCL's vector literal ``(float3)(a,b,c)`` as a cast of a comma
expression:

    (floatN|int3|uint2|uchar4)(args...)   ->   mk_<type>(args...)

Everything else (types, swizzles, builtins, address-space qualifiers) is supplied by
``oracle/clshim.h``.  No arithmetic line is altered.

``--hooks`` additionally inserts instrumentation macro calls at NON-arithmetic places of
the grid-walk kernels so the oracle can report per-ray cells visited / primitive tests /
winning reference index.  The macros expand to nothing unless the driver
is built with -DREF_INSTRUMENT; tests assert both builds produce identical buffers.
"""
import re
import sys

BOM = chr(0xFEFF)
VEC_LITERAL = re.compile(r"\((float2|float3|float4|float16|int3|uint2|uchar4)\)\s*\(")

HOOKS = [
    (re.compile(r"(while\(true\)\{)"), r"\1 REF_HOOK_CELL"),
    (re.compile(r"(inter = interTriangle\(ray, ?tri\);)"), r"\1 REF_HOOK_TEST"),
    (re.compile(r"(inter = interSphere\(ray, ?s\);)"), r"\1 REF_HOOK_TEST"),
    (re.compile(r"(if\(champ_i < UINT_MAX\)\{)"), r"\1 REF_HOOK_HIT(champ_i)"),
    (re.compile(r"(if\(champ_slab\.[xz] < n_slabs\)\{)"), r"\1 REF_HOOK_HIT(champ_i)"),
]


def convert(text, hooks=False):
    if text.startswith(BOM):
        text = text[1:]
    out, n = VEC_LITERAL.subn(lambda m: "mk_%s(" % m.group(1), text)
    if hooks:
        for rx, rep in HOOKS:
            out = rx.sub(rep, out)
    return out, n


def main(argv):
    hooks = "--hooks" in argv
    argv = [a for a in argv if a != "--hooks"]
    src, dst = argv[1], argv[2]
    with open(src, "r", encoding="utf-8-sig") as f:
        text = f.read()
    out, n = convert(text, hooks)
    with open(dst, "w", encoding="utf-8") as f:
        f.write("// GENERATED from %s by oracle/cl2cpp.py (%d vector literals rewritten%s); do not commit.\n"
                % (src, n, ", hooks" if hooks else ""))
        f.write("#line 1 \"%s\"\n" % src)
        f.write(out)
    print("cl2cpp: %s -> %s (%d literals)" % (src, dst, n))


if __name__ == "__main__":
    main(sys.argv)
