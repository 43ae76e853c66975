#!/usr/bin/env python
"""ORACLE TOOLING (test infrastructure, NOT product code).

Turns a reference shader (read from the reference tree at build time, never copied into the repo) into a C++ struct
whose members are the shader's globals and functions, to be compiled against glsl_compat.hpp:

    python transpile.py <reference>/glsl <out_dir> pre_morph.fs pre_depth.fs ...

Writes <out_dir>/<stem>.inc (git-ignored, under oracle/_ref/). The shader text is kept verbatim except for what C++
cannot parse:
  * `#version` / `#extension` lines are dropped; `#include </name>` is replaced by the named file's text, each file
    once per shader (the names are the ones NetKinectArray.cpp:90,208-209 registers with globjects::NamedString);
  * `layout(...)`, `noperspective`, `flat`, `highp`, and the storage qualifiers `uniform` / `in` / `out` / `buffer` of global
    declarations are removed (globals become struct members); interface blocks lose their braces (their members are
    global names in GLSL too); parameter qualifiers `in` / `const in` are removed;
  * GLSL array declarators `T[n] name` become `T name[n]`, an unsized SSBO array `T[] name` becomes `ssbo_array<T> name`
    (bounds-checked: the shaders index it out of range at the brick-grid border), a geometry shader's `in T name[]` becomes
    `T name[3]`;
  * swizzles `.xy .rg .xyz .rgb .xx .yz` become calls (`.xy()` ...), the only multi-component swizzles these shaders
    use, always as r-values;
  * global `const int` become `static constexpr int` (they size arrays);
  * function prototypes are dropped (class members need none), `out` / `inout` parameters become references,
    `T[5]` array values become `arr5<T>`, `discard;` sets a flag and returns.
Everything else - every expression, constant, loop and branch - is compiled as the reference wrote it.
"""
import os
import re
import sys

NAMED = {"/bricks.glsl": "inc_bricks.glsl", "/inc_bbox_test.glsl": "inc_bbox_test.glsl", "/inc_color.glsl": "inc_color.glsl",
         "/shading.glsl": "shading.glsl"}          # reconstruction.cpp:24 registers glsl/shading.glsl under this name


def load(glsl_dir, name, seen):
    """Shader text with includes expanded (each file at most once) and include guards removed."""
    text = open(os.path.join(glsl_dir, name)).read()
    lines = text.split("\n")
    # an include guard: first directive `#ifndef X` directly followed by `#define X`, closed by the file's last `#endif`
    idx = [i for i, l in enumerate(lines) if l.strip().startswith("#")]
    ends = [i for i, l in enumerate(lines) if l.strip().startswith("#endif")]
    if len(idx) >= 3 and ends:
        a, b = lines[idx[0]].split(), lines[idx[1]].split()
        last = max(ends)
        if a[0] == "#ifndef" and b[0] == "#define" and len(a) > 1 and len(b) > 1 and a[1] == b[1]:
            lines[idx[0]] = lines[idx[1]] = ""
            lines[last] = lines[last].replace("#endif", "", 1)
    out = []
    for l in lines:
        m = re.match(r"\s*#include\s*<([^>]+)>", l)
        if m:
            inc = NAMED[m.group(1)]
            if inc not in seen:
                seen.add(inc)
                out.append(f"// ---- #include <{m.group(1)}> -> glsl/{inc}")
                out.extend(load(glsl_dir, inc, seen))
                out.append(f"// ---- end of {inc}")
            continue
        out.append(l)
    return out


def transpile(lines):
    out = []
    depth = 0                 # brace depth in the emitted C++
    in_block = False          # inside a GLSL interface block (uniform X { ... }; / buffer X { ... };)
    for l in lines:
        s = l.strip()
        if s.startswith("#version") or s.startswith("#extension"):
            continue
        code = l
        code = re.sub(r"layout\s*\([^)]*\)\s*", "", code)
        code = re.sub(r"\bhighp\s+", "", code)                   # precision qualifiers have no effect on desktop GL
        code = re.sub(r"^(\s*)flat\s+", r"\1", code)            # interpolation qualifier of a global declaration
        if re.match(r"^\s*(in|out)\s*;\s*$", code):             # geometry shader: layout(triangles) in; / layout(...) out;
            out.append("// " + l.strip())
            continue
        if depth == 0 and not in_block:
            if re.match(r"\s*(uniform|buffer)\s+\w+\s*\{\s*$", code):
                in_block = True
                out.append("// " + l.strip())
                continue
            code = re.sub(r"^(\s*)noperspective\s+", r"\1", code)
            code = re.sub(r"^(\s*)(uniform|in|out)\s+", r"\1", code)
        elif in_block:
            if re.match(r"\s*\}\s*;\s*$", code):
                in_block = False
                out.append("// " + l.strip())
                continue
        # integer constants that size arrays must be constant expressions in C++ too
        if depth == 0:
            code = re.sub(r"^(\s*)const\s+int\s+(\w+\s*=)", r"\1static constexpr int \2", code)
        # function prototypes (GLSL needs them for forward references; members of a C++ class do not, and may not repeat)
        if depth == 0 and re.match(r"^\s*[\w\[\]]+\s+\w+\s*\([^()=]*\)\s*;\s*(//.*)?$", code):
            out.append("// " + l.strip())
            continue
        code = re.sub(r"\bdiscard\s*;", "{ discarded = true; return; }", code)
        # parameter qualifiers
        code = re.sub(r"([(,]\s*)const\s+in\s+", r"\1const ", code)
        code = re.sub(r"([(,]\s*)in\s+", r"\1", code)
        code = re.sub(r"([(,]\s*)(?:out|inout)\s+(\w+)\s+", r"\1\2& ", code)
        # first-class arrays of 5: T[5] f(...) / T[5] name = ... / T name[5] = T[5](...)
        code = re.sub(r"\b(\w+)\s+(\w+)\[5\]\s*=", r"arr5<\1> \2 =", code)
        code = re.sub(r"\b(\w+)\[5\]\s+(\w+)\s*(=|\()", r"arr5<\1> \2 \3", code)
        code = re.sub(r"\b(\w+)\[5\]\s*\(", r"arr5<\1>(", code)
        # array declarators: T[n] name; -> T name[n];   T[] name; -> T* name;
        code = re.sub(r"\b(\w+)\[(\d+)\]\s+(\w+)\s*;", r"\1 \3[\2];", code)
        code = re.sub(r"\b(\w+)\[\]\s+(\w+)\s*;", r"ssbo_array<\1> \2;", code)
        code = re.sub(r"\b(\w+)\s+(\w+)\[\]\s*;", r"\1 \2[3];", code)      # geometry-shader inputs: one entry per primitive vertex (<= 3)
        # r-value swizzles
        code = re.sub(r"\.(xyz|rgb|xy|rg|xx|yz)\b(?!\s*\()", r".\1()", code)
        if not in_block:
            depth += code.count("{") - code.count("}")
        out.append(code)
    return out


def main():
    glsl_dir, out_dir = sys.argv[1], sys.argv[2]
    os.makedirs(out_dir, exist_ok=True)
    for name in sys.argv[3:]:
        name, _, stem = name.partition(":")                       # "bricks.vs:bricks_vs" names the output; default = the file's stem
        stem = stem or os.path.splitext(name)[0]
        body = transpile(load(glsl_dir, name, set()))
        with open(os.path.join(out_dir, stem + ".inc"), "w") as f:
            f.write(f"// GENERATED from the reference's glsl/{name} by oracle/glsl_host/transpile.py - do not commit\n")
            f.write("\n".join(body) + "\n")


if __name__ == "__main__":
    main()
