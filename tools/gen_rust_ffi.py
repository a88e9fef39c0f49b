#!/usr/bin/env python
"""Regenerates the `extern "C"` block of ffi/muopdb_gpu.rs from include/muopdb_gpu.h (every prototype, same order), so the Rust
binding can never lag behind the header; tests/test_abi.py checks that the committed file matches.
usage: python tools/gen_rust_ffi.py [--check]"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BEGIN, END = "    // ---- BEGIN GENERATED (tools/gen_rust_ffi.py) ----\n", "    // ---- END GENERATED ----\n"

BASE = {"int": "c_int", "void": "c_void", "float": "c_float", "char": "c_char", "uint8_t": "u8", "uint16_t": "u16", "uint32_t": "u32",
        "uint64_t": "u64", "int64_t": "i64", "int32_t": "i32"}
RUST_KEYWORDS = {"type", "ref", "in", "fn", "mod", "use", "box", "move", "match", "loop", "impl", "self", "where"}


def rust_type(ctype):
    c = ctype.strip()
    const = False
    stars = c.count("*")
    c = c.replace("*", " ").strip()
    toks = [t for t in c.split() if t not in ("struct",)]
    if "const" in toks:
        const = True
        toks.remove("const")
    base = " ".join(toks)
    r = BASE.get(base, base)   # mgpu_* opaque structs keep their names
    for _ in range(stars):
        r = ("*const " if const else "*mut ") + r
        const = False if stars > 1 else const
    return r


def prototypes():
    src = open(os.path.join(ROOT, "include", "muopdb_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = []
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(mgpu_[a-z0-9_]+)\s*\(([^;{]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        out.append((ret, name, args))
    return out


def emit():
    lines = []
    for ret, name, args in prototypes():
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                arr = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)\[(\d*)\]$", a)
                if arr:   # array parameter decays to a pointer
                    ctype, pname = arr.group(1) + "*", arr.group(2)
                else:
                    mm = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)$", a)
                    ctype, pname = mm.group(1), mm.group(2)
                if pname in RUST_KEYWORDS:
                    pname += "_"
                params.append(f"{pname}: {rust_type(ctype)}")
        r = "" if ret == "void" else f" -> {rust_type(ret)}"
        decl = f"    pub fn {name}({', '.join(params)}){r};"
        while len(decl) > 124:   # wrap
            cut = decl.rfind(", ", 0, 124)
            lines.append(decl[:cut + 1])
            decl = " " * 8 + decl[cut + 2:]
        lines.append(decl)
    return "\n".join(lines) + "\n"


def main():
    path = os.path.join(ROOT, "ffi", "muopdb_gpu.rs")
    txt = open(path).read()
    a, b = txt.index(BEGIN) + len(BEGIN), txt.index(END)
    new = txt[:a] + emit() + txt[b:]
    if "--check" in sys.argv:
        sys.exit(0 if new == txt else 1)
    open(path, "w").write(new)
    print(f"{len(prototypes())} prototypes bound")


if __name__ == "__main__":
    main()
