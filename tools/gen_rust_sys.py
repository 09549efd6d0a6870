#!/usr/bin/env python
"""Generate camera-intrinsic-calibration-rs_b200/rust/ccrs-b200-sys/src/lib.rs from include/ccrs_b200.h.

The header is the single source of truth for the C ABI; the Rust `-sys` crate (north_star: "a thin extern "C" FFI
crate built by build.rs/nvcc") is derived from it so that it cannot drift: every function, the option / summary /
backend structs, the model and status constants. tests/test_rust_sys.py parses both files independently and fails on
any difference in name or arity. No Rust toolchain exists in the build image: the crate is source only."""
from __future__ import annotations

import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ccrs_b200.h")
OUT = os.path.join(ROOT, "camera-intrinsic-calibration-rs_b200", "rust", "ccrs-b200-sys", "src", "lib.rs")

SCALARS = {"int": "c_int", "double": "c_double", "float": "c_float", "char": "c_char", "void": "c_void",
           "int32_t": "i32", "int64_t": "i64", "unsigned char": "c_uchar", "ccrs_problem": "ccrs_problem",
           "ccrs_joint": "ccrs_joint", "ccrs_options": "ccrs_options", "ccrs_summary": "ccrs_summary",
           "ccrs_backend": "ccrs_backend"}


def strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def rust_type(ctype: str) -> str:
    """'const double*' -> '*const c_double', 'ccrs_problem**' -> '*mut *mut ccrs_problem', 'int' -> 'c_int'."""
    t = ctype.strip()
    stars = t.count("*")
    t = t.replace("*", " ").strip()
    const = bool(re.search(r"\bconst\b", t))
    t = re.sub(r"\b(const|struct|volatile)\b", " ", t)
    t = " ".join(t.split())
    base = SCALARS[t]
    out = base
    for i in range(stars):
        out = ("*const " if (const and i == 0) else "*mut ") + out
    return out


def split_params(s: str):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        if ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [p.strip() for p in out]


def parse_param(p: str):
    """'const double* x' -> ('x', '*const c_double'); function pointers are handled by the caller."""
    if p == "void":
        return None
    m = re.match(r"(.*?)(\w+)\s*(\[\s*\w*\s*\])?$", p.strip())
    ctype, name, arr = m.group(1), m.group(2), m.group(3)
    if not ctype.strip():           # unnamed parameter
        ctype, name = p, "_arg"
    if arr:
        ctype += "*"
    if name in ("type", "fn", "in", "mod", "ref", "box"):
        name += "_"
    return name, rust_type(ctype)


def parse_header(text: str):
    src = strip_comments(text)
    consts = []
    for m in re.finditer(r"enum\s+\w+\s*\{(.*?)\}", src, flags=re.S):
        for item in m.group(1).split(","):
            if "=" in item:
                k, v = item.split("=")
                consts.append((k.strip(), v.strip()))
    for m in re.finditer(r"#define\s+(CCRS_\w+)\s+(-?\d+)", src):
        consts.append((m.group(1), m.group(2)))
    structs = []
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            fp = re.match(r"(.+?)\(\s*\*\s*(\w+)\s*\)\s*\((.*)\)$", decl)
            if fp:   # function pointer field
                ret, name, params = fp.group(1).strip(), fp.group(2), split_params(fp.group(3))
                ps = [parse_param(p) for p in params]
                sig = ", ".join(f"{n}: {t}" for n, t in ps if n)
                r = "" if ret == "void" else f" -> {rust_type(ret)}"
                fields.append((name, f"Option<unsafe extern \"C\" fn({sig}){r}>"))
            else:
                # 'int n_accepted, n_rejected' -> two fields
                m2 = re.match(r"(.*?)(\w+(?:\s*,\s*\w+)*)$", decl)
                ctype, names = m2.group(1), [n.strip() for n in m2.group(2).split(",")]
                for n in names:
                    fields.append((n, rust_type(ctype)))
        structs.append((m.group(3), fields))
    funcs = []
    body = re.sub(r"typedef\s+struct\s+\w+\s*\{.*?\}\s*\w+\s*;", " ", src, flags=re.S)
    body = re.sub(r"enum\s+\w+\s*\{.*?\}\s*;", " ", body, flags=re.S)
    for m in re.finditer(r"([\w\s\*]+?)\b(ccrs_\w+)\s*\(([^;{}]*?)\)\s*;", body, flags=re.S):
        ret, name, params = " ".join(m.group(1).split()), m.group(2), " ".join(m.group(3).split())
        if ret.startswith("typedef") or not ret:
            continue
        ps = [parse_param(p) for p in split_params(params)] if params.strip() else []
        funcs.append((name, ret, [p for p in ps if p]))
    return consts, structs, funcs


def render(consts, structs, funcs) -> str:
    o = []
    o.append("//! Raw bindings to `include/ccrs_b200.h` — GENERATED by tools/gen_rust_sys.py, do not edit by hand.")
    o.append("//! SOURCE ONLY: there is no Rust toolchain in the build image, so this crate has never been compiled there;")
    o.append("//! tests/test_rust_sys.py checks it against the header (every function, same arity).")
    o.append("#![allow(non_camel_case_types, clippy::too_many_arguments)]")
    o.append("use libc::{c_char, c_double, c_float, c_int, c_uchar, c_void};")
    o.append("")
    for opaque in ("ccrs_problem", "ccrs_joint"):
        o.append("#[repr(C)]")
        o.append(f"pub struct {opaque} {{\n    _private: [u8; 0],\n}}")
        o.append("")
    for k, v in consts:
        o.append(f"pub const {k}: c_int = {v};")
    o.append("")
    for name, fields in structs:
        has_fp = any(t.startswith("Option<") for _, t in fields)
        o.append("#[repr(C)]")
        o.append("#[derive(Clone, Copy)]" if has_fp else "#[derive(Clone, Copy, Debug)]")
        o.append(f"pub struct {name} {{")
        for n, t in fields:
            o.append(f"    pub {n}: {t},")
        o.append("}")
        o.append("")
    o.append('extern "C" {')
    for name, ret, ps in funcs:
        sig = ", ".join(f"{n}: {t}" for n, t in ps)
        r = "" if ret == "void" else f" -> {rust_type(ret)}"
        o.append(f"    pub fn {name}({sig}){r};")
    o.append("}")
    o.append("")
    return "\n".join(o)


def main():
    consts, structs, funcs = parse_header(open(HEADER).read())
    text = render(consts, structs, funcs)
    if "--check" in sys.argv:
        if open(OUT).read() != text:
            print("lib.rs is out of date: run tools/gen_rust_sys.py", file=sys.stderr)
            sys.exit(1)
        return
    open(OUT, "w").write(text)
    print(f"wrote {OUT}: {len(funcs)} functions, {len(structs)} structs, {len(consts)} constants")


if __name__ == "__main__":
    main()
