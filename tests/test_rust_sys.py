"""The Rust `-sys` crate (north_star: "a thin extern "C" FFI crate built by build.rs/nvcc") must be true to the header:
every function of include/ccrs_b200.h with the same arity, the option / summary / backend structs field by field, the
model and status constants, and a build.rs that compiles every translation unit of csrc/Makefile. No Rust toolchain
exists in this image, so this is the check that the (source-only) crate cannot drift. Both files are parsed here
independently of tools/gen_rust_sys.py."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ccrs_b200.h")
CRATE = os.path.join(ROOT, "camera-intrinsic-calibration-rs_b200", "rust", "ccrs-b200-sys")


def _no_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        depth += ch in "(<"
        depth -= ch in ")>"
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    return [p for p in (x.strip() for x in out + [cur]) if p]


def header_functions():
    src = _no_comments(open(HEADER).read())
    src = re.sub(r"typedef\s+struct\s+\w+\s*\{.*?\}\s*\w+\s*;", " ", src, flags=re.S)
    src = re.sub(r"enum\s+\w+\s*\{.*?\}\s*;", " ", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(ccrs_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        params = " ".join(m.group(2).split())
        out[m.group(1)] = 0 if params in ("", "void") else len(_split_top(params))
    return out


def header_structs():
    src = _no_comments(open(HEADER).read())
    out = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        names = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            fp = re.match(r".+?\(\s*\*\s*(\w+)\s*\)\s*\(", decl)
            if fp:
                names.append(fp.group(1))
            else:
                names += [n.strip(" *") for n in re.match(r".*?(\w+(?:\s*,\s*\w+)*)$", decl).group(1).split(",")]
        out[m.group(3)] = names
    return out


def header_constants():
    src = _no_comments(open(HEADER).read())
    out = {}
    for m in re.finditer(r"enum\s+\w+\s*\{(.*?)\}", src, flags=re.S):
        for item in m.group(1).split(","):
            if "=" in item:
                k, v = item.split("=")
                out[k.strip()] = int(v)
    return out


def rust_items():
    src = _no_comments(open(os.path.join(CRATE, "src", "lib.rs")).read())
    funcs = {}
    ext = re.search(r'extern\s+"C"\s*\{(.*)\}', src, flags=re.S).group(1)
    for m in re.finditer(r"pub\s+fn\s+(\w+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", ext, flags=re.S):
        funcs[m.group(1)] = len(_split_top(m.group(2)))
    structs = {}
    for m in re.finditer(r"pub\s+struct\s+(\w+)\s*\{(.*?)\n\}", src, flags=re.S):
        structs[m.group(1)] = re.findall(r"pub\s+(\w+)\s*:", m.group(2))
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"pub\s+const\s+(\w+)\s*:\s*c_int\s*=\s*(-?\d+)\s*;", src)}
    return funcs, structs, consts


def test_every_header_function_is_bound_with_the_same_arity():
    h = header_functions()
    r, _, _ = rust_items()
    assert len(h) >= 55
    assert set(h) == set(r), f"missing in lib.rs: {sorted(set(h) - set(r))}; not in the header: {sorted(set(r) - set(h))}"
    assert {k: (h[k], r[k]) for k in h if h[k] != r[k]} == {}


def test_structs_and_constants_match():
    hs, hc = header_structs(), header_constants()
    _, rs, rc = rust_items()
    for name in ("ccrs_options", "ccrs_summary", "ccrs_backend"):
        assert hs[name] == rs[name], name
    for k, v in hc.items():
        assert rc.get(k) == v, k


def test_python_abi_lists_every_header_function(pkg):
    assert set(pkg.SYMBOLS) == set(header_functions())


def test_build_rs_compiles_every_translation_unit():
    mk = open(os.path.join(ROOT, "camera-intrinsic-calibration-rs_b200", "csrc", "Makefile")).read()
    objs = re.search(r"^OBJ\s*:=\s*(.*)$", mk, flags=re.M).group(1).split()
    units = {os.path.basename(o)[:-2] for o in objs}                      # build/x.o -> x
    build = open(os.path.join(CRATE, "build.rs")).read()
    listed = {os.path.splitext(f)[0] for f in re.findall(r'"(ccrs_\w+\.(?:cu|cpp))"', build)}
    assert units == listed, (units, listed)
    on_disk = {os.path.splitext(f)[0] for f in os.listdir(os.path.dirname(os.path.join(ROOT, "camera-intrinsic-calibration-rs_b200", "csrc", "x")))
               if f.endswith((".cu", ".cpp"))}
    assert on_disk == listed


def test_generated_binding_is_current():
    import subprocess, sys
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"], check=True)
