#!/usr/bin/env python3
"""Generates rs-doko-cuda/src/ffi.rs from include/doko_cuda.h: every DK_API function, every struct, every #define / enum constant.

    python rs-doko-cuda/tools/gen_ffi.py            # rewrite src/ffi.rs
    python rs-doko-cuda/tools/gen_ffi.py --check    # exit 1 if src/ffi.rs differs from what the header gives (tests/test_rust_ffi_drift.py)

The binding is generated, never edited by hand, so the Rust side cannot drift from the C ABI."""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
HEADER = os.path.join(ROOT, "include", "doko_cuda.h")
OUT = os.path.join(os.path.dirname(HERE), "src", "ffi.rs")

SCALARS = {"int": "c_int", "size_t": "usize", "uint8_t": "u8", "uint16_t": "u16", "uint32_t": "u32", "uint64_t": "u64", "int8_t": "i8", "int16_t": "i16",
           "int32_t": "i32", "int64_t": "i64", "float": "f32", "double": "f64", "char": "c_char", "void": "c_void", "dk_status": "dk_status",
           "dk_stream": "dk_stream"}
OPAQUE = {"dk_ctx", "dk_selfplay"}
STRUCTS = {"dk_state", "dk_rng", "dk_sp_buffers", "dk_playout_stats", "dk_nccl_id"}


def strip_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def rust_type(ctype):
    """'const dk_state*' -> '*const dk_state', 'dk_ctx**' -> '*mut *mut dk_ctx', 'const uint64_t**' -> '*mut *const u64'."""
    t = ctype.strip()
    stars = t.count("*")
    t = t.replace("*", " ").split()
    const = "const" in t
    base = [w for w in t if w != "const"][0]
    r = SCALARS.get(base, base)
    if base not in SCALARS and base not in OPAQUE and base not in STRUCTS:
        raise SystemExit(f"gen_ffi: unknown C type {ctype!r}")
    for level in range(stars):
        # the innermost pointer carries the const of the pointee; outer levels are out-parameters
        r = ("*const " if (const and level == 0) else "*mut ") + r
    return r


def parse_functions(text):
    fns = []
    for m in re.finditer(r"DK_API\s+([^;(]+?)\s*\b(dk_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.+?)\b(\w+)$", a)
                params.append((mm.group(2), mm.group(1).strip()))
        fns.append((name, ret, params))
    return fns


def parse_structs(text):
    out = []
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            mm = re.match(r"(.+?)\b(\w+)\s*(\[(\d+)\])?$", decl)
            ctype, name, dim = mm.group(1).strip(), mm.group(2), mm.group(4)
            fields.append((name, ctype, int(dim) if dim else None))
        out.append((m.group(3), fields))
    return out


def parse_constants(text):
    consts = []
    for m in re.finditer(r"^#define\s+(DK_\w+)\s+(\(?-?[0-9xXa-fA-F]+[uU]?[lL]*\)?)\s*$", text, flags=re.M):
        name, val = m.group(1), m.group(2).strip("()")
        if name in ("DK_API",):
            continue
        unsigned = val.lower().endswith("u")
        consts.append((name, val.rstrip("uUlL"), "u32" if unsigned else "i32"))
    for m in re.finditer(r"enum\s*\{(.*?)\}\s*;", text, flags=re.S):
        nxt = 0
        for item in m.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, val = (x.strip() for x in item.split("="))
                nxt = int(val, 0)
            else:
                name = item
            consts.append((name, str(nxt) if nxt < 10 else hex(nxt), "i32"))
            nxt += 1
    return consts


def generate():
    text = strip_comments(open(HEADER).read())
    lines = ["//! GENERATED from include/doko_cuda.h by rs-doko-cuda/tools/gen_ffi.py — do not edit; `gen_ffi.py --check` runs in the test suite.",
             "//! Raw FFI of libdoko_cuda.so: every DK_API entry point, the POD structs and the constants of the C ABI.",
             "#![allow(non_camel_case_types, non_upper_case_globals, dead_code)]",
             "use std::ffi::{c_char, c_int, c_void};", "",
             "pub type dk_status = i32;", "pub type dk_stream = *mut c_void;", ""]
    for name in sorted(OPAQUE):
        lines += ["#[repr(C)]", f"pub struct {name} {{", "    _private: [u8; 0],", "}", ""]
    sizes = {"dk_state": 128, "dk_rng": 24, "dk_playout_stats": 2160, "dk_nccl_id": 128}
    for sname, fields in parse_structs(text):
        align = ", align(16)" if sname == "dk_state" else ""
        lines += [f"#[repr(C{align})]", "#[derive(Clone, Copy)]", f"pub struct {sname} {{"]
        for fname, ctype, dim in fields:
            rt = rust_type(ctype)
            lines.append(f"    pub {fname}: " + (f"[{rt}; {dim}]," if dim else f"{rt},"))
        lines += ["}"]
        if sname in sizes:
            lines.append(f"const _: () = assert!(std::mem::size_of::<{sname}>() == {sizes[sname]});")
        lines.append("")
    for name, val, ty in parse_constants(text):
        lines.append(f"pub const {name}: {ty} = {val};")
    lines += ["", 'extern "C" {']
    for name, ret, params in parse_functions(text):
        ps = ", ".join(f"{'r#' + p if p in ('in', 'type', 'move', 'ref') else p}: {rust_type(t)}" for p, t in params)
        rr = "" if ret == "void" else f" -> {rust_type(ret)}"
        lines.append(f"    pub fn {name}({ps}){rr};")
    lines += ["}", ""]
    return "\n".join(lines)


def main():
    want = generate()
    if "--check" in sys.argv:
        have = open(OUT).read() if os.path.exists(OUT) else ""
        if have != want:
            sys.stderr.write("rs-doko-cuda/src/ffi.rs is out of date: run python rs-doko-cuda/tools/gen_ffi.py\n")
            sys.exit(1)
        return
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    open(OUT, "w").write(want)
    print(f"wrote {OUT}: {want.count('pub fn ')} functions")


if __name__ == "__main__":
    main()
