"""The Rust binding cannot be compiled here (no cargo / rustc in the image), so it is pinned structurally: src/ffi.rs must be exactly
what tools/gen_ffi.py generates from include/doko_cuda.h (every DK_API function, struct and constant), and every function it declares
must be exported by the built library.  Any signature drift between header, Rust and .so fails here."""
import ctypes
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CRATE = os.path.join(ROOT, "rs-doko-cuda")


def test_ffi_rs_is_the_generated_binding_of_the_header():
    r = subprocess.run([sys.executable, os.path.join(CRATE, "tools", "gen_ffi.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_every_rust_extern_fn_is_declared_in_the_header_and_exported():
    import master_doko_reinforcement_learning_b200 as pkg

    ffi = open(os.path.join(CRATE, "src", "ffi.rs")).read()
    header = open(os.path.join(ROOT, "include", "doko_cuda.h")).read()
    rust_fns = re.findall(r"pub fn (dk_\w+)\(", ffi)
    header_fns = re.findall(r"DK_API[^;(]*?\b(dk_\w+)\s*\(", header)
    assert sorted(rust_fns) == sorted(header_fns) and len(rust_fns) == len(set(rust_fns)) >= 47
    lib = ctypes.CDLL(pkg.library_path())
    for name in rust_fns:
        assert hasattr(lib, name), name


def test_rust_argument_counts_match_the_header():
    ffi = open(os.path.join(CRATE, "src", "ffi.rs")).read()
    header = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "doko_cuda.h")).read(), flags=re.S)
    for m in re.finditer(r"DK_API\s+[^;(]+?\b(dk_\w+)\s*\(([^;]*?)\)\s*;", header, flags=re.S):
        name, args = m.group(1), " ".join(m.group(2).split())
        n_c = 0 if args in ("", "void") else args.count(",") + 1
        rm = re.search(r"pub fn " + name + r"\(([^)]*)\)", ffi)
        n_r = 0 if not rm.group(1).strip() else rm.group(1).count(",") + 1
        assert n_c == n_r, name


def test_crate_uses_only_declared_ffi_symbols():
    """Every `ffi::dk_*` / bare dk_* call in the hand-written modules names a generated function (a typo would only show up at link time)."""
    ffi = open(os.path.join(CRATE, "src", "ffi.rs")).read()
    declared = set(re.findall(r"pub fn (dk_\w+)\(", ffi)) | set(re.findall(r"pub (?:struct|type) (dk_\w+)", ffi)) | {"dk_unpack_points"}   # (a static inline of the header, restated in lib.rs)
    for mod in ("lib.rs", "env.rs", "convert.rs"):
        src = open(os.path.join(CRATE, "src", mod)).read()
        for name in set(re.findall(r"\b(dk_[a-z_0-9]+)\b", src)):
            assert name in declared, (mod, name)
