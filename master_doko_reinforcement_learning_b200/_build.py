"""Builds libdoko_cuda.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc."""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libdoko_cuda.so")
SOURCES = ["cabi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libdoko_cuda.so cannot be built (there is no CPU fallback)")


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(PKG_DIR), "include", "doko_cuda.h")]
    return any(os.path.getmtime(d) > t for d in deps)


SASS_HASHES = os.path.join(PKG_DIR, "libdoko_cuda.sass.json")


def write_sass_hashes():
    """sha256 of every kernel's SASS in the built library → libdoko_cuda.sass.json (next to the .so, a build artefact).
    bench.py compares them with the hashes recorded in profiles/kernel_counters.json: ncu counters that were taken on another build of a
    kernel are reported as stale instead of silently feeding the roofline."""
    import hashlib
    import json

    cuobjdump = os.path.join(os.path.dirname(find_nvcc()), "cuobjdump")
    if not os.path.exists(cuobjdump):
        cuobjdump = shutil.which("cuobjdump")
    if not cuobjdump:
        return None
    txt = subprocess.run([cuobjdump, "-sass", LIB_PATH], capture_output=True, text=True, check=True).stdout
    hashes, name, body = {}, None, []
    for line in txt.splitlines():
        t = line.strip()
        if t.startswith("Function :"):
            if name:
                hashes[name] = hashlib.sha256("\n".join(body).encode()).hexdigest()
            name, body = t.split(":", 1)[1].strip(), []
        elif name and t.startswith("/*"):
            body.append(t)
    if name:
        hashes[name] = hashlib.sha256("\n".join(body).encode()).hexdigest()
    with open(SASS_HASHES, "w") as f:
        json.dump(hashes, f, indent=0, sort_keys=True)
    return hashes


def sass_hashes():
    """The hashes written by the last build ({} when the file is missing)."""
    import json

    try:
        if os.path.getmtime(SASS_HASHES) + 1 < os.path.getmtime(LIB_PATH):
            return {}
        with open(SASS_HASHES) as f:
            return json.load(f)
    except OSError:
        return {}


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu → libdoko_cuda.so for sm_100a.  Returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES + ["-ldl"]
    subprocess.check_call(cmd, cwd=CSRC)
    write_sass_hashes()
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
