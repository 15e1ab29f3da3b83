"""The C-ABI library loads on a CPU-only box and exports every symbol include/doko_cuda.h declares; compute calls fail loudly."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "doko_cuda.h")).read()
    return sorted(set(re.findall(r"DK_API\s+[\w\s\*]+?\b(dk_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import master_doko_reinforcement_learning_b200 as pkg

    lib = pkg.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in doko_cuda.h but not exported"


def test_no_cpu_fallback():
    """Without a GPU dk_init must fail (DK_ERR_NO_DEVICE) and the Python layer must raise — never compute on the CPU."""
    import torch

    import master_doko_reinforcement_learning_b200 as pkg

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.DokoCudaError):
        pkg.DokoCuda(0)


def test_product_does_not_reference_the_oracle():
    """Nothing under the package or include/ may import, include or link oracle/ or tests/hostsim."""
    pkg_dir = os.path.join(ROOT, "master_doko_reinforcement_learning_b200")
    for base, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle/" not in text and "liboracle" not in text and "oracle_lib" not in text, f"{f} references the oracle"
                assert "libhostsim" not in text and "hostsim.cpp" not in text


def test_dk_state_is_128_bytes():
    import master_doko_reinforcement_learning_b200 as pkg

    assert pkg.DK_STATE_DTYPE.itemsize == 128
