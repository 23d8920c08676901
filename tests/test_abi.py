"""CPU tests: the C-ABI library loads without a GPU and exports every symbol include/blu_b200.h
declares; compute calls fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import blu_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "blu_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(blu_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = blu_b200.library_path()
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    L = ctypes.CDLL(lib)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/blu_b200.h but not exported"
    L.blu_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.blu_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        blu_b200.BLU(10, 32)          # blu_create -> BLU_ERROR_CUDA, there is no CPU path


def test_product_does_not_reference_oracle():
    """Nothing under blu_b200/ (the product) may import, link or execute oracle/."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "blu_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c")) or f == "Makefile":
                s = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in s.lower() or f in ("gen.py", "blugen.c", "Makefile", "blu_types.h"), f
                assert "libblo" not in s and "blo_" not in s, f


def test_c_program_links_and_fails_loudly(tmp_path):
    """tests/abi/abi_check.c: plain C99 against include/blu_b200.h and libblu_b200.so.  Without a GPU the first
    call answers BLU_ERROR_CUDA; with one the program solves the reference's 10 x 10 example."""
    import subprocess
    lib = blu_b200.library_path()
    exe = str(tmp_path / "abi_check")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi", "abi_check.c"),
                           "-L", os.path.dirname(lib), "-lblu_b200", "-Wl,-rpath," + os.path.dirname(lib), "-lm", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "sm_100a" in out.stdout
    assert ("BLU_ERROR_CUDA" in out.stdout) or ("solution 0.1..1.0 recovered" in out.stdout)
