"""The C-ABI libraries load and export every symbol the headers declare (no compute, no GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import CUDA_HOST_LIB, CUDA_LIB, ORACLE_LIB, ROOT


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kmlh?_[a-z0-9_A-Z]+)\s*\(", text)))


def exported(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    return {ln.split()[-1] for ln in out.splitlines() if " T " in ln}


def ensure_cuda_build():
    if not (os.path.exists(CUDA_LIB) and os.path.exists(CUDA_HOST_LIB)):
        import __graft_entry__
        __graft_entry__.build()


def test_kml_h_symbols_exported_by_cuda_library():
    ensure_cuda_build()
    names = declared("kml.h")
    assert len(names) > 35
    missing = [n for n in names if n not in exported(CUDA_LIB)]
    assert not missing, missing


def test_kml_host_h_symbols_exported():
    ensure_cuda_build()
    missing = [n for n in declared("kml_host.h") if n not in exported(CUDA_HOST_LIB)]
    assert not missing, missing


def test_oracle_implements_the_same_abi(oracle_lib):
    missing = [n for n in declared("kml.h") if n not in exported(ORACLE_LIB)]
    assert not missing, missing


def test_cuda_library_loads_and_reports_backend():
    ensure_cuda_build()
    lib = ctypes.CDLL(CUDA_LIB)
    lib.kml_backend.restype = ctypes.c_char_p
    assert lib.kml_backend() == b"cuda-sm_100a"


def test_product_fails_loudly_without_a_gpu():
    """No CPU fallback: creating a context without a CUDA device is an error, not a silent port."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ensure_cuda_build()
    from karamelo_b200.api import Engine, KmlError
    from cases import two_disks
    e = Engine()
    with pytest.raises(KmlError, match="no CUDA device|cuda"):
        e.script(two_disks())


def test_product_never_references_the_oracle():
    for dp, _, fns in os.walk(os.path.join(ROOT, "karamelo_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cpp", ".h", ".cu", ".cuh", "Makefile")):
                assert "oracle" not in open(os.path.join(dp, fn), errors="ignore").read().lower(), os.path.join(dp, fn)
