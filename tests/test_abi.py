"""CPU tests of the drop-in boundary: libpedfac_cuda.so loads, exports exactly what
include/pedfac_cuda.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from pedfac_b200 import abi

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    txt = (ROOT / "include" / "pedfac_cuda.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", txt))
    names -= {"pf_compose_ll", "pf_compose_swap_ll"}      # static inline helpers
    return names


def test_header_and_binding_list_agree():
    assert header_symbols() == {"pf_" + n for n in abi.EXPORTS}


def test_library_exports_every_declared_symbol():
    lib = abi.load_cuda_library()
    for name in sorted(header_symbols()):
        assert hasattr(lib, name), f"{name} declared in include/pedfac_cuda.h but not exported"


def test_no_cpu_fallback_without_gpu():
    lib = abi.load_cuda_library()
    n = C.c_int(0)
    cudart = C.CDLL("libcudart.so")
    if cudart.cudaGetDeviceCount(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(abi.PedfacError) as ei:
        abi.CEngine(lib, "pf_", 8, 4, 2)
    assert ei.value.code == abi.PF_ENODEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_never_references_the_oracle():
    banned = re.compile(r"import\s+oracle|from\s+oracle|libpedfac_oracle|\bpfo_|pedfac_oracle\.h")
    for p in (ROOT / "pedfac_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".cpp", ".h", ".c"} or p.name == "Makefile":
            assert not banned.search(p.read_text()), f"{p} references the test-only oracle"
