"""CPU tests: the C-ABI library loads and exports every symbol include/libdamgpu.h declares;
compute entry points fail loudly without a GPU (no CPU fallback)."""
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "libdamgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(damgpu_[A-Za-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from damapper_b200 import api
    assert os.path.exists(api.LIB_PATH), "libdamgpu.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(api.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    # the Python binding covers the same surface
    assert set(names) == set(api.SYMBOLS)


def test_map_h_quartet_is_present():
    names = _declared_symbols()
    for n in ("damgpu_Set_Filter_Params", "damgpu_Sort_Kmers", "damgpu_Match_Filter", "damgpu_Reporter"):
        assert n in names


def test_set_filter_params_semantics():
    """Set_Filter_Params (map.c:124-150): 1 for kmer <= 1, else 0 (no GPU needed)."""
    from damapper_b200 import api
    L = api.load()
    assert L.damgpu_Set_Filter_Params(1, 0, 4) == 1
    assert L.damgpu_Set_Filter_Params(20, 0, 4) == 0


def test_no_cpu_fallback():
    """Without a CUDA device the library refuses to initialise instead of computing on the host."""
    import torch
    from damapper_b200 import api
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        api.init()


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "damapper_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower(), (dirpath, f)
