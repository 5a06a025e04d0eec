"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hyptok_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hyp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from hyptokenizer_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    h = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(h, name), f"{name} declared in include/hyptok_b200.h but not exported"
    # the Python binding covers the same set
    assert sorted(_lib.exported_symbols()) == names


def test_loads_without_gpu_and_reports_version():
    from hyptokenizer_b200 import _lib
    assert _lib.lib().hyp_abi_version() == 1


def test_struct_layouts():
    from hyptokenizer_b200 import _lib
    assert ctypes.sizeof(_lib.HypBest) == 32
    assert ctypes.sizeof(_lib.HypMergeState) == 40


def test_no_cpu_fallback():
    import torch
    from hyptokenizer_b200.embedding import lorentz_model as LM
    x = torch.randn(3, 6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        LM.distance(x, x)
    if not torch.cuda.is_available():
        from hyptokenizer_b200.tokenizer.hyperbolic_merge import HyperbolicTokenizer
        with pytest.raises(RuntimeError):
            HyperbolicTokenizer(["a", "b"], torch.nn.Parameter(torch.randn(2, 4)))


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "hyptokenizer_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), (dirpath, fn)
