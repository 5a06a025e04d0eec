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


def test_module_alias_drop_in():
    """INTEGRATION.md option A: the package can be aliased onto the reference's module names."""
    import importlib
    import sys
    import hyptokenizer_b200.embedding as E
    import hyptokenizer_b200.tokenizer as T
    for name in ("lorentz_model",):
        importlib.import_module(f"hyptokenizer_b200.embedding.{name}")
    mods = ("hyperbolic_merge", "fast_hyperbolic_merge", "frequency_aware_hyperbolic_merge",
            "hierarchical_hyperbolic_merge", "compression_aware_tokenizer", "adaptive_curvature_tokenizer",
            "enhanced_fast_hyperbolic_merge")
    for name in mods:
        importlib.import_module(f"hyptokenizer_b200.tokenizer.{name}")
    saved = {k: sys.modules.get(k) for k in ("embedding", "embedding.lorentz_model", "tokenizer",
                                             "tokenizer.hyperbolic_merge", "tokenizer.fast_hyperbolic_merge",
                                             *(f"tokenizer.{m}" for m in mods))}
    try:
        sys.modules["embedding"] = E
        sys.modules["embedding.lorentz_model"] = E.lorentz_model
        sys.modules["tokenizer"] = T
        for m in mods:
            sys.modules[f"tokenizer.{m}"] = getattr(T, m)
        from embedding.lorentz_model import (batch_distance, distance, exp_map, log_map, minkowski_dot,  # noqa: F401
                                             minkowski_norm, parallel_transport, project_to_hyperboloid)
        from tokenizer.fast_hyperbolic_merge import AdaptiveMergeCache, FastHyperbolicTokenizer, MergeCandidate  # noqa: F401
        from tokenizer.frequency_aware_hyperbolic_merge import FrequencyAwareHyperbolicTokenizer  # noqa: F401
        from tokenizer.enhanced_fast_hyperbolic_merge import (EnhancedFastHyperbolicTokenizer,  # noqa: F401
                                                              EnhancedMergeCandidate)
        from tokenizer.adaptive_curvature_tokenizer import AdaptiveCurvatureTokenizer  # noqa: F401
        from tokenizer.hyperbolic_merge import HyperbolicTokenizer  # noqa: F401
        import inspect
        sig = inspect.signature(HyperbolicTokenizer.__init__)
        assert list(sig.parameters)[1:9] == ["vocab", "embeddings", "curvature", "merge_threshold", "lr", "device",
                                             "max_vocab_size", "use_approximate_search"]
        sig = inspect.signature(FastHyperbolicTokenizer.__init__)
        for name in ("cache_size", "rebuild_frequency", "hnsw_m", "hnsw_ef_construction", "hnsw_ef_search"):
            assert name in sig.parameters
        c = AdaptiveMergeCache(max_size=3)
        c.add_batch([MergeCandidate(0.3, 1, 2), MergeCandidate(0.1, 0, 5), MergeCandidate(0.1, 0, 4), MergeCandidate(0.9, 7, 8)])
        assert [(m.token_i, m.token_j) for m in c.get_best(2)] == [(0, 5), (0, 4)]     # stable on distance only
        assert c.get_stats()["size"] == 1 and c.get_stats()["hit_count"] == 2
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
