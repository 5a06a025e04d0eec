"""ctypes binding of libhyptok_b200.so (the C ABI declared in include/hyptok_b200.h).

This is the only place the package touches native code.  There is NO fallback: if the
shared library is missing or the device is not a CUDA device, calls raise.
torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# HYPTOK_B200_LIB points at another build of the same library (used to A/B kernel variants on one GPU box)
LIB_PATH = os.environ.get("HYPTOK_B200_LIB") or os.path.join(_HERE, "lib", "libhyptok_b200.so")
CSRC = os.path.join(_HERE, "csrc")

HYP_OK = 0
HYP_ERR_ARG, HYP_ERR_CUDA, HYP_ERR_FULL, HYP_ERR_UNSUPPORTED, HYP_ERR_WORKSPACE = -1, -2, -3, -4, -5
SEM = {"reference": 0, "lorentz": 1}


class HypBest(C.Structure):
    _fields_ = [("d", C.c_float), ("i", C.c_int32), ("j", C.c_int32),
                ("count_lo", C.c_uint32), ("count_hi", C.c_uint32), ("pad", C.c_uint32 * 3)]


class HypMergeState(C.Structure):
    _fields_ = [("threshold", C.c_double), ("n", C.c_int32), ("capacity", C.c_int32),
                ("best_d", C.c_float), ("best_i", C.c_int32), ("best_j", C.c_int32),
                ("steps_done", C.c_int32), ("stop", C.c_int32), ("pad", C.c_int32)]


assert C.sizeof(HypBest) == 32 and C.sizeof(HypMergeState) == 40

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_f = C.c_float

# name -> argtypes; every function returns int except the *_bytes / version helpers
_SIGNATURES = {
    "hyp_abi_version": ([], C.c_int),
    "hyp_last_error": ([], C.c_char_p),
    "hyp_check_device": ([], C.c_int),
    "hyp_minkowski_dot": ([_p, _i64, _p, _i64, _p, _i64, _i32, _p], C.c_int),
    "hyp_distance": ([_p, _i64, _p, _i64, _p, _i64, _i32, _f, _i32, _p], C.c_int),
    "hyp_log_map": ([_p, _i64, _p, _i64, _p, _i64, _i64, _i32, _i32, _p], C.c_int),
    "hyp_exp_map": ([_p, _i64, _p, _i64, _p, _i64, _i64, _i32, _p], C.c_int),
    "hyp_project": ([_p, _i64, _p, _i64, _i64, _i32, _f, _p], C.c_int),
    "hyp_midpoint": ([_p, _i64, _p, _p, _p, _p, _p, _i64, _i64, _i32, _f, _i32, _i32, _p], C.c_int),
    "hyp_rescore_pairs": ([_p, _i64, _p, _p, _p, _p, _i64, _i32, _f, _i32, _p], C.c_int),
    "hyp_batch_distance": ([_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i32, _f, _i32, _p], C.c_int),
    "hyp_allpairs_workspace_bytes": ([_i64], _i64),
    "hyp_allpairs_min": ([_p, _i64, _i64, _i32, _f, _i32, _f, _p, _p, _i64, _p], C.c_int),
    "hyp_allpairs_emit": ([_p, _i64, _i64, _i32, _f, _i32, _f, _p, _p, _p, _i64, _p, _p], C.c_int),
    "hyp_allpairs_hist": ([_p, _i64, _i64, _i32, _f, _i32, _f, C.c_uint32, _i32, _i32, _p, _p], C.c_int),
    "hyp_allpairs_row_ties": ([_p, _i64, _i64, _i32, _f, _i32, _f, C.c_uint32, _p, _p], C.c_int),
    "hyp_allpairs_emit_cut": ([_p, _i64, _i64, _i32, _f, _i32, _f, C.c_uint32, _i64, _p, _p, _p, _i64, _p, _p], C.c_int),
    "hyp_allpairs_topk": ([_p, _i64, _i64, _i64, _i64, _i32, _f, _i32, _i32, _p, _p, _p], C.c_int),
    "hyp_gram_topk_workspace_bytes": ([_i64, _i64, _i32], _i64),
    "hyp_gram_topk": ([_p, _i64, _i64, _i64, _i64, _i32, _f, _i32, _i32, _p, _p, _p, _p, _i64, _p], C.c_int),
    "hyp_ctx_create": ([C.POINTER(_p), _i32, _i32, _i64], C.c_int),
    "hyp_ctx_export": ([_p, _p], C.c_int),
    "hyp_ctx_connect": ([_p, _p], C.c_int),
    "hyp_ctx_status": ([_p, C.POINTER(C.c_int)], C.c_int),
    "hyp_ctx_destroy": ([_p], C.c_int),
    "hyp_allgather_topk": ([_p, _p, _i64, C.POINTER(_p), _p], C.c_int),
    "hyp_gram_topk_allgather": ([_p, _p, _i64, _i64, _i32, _f, _i32, _i32, _p, _p, _i64, C.POINTER(_p), _p], C.c_int),
    "hyp_distance_backward": ([_p, _i64, _p, _i64, _p, _p, _p, _i64, _i32, _f, _i32, _p], C.c_int),
    "hyp_batch_distance_backward_coef": ([_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _p, _i64, _i32, _f, _i32, _p],
                                         C.c_int),
    "hyp_merge_workspace_bytes": ([], _i64),
    "hyp_merge_state_init": ([_p, _p, _i32, _i32, C.c_double, _p], C.c_int),
    "hyp_merge_steps": ([_p, _i64, _p, _i32, _f, _i32, _p, _p, _i32, _i32, _i32, C.c_double, _i32, _p, _i64, _p],
                        C.c_int),
    "hyp_row_min": ([_p, _i64, _i64, _i64, _i32, _f, _i32, _f, _p, _p, _i64, _p], C.c_int),
    "hyp_gemv_topk_workspace_bytes": ([_i64], _i64),
    "hyp_gemv_topk": ([_p, _i64, _i64, _p, _i64, _i32, _f, _i32, _i32, _p, _p, _p, _i64, _p], C.c_int),
    "hyp_score_candidates": ([_p, _i64, _p, _p, _p, _p, _p, _i32, _p, _p, C.c_double, C.c_double, C.c_double, C.c_double,
                              _p, _p, _i64, _i32, _f, _i32, _p], C.c_int),
    "hyp_coherence_distances": ([_p, _i64, _p, _p, _p, _p, _p, _i32, _p, _i64, _i32, _f, _i32, _p], C.c_int),
    "hyp_apply_merges": ([_p, _p, _i64, _p, _p, _p, _i32, _p, _p, _i64, _p, _p, _p], C.c_int),
    "hyp_pair_count": ([_p, _i64, _p, _p, _p, _i64, _p, _p], C.c_int),
    "hyp_pair_count_sorted_workspace_bytes": ([_i64], _i64),
    "hyp_pair_count_sorted": ([_p, _p, _p, _i64, _i64, _p, _p, _i64, _p, _p, _i64, _p], C.c_int),
}

_lib: Optional[C.CDLL] = None


def build(verbose: bool = False) -> str:
    """Compile every kernel for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building libhyptok_b200.so failed (see output above)")
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) and shutil.which("nvcc") or (
                not os.path.exists(LIB_PATH) and os.path.exists("/usr/local/cuda/bin/nvcc")):
            build()                      # building the product is not a fallback; a missing compiler still raises below
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA library has not been built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "There is no CPU fallback.")
        h = C.CDLL(LIB_PATH)
        for name, (args, res) in _SIGNATURES.items():
            fn = getattr(h, name)        # AttributeError here == header/library drift
            fn.argtypes = args
            fn.restype = res
        if h.hyp_abi_version() != 1:
            raise RuntimeError("libhyptok_b200.so ABI version mismatch")
        _lib = h
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc: int) -> None:
    if rc == HYP_OK:
        return
    msg = lib().hyp_last_error().decode("utf-8", "replace")
    if rc == HYP_ERR_FULL:
        raise ValueError(msg)
    if rc == HYP_ERR_ARG:
        raise ValueError(f"hyptok_b200: {msg}")
    raise RuntimeError(f"hyptok_b200 error {rc}: {msg}")


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """The product path is CUDA only; anything else is an error, never a fallback."""
    dev = None
    for t in tensors:
        if not isinstance(t, torch.Tensor):
            raise TypeError("expected a torch.Tensor")
        if not t.is_cuda:
            raise RuntimeError("hyptokenizer_b200 runs on CUDA (sm_100a) tensors only; got a "
                               f"{t.device.type} tensor and there is no CPU fallback")
        if t.dtype != torch.float32:
            raise TypeError(f"fp32 tensors only (reference arithmetic is fp32), got {t.dtype}")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("tensors live on different devices")
    return dev


_device_checked = set()


def check_device(dev: torch.device) -> None:
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx in _device_checked:
        return
    with torch.cuda.device(idx):
        check(lib().hyp_check_device())
    _device_checked.add(idx)


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()
