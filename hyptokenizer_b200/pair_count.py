"""K6 host wrapper: adjacent code-point pair counts of a UTF-8 corpus on the device.

Replaces the Python dict loop of the reference's
tokenizer/frequency_aware_hyperbolic_merge.py:92-112 (see csrc/pair_count.cu for the exact
line/strip semantics that are reproduced).
"""
from __future__ import annotations

from typing import Dict, Tuple, Union

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def to_device_bytes(data: Union[bytes, bytearray, memoryview, np.ndarray, torch.Tensor],
                    device: torch.device) -> torch.Tensor:
    """uint8 CUDA tensor holding the corpus (pinned staging for host inputs)."""
    if isinstance(data, torch.Tensor):
        t = data
        if t.dtype != torch.uint8:
            raise TypeError("corpus tensor must be uint8")
        return t.to(device).contiguous()
    arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    arr = np.ascontiguousarray(arr)
    if not arr.flags.writeable:
        arr = arr.copy()             # torch refuses read-only buffers (bytes objects)
    host = torch.from_numpy(arr)
    if host.numel() and torch.cuda.is_available():
        host = host.pin_memory()
    return host.to(device, non_blocking=True)


def count_pairs_device(text: torch.Tensor, hash_capacity: int = 1 << 20):
    """Run the kernel on a uint8 CUDA tensor.  Returns (ascii_counts[128*128], keys, vals) on the
    device; raises if the non-ASCII table overflowed."""
    if not text.is_cuda or text.dtype != torch.uint8:
        raise RuntimeError("count_pairs_device needs a uint8 CUDA tensor (no CPU fallback)")
    dev = text.device
    _lib.check_device(dev)
    text = text.contiguous()
    if text.data_ptr() % 16:
        text = text.clone()          # the kernel stages the stream with 16-byte vector loads
    ascii_counts = torch.empty(128 * 128, dtype=torch.int64, device=dev)
    keys = torch.empty(hash_capacity, dtype=torch.int64, device=dev)
    vals = torch.empty(hash_capacity, dtype=torch.int64, device=dev)
    overflow = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().hyp_pair_count(ptr(text) if text.numel() else None, text.numel(), ptr(ascii_counts),
                                        ptr(keys), ptr(vals), hash_capacity, ptr(overflow), stream_ptr()))
    return ascii_counts, keys, vals, overflow


def count_pairs(data, device: torch.device = None, hash_capacity: int = 1 << 20) -> Dict[Tuple[str, str], int]:
    """Corpus bytes -> {(a, b): count}, the reference's `pair_frequencies` dict."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    text = to_device_bytes(data, device)
    while True:
        ascii_counts, keys, vals, overflow = count_pairs_device(text, hash_capacity)
        if int(overflow.item()) == 0:
            break
        hash_capacity *= 4
    return pairs_to_dict(ascii_counts, keys, vals)


def pairs_to_dict(ascii_counts: torch.Tensor, keys: torch.Tensor, vals: torch.Tensor) -> Dict[Tuple[str, str], int]:
    out: Dict[Tuple[str, str], int] = {}
    a = ascii_counts.cpu().numpy()
    nz = np.nonzero(a)[0]
    for k in nz.tolist():
        out[(chr(k >> 7), chr(k & 127))] = int(a[k])
    used = (keys != -1).nonzero(as_tuple=True)[0]
    if used.numel():
        kk = keys[used].cpu().numpy().astype(np.uint64)
        vv = vals[used].cpu().numpy()
        for key, v in zip(kk.tolist(), vv.tolist()):
            out[(chr(key >> 32), chr(key & 0xffffffff))] = int(v)
    return out
