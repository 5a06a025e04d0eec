"""K6 host wrapper: adjacent code-point pair counts of a UTF-8 corpus on the device.

Replaces the Python dict loop of the reference's
tokenizer/frequency_aware_hyperbolic_merge.py:92-112 (see csrc/pair_count.cu for the exact
line/strip semantics that are reproduced).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple, Union

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


# ---------------------------------------------------------------------------------------------------------
# multi-GPU: byte-range shards aligned to line boundaries
# ---------------------------------------------------------------------------------------------------------
def shard_byte_range(buf: np.ndarray, rank: int, world: int) -> Tuple[int, int]:
    """[start, end) of `rank`'s share of a corpus: the nominal cuts k*n//world, each moved forward to just behind
    the next '\n' (or to n).  The reference counts pairs inside each line only (frequency_aware_hyperbolic_merge.py:
    92-112), and a cut behind '\n' never splits a line (nor a "\r\n"), so the shards count independently."""
    n = int(buf.size)

    def cut(k: int) -> int:
        pos = (k * n) // world
        if pos <= 0 or pos >= n:
            return max(0, min(pos, n))
        if buf[pos - 1] == 0x0A:
            return pos
        nl = np.flatnonzero(buf[pos:] == 0x0A) if n - pos <= (1 << 20) else None
        if nl is None:                                     # long corpus: look in growing windows
            width = 1 << 16
            while True:
                w = np.flatnonzero(buf[pos:pos + width] == 0x0A)
                if w.size or pos + width >= n:
                    nl = w
                    break
                width *= 16
        return pos + int(nl[0]) + 1 if nl.size else n

    return cut(rank), cut(rank + 1)


def reduce_pair_counts(ascii_counts: torch.Tensor, extra: Dict[Tuple[int, int], int], group=None):
    """Sum the per-shard results over the process group: ONE all_reduce of the dense 128x128 histogram (128 KB,
    the only data-path collective) and an object all-gather of the (rare) non-ASCII pairs.  Every rank returns
    the same (ascii_counts, extra)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return ascii_counts, dict(extra)
    dist.all_reduce(ascii_counts, op=dist.ReduceOp.SUM, group=group)
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, extra, group=group)
    merged: Dict[Tuple[int, int], int] = {}
    for part in parts:
        for key, v in part.items():
            merged[key] = merged.get(key, 0) + int(v)
    return ascii_counts, merged


def _device_counter(device: torch.device, hash_capacity: int):
    def count(shard: np.ndarray):
        text = to_device_bytes(shard, device)
        cap = hash_capacity
        while True:
            ascii_counts, keys, vals, overflow = count_pairs_device(text, cap)
            if int(overflow.item()) == 0:
                break
            cap *= 4
        extra: Dict[Tuple[int, int], int] = {}
        used = (keys != -1).nonzero(as_tuple=True)[0]
        if used.numel():
            kk = keys[used].cpu().numpy().astype(np.uint64)
            vv = vals[used].cpu().numpy()
            for key, v in zip(kk.tolist(), vv.tolist()):
                extra[(key >> 32, key & 0xffffffff)] = int(v)
        return ascii_counts, extra
    return count


def count_pairs_sharded(data, group=None, device: Optional[torch.device] = None, hash_capacity: int = 1 << 20,
                        counter: Optional[Callable] = None) -> Dict[Tuple[str, str], int]:
    """`count_pairs` over a process group: every rank holds (or can read) the corpus, counts its line-aligned
    byte range on its own GPU and the histograms are summed.  `counter(shard_bytes) -> (int64[16384] tensor,
    {(cp_a, cp_b): n})` replaces the device kernel in the CPU (gloo) tests of this host logic."""
    import torch.distributed as dist
    arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    if counter is None:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        counter = _device_counter(device, hash_capacity)
    start, end = shard_byte_range(arr, rank, world)
    ascii_counts, extra = counter(arr[start:end])
    ascii_counts, extra = reduce_pair_counts(ascii_counts, extra, group)
    out: Dict[Tuple[str, str], int] = {}
    a = ascii_counts.cpu().numpy()
    for k in np.nonzero(a)[0].tolist():
        out[(chr(k >> 7), chr(k & 127))] = int(a[k])
    for (ca, cb), v in extra.items():
        out[(chr(ca), chr(cb))] = int(v)
    return out
