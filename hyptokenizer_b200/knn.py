"""All-pairs Lorentz distance + per-row top-k (BASELINE configs[2]), row-sharded across GPUs.

This is the exact replacement for what the reference asks FAISS for (k nearest neighbours of every
token, fast_hyperbolic_merge.py:301-304 / hyperbolic_merge.py:217): true Lorentz distance, every row,
no sampling.  Rows are independent, so rank r of G owns the contiguous block
[r*ceil(n/G), ...) and scores it against ALL n columns (every rank holds the full table: 40 MB at
V=100k); the only exchange is one all-gather of the per-shard (n/G, k) results over NCCL/NVLink,
after which every rank holds the full (n, k) lists and derives the same global argmin -- the
"chosen merge broadcast" of the north star is a replicated deterministic reduction.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import SEM, check, ptr, stream_ptr
from .embedding import lorentz_model as LM


def shard_rows(n: int, world: int, rank: int) -> Tuple[int, int, int]:
    """(row0, nrows, rows_per_rank): contiguous blocks of ceil(n/world) rows; trailing ranks may be short/empty."""
    per = (n + world - 1) // world
    row0 = min(rank * per, n)
    return row0, min(per, n - row0), per


def lorentz_topk(E: torch.Tensor, k: int = 32, c: float = 1.0, semantics: Optional[str] = None,
                 n: Optional[int] = None, row0: int = 0, nrows: Optional[int] = None,
                 engine: str = "exact") -> Tuple[torch.Tensor, torch.Tensor]:
    """k nearest rows (by Lorentz distance, ties on index) of rows [row0, row0+nrows) among rows [0, n) of E.
    Returns (idx int32 [nrows, k], dist fp32 [nrows, k]), ascending.  CUDA only."""
    if not E.is_cuda or E.dtype != torch.float32 or E.dim() != 2:
        raise RuntimeError("lorentz_topk needs a 2-D float32 CUDA tensor (no CPU fallback)")
    _lib.check_device(E.device)
    E = E.detach()
    if E.stride(1) != 1:
        E = E.contiguous()
    n = E.shape[0] if n is None else n
    nrows = n - row0 if nrows is None else nrows
    sem = SEM[LM.get_semantics() if semantics is None else semantics]
    idx = torch.empty((nrows, k), dtype=torch.int32, device=E.device)
    d = torch.empty((nrows, k), dtype=torch.float32, device=E.device)
    L = _lib.lib()
    if engine == "auto":
        engine = "tc" if (n >= 8192 and k <= 32 and E.shape[1] - 1 <= 124 and sem == SEM["lorentz"]) else "exact"
    if engine == "exact":
        with torch.cuda.device(E.device):
            check(L.hyp_allpairs_topk(ptr(E), E.stride(0), n, row0, nrows, E.shape[1], float(c), sem, k,
                                      ptr(idx), ptr(d), stream_ptr()))
        return idx, d
    if engine != "tc":
        raise ValueError("engine must be 'exact', 'tc' or 'auto'")
    nbytes = L.hyp_gram_topk_workspace_bytes(n, nrows, E.shape[1])
    if nbytes < 0:
        raise ValueError("tensor-core path supports d <= 124 (d + 4 operand columns fit one 128-column tile)")
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=E.device)
    off = (-ws.data_ptr()) % 256
    flags = torch.empty(nrows, dtype=torch.int32, device=E.device)
    with torch.cuda.device(E.device):
        check(L.hyp_gram_topk(ptr(E), E.stride(0), n, row0, nrows, E.shape[1], float(c), sem, k, ptr(idx), ptr(d),
                              ptr(flags), ws.data_ptr() + off, nbytes, stream_ptr()))
        bad = flags.nonzero(as_tuple=True)[0]
        lorentz_topk.last_flagged = int(bad.numel())
        if bad.numel() > max(64, nrows // 8):
            # ties everywhere (e.g. the shipped semantics, where every distance is 0): exact path for the shard
            check(L.hyp_allpairs_topk(ptr(E), E.stride(0), n, row0, nrows, E.shape[1], float(c), sem, k,
                                      ptr(idx), ptr(d), stream_ptr()))
        else:
            for r in bad.tolist():
                check(L.hyp_allpairs_topk(ptr(E), E.stride(0), n, row0 + r, 1, E.shape[1], float(c), sem, k,
                                          idx[r].data_ptr(), d[r].data_ptr(), stream_ptr()))
    return idx, d


def gather_topk(local_idx: torch.Tensor, local_d: torch.Tensor, n: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather per-shard (nrows_r, k) results into the full (n, k) lists on every rank.
    Works on CUDA tensors over NCCL and on CPU tensors over gloo (host-logic tests)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    _, nrows, per = shard_rows(n, world, rank)
    k = local_idx.shape[1]
    assert local_idx.shape[0] == nrows and local_d.shape == local_idx.shape
    pad_i = torch.full((per, k), -1, dtype=local_idx.dtype, device=local_idx.device)
    pad_d = torch.full((per, k), float("inf"), dtype=local_d.dtype, device=local_d.device)
    pad_i[:nrows] = local_idx
    pad_d[:nrows] = local_d
    all_i = torch.empty((world * per, k), dtype=local_idx.dtype, device=local_idx.device)
    all_d = torch.empty((world * per, k), dtype=local_d.dtype, device=local_d.device)
    dist.all_gather_into_tensor(all_i, pad_i, group=group)
    dist.all_gather_into_tensor(all_d, pad_d, group=group)
    return all_i[:n], all_d[:n]


def best_pair_from_topk(idx: torch.Tensor, d: torch.Tensor) -> Tuple[int, int, float]:
    """argmin over (d, i, j), i < j, from full per-row lists -- the pair the merge loop would pick
    (hyperbolic_merge.py:378).  Deterministic, so every rank computes the same answer."""
    n, k = idx.shape
    rows = torch.arange(n, device=idx.device).unsqueeze(1).expand(n, k)
    valid = (idx > rows) & torch.isfinite(d)          # each unordered pair appears in the row of its smaller index
    if not bool(valid.any()):                          # or only in the other row's list: fall back to i > j entries
        valid = (idx >= 0) & torch.isfinite(d)
    dd = torch.where(valid, d, torch.full_like(d, float("inf")))
    lo = torch.minimum(rows, idx.to(rows.dtype))
    hi = torch.maximum(rows, idx.to(rows.dtype))
    best_d = dd.min()
    cand = dd == best_d
    key = torch.where(cand, lo * n + hi, torch.full_like(lo, n * n))
    flat = int(key.min().item())
    return flat // n, flat % n, float(best_d.item())


def lorentz_topk_sharded(E: torch.Tensor, k: int = 32, c: float = 1.0, semantics: Optional[str] = None,
                         n: Optional[int] = None, group=None, engine: str = "exact") -> Tuple[torch.Tensor, torch.Tensor]:
    """Row-sharded all-pairs top-k: local shard on this GPU, one all-gather, full lists everywhere."""
    n = E.shape[0] if n is None else n
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return lorentz_topk(E, k, c, semantics, n, engine=engine)
    row0, nrows, _ = shard_rows(n, dist.get_world_size(group), dist.get_rank(group))
    li, ld = lorentz_topk(E, k, c, semantics, n, row0, nrows, engine=engine)
    return gather_topk(li, ld, n, group)


# ---------------------------------------------------------------------------------------------------------
# recall against what the reference's FAISS indexes rank by (evaluation only, never on a hot path)
# ---------------------------------------------------------------------------------------------------------
def klein_l2_topk(E: torch.Tensor, k: int, rows: torch.Tensor, n: Optional[int] = None) -> torch.Tensor:
    """The k nearest rows of each of `rows` by squared L2 distance between Klein coordinates `xs / (x0 + 1e-8)`, self
    excluded, ties on index: EXACTLY what the reference's `IndexFlatL2` returns and what its HNSW index approximates
    (fast_hyperbolic_merge.py:195-240 builds the index over these coordinates, :286-304 queries it).  FAISS itself is
    not installable here, so this is the stand-in for "the reference FAISS path" in recall figures: a perfect HNSW
    search returns this list.  Plain torch ops (a reporting helper, works on any device); returns int64 [len(rows), k]."""
    n = E.shape[0] if n is None else n
    K = (E[:n, 1:] / (E[:n, 0:1] + 1e-8)).to(torch.float64)
    q = K[rows]
    d2 = (q * q).sum(1, keepdim=True) - 2.0 * (q @ K.T) + (K * K).sum(1)[None, :]
    d2[torch.arange(len(rows), device=E.device), rows] = float("inf")
    # ascending by (distance, index): stable sort of the index-ordered columns
    return torch.sort(d2, dim=1, stable=True).indices[:, :k]


def recall_at_k(found: torch.Tensor, truth: torch.Tensor) -> float:
    """Mean over rows of |found[r] ∩ truth[r]| / k for two [rows, k] index lists."""
    f = found.to(torch.int64).unsqueeze(2)
    t = truth.to(torch.int64).unsqueeze(1)
    return float((f == t).any(dim=2).to(torch.float64).mean().item())
