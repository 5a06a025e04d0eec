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
    E, n, nrows, sem = _check_table(E, n, row0, nrows, semantics)
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
    ws, wsp, nbytes = _tc_workspace(E, n, nrows)
    flags = torch.empty(nrows, dtype=torch.int32, device=E.device)
    with torch.cuda.device(E.device):
        # rows the filter cannot certify (massive ties) are recomputed exactly inside the call; flags only report them
        check(L.hyp_gram_topk(ptr(E), E.stride(0), n, row0, nrows, E.shape[1], float(c), sem, k, ptr(idx), ptr(d),
                              ptr(flags), wsp, nbytes, stream_ptr()))
    lorentz_topk.last_flags = flags
    return idx, d


def row_topk(E: torch.Tensor, row: int, k: int = 32, c: float = 1.0, semantics: Optional[str] = None,
             n: Optional[int] = None, q: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The k nearest rows of ONE row (or of an external query `q` of D floats; then nothing is excluded unless `row`
    >= 0) among rows [0, n): the per-merge incremental query (hyp_gemv_topk, K4).  Returns (idx int32 [k], d fp32 [k]),
    bit-identical to the corresponding row of lorentz_topk."""
    E, n, _, sem = _check_table(E, n, 0, None, semantics)
    qv = E[row] if q is None else q.detach().to(torch.float32).contiguous()
    idx = torch.empty(k, dtype=torch.int32, device=E.device)
    d = torch.empty(k, dtype=torch.float32, device=E.device)
    L = _lib.lib()
    ws = torch.empty(max(L.hyp_gemv_topk_workspace_bytes(n), 8), dtype=torch.uint8, device=E.device)
    with torch.cuda.device(E.device):
        check(L.hyp_gemv_topk(ptr(E), E.stride(0), n, ptr(qv), int(row), E.shape[1], float(c), sem, k, ptr(idx), ptr(d),
                              ptr(ws), ws.numel(), stream_ptr()))
    return idx, d


def _check_table(E, n, row0, nrows, semantics):
    if not E.is_cuda or E.dtype != torch.float32 or E.dim() != 2:
        raise RuntimeError("lorentz_topk needs a 2-D float32 CUDA tensor (no CPU fallback)")
    _lib.check_device(E.device)
    E = E.detach()
    if E.stride(1) != 1:
        E = E.contiguous()
    n = E.shape[0] if n is None else n
    nrows = n - row0 if nrows is None else nrows
    return E, n, nrows, SEM[LM.get_semantics() if semantics is None else semantics]


def _tc_workspace(E, n, nrows):
    nbytes = _lib.lib().hyp_gram_topk_workspace_bytes(n, nrows, E.shape[1])
    if nbytes < 0:
        raise ValueError("tensor-core path supports d <= 124 (d + 4 operand columns fit one 128-column tile)")
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=E.device)
    return ws, ws.data_ptr() + ((-ws.data_ptr()) % 256), nbytes


def last_flagged() -> int:
    """Rows of the last tensor-core call that took the exact redo (synchronises; diagnostics and tests)."""
    f = getattr(lorentz_topk, "last_flags", None)
    return 0 if f is None else int(f.sum().item())


def split_records(rec: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[rows, k] int64 records {int32 idx (low word), fp32 d bits (high word)} -> (idx int32, d fp32) views."""
    w = rec.view(torch.int32).view(rec.shape[0], rec.shape[1], 2)
    return w[..., 0], w[..., 1].view(torch.float32)


def join_records(idx: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    w = torch.stack((idx.to(torch.int32), d.to(torch.float32).view(torch.int32)), dim=2).contiguous()
    return w.view(torch.int64).view(idx.shape[0], idx.shape[1])


class TopkContext:
    """hyp_ctx (include/hyptok_b200.h, K8): per-rank gather buffers mapped into every peer with CUDA IPC, so the
    kernels that finish a row of the top-k lists store it into ALL ranks' buffers over NVLink and the all-gather
    costs one barrier kernel.  The 64-byte IPC handles are exchanged with one torch.distributed all_gather."""

    def __init__(self, n: int, k: int, device: torch.device, group=None):
        import ctypes as C
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n, self.k, self.device = n, k, device
        self.per = (n + self.world - 1) // self.world
        self.slot_bytes = ((max(self.per * k * 8, 16) + 255) // 256) * 256
        L = _lib.lib()
        self._h = C.c_void_p()
        with torch.cuda.device(device):
            check(L.hyp_ctx_create(C.byref(self._h), self.rank, self.world, self.slot_bytes))
            if self.world > 1:
                mine = (C.c_uint8 * 64)()
                check(L.hyp_ctx_export(self._h, mine))
                t = torch.tensor(list(mine), dtype=torch.uint8, device=device)
                allh = torch.empty(self.world * 64, dtype=torch.uint8, device=device)
                dist.all_gather_into_tensor(allh, t, group=group)
                buf = (C.c_uint8 * (self.world * 64))(*allh.cpu().tolist())
                rc = L.hyp_ctx_connect(self._h, buf)
                # every rank must agree on whether peer memory is usable, or the barriers would not match up
                ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
                if int(ok.item()) == 0:
                    msg = L.hyp_last_error().decode("utf-8", "replace") if rc else "a peer could not map this rank's buffer"
                    self.close()
                    raise RuntimeError(f"hyp_ctx_connect failed: {msg}")

    def gram_topk(self, E: torch.Tensor, c: float = 1.0, semantics: Optional[str] = None) -> torch.Tensor:
        """Sharded tensor-core top-k + fused all-gather.  Returns the [world*per, k] int64 record view of this rank's
        gather buffer (valid until the second next call); use split_records()."""
        import ctypes as C
        E, n, _, sem = _check_table(E, self.n, 0, None, semantics)
        nrows = max(0, min(self.per, n - self.rank * self.per))
        if getattr(self, "_ws_key", None) != (n, E.shape[1]):
            self._ws, self._wsp, self._wsb = _tc_workspace(E, n, max(nrows, 1))
            self._flags = torch.empty(max(self.per, 1), dtype=torch.int32, device=E.device)
            self._ws_key = (n, E.shape[1])
        out = C.c_void_p()
        with torch.cuda.device(E.device):
            check(_lib.lib().hyp_gram_topk_allgather(self._h, ptr(E), E.stride(0), n, E.shape[1], float(c), sem, self.k,
                                                     ptr(self._flags), self._wsp, self._wsb, C.byref(out), stream_ptr()))
        return self._view(out.value)

    def allgather(self, idx: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
        """All-gather of per-shard (<= per, k) lists computed elsewhere (e.g. by the exact engine)."""
        import ctypes as C
        rec = torch.full((self.slot_bytes // 8,), -1, dtype=torch.int64, device=idx.device)
        rec[:idx.shape[0] * self.k] = join_records(idx, d).reshape(-1)
        out = C.c_void_p()
        with torch.cuda.device(idx.device):
            check(_lib.lib().hyp_allgather_topk(self._h, ptr(rec), self.slot_bytes, C.byref(out), stream_ptr()))
        rec.record_stream(torch.cuda.current_stream())   # the push kernel reads `rec` on this stream
        return self._view(out.value)

    def _view(self, address: int) -> torch.Tensor:
        """[world*per, k] int64 records of the gather buffer at `address` (slot g = rank g's rows, slot_bytes apart)."""
        class _Mem:
            pass
        m = _Mem()
        m.owner = self      # the tensor's storage keeps `m`, hence the context (which owns the memory), alive
        m.__cuda_array_interface__ = {"shape": (self.world, self.slot_bytes // 8), "typestr": "<i8",
                                      "data": (address, False), "version": 3}
        flat = torch.as_tensor(m, device=self.device)
        return flat[:, :self.per * self.k].reshape(self.world * self.per, self.k)

    def status(self) -> int:
        import ctypes as C
        v = C.c_int(0)
        check(_lib.lib().hyp_ctx_status(self._h, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().hyp_ctx_destroy(self._h)
            self._h.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_topk(local_idx: torch.Tensor, local_d: torch.Tensor, n: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather per-shard (nrows_r, k) results into the full (n, k) lists on every rank with ONE collective over
    interleaved {idx, d} records.  Library path (NCCL on CUDA tensors, gloo on CPU tensors for the host-logic tests);
    the peer-memory path is TopkContext."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    _, nrows, per = shard_rows(n, world, rank)
    k = local_idx.shape[1]
    assert local_idx.shape[0] == nrows and local_d.shape == local_idx.shape
    pad = join_records(torch.full((per, k), -1, dtype=torch.int32, device=local_idx.device),
                       torch.full((per, k), float("inf"), dtype=torch.float32, device=local_d.device))
    pad[:nrows] = join_records(local_idx, local_d)
    allr = torch.empty((world * per, k), dtype=torch.int64, device=local_idx.device)
    dist.all_gather_into_tensor(allr, pad, group=group)
    gi, gd = split_records(allr[:n])
    return gi.contiguous(), gd.contiguous()


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


_contexts = {}


def topk_context(n: int, k: int, device: torch.device, group=None) -> Optional[TopkContext]:
    """The cached peer-memory context for (n, k) on this device, or None when CUDA IPC between the ranks is not
    available (then the caller uses the NCCL all-gather).  Collective: every rank must call it."""
    key = (n, k, device.index, id(group))
    if key not in _contexts:
        try:
            _contexts[key] = TopkContext(n, k, device, group)
        except RuntimeError as e:
            import warnings
            warnings.warn(f"peer-memory all-gather unavailable ({e}); using the NCCL all-gather")
            _contexts[key] = None
    return _contexts[key]


def lorentz_topk_sharded(E: torch.Tensor, k: int = 32, c: float = 1.0, semantics: Optional[str] = None,
                         n: Optional[int] = None, group=None, engine: str = "exact",
                         exchange: str = "auto") -> Tuple[torch.Tensor, torch.Tensor]:
    """Row-sharded all-pairs top-k: local shard on this GPU, full lists everywhere.  `exchange`: "p2p" = the
    kernels write into every rank's gather buffer over NVLink (hyp_ctx), "nccl" = one all_gather_into_tensor of
    interleaved records, "auto" = p2p when the ranks can map each other's memory."""
    n = E.shape[0] if n is None else n
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return lorentz_topk(E, k, c, semantics, n, engine=engine)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    row0, nrows, _ = shard_rows(n, world, rank)
    ctx = topk_context(n, k, E.device, group) if exchange in ("auto", "p2p") and E.is_cuda else None
    if exchange == "p2p" and ctx is None:
        raise RuntimeError("exchange='p2p' requested but the ranks cannot map each other's memory")
    if ctx is not None:
        if engine == "tc":
            rec = ctx.gram_topk(E, c, semantics)
        else:
            li, ld = lorentz_topk(E, k, c, semantics, n, row0, nrows, engine=engine)
            rec = ctx.allgather(li, ld)
        gi, gd = split_records(rec[:n])
        return gi.contiguous(), gd.contiguous()
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
