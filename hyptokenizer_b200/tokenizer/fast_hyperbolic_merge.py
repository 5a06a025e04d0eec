"""Drop-in for the reference's ``tokenizer/fast_hyperbolic_merge.py`` on B200.

The reference makes the per-step all-pairs search affordable with a FAISS-HNSW index over Klein
coordinates (1000 sampled queries) and a 10 000-entry sorted Python cache that is popped 100 at a
time.  Here the search is EXACT and device resident: a running argmin over all pairs (the
"candidate heap" degenerates to its head because the reference never removes a pair) that is
updated by scoring each new row against the table.  That is the merge sequence of the
reference's own brute-force `HyperbolicTokenizer` -- the stated correctness target -- and is what
`cache_semantics="fresh"` (default) produces.  `cache_semantics="snapshot"` replays the shipped
class's stale pop-100 cache bit for bit for strict trace parity.
"""
from __future__ import annotations

import logging
import random
import time
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from .._lib import SEM, HypBest, check, ptr, stream_ptr
from ..embedding import lorentz_model as LM
from .hyperbolic_merge import HyperbolicTokenizer, _threshold_f32

logger = logging.getLogger(__name__)
FAISS_AVAILABLE = False


@dataclass
class MergeCandidate:
    """reference fast_hyperbolic_merge.py:52-60 (ordering on distance only)."""
    distance: float
    token_i: int
    token_j: int

    def __lt__(self, other):
        return self.distance < other.distance


class AdaptiveMergeCache:
    """reference fast_hyperbolic_merge.py:63-133.  Kept as parallel numpy arrays sorted by
    distance only (stable), which is what the reference's list of dataclasses amounts to."""

    def __init__(self, max_size: int = 10000):
        self.max_size = max_size
        self._d = np.empty(0, np.float32)
        self._i = np.empty(0, np.int64)
        self._j = np.empty(0, np.int64)
        self._hits = 0
        self.miss_count = 0

    @property
    def candidates(self) -> List[MergeCandidate]:
        return [MergeCandidate(float(d), int(i), int(j)) for d, i, j in zip(self._d, self._i, self._j)]

    def add_arrays(self, ii, jj, dd) -> None:
        d = np.concatenate([self._d, np.asarray(dd, np.float32)])
        i = np.concatenate([self._i, np.asarray(ii, np.int64)])
        j = np.concatenate([self._j, np.asarray(jj, np.int64)])
        order = np.argsort(d, kind="stable")[: self.max_size]
        self._d, self._i, self._j = d[order], i[order], j[order]

    def add_batch(self, new_candidates: List[MergeCandidate]) -> None:
        self.add_arrays([c.token_i for c in new_candidates], [c.token_j for c in new_candidates],
                        [c.distance for c in new_candidates])

    def pop_arrays(self, n: int):
        if len(self._d) == 0:
            self.miss_count += 1
            return None
        out = (self._i[:n], self._j[:n], self._d[:n])
        self._hits += len(out[2])
        self._i, self._j, self._d = self._i[n:], self._j[n:], self._d[n:]
        return out

    def get_best(self, n: int = 1) -> List[MergeCandidate]:
        got = self.pop_arrays(n)
        if got is None:
            return []
        return [MergeCandidate(float(d), int(i), int(j)) for i, j, d in zip(*got)]

    def get_stats(self) -> Dict[str, Any]:
        return {"size": len(self._d), "max_size": self.max_size, "hit_count": self._hits,
                "miss_count": self.miss_count,
                "hit_ratio": self._hits / (self._hits + self.miss_count + 1e-10)}


class FastHyperbolicTokenizer(HyperbolicTokenizer):
    """reference fast_hyperbolic_merge.py:136-576."""

    def __init__(self, vocab: List[str], embeddings: torch.nn.Parameter, curvature: float = 1.0,
                 merge_threshold: float = 0.1, lr: float = 1e-3, device: Optional[torch.device] = None,
                 max_vocab_size: int = 100000, use_approximate_search: bool = True, cache_size: int = 10000,
                 rebuild_frequency: int = 100, hnsw_m: int = 32, hnsw_ef_construction: int = 200,
                 hnsw_ef_search: int = 100, semantics: Optional[str] = None, cache_semantics: str = "fresh"):
        super().__init__(vocab=vocab, embeddings=embeddings, curvature=curvature, merge_threshold=merge_threshold,
                         lr=lr, device=device, max_vocab_size=max_vocab_size,
                         use_approximate_search=use_approximate_search, semantics=semantics)
        if cache_semantics not in ("fresh", "snapshot"):
            raise ValueError("cache_semantics must be 'fresh' or 'snapshot'")
        self.cache_semantics = cache_semantics
        self.index = None
        self.index_outdated = True
        self.cache = AdaptiveMergeCache(max_size=cache_size)
        self.rebuild_frequency = rebuild_frequency
        self.merges_since_rebuild = 0
        # accepted for signature compatibility; the HNSW index they configure is replaced by exact search
        self.hnsw_m = hnsw_m
        self.hnsw_ef_construction = hnsw_ef_construction
        self.hnsw_ef_search = hnsw_ef_search

    def _build_faiss_index(self) -> None:
        """reference :195-240.  Nothing to build: the device search is exact."""
        self.use_approximate_search = False

    # ---- candidate search ------------------------------------------------------------------------
    # _candidate_arrays: lists up to this many candidates are materialised and sorted whole; longer ones go through the
    # device radix select, which materialises at most _SELECT_EMIT_LIMIT + n of them
    _EMIT_ALL_LIMIT = 1 << 21
    _SELECT_EMIT_LIMIT = 1 << 20
    # subclasses that re-rank the WHOLE list of a refill by something else than distance (the enhanced tokenizer's
    # combined score, reference enhanced_fast_hyperbolic_merge.py:992-1013) cannot take the distance-ordered head of it
    _refill_needs_full_list = False

    def _sorted_candidates(self, oi, oj, od, n, keep=None):
        """(d, i, j) lexicographic == stable sort on d of the row-major list; first `keep` entries as numpy arrays."""
        order = torch.argsort(oi.to(torch.int64) * n + oj.to(torch.int64))
        od = od[order]
        o2 = torch.sort(od, stable=True).indices
        if keep is not None:
            o2 = o2[:keep]
        order = order[o2]
        return oi[order].cpu().numpy().astype(np.int64), oj[order].cpu().numpy().astype(np.int64), od[o2].cpu().numpy()

    def _candidate_topk(self, keep: int, thr: float):
        """The first `keep` candidates in (d, i, j) order of a list too long to materialise -- all AdaptiveMergeCache
        .add_batch keeps of it (reference :91-95) -- by a radix select on the device (include/hyptok_b200.h,
        hyp_allpairs_hist / _row_ties / _emit_cut): <= 3 histogram sweeps over the 31 bits of d narrow down the
        keep-th distance, a per-row count of the pairs AT that distance finds the row where the keep-th pair sits
        (ties are row-major), one sweep emits the <= keep + n survivors."""
        E, ws = self._table(), self._workspace()
        n, D = self.current_vocab_size, E.shape[1]
        L, c, sem = _lib.lib(), LM._curv(self.curvature), SEM[self.semantics]
        hist = torch.empty(1 << 12, dtype=torch.int64, device=E.device)
        prefix, plen, below = 0, 0, 0
        cut_bits, cut_row, n_emit = None, -1, 0
        with torch.cuda.device(E.device):
            for bin_bits in (12, 12, 7):
                check(L.hyp_allpairs_hist(ptr(E), E.stride(0), n, D, c, sem, thr, prefix, plen, bin_bits, ptr(hist),
                                          stream_ptr()))
                h = hist[: 1 << bin_bits].cpu().numpy()
                cum = below + np.cumsum(h)
                b = min(int(np.searchsorted(cum, keep, side="left")), len(h) - 1)
                below = int(cum[b - 1]) if b > 0 else below
                prefix, plen = (prefix << bin_bits) | b, plen + bin_bits
                if int(cum[b]) <= self._SELECT_EMIT_LIMIT:      # everything up to the end of this bin: few enough to sort
                    cut_bits, n_emit = (prefix + 1) << (31 - plen), int(cum[b])
                    break
            else:
                # `prefix` is the bit pattern of the keep-th distance: `below` pairs are strictly closer, the rest of
                # the list is decided among the pairs exactly AT it, in row-major order
                ties = torch.empty(n, dtype=torch.int32, device=E.device)
                check(L.hyp_allpairs_row_ties(ptr(E), E.stride(0), n, D, c, sem, thr, prefix, ptr(ties), stream_ptr()))
                cs = np.cumsum(ties.cpu().numpy().astype(np.int64))
                cut_row = min(int(np.searchsorted(cs, keep - below, side="left")), n - 1)
                cut_bits, n_emit = prefix, below + int(cs[cut_row])
            oi = torch.empty(n_emit, dtype=torch.int32, device=E.device)
            oj = torch.empty(n_emit, dtype=torch.int32, device=E.device)
            od = torch.empty(n_emit, dtype=torch.float32, device=E.device)
            check(L.hyp_allpairs_emit_cut(ptr(E), E.stride(0), n, D, c, sem, thr, cut_bits, cut_row, ptr(oi), ptr(oj),
                                          ptr(od), n_emit, ptr(ws["count"]), stream_ptr()))
            assert int(ws["count"].item()) == n_emit, "radix select: emitted count differs from the histogram's"
        return self._sorted_candidates(oi, oj, od, n, keep)

    def _candidate_arrays(self):
        """All (i, j, d) below the threshold, sorted by distance (stable over row-major order) -- or, when that list
        would not fit (`_last_candidate_total` > _EMIT_ALL_LIMIT), its first cache.max_size entries, which is all
        the cache keeps of it."""
        E, ws = self._table(), self._workspace()
        n, D = self.current_vocab_size, E.shape[1]
        thr = _threshold_f32(self.merge_threshold, n)
        head = self._global_best(thr)
        total = head.count_lo | (head.count_hi << 32)
        self._last_candidate_total = total
        if total > self._EMIT_ALL_LIMIT and total > self.cache.max_size and not self._refill_needs_full_list:
            return self._candidate_topk(self.cache.max_size, thr)
        if total > 200_000_000:
            raise NotImplementedError(f"this tokenizer scores every candidate of a refill and {total} of them would be "
                                      "materialised; lower merge_threshold")
        oi = torch.empty(total, dtype=torch.int32, device=E.device)
        oj = torch.empty(total, dtype=torch.int32, device=E.device)
        od = torch.empty(total, dtype=torch.float32, device=E.device)
        if total:
            with torch.cuda.device(E.device):
                check(_lib.lib().hyp_allpairs_emit(ptr(E), E.stride(0), n, D, LM._curv(self.curvature),
                                                   SEM[self.semantics], thr, ptr(oi), ptr(oj), ptr(od), total,
                                                   ptr(ws["count"]), stream_ptr()))
        return self._sorted_candidates(oi, oj, od, n)

    def _find_merge_candidates_fast_arrays(self):
        """reference :253-376 (cache first, else full search + cache.add_batch)."""
        got = self.cache.pop_arrays(100)
        if got is not None:
            self._last_candidate_total = len(got[2])
            return got
        ii, jj, dd = self._candidate_arrays()
        self.cache.add_arrays(ii, jj, dd)
        return ii, jj, dd

    def _find_merge_candidates_fast(self) -> List[MergeCandidate]:
        ii, jj, dd = self._find_merge_candidates_fast_arrays()
        return [MergeCandidate(float(d), int(i), int(j)) for i, j, d in zip(ii, jj, dd)]

    def _find_merge_candidates(self) -> List[Tuple[int, int, float]]:
        """reference :242-251."""
        return [(c.token_i, c.token_j, c.distance) for c in self._find_merge_candidates_fast()]

    def _merge_tokens(self, i: int, j: int) -> None:
        """reference :378-392."""
        super()._merge_tokens(i, j)
        self.merges_since_rebuild += 1
        if self.merges_since_rebuild >= self.rebuild_frequency:
            self.index_outdated = True

    # ---- statistics ----------------------------------------------------------------------------------
    def _compute_distance_statistics(self, sample_size: int = 1000) -> Dict[str, float]:
        """reference :433-465: same `random.sample` draws on the host, one batched exact re-score."""
        n = self.current_vocab_size
        k = min(sample_size, n * (n - 1) // 2)
        pairs = [random.sample(range(n), 2) for _ in range(k)]
        if not pairs:
            return {"min": 0.0, "max": 0.0, "mean": 0.0, "std": 0.0}
        E = self._table()
        idx = torch.tensor(pairs, dtype=torch.int32).t().contiguous().to(E.device)
        out = torch.empty(k, dtype=torch.float32, device=E.device)
        with torch.cuda.device(E.device):
            check(_lib.lib().hyp_rescore_pairs(ptr(E), E.stride(0), idx[0].data_ptr(), idx[1].data_ptr(), ptr(out),
                                               None, k, E.shape[1], LM._curv(self.curvature), SEM[self.semantics],
                                               stream_ptr()))
        d = out.tolist()
        return {"min": min(d), "max": max(d), "mean": np.mean(d), "std": np.std(d)}

    # ---- loop --------------------------------------------------------------------------------------------
    def optimize_merges(self, steps: int = 10000, log_every: int = 1000, text_sample: Optional[List[str]] = None,
                        adaptive_threshold: bool = True) -> None:
        """reference :467-576."""
        self.stats = {"step": [], "vocab_size": [], "min_dist": [], "max_dist": [], "mean_dist": [],
                      "num_candidates": []}
        if adaptive_threshold:
            st = self._compute_distance_statistics()
            logger.info(f"Initial distance statistics: min={st['min']:.6f}, max={st['max']:.6f}, mean={st['mean']:.6f}")
            if st["max"] < 1e-6:
                logger.warning("WARNING: Maximum distance is near zero! This will prevent finding merge candidates.")
                self.merge_threshold = 1e-5
            if st["max"] > 0 and self.merge_threshold > st["max"]:
                self.merge_threshold = min(self.merge_threshold, st["mean"] * 1.5)
        if self.cache_semantics == "snapshot":
            return self._optimize_snapshot(steps, log_every, adaptive_threshold)
        return self._optimize_fresh(steps, log_every, adaptive_threshold)

    def _log_stats(self, step: int, n_candidates: int) -> None:
        st = self._compute_distance_statistics()
        s = self.stats
        s["step"].append(step)
        s["vocab_size"].append(self.current_vocab_size)
        s["min_dist"].append(st["min"])
        s["max_dist"].append(st["max"])
        s["mean_dist"].append(st["mean"])
        s["num_candidates"].append(n_candidates)

    def _optimize_fresh(self, steps: int, log_every: int, adaptive_threshold: bool) -> None:
        """Exact, always-fresh search: device loop in segments that end where the reference's loop
        does host-visible work (statistics every `log_every` steps, empty-candidate back-off)."""
        traces = []
        step = 0
        misses = 0
        best: Optional[HypBest] = None
        every, mul = (1000, 1.1) if adaptive_threshold else (0, 1.0)
        while step < steps:
            if step % log_every == 0:
                self._log_stats(step, -1)
            seg = min(steps, (step // log_every + 1) * log_every) - step
            rec, stop = self._device_loop(seg, step, every, mul, best)
            traces.append(rec)
            last = self._last_state
            best = HypBest(d=last.best_d, i=last.best_i, j=last.best_j)
            step += len(rec)
            self.merges_since_rebuild += len(rec)
            if len(rec):
                misses = 0
            if stop == 2:
                self.last_trace = np.concatenate(traces)
                raise ValueError(f"Maximum vocabulary size {self.max_vocab_size} reached. Cannot merge more tokens.")
            if stop == 1 and step < steps:
                # reference :529-541: an empty step costs a step and a statistics call
                if step % log_every != 0:
                    self._log_stats(step, 0)
                misses += 1
                step += 1
                if misses > 5 and adaptive_threshold:
                    self.merge_threshold *= 1.5
                    misses = 0
                elif misses > 10:
                    break
        self.last_trace = np.concatenate(traces) if traces else np.empty(0)

    def _optimize_snapshot(self, steps: int, log_every: int, adaptive_threshold: bool) -> None:
        """The shipped control flow, step for step (stale pop-100 cache included)."""
        misses = 0
        trace = []
        for step in range(steps):
            ii, jj, dd = self._find_merge_candidates_fast_arrays()
            if step % log_every == 0 or len(dd) == 0:
                self._log_stats(step, self._last_candidate_total)      # the reference logs the length of the FULL list
            if len(dd) == 0:
                misses += 1
                if misses > 5 and adaptive_threshold:
                    self.merge_threshold *= 1.5
                    misses = 0
                    continue
                elif misses > 10:
                    break
                continue
            misses = 0
            trace.append((int(ii[0]), int(jj[0]), float(dd[0])))
            self._merge_tokens(int(ii[0]), int(jj[0]))
            if adaptive_threshold and step > 0 and step % 1000 == 0:
                self.merge_threshold *= 1.1
        self.last_trace = trace
