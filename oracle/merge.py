"""Oracle restatement of the reference's merge loops (CPU, torch eager fp32 + numpy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows, step for step:

* tokenizer/hyperbolic_merge.py      HyperbolicTokenizer  (:96-412)
* scripts/train_hyperbolic_tokenizer.py:236-283  (the loop that honours target_vocab_size)
* tokenizer/fast_hyperbolic_merge.py  AdaptiveMergeCache (:63-133), FastHyperbolicTokenizer (:136-576)
* tokenizer/frequency_aware_hyperbolic_merge.py  (:92-313)

The algorithm is the reference's (full all-pairs recompute every step, strict `<`
threshold, stable sort on distance => argmin over (d, i, j)).  The one deliberate
difference is mechanical: candidate tuples are extracted with vectorised numpy
instead of a Python loop of `.item()` calls (hyperbolic_merge.py:266-269), which
yields the same list in the same order, only faster -- so as a CPU baseline this
port FLATTERS the reference.  The n<=100 per-pair regime (:270-289) is folded into
the batch regime; both produce bit-identical distances,
which tests/test_oracle_golden.py re-checks against the real reference.
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import lorentz as L


class OracleTokenizer:
    """HyperbolicTokenizer (tokenizer/hyperbolic_merge.py:96-412) restated."""

    def __init__(self, vocab: List[str], embeddings: torch.Tensor, curvature: float = 1.0,
                 merge_threshold: float = 0.1, max_vocab_size: int = 100000,
                 semantics: str = "reference"):
        self.semantics = semantics
        self.vocab = list(vocab)
        self.n = len(vocab)
        self.max_vocab_size = max_vocab_size
        self.curvature = curvature
        self.merge_threshold = merge_threshold
        emb = embeddings.detach().to(torch.float32)
        self.E = torch.zeros((max_vocab_size, emb.shape[1]), dtype=torch.float32)   # :146-152
        self.E[: self.n] = emb
        self.token2idx = {t: k for k, t in enumerate(self.vocab)}
        self.merge_history: List[Tuple[str, str, str]] = []
        self.trace: List[Tuple[int, int, float]] = []       # (i, j, d) of every merge
        self.n_candidates: List[int] = []

    # ---- candidate search: hyperbolic_merge.py:192-291 -------------------------------
    def all_distances(self) -> torch.Tensor:
        a = self.E[: self.n]
        return L.batch_distance(a, a, self.curvature, self.semantics)

    def find_candidates(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """All (i, j, d) with i < j and d < threshold, row-major (:259-269)."""
        n = self.n
        dist = self.all_distances()
        keep = (dist < self.merge_threshold) & torch.triu(torch.ones(n, n, dtype=torch.bool), diagonal=1)
        ii, jj = keep.nonzero(as_tuple=True)
        return ii.numpy(), jj.numpy(), dist[ii, jj].numpy()

    @staticmethod
    def pick(ii, jj, dd) -> Optional[Tuple[int, int, float]]:
        """candidates.sort(key=d) is stable (:378) => first row-major entry of min d."""
        if len(dd) == 0:
            return None
        k = int(np.argmin(dd))          # first occurrence == stable-sort head; NaN never enters
        return int(ii[k]), int(jj[k]), float(dd[k])

    # ---- merge: hyperbolic_merge.py:309-355 ------------------------------------------
    def merge_tokens(self, i: int, j: int) -> None:
        ti, tj = self.vocab[i], self.vocab[j]
        row = L.midpoint(self.E[i], self.E[j], len(ti), len(tj), self.curvature, self.semantics)
        if self.n >= self.max_vocab_size:
            raise ValueError(f"Maximum vocabulary size {self.max_vocab_size} reached. Cannot merge more tokens.")
        self.vocab.append(ti + tj)
        self.token2idx[ti + tj] = self.n
        self.E[self.n] = row
        self.n += 1
        self.merge_history.append((ti, tj, ti + tj))

    # ---- loops -------------------------------------------------------------------------
    def optimize_merges(self, steps: int) -> None:
        """hyperbolic_merge.py:357-412 (CPU branch: no parallel eval, sample_ratio 1)."""
        for _ in range(steps):
            ii, jj, dd = self.find_candidates()
            best = self.pick(ii, jj, dd)
            if best is None:
                break
            self.n_candidates.append(len(dd))
            self.trace.append(best)
            self.merge_tokens(best[0], best[1])

    def optimize_script_loop(self, merge_steps: int, target_vocab_size: Optional[int] = None) -> None:
        """scripts/train_hyperbolic_tokenizer.py:236-283 without the logging callback."""
        for step in range(merge_steps):
            if target_vocab_size is not None and len(self.vocab) >= target_vocab_size:
                break
            ii, jj, dd = self.find_candidates()
            best = self.pick(ii, jj, dd)
            if best is None:
                break
            self.n_candidates.append(len(dd))
            self.trace.append(best)
            self.merge_tokens(best[0], best[1])
            if step > 0 and step % 1000 == 0:
                self.merge_threshold *= 1.05

    # ---- tokenize / encode / decode: hyperbolic_merge.py:414-471 ------------------------
    def tokenize(self, text: str) -> List[str]:
        if not hasattr(self, "_merge_rules"):
            self._merge_rules = {(a, b): m for a, b, m in self.merge_history}
        toks = list(text)
        again = True
        while again:
            again = False
            k = 0
            while k < len(toks) - 1:
                m = self._merge_rules.get((toks[k], toks[k + 1]))
                if m is not None:
                    toks[k] = m
                    del toks[k + 1]
                    again = True
                else:
                    k += 1
        return toks

    def encode(self, text: str) -> List[int]:
        unk = self.token2idx.get("<unk>", 3)
        return [self.token2idx.get(t, unk) for t in self.tokenize(text)]

    def decode(self, ids: List[int]) -> str:
        return "".join(self.vocab[k] for k in ids)


class OracleFastTokenizer(OracleTokenizer):
    """FastHyperbolicTokenizer without FAISS (tokenizer/fast_hyperbolic_merge.py:136-576).

    The cache is kept as parallel arrays sorted by distance only (MergeCandidate.__lt__,
    :59-60); `add_batch` = concat(old, new) -> stable sort -> truncate (:91-95);
    `get_best(100)` pops (:111-117)."""

    def __init__(self, *a, cache_size: int = 10000, **kw):
        super().__init__(*a, **kw)
        self.cache_size = cache_size
        self.c_d = np.empty(0, np.float32)
        self.c_i = np.empty(0, np.int64)
        self.c_j = np.empty(0, np.int64)
        self.hits = 0
        self.misses = 0
        self.refills: List[int] = []

    def _cache_get_best(self, k: int):
        if len(self.c_d) == 0:
            self.misses += 1
            return None
        out = (self.c_i[:k], self.c_j[:k], self.c_d[:k])
        self.hits += len(out[2])
        self.c_i, self.c_j, self.c_d = self.c_i[k:], self.c_j[k:], self.c_d[k:]
        return out

    def _cache_add_batch(self, ii, jj, dd) -> None:
        d = np.concatenate([self.c_d, dd.astype(np.float32)])
        i = np.concatenate([self.c_i, ii])
        j = np.concatenate([self.c_j, jj])
        order = np.argsort(d, kind="stable")[: self.cache_size]
        self.c_d, self.c_i, self.c_j = d[order], i[order], j[order]

    def find_candidates_fast(self):
        """:253-376, n>100 / n<=100 branches (identical values); returns sorted arrays."""
        got = self._cache_get_best(100)
        if got is not None:
            return got
        ii, jj, dd = self.find_candidates()
        order = np.argsort(dd, kind="stable")
        ii, jj, dd = ii[order], jj[order], dd[order]
        self._cache_add_batch(ii, jj, dd)
        return ii, jj, dd

    def distance_statistics(self, sample_size: int = 1000) -> Dict[str, float]:
        """:433-465 -- identical `random.sample` consumption, distances gathered in one batch."""
        n = self.n
        k = min(sample_size, n * (n - 1) // 2)
        pairs = [random.sample(range(n), 2) for _ in range(k)]
        if not pairs:
            return {"min": 0.0, "max": 0.0, "mean": 0.0, "std": 0.0}
        a = torch.tensor([p[0] for p in pairs])
        b = torch.tensor([p[1] for p in pairs])
        d = L.distance(self.E[a], self.E[b], self.curvature, self.semantics).tolist()
        return {"min": min(d), "max": max(d), "mean": np.mean(d), "std": np.std(d)}

    def optimize_merges(self, steps: int, log_every: int = 1000, adaptive_threshold: bool = True) -> None:
        """:467-576."""
        misses_in_a_row = 0
        if adaptive_threshold:
            st = self.distance_statistics()
            if st["max"] < 1e-6:
                self.merge_threshold = 1e-5
            if st["max"] > 0 and self.merge_threshold > st["max"]:
                self.merge_threshold = min(self.merge_threshold, st["mean"] * 1.5)
        for step in range(steps):
            was_empty = len(self.c_d) == 0
            ii, jj, dd = self.find_candidates_fast()
            if was_empty:
                self.refills.append(step)
            if step % log_every == 0 or len(dd) == 0:
                self.distance_statistics()
            if len(dd) == 0:
                misses_in_a_row += 1
                if misses_in_a_row > 5 and adaptive_threshold:
                    self.merge_threshold *= 1.5
                    misses_in_a_row = 0
                    continue
                elif misses_in_a_row > 10:
                    break
                continue
            misses_in_a_row = 0
            best = (int(ii[0]), int(jj[0]), float(dd[0]))
            self.trace.append(best)
            self.merge_tokens(best[0], best[1])
            if adaptive_threshold and step > 0 and step % 1000 == 0:
                self.merge_threshold *= 1.1


def count_pairs_py(lines) -> Dict[Tuple[str, str], int]:
    """tokenizer/frequency_aware_hyperbolic_merge.py:92-112 with `_merge_rules == {}`
    (always the case inside __init__): adjacent code-point pairs of
    `line.strip()`, never across lines."""
    freq: Dict[Tuple[str, str], int] = {}
    for line in lines:
        toks = list(line.strip())
        for k in range(len(toks) - 1):
            key = (toks[k], toks[k + 1])
            freq[key] = freq.get(key, 0) + 1
    return freq


class OracleFrequencyAwareTokenizer(OracleTokenizer):
    """FrequencyAwareHyperbolicTokenizer (tokenizer/frequency_aware_hyperbolic_merge.py:29-313)."""

    def __init__(self, vocab, embeddings, corpus_lines=None, alpha=0.4, beta=0.4, gamma=0.2,
                 curvature=1.0, merge_threshold=1.0, max_vocab_size=100000, semantics="reference"):
        super().__init__(vocab, embeddings, curvature, merge_threshold, max_vocab_size, semantics)
        self.alpha, self.beta, self.gamma = alpha, beta, gamma
        self.pair_frequencies = count_pairs_py(corpus_lines) if corpus_lines is not None else {}

    def semantic_coherence(self, i: int, j: int) -> float:
        """:114-166 -- un-projected midpoint, torch.randperm(n)[:50] from the global generator."""
        ti, tj = self.vocab[i], self.vocab[j]
        merged = L.midpoint(self.E[i], self.E[j], len(ti), len(tj), self.curvature,
                            self.semantics, project=False)
        k = min(50, self.n)
        idx = torch.randperm(self.n)[:k]
        keep = [int(t) for t in idx if int(t) != i and int(t) != j]
        if not keep:
            return 0.0
        d = L.distance(merged, self.E[torch.tensor(keep)], self.curvature, self.semantics).tolist()
        avg = np.mean(d)
        return 1.0 / (1.0 + np.exp(avg - self.merge_threshold))

    def score(self, i: int, j: int, dist: float) -> float:
        """:168-199."""
        dist_score = 1.0 / (1.0 + dist)
        f = self.pair_frequencies.get((self.vocab[i], self.vocab[j]), 0)
        fs = np.log1p(f)
        fmax = max(self.pair_frequencies.values()) if self.pair_frequencies else 1
        fs = fs / np.log1p(fmax) if fmax > 0 else 0
        return self.alpha * dist_score + self.beta * fs + self.gamma * self.semantic_coherence(i, j)

    def find_scored(self) -> List[Tuple[int, int, float]]:
        """:201-234 -- returns (i, j, -score) sorted ascending (stable)."""
        ii, jj, dd = self.find_candidates()
        cands = [(int(a), int(b), float(c)) for a, b, c in zip(ii, jj, dd)]
        if not cands or (self.beta > 0 and not self.pair_frequencies):
            return cands
        scored = [(a, b, -self.score(a, b, c)) for a, b, c in cands]
        scored.sort(key=lambda t: t[2])
        return scored

    def optimize_merges(self, steps: int) -> None:
        """:236-313."""
        misses = 0
        for step in range(steps):
            cands = self.find_scored()
            if not cands:
                misses += 1
                if misses > 5:
                    self.merge_threshold *= 1.5
                    misses = 0
                    continue
                elif misses > 10:
                    break
                continue
            misses = 0
            i, j, s = cands[0]
            self.trace.append((i, j, s))
            self.merge_tokens(i, j)
            if step > 0 and step % 1000 == 0:
                self.merge_threshold *= 1.1
