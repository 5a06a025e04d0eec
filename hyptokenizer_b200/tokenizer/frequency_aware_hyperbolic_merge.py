"""Drop-in for the reference's ``tokenizer/frequency_aware_hyperbolic_merge.py`` on B200.

Device work: pair counting over the corpus byte stream (K6), the candidate list (K2), and for the
whole candidate list at once the un-projected midpoints and their distances to the sampled rows
(K7).  Host work, kept exactly as in the reference because it is Python-float arithmetic on a few
numbers per candidate: the torch.randperm draws (same global CPU generator, same order), the
float64 mean / sigmoid, the log1p frequency score and the stable sort on -score.
"""
from __future__ import annotations

import json
import logging
import time
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from .._lib import SEM, check, ptr, stream_ptr
from ..embedding import lorentz_model as LM
from ..pair_count import count_pairs, count_pairs_sharded
from .hyperbolic_merge import HyperbolicTokenizer

logger = logging.getLogger(__name__)


class FrequencyAwareHyperbolicTokenizer(HyperbolicTokenizer):
    """reference frequency_aware_hyperbolic_merge.py:29-396."""

    # Set on the class or an instance when every rank of an initialised torch.distributed group constructs the
    # tokenizer over the SAME corpus: each rank then counts a line-aligned byte range and the histograms are summed
    # (pair_count.count_pairs_sharded).  Off by default: replicas over different corpora must not be mixed.
    shard_pair_counts = False

    def __init__(self, vocab: List[str], embeddings: torch.nn.Parameter, corpus_path: Optional[str] = None,
                 alpha: float = 0.4, beta: float = 0.4, gamma: float = 0.2, curvature: float = 1.0,
                 merge_threshold: float = 1.0, lr: float = 1e-3, device: Optional[torch.device] = None,
                 max_vocab_size: int = 100000, use_approximate_search: bool = True,
                 semantics: Optional[str] = None):
        super().__init__(vocab=vocab, embeddings=embeddings, curvature=curvature, merge_threshold=merge_threshold,
                         lr=lr, device=device, max_vocab_size=max_vocab_size,
                         use_approximate_search=use_approximate_search, semantics=semantics)
        self.alpha = alpha
        self.beta = beta
        self.gamma = gamma
        self.pair_frequencies: Dict[Tuple[str, str], int] = {}
        if corpus_path:
            self._compute_pair_frequencies(corpus_path)

    # ---- K6 ------------------------------------------------------------------------------------------
    def _compute_pair_frequencies(self, corpus_path: str) -> None:
        """reference :92-112.  The reference tokenizes every line with `self.tokenize`; until merges
        exist (always the case in __init__, and for the life of the object once `_merge_rules` has
        been built empty) that is `list(line.strip())`, which the device kernel
        counts directly from the bytes.  With non-empty merge rules the host path is used."""
        if getattr(self, "_merge_rules", None) or (not hasattr(self, "_merge_rules") and self.merge_history):
            return self._compute_pair_frequencies_host(corpus_path)
        if not hasattr(self, "_merge_rules"):
            self._merge_rules = {}          # what the first self.tokenize() call would have done (:425-428)
        with open(corpus_path, "rb") as f:
            data = f.read()
        counts = (count_pairs_sharded(data, device=self.device) if self.shard_pair_counts
                  else count_pairs(data, self.device))
        total = 0
        for pair, cnt in counts.items():
            self.pair_frequencies[pair] = self.pair_frequencies.get(pair, 0) + cnt
            total += cnt
        logger.info(f"Computed frequencies for {len(self.pair_frequencies)} unique token pairs "
                    f"from {total} total pairs")

    def _compute_pair_frequencies_host(self, corpus_path: str) -> None:
        with open(corpus_path, "r", encoding="utf-8") as f:
            for line in f:
                tokens = self.tokenize(line.strip())
                for k in range(len(tokens) - 1):
                    pair = (tokens[k], tokens[k + 1])
                    self.pair_frequencies[pair] = self.pair_frequencies.get(pair, 0) + 1

    # ---- K7 ------------------------------------------------------------------------------------------
    def _coherence_batch(self, cands: List[Tuple[int, int, float]]) -> List[float]:
        """reference :114-166 for every candidate, in candidate order (RNG consumption preserved)."""
        n = self.current_vocab_size
        C = len(cands)
        S = min(50, n)
        E = self._table()
        samples = torch.empty((C, S), dtype=torch.int64)
        for c in range(C):
            samples[c] = torch.randperm(n)[:S]                      # :144-145, global CPU generator
        ii = torch.tensor([c[0] for c in cands], dtype=torch.int32)
        jj = torch.tensor([c[1] for c in cands], dtype=torch.int32)
        li = torch.tensor([len(self.vocab[c[0]]) for c in cands], dtype=torch.int32)
        lj = torch.tensor([len(self.vocab[c[1]]) for c in cands], dtype=torch.int32)
        dev = E.device
        d_ii, d_jj, d_li, d_lj = (t.to(dev) for t in (ii, jj, li, lj))
        d_samples = samples.to(torch.int32).to(dev)
        out = torch.empty((C, S), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().hyp_coherence_distances(ptr(E), E.stride(0), ptr(d_ii), ptr(d_jj), ptr(d_li), ptr(d_lj),
                                                     ptr(d_samples), S, ptr(out), C, E.shape[1],
                                                     LM._curv(self.curvature), SEM[self.semantics], stream_ptr()))
        dist = out.cpu().numpy()
        samp = samples.numpy()
        keep = (samp != ii.numpy()[:, None].astype(np.int64)) & (samp != jj.numpy()[:, None].astype(np.int64))
        coh: List[float] = []
        thr = self.merge_threshold
        with np.errstate(over="ignore", invalid="ignore"):
            for c in range(C):
                vals = dist[c][keep[c]].tolist()                      # Python floats, as `.item()` gives (:152-153)
                if not vals:
                    coh.append(0.0)
                    continue
                avg = np.mean(vals)
                coh.append(1.0 / (1.0 + np.exp(avg - thr)))
        return coh

    def _compute_semantic_coherence(self, i: int, j: int) -> float:
        """reference :114-166 (single candidate)."""
        return self._coherence_batch([(i, j, 0.0)])[0]

    def _score_merge_candidate(self, i: int, j: int, dist: float) -> float:
        """reference :168-199."""
        return self._score_batch([(i, j, dist)])[0]

    def _score_batch(self, cands: List[Tuple[int, int, float]]) -> List[float]:
        coh = self._coherence_batch(cands)
        max_freq = max(self.pair_frequencies.values()) if self.pair_frequencies else 1
        scores = []
        for (i, j, dist), semantic_score in zip(cands, coh):
            dist_score = 1.0 / (1.0 + dist)
            pair_freq = self.pair_frequencies.get((self.vocab[i], self.vocab[j]), 0)
            freq_score = np.log1p(pair_freq)
            freq_score = freq_score / np.log1p(max_freq) if max_freq > 0 else 0
            scores.append(self.alpha * dist_score + self.beta * freq_score + self.gamma * semantic_score)
        return scores

    def _score_batch_device(self, cands: List[Tuple[int, int, float]]) -> List[float]:
        """`_score_batch` with the float64 part on the device too (hyp_score_candidates): one launch computes the
        coherence distances, their numpy-order mean, the sigmoid and the weighted score.  Same RNG consumption; the
        scores agree with the host path to the last bits of `exp` (not used by default: the host path IS numpy)."""
        n = self.current_vocab_size
        C = len(cands)
        if C == 0:
            return []
        S = min(50, n)
        E = self._table()
        samples = torch.empty((C, S), dtype=torch.int64)
        for c in range(C):
            samples[c] = torch.randperm(n)[:S]
        max_freq = max(self.pair_frequencies.values()) if self.pair_frequencies else 1
        fs = np.empty(C, dtype=np.float64)
        for c, (i, j, _) in enumerate(cands):
            f = np.log1p(self.pair_frequencies.get((self.vocab[i], self.vocab[j]), 0))
            fs[c] = f / np.log1p(max_freq) if max_freq > 0 else 0
        dev = E.device
        as_i32 = lambda v: torch.tensor(v, dtype=torch.int32).to(dev)
        d_ii, d_jj = as_i32([c[0] for c in cands]), as_i32([c[1] for c in cands])
        d_li, d_lj = as_i32([len(self.vocab[c[0]]) for c in cands]), as_i32([len(self.vocab[c[1]]) for c in cands])
        d_dist = torch.tensor([c[2] for c in cands], dtype=torch.float32).to(dev)
        d_fs = torch.from_numpy(fs).to(dev)
        d_samples = samples.to(torch.int32).to(dev)
        out = torch.empty(C, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().hyp_score_candidates(ptr(E), E.stride(0), ptr(d_ii), ptr(d_jj), ptr(d_li), ptr(d_lj),
                                                  ptr(d_samples), S, ptr(d_dist), ptr(d_fs), float(self.alpha),
                                                  float(self.beta), float(self.gamma), float(self.merge_threshold),
                                                  ptr(out), None, C, E.shape[1], LM._curv(self.curvature),
                                                  SEM[self.semantics], stream_ptr()))
        return out.cpu().tolist()

    def _find_merge_candidates(self) -> List[Tuple[int, int, float]]:
        """reference :201-234: (i, j, -score), ascending (stable)."""
        candidates = super()._find_merge_candidates()
        if not candidates or (self.beta > 0 and not self.pair_frequencies):
            return candidates
        scores = self._score_batch(candidates)
        scored = [(i, j, -s) for (i, j, _), s in zip(candidates, scores)]
        scored.sort(key=lambda x: x[2])
        return scored

    def optimize_merges(self, steps: int = 10000, log_every: int = 1000, parallel_eval: bool = True,
                        sample_ratio: float = 1.0, corpus_path: Optional[str] = None) -> None:
        """reference :236-313."""
        if corpus_path:
            self._compute_pair_frequencies(corpus_path)
        no_candidate_count = 0
        self.last_trace = []
        for step in range(steps):
            candidates = self._find_merge_candidates()
            if step % log_every == 0:
                logger.info(f"Step {step}: vocab_size={self.current_vocab_size}")
                logger.info(f"  Merge candidates: {len(candidates)}")
                logger.info(f"  Merge threshold: {self.merge_threshold:.6f}")
            if not candidates:
                no_candidate_count += 1
                if no_candidate_count > 5:
                    self.merge_threshold *= 1.5
                    no_candidate_count = 0
                    continue
                elif no_candidate_count > 10:
                    break
                continue
            else:
                no_candidate_count = 0
            i, j, score = candidates[0]
            self.last_trace.append((i, j, score))
            self._merge_tokens(i, j)
            if step > 0 and step % 1000 == 0:
                self.merge_threshold *= 1.1

    # ---- persistence ---------------------------------------------------------------------------------------
    def save(self, path: str) -> None:
        """reference :315-342 (including the "a|b" key format)."""
        super().save(path)
        frequencies_json = {f"{k[0]}|{k[1]}": v for k, v in self.pair_frequencies.items()}
        with open(f"{path}/frequencies.json", "w") as f:
            json.dump(frequencies_json, f)
        with open(f"{path}/freq_hyperparams.json", "w") as f:
            json.dump({"alpha": self.alpha, "beta": self.beta, "gamma": self.gamma}, f)

    @classmethod
    def load(cls, path: str, device: Optional[torch.device] = None) -> "FrequencyAwareHyperbolicTokenizer":
        """reference :344-396."""
        tokenizer = super().load(path, device)
        try:
            with open(f"{path}/freq_hyperparams.json", "r") as f:
                hp = json.load(f)
                tokenizer.alpha = hp.get("alpha", 0.4)
                tokenizer.beta = hp.get("beta", 0.4)
                tokenizer.gamma = hp.get("gamma", 0.2)
        except FileNotFoundError:
            logger.warning("Hyperparameters file not found, using defaults")
        try:
            with open(f"{path}/frequencies.json", "r") as f:
                tokenizer.pair_frequencies = {tuple(k.split("|")): v for k, v in json.load(f).items()}
        except FileNotFoundError:
            logger.warning("Frequencies file not found")
        return tokenizer
