"""Drop-in for the reference's ``tokenizer/compression_aware_tokenizer.py`` on B200.

A HOST policy over the device candidate list, like the hierarchical tokenizer: every step takes all pairs under the
threshold from the K2 min + emit kernels, re-scores the first `sample_size` of them by how much merging the pair would
shorten a greedy longest-match tokenization of a corpus sample (pure string work), sorts by the negated score and
merges the head (K1 midpoint).  The greedy tokenizer here looks the longest match up by length in per-length sets
instead of scanning the whole vocabulary at every position -- the same token boundaries (the reference scans the
vocabulary sorted by length, so the first hit is the longest match), which is all its caller uses: the count.
The reference's cache quirk is kept: counts are cached under `merge_{i}_{j}_{first 20 characters of the text}`, so
texts that share their first 20 characters share a count.
"""
from __future__ import annotations

import json
import logging
from typing import Dict, List, Optional, Tuple

import torch

from .hyperbolic_merge import HyperbolicTokenizer

logger = logging.getLogger(__name__)

Candidate = Tuple[int, int, float]


class CompressionAwareTokenizer(HyperbolicTokenizer):
    """reference compression_aware_tokenizer.py:28-340."""

    def __init__(self, vocab: List[str], embeddings: torch.nn.Parameter, corpus_sample: Optional[List[str]] = None,
                 compression_weight: float = 0.7, distance_weight: float = 0.3, sample_size: int = 100,
                 curvature: float = 1.0, merge_threshold: float = 0.1, lr: float = 1e-3,
                 device: Optional[torch.device] = None, max_vocab_size: int = 100000,
                 use_approximate_search: bool = True, semantics: Optional[str] = None):
        super().__init__(vocab=vocab, embeddings=embeddings, curvature=curvature, merge_threshold=merge_threshold,
                         lr=lr, device=device, max_vocab_size=max_vocab_size,
                         use_approximate_search=use_approximate_search, semantics=semantics)
        self.compression_weight = compression_weight
        self.distance_weight = distance_weight
        self.sample_size = sample_size
        self.corpus_sample = corpus_sample or []
        self.tokenize_cache: Dict[str, int] = {}

    # ---- greedy longest match (reference :91-122) ---------------------------------------------------------------
    @staticmethod
    def _length_index(vocab: List[str]) -> List[Tuple[int, set]]:
        by_len: Dict[int, set] = {}
        for tok in vocab:
            if tok:                                   # an empty token never advances the reference's scan either:
                by_len.setdefault(len(tok), set()).add(tok)   # it sorts last, behind every one-character token
        return sorted(by_len.items(), reverse=True)

    def _tokenize_with_vocab(self, text: str, vocab: List[str]) -> List[str]:
        index = self._length_index(vocab)
        tokens: List[str] = []
        i, n = 0, len(text)
        while i < n:
            for length, toks in index:
                piece = text[i:i + length]
                if len(piece) == length and piece in toks:
                    tokens.append(piece)
                    i += length
                    break
            else:
                tokens.append(text[i])                # no vocabulary entry starts here: the character itself
                i += 1
        return tokens

    # ---- scoring (reference :124-190) ------------------------------------------------------------------------------
    def _compression_aware_scoring(self, candidates: List[Candidate]) -> List[float]:
        if not self.corpus_sample:
            return [1.0 / (1.0 + dist) for _, _, dist in candidates]
        if "original" not in self.tokenize_cache:
            self.tokenize_cache["original"] = sum(len(self.tokenize(text)) for text in self.corpus_sample)
        original_tokens = self.tokenize_cache["original"]
        head = min(self.sample_size, len(candidates))
        scores: List[float] = []
        for i, j, dist in candidates[:head]:
            trial_vocab = self.vocab + [self.vocab[i] + self.vocab[j]]
            merged_tokens = 0
            for text in self.corpus_sample:
                key = f"merge_{i}_{j}_{text[:20]}"
                if key not in self.tokenize_cache:
                    self.tokenize_cache[key] = len(self._tokenize_with_vocab(text, trial_vocab))
                merged_tokens += self.tokenize_cache[key]
            ratio = 1.0 if merged_tokens == 0 else original_tokens / merged_tokens
            scores.append(self.compression_weight * ratio + self.distance_weight * (1.0 / (1.0 + dist)))
        scores.extend(1.0 / (1.0 + dist) for _, _, dist in candidates[head:])
        return scores

    def _find_merge_candidates(self) -> List[Candidate]:
        """reference :192-215: (i, j, -score), best first (stable sort)."""
        candidates = super()._find_merge_candidates()
        if not candidates:
            return []
        scores = self._compression_aware_scoring(candidates)
        scored = [(i, j, -score) for (i, j, _), score in zip(candidates, scores)]
        scored.sort(key=lambda c: c[2])
        return scored

    def optimize_merges(self, steps: int = 10000, log_every: int = 1000,
                        corpus_sample: Optional[List[str]] = None) -> None:
        """reference :217-276."""
        if corpus_sample:
            self.corpus_sample = corpus_sample
            self.tokenize_cache = {}
        for step in range(steps):
            candidates = self._find_merge_candidates()
            if not candidates:
                logger.info(f"No more merge candidates found after {step} steps")
                break
            i, j, neg_score = candidates[0]
            self._merge_tokens(i, j)
            for key in [k for k in self.tokenize_cache if k.startswith("merge_")]:
                self.tokenize_cache.pop(key)
            if (step + 1) % log_every == 0:
                logger.info(f"Step {step + 1}: merged '{self.vocab[i]}' + '{self.vocab[j]}' -> '{self.vocab[-1]}' "
                            f"(score: {-neg_score:.4f})")
            if step > 0 and step % 1000 == 0:
                self.merge_threshold *= 1.1

    # ---- persistence (reference :278-340) -------------------------------------------------------------------------
    def save(self, path: str) -> None:
        super().save(path)
        with open(f"{path}/compression_config.json", "w") as f:
            json.dump({"compression_weight": self.compression_weight, "distance_weight": self.distance_weight,
                       "sample_size": self.sample_size}, f)

    @classmethod
    def load(cls, path: str, device: Optional[torch.device] = None) -> "CompressionAwareTokenizer":
        tokenizer = super().load(path, device)
        try:
            with open(f"{path}/compression_config.json", "r") as f:
                cfg = json.load(f)
        except FileNotFoundError:
            logger.warning("Compression config file not found, using defaults")
            cfg = {}
        tokenizer.compression_weight = cfg.get("compression_weight", 0.7)
        tokenizer.distance_weight = cfg.get("distance_weight", 0.3)
        tokenizer.sample_size = cfg.get("sample_size", 100)
        return tokenizer
