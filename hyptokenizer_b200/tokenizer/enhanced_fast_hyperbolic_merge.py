"""Drop-in for the reference's ``tokenizer/enhanced_fast_hyperbolic_merge.py`` on B200 (BASELINE config 5).

`EnhancedFastHyperbolicTokenizer` is a HOST policy over the device kernels, like the hierarchical and the
compression-aware tokenizers.  Per step the candidate list comes from the Fast class's search (K2 min + emit, through
its pop-100 cache: `cache_semantics="snapshot"` is what the shipped class does, `"fresh"` searches every step), and
every candidate gets a combined score (reference :903-990):

* distance score `1 / (1 + d)`;
* frequency score from the pair counts of the corpus (K6 pair counting over the byte stream);
* semantic coherence: un-projected midpoint against <= 50 `torch.randperm` rows, for all candidates of a step in ONE
  launch (K7), the draws consumed from the global CPU generator in the reference's order;
* compression score: greedy longest-match token count of a corpus sample with the merged token added (string work);
* morphology score of the current phase (string work; phase thresholds 0.05 / 0.1 / 0.2, transitions at steps
  {2: 1000, 3: 6000} by default).

The list is ordered with Python's own sort over the negated score, so a NaN score (every coherence is NaN in the shipped
arithmetic) leaves the order exactly as the reference's `list.sort()` leaves it.

Adaptive curvature.  As shipped, the curvature step cannot run: `distance` detaches `c` (`torch.tensor(c)`,
reference lorentz_model.py:137), so `loss.backward()` raises at the first step that is a multiple of
`optimize_curvature_freq` (probed).  `semantics="reference"` keeps that behaviour, error included.  In
`semantics="lorentz"` the step is the corrected one of `adaptive_curvature_tokenizer.CurvatureStepMixin` (SURVEY.md
8f-4), shared with `AdaptiveCurvatureTokenizer`.
"""
from __future__ import annotations

import json
import logging
import os
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Set, Tuple, Union

import numpy as np
import torch

from ..pair_count import count_pairs
from .adaptive_curvature_tokenizer import CurvatureStepMixin
from .fast_hyperbolic_merge import AdaptiveMergeCache, FastHyperbolicTokenizer, MergeCandidate
from .frequency_aware_hyperbolic_merge import FrequencyAwareHyperbolicTokenizer as _FreqAware
from .hierarchical_hyperbolic_merge import NLTK_AVAILABLE, HierarchicalHyperbolicTokenizer as _Hier  # noqa: F401

logger = logging.getLogger(__name__)

__all__ = ["EnhancedMergeCandidate", "EnhancedFastHyperbolicTokenizer", "MergeCandidate", "AdaptiveMergeCache"]


@dataclass
class EnhancedMergeCandidate(MergeCandidate):
    """reference enhanced_fast_hyperbolic_merge.py:52-63 (ordering on the negated combined score)."""
    frequency_score: float = 0.0
    semantic_score: float = 0.0
    compression_score: float = 0.0
    morphology_score: float = 0.0
    combined_score: float = 0.0

    def __lt__(self, other):
        return self.combined_score < other.combined_score


class _LengthIndex:
    """Greedy longest-match tokenization (reference :813-847) by per-length sets.  The reference scans the vocabulary
    sorted by length at every position, so the first hit is the longest match; only the token COUNT is used."""

    def __init__(self, vocab: List[str]):
        by_len: Dict[int, Set[str]] = {}
        for tok in vocab:
            if tok:                     # an empty token sorts last and never advances the reference's scan either
                by_len.setdefault(len(tok), set()).add(tok)
        self.by_len = by_len
        self.lengths = sorted(by_len, reverse=True)

    def tokenize(self, text: str, extra: str = "") -> List[str]:
        """Greedy longest-match tokens of `text` under vocabulary + [extra]."""
        by_len = self.by_len
        lengths = self.lengths
        le = len(extra)
        if le and le not in by_len:
            lengths = sorted(set(lengths) | {le}, reverse=True)
        tokens: List[str] = []
        i, n = 0, len(text)
        while i < n:
            for length in lengths:
                piece = text[i:i + length]
                if len(piece) == length and ((length == le and piece == extra) or piece in by_len.get(length, ())):
                    break
            else:
                piece = text[i]         # no vocabulary entry starts here: the character itself
            tokens.append(piece)
            i += len(piece)
        return tokens

    def count(self, text: str, extra: str) -> int:
        """Number of tokens of `text` under vocabulary + [extra].  A token that does not occur in the text cannot be
        matched anywhere in it, so the count is the one without it -- computed once per text and index (at size this is
        the common case: almost no candidate's merged string occurs in the ten sample lines, and re-tokenising them for
        each of 10^5 candidates of a cache refill took minutes of Python)."""
        if extra and extra not in text:
            base = self.__dict__.setdefault("_base_counts", {})
            got = base.get(text)
            if got is None:
                got = base[text] = len(self.tokenize(text, ""))
            return got
        return len(self.tokenize(text, extra))


class EnhancedFastHyperbolicTokenizer(CurvatureStepMixin, FastHyperbolicTokenizer):
    """reference enhanced_fast_hyperbolic_merge.py:66-1427."""

    def __init__(self, vocab: List[str], embeddings: torch.nn.Parameter,
                 curvature: float = 1.0, merge_threshold: float = 0.5, lr: float = 1e-3,
                 device: Optional[torch.device] = None, max_vocab_size: int = 100000,
                 use_approximate_search: bool = True, cache_size: int = 10000, rebuild_frequency: int = 100,
                 hnsw_m: int = 32, hnsw_ef_construction: int = 200, hnsw_ef_search: int = 100,
                 use_frequency_aware: bool = True, use_hierarchical: bool = True,
                 use_adaptive_curvature: bool = True, use_compression_aware: bool = True,
                 corpus_path: Optional[str] = None, alpha: float = 0.4, beta: float = 0.4, gamma: float = 0.2,
                 language: str = "english",
                 curvature_lr: float = 0.01, hierarchy_weight: float = 1.0, distortion_weight: float = 0.1,
                 optimize_curvature_freq: int = 100,
                 corpus_sample: Optional[List[str]] = None, compression_weight: float = 0.7,
                 distance_weight: float = 0.3, sample_size: int = 100,
                 semantics: Optional[str] = None, cache_semantics: str = "snapshot"):
        super().__init__(vocab=vocab, embeddings=embeddings, curvature=curvature, merge_threshold=merge_threshold,
                         lr=lr, device=device, max_vocab_size=max_vocab_size,
                         use_approximate_search=use_approximate_search, cache_size=cache_size,
                         rebuild_frequency=rebuild_frequency, hnsw_m=hnsw_m,
                         hnsw_ef_construction=hnsw_ef_construction, hnsw_ef_search=hnsw_ef_search,
                         semantics=semantics, cache_semantics=cache_semantics)
        self.use_frequency_aware = use_frequency_aware
        self.use_hierarchical = use_hierarchical
        self.use_adaptive_curvature = use_adaptive_curvature
        self.use_compression_aware = use_compression_aware
        self.current_phase = 1

        have_corpus = bool(corpus_path) and os.path.exists(corpus_path)
        if self.use_frequency_aware:                                   # reference :196-203
            self.alpha, self.beta, self.gamma = alpha, beta, gamma
            self.pair_frequencies: Dict[Tuple[str, str], int] = {}
            if have_corpus:
                self._compute_pair_frequencies(corpus_path)
        if self.use_hierarchical:                                      # reference :205-222
            self.language = language
            self.token_frequencies: Dict[str, int] = {}
            self.common_morphemes: Set[str] = set()
            self.common_words: Set[str] = set()
            if have_corpus:
                self._compute_corpus_statistics(corpus_path)
        if self.use_adaptive_curvature:                                # reference :224-243
            self.static_curvature = self.curvature
            self.curvature = torch.nn.Parameter(torch.tensor(float(curvature), device=self.device))
            self.curvature_optimizer = torch.optim.Adam([self.curvature], lr=curvature_lr)
            self.hierarchy_weight = hierarchy_weight
            self.distortion_weight = distortion_weight
            self.optimize_curvature_freq = optimize_curvature_freq
            self.merge_pairs: List[Tuple[int, int]] = []
            self._project_embeddings()
        if self.use_compression_aware:                                 # reference :245-254
            self.compression_weight = compression_weight
            self.distance_weight = distance_weight
            self.sample_size = sample_size
            self.corpus_sample = corpus_sample or []
            self.tokenize_cache: Dict[str, Any] = {}

    # ---- frequency-aware pieces (reference :266-372) ---------------------------------------------------------------
    def _compute_pair_frequencies(self, corpus_path: str) -> None:
        """reference :266-289: K6 over the byte stream while no merge rule exists (then `self.tokenize(line.strip())`
        is `list(line.strip())`), the reference's host loop otherwise."""
        if not self.use_frequency_aware:
            return
        if getattr(self, "_merge_rules", None) or (not hasattr(self, "_merge_rules") and self.merge_history):
            return _FreqAware._compute_pair_frequencies_host(self, corpus_path)
        if not hasattr(self, "_merge_rules"):
            self._merge_rules = {}          # what the first self.tokenize() call would have done
        with open(corpus_path, "rb") as f:
            data = f.read()
        for pair, cnt in count_pairs(data, self.device).items():
            self.pair_frequencies[pair] = self.pair_frequencies.get(pair, 0) + cnt

    def _coherence_batch(self, cands: List[Tuple[int, int, float]]) -> List[float]:
        """reference :291-346 for every candidate of a step, in order (one K7 launch)."""
        return _FreqAware._coherence_batch(self, cands)

    def _compute_semantic_coherence(self, i: int, j: int) -> float:
        if not self.use_frequency_aware:
            return 0.0
        return self._coherence_batch([(i, j, 0.0)])[0]

    def _compute_frequency_score(self, i: int, j: int, max_freq: Optional[int] = None) -> float:
        """reference :348-372."""
        if not self.use_frequency_aware or not self.pair_frequencies:
            return 0.0
        pair_freq = self.pair_frequencies.get((self.vocab[i], self.vocab[j]), 0)
        freq_score = np.log1p(pair_freq)
        if max_freq is None:
            max_freq = max(self.pair_frequencies.values())
        return freq_score / np.log1p(max_freq) if max_freq > 0 else 0

    def get_curvature(self) -> Union[float, torch.Tensor]:
        """reference :374-384."""
        if self.use_adaptive_curvature:
            return self.curvature
        return self.static_curvature if hasattr(self, "static_curvature") else self.curvature

    # ---- hierarchical pieces (reference :388-635) ------------------------------------------------------------------
    def _compute_corpus_statistics(self, corpus_path: str) -> None:
        if self.use_hierarchical:
            _Hier._compute_corpus_statistics(self, corpus_path)

    def _is_potential_morpheme(self, token: str) -> bool:
        return True if not self.use_hierarchical else _Hier._is_potential_morpheme(self, token)

    def _is_valid_word(self, token: str) -> bool:
        return True if not self.use_hierarchical else _Hier._is_valid_word(self, token)

    def _get_merge_phase_threshold(self) -> float:
        """reference :514-530."""
        if not self.use_hierarchical:
            return self.merge_threshold
        return {1: 0.05, 2: 0.1}.get(self.current_phase, 0.2)

    def _morphology_score(self, token_i: str, token_j: str) -> float:
        """reference :936-945."""
        if self.current_phase == 1:
            return 0.8 if len(token_i) <= 2 and len(token_j) <= 2 else 0.2
        if self.current_phase == 2:
            return 0.9 if self._is_potential_morpheme(token_i + token_j) else 0.3
        return 1.0 if self._is_valid_word(token_i + token_j) else 0.4

    def _filter_by_current_phase(self, candidates: List[MergeCandidate]) -> List[MergeCandidate]:
        """reference :532-633 (defined there, never called by the loop): every candidate is kept; the ones the phase
        favours get their distance scaled by 0.9 / 0.8 / 0.7."""
        if not self.use_hierarchical or not candidates:
            return candidates
        factor = {1: 0.9, 2: 0.8}.get(self.current_phase, 0.7)
        favoured = {1: 0.8, 2: 0.9}.get(self.current_phase, 1.0)
        out: List[MergeCandidate] = []
        for c in candidates:
            score = self._morphology_score(self.vocab[c.token_i], self.vocab[c.token_j])
            dist = c.distance * factor if score == favoured else c.distance
            out.append(EnhancedMergeCandidate(distance=dist, token_i=c.token_i, token_j=c.token_j,
                                              morphology_score=score))
        return out

    # ---- adaptive curvature (reference :637-811) ---------------------------------------------------------------------
    def _optimize_curvature(self, embeddings: Optional[torch.Tensor] = None) -> None:
        """reference :753-782.  `embeddings` is accepted for signature compatibility; the table is read in place."""
        if self.use_adaptive_curvature:
            self._curvature_step()

    def _project_embeddings(self) -> None:
        """reference :784-792 and :243: the whole `[max_vocab_size, D]` table (unused rows become the origin)."""
        if self.use_adaptive_curvature:
            self._project_table()

    def _merge_tokens(self, i: int, j: int) -> None:
        """reference :794-809."""
        if self.use_adaptive_curvature:
            self.merge_pairs.append((i, j))
        super()._merge_tokens(i, j)

    # ---- compression-aware pieces (reference :813-899) -----------------------------------------------------------------
    def _tokenize_with_vocab(self, text: str, vocab: List[str]) -> List[str]:
        """reference :813-847."""
        if not self.use_compression_aware:
            return self.tokenize(text)
        return _LengthIndex(vocab).tokenize(text)

    def _compute_compression_score(self, i: int, j: int, index: Optional[_LengthIndex] = None) -> float:
        """reference :849-899, including the `merge_{i}_{j}_{text[:20]}` cache key and the 10-text cap."""
        if not self.use_compression_aware or not self.corpus_sample:
            return 0.0
        cache = self.tokenize_cache
        if "original" not in cache:
            cache["original"] = sum(len(self.tokenize(text)) for text in self.corpus_sample)
        original_tokens = cache["original"]
        merged_token = self.vocab[i] + self.vocab[j]
        if index is None:
            index = _LengthIndex(self.vocab)
        merged_tokens = 0
        sample_size = min(len(self.corpus_sample), 10)
        for text in self.corpus_sample[:sample_size]:
            key = f"merge_{i}_{j}_{text[:20]}"
            if key not in cache:
                cache[key] = index.count(text, merged_token)
            merged_tokens += cache[key]
        if sample_size < len(self.corpus_sample):
            merged_tokens = merged_tokens * (len(self.corpus_sample) / sample_size)
        ratio = 1.0 if merged_tokens == 0 else original_tokens / merged_tokens
        return max(0.0, min(1.0, (ratio - 1.0) / 1.0))

    # ---- scoring and search (reference :903-1013) --------------------------------------------------------------------
    def _weights(self) -> Tuple[float, float, float, float, float]:
        """reference :947-969, the same float operations in the same order."""
        alpha, beta, gamma = 0.7, 0.0, 0.0
        if self.use_frequency_aware:
            alpha, beta, gamma = self.alpha, self.beta, self.gamma
        compression_weight = 0.0
        if self.use_compression_aware:
            compression_weight = self.compression_weight
            alpha *= (1 - compression_weight)
            beta *= (1 - compression_weight)
            gamma *= (1 - compression_weight)
        morphology_weight = 0.0
        if self.use_hierarchical:
            morphology_weight = 0.3
            alpha *= (1 - morphology_weight)
            beta *= (1 - morphology_weight)
            gamma *= (1 - morphology_weight)
            if self.use_compression_aware:
                compression_weight *= (1 - morphology_weight)
        return alpha, beta, gamma, compression_weight, morphology_weight

    def _score_arrays(self, ii, jj, dd) -> List[EnhancedMergeCandidate]:
        """`[_score_candidate(c) for c in candidates]` with the device work of all candidates in one launch."""
        cands = [(int(i), int(j), float(d)) for i, j, d in zip(ii, jj, dd)]
        if not cands:
            return []
        coh = self._coherence_batch(cands) if self.use_frequency_aware else [0.0] * len(cands)
        alpha, beta, gamma, cw, mw = self._weights()
        max_freq = (max(self.pair_frequencies.values())
                    if self.use_frequency_aware and self.pair_frequencies else None)
        index = _LengthIndex(self.vocab) if self.use_compression_aware and self.corpus_sample else None
        out: List[EnhancedMergeCandidate] = []
        for (i, j, dist), semantic_score in zip(cands, coh):
            distance_score = 1.0 / (1.0 + dist)
            frequency_score = self._compute_frequency_score(i, j, max_freq) if self.use_frequency_aware else 0.0
            compression_score = self._compute_compression_score(i, j, index) if self.use_compression_aware else 0.0
            morphology_score = (self._morphology_score(self.vocab[i], self.vocab[j])
                                if self.use_hierarchical else 0.0)
            combined = (alpha * distance_score + beta * frequency_score + gamma * semantic_score +
                        cw * compression_score + mw * morphology_score)
            out.append(EnhancedMergeCandidate(distance=dist, token_i=i, token_j=j, frequency_score=frequency_score,
                                              semantic_score=semantic_score, compression_score=compression_score,
                                              morphology_score=morphology_score, combined_score=-combined))
        return out

    def _score_candidate(self, candidate: MergeCandidate) -> EnhancedMergeCandidate:
        """reference :903-990."""
        return self._score_arrays([candidate.token_i], [candidate.token_j], [candidate.distance])[0]

    _refill_needs_full_list = True      # every candidate of a refill is scored and re-ranked (reference :992-1013)

    def _basic_candidate_arrays(self):
        if self.cache_semantics == "snapshot":
            return self._find_merge_candidates_fast_arrays()
        return self._candidate_arrays()

    def _find_merge_candidates_fast(self) -> List[MergeCandidate]:
        """reference :992-1013."""
        ii, jj, dd = self._basic_candidate_arrays()
        if not (self.use_frequency_aware or self.use_hierarchical or self.use_compression_aware or
                self.use_adaptive_curvature):
            return [MergeCandidate(float(d), int(i), int(j)) for i, j, d in zip(ii, jj, dd)]
        enhanced = self._score_arrays(ii, jj, dd)
        enhanced.sort()                   # Python's own sort over __lt__: NaN scores behave as in the reference
        return enhanced

    # ---- loop (reference :1015-1209) ------------------------------------------------------------------------------------
    def _sample_statistics(self, step: int, stats: Dict[int, Dict[str, Any]]) -> None:
        """reference :1075-1112: `random.sample` draws on the host, one batched exact re-score (K3)."""
        st = self._compute_distance_statistics(1000)
        n = self.current_vocab_size
        if min(1000, n * (n - 1) // 2) <= 0:
            return
        stats[step] = {"vocab_size": n, "min_dist": st["min"], "max_dist": st["max"], "mean_dist": st["mean"],
                       "phase": self.current_phase if self.use_hierarchical else 0}

    def optimize_merges(self, steps: int = 10000, log_every: int = 1000, corpus_sample: Optional[List[str]] = None,
                        adaptive_threshold: bool = True,
                        phase_transition_steps: Optional[Dict[int, int]] = None) -> None:
        if corpus_sample and self.use_compression_aware:
            self.corpus_sample = corpus_sample
            self.tokenize_cache = {}
        if self.use_hierarchical and phase_transition_steps is None:
            phase_transition_steps = {2: 1000, 3: 6000}
        no_candidate_count = 0
        stats: Dict[int, Dict[str, Any]] = {}
        self.last_trace: List[Tuple[int, int, float]] = []
        if self.use_hierarchical:
            self.merge_threshold = self._get_merge_phase_threshold()
        for step in range(steps):
            if self.use_hierarchical and step in phase_transition_steps.values():
                for phase, transition_step in phase_transition_steps.items():
                    if step == transition_step:
                        self.current_phase = phase
                        self.merge_threshold = self._get_merge_phase_threshold()
                        logger.info(f"Transitioning to phase {self.current_phase} with threshold: "
                                    f"{self.merge_threshold:.4f}")
                        if hasattr(self, "tokenize_cache"):
                            self.tokenize_cache = {}
            if self.use_adaptive_curvature and step > 0 and step % self.optimize_curvature_freq == 0:
                self._optimize_curvature()
                self._project_embeddings()
            if step % log_every == 0 and adaptive_threshold:
                self._sample_statistics(step, stats)
            candidates = self._find_merge_candidates_fast()
            if not candidates:
                no_candidate_count += 1
                if no_candidate_count > 5 and adaptive_threshold:
                    self.merge_threshold *= 1.5
                    no_candidate_count = 0
                    continue
                elif no_candidate_count > 10:
                    logger.info(f"No more merge candidates found after {step} steps")
                    break
                continue
            no_candidate_count = 0
            best = candidates[0]
            self.last_trace.append((best.token_i, best.token_j, best.distance))
            self._merge_tokens(best.token_i, best.token_j)
            if hasattr(self, "tokenize_cache") and self.tokenize_cache:
                for key in [k for k in self.tokenize_cache if k.startswith("merge_")]:
                    self.tokenize_cache.pop(key, None)
            if adaptive_threshold and step > 0 and step % 1000 == 0:
                if self.use_hierarchical:
                    self.merge_threshold = self._get_merge_phase_threshold() * (1.1 ** (step // 1000))
                else:
                    self.merge_threshold *= 1.1
        if stats:
            self.training_stats = stats

    # ---- persistence (reference :1211-1427) ------------------------------------------------------------------------------
    def save(self, path: str) -> None:
        """Same file set as the reference: vocab.json, embeddings.pt (the full Parameter, as the reference saves it),
        merges.json, enhanced_config.json, and per feature curvature.pt / merge_pairs.pt, frequencies.json,
        hierarchical_data.json, training_stats.json."""
        os.makedirs(path, exist_ok=True)
        with open(f"{path}/vocab.json", "w") as f:
            json.dump(self.vocab, f)
        torch.save(self.embeddings, f"{path}/embeddings.pt")
        with open(f"{path}/merges.json", "w") as f:
            json.dump(self.merge_history, f)
        if self.use_adaptive_curvature:
            torch.save(self.curvature, f"{path}/curvature.pt")
        c = self.get_curvature()
        config = {
            "curvature": c.item() if hasattr(c, "item") else c,
            "merge_threshold": self.merge_threshold,
            "max_vocab_size": self.max_vocab_size,
            "use_approximate_search": self.use_approximate_search,
            "use_frequency_aware": self.use_frequency_aware,
            "use_hierarchical": self.use_hierarchical,
            "use_adaptive_curvature": self.use_adaptive_curvature,
            "use_compression_aware": self.use_compression_aware,
            "alpha": getattr(self, "alpha", 0.4),
            "beta": getattr(self, "beta", 0.4),
            "gamma": getattr(self, "gamma", 0.2),
            "language": getattr(self, "language", "english"),
            "hierarchy_weight": getattr(self, "hierarchy_weight", 1.0),
            "distortion_weight": getattr(self, "distortion_weight", 0.1),
            "compression_weight": getattr(self, "compression_weight", 0.7),
            "distance_weight": getattr(self, "distance_weight", 0.3),
            "current_phase": getattr(self, "current_phase", 1),
            "current_vocab_size": self.current_vocab_size,
        }
        with open(f"{path}/enhanced_config.json", "w") as f:
            json.dump(config, f, indent=2)
        if getattr(self, "training_stats", None):
            with open(f"{path}/training_stats.json", "w") as f:
                json.dump({str(k): v for k, v in self.training_stats.items()}, f, indent=2)
        if self.use_frequency_aware and getattr(self, "pair_frequencies", None):
            with open(f"{path}/frequencies.json", "w") as f:
                json.dump({f"{k[0]}|{k[1]}": v for k, v in self.pair_frequencies.items()}, f)
        if self.use_hierarchical:
            with open(f"{path}/hierarchical_data.json", "w") as f:
                json.dump({"language": getattr(self, "language", "english"),
                           "common_morphemes": list(getattr(self, "common_morphemes", set())),
                           "common_words": list(getattr(self, "common_words", set()))}, f)
        if self.use_adaptive_curvature and hasattr(self, "merge_pairs"):
            torch.save(self.merge_pairs, f"{path}/merge_pairs.pt")

    @classmethod
    def load(cls, path: str, device: Optional[torch.device] = None) -> "EnhancedFastHyperbolicTokenizer":
        """reference :1300-1427.  The saved `embeddings.pt` is the full table; the active rows are the first
        `len(vocab)` (the reference passes the full table on and fails on the shape)."""
        with open(f"{path}/vocab.json", "r") as f:
            vocab = json.load(f)
        embeddings = torch.load(f"{path}/embeddings.pt", map_location="cpu")
        try:
            with open(f"{path}/enhanced_config.json", "r") as f:
                config = json.load(f)
        except FileNotFoundError:
            with open(f"{path}/config.json", "r") as f:
                config = json.load(f)
            config.update({"use_frequency_aware": False, "use_hierarchical": False,
                           "use_adaptive_curvature": False, "use_compression_aware": False})
        tokenizer = cls(
            vocab=vocab, embeddings=torch.nn.Parameter(embeddings.detach()[: len(vocab)].clone()),
            curvature=config.get("curvature", 1.0), merge_threshold=config.get("merge_threshold", 0.1),
            device=device, max_vocab_size=config.get("max_vocab_size", 100000),
            use_approximate_search=config.get("use_approximate_search", True),
            use_frequency_aware=config.get("use_frequency_aware", False),
            use_hierarchical=config.get("use_hierarchical", False),
            use_adaptive_curvature=config.get("use_adaptive_curvature", False),
            use_compression_aware=config.get("use_compression_aware", False),
            alpha=config.get("alpha", 0.4), beta=config.get("beta", 0.4), gamma=config.get("gamma", 0.2),
            language=config.get("language", "english"), hierarchy_weight=config.get("hierarchy_weight", 1.0),
            distortion_weight=config.get("distortion_weight", 0.1),
            compression_weight=config.get("compression_weight", 0.7),
            distance_weight=config.get("distance_weight", 0.3))
        with open(f"{path}/merges.json", "r") as f:
            tokenizer.merge_history = [tuple(m) for m in json.load(f)]
        tokenizer.current_phase = config.get("current_phase", 1)
        tokenizer.current_vocab_size = config.get("current_vocab_size", len(tokenizer.vocab))
        if tokenizer.use_adaptive_curvature:
            try:
                c = torch.load(f"{path}/curvature.pt", map_location="cpu")
                tokenizer.curvature = torch.nn.Parameter(c.detach().to(tokenizer.device))
                tokenizer.merge_pairs = [tuple(p) for p in torch.load(f"{path}/merge_pairs.pt")]
                tokenizer.curvature_optimizer = torch.optim.Adam([tokenizer.curvature],
                                                                 lr=config.get("curvature_lr", 0.01))
            except FileNotFoundError:
                logger.warning("Could not load adaptive curvature data")
        if tokenizer.use_frequency_aware:
            try:
                with open(f"{path}/frequencies.json", "r") as f:
                    tokenizer.pair_frequencies = {tuple(k.split("|")): v for k, v in json.load(f).items()}
            except FileNotFoundError:
                logger.warning("Could not load frequency data")
        if tokenizer.use_hierarchical:
            try:
                with open(f"{path}/hierarchical_data.json", "r") as f:
                    data = json.load(f)
                tokenizer.language = data.get("language", "english")
                tokenizer.common_morphemes = set(data.get("common_morphemes", []))
                tokenizer.common_words = set(data.get("common_words", []))
            except FileNotFoundError:
                logger.warning("Could not load hierarchical data")
        try:
            with open(f"{path}/training_stats.json", "r") as f:
                tokenizer.training_stats = {int(k): v for k, v in json.load(f).items()}
        except FileNotFoundError:
            pass
        return tokenizer
