"""Drop-in for the reference's ``tokenizer/hierarchical_hyperbolic_merge.py`` on B200.

The hierarchical strategy is a HOST policy over the device candidate list: every step asks the base class for
all pairs under the current threshold (K2 min + emit kernels), keeps or re-weights them by string criteria, and
merges the first minimum (K1 midpoint).  Three phases -- character merges, morpheme-biased merges, word-biased
merges -- each with its own threshold, step budget and quota (reference :279-428).  The string heuristics
(:156-225) depend on corpus statistics only; the optional NLTK/WordNet lookups of the reference are honoured when
`nltk` is importable (it is not in this image, nor in the container that produced the golden traces).
"""
from __future__ import annotations

import json
import logging
import re
from collections import Counter
from typing import Callable, Dict, List, Optional, Set, Tuple

import numpy as np
import torch

from .hyperbolic_merge import HyperbolicTokenizer

try:                                            # reference :28-36
    from nltk.corpus import wordnet             # type: ignore
    NLTK_AVAILABLE = True
except ImportError:
    wordnet = None
    NLTK_AVAILABLE = False

logger = logging.getLogger(__name__)

Candidate = Tuple[int, int, float]

_PREFIXES = {"re", "un", "in", "im", "il", "ir", "dis", "en", "em", "non", "de", "pre", "pro", "mis"}
_SUFFIXES = {"ing", "ed", "er", "est", "ly", "ity", "ment", "ness", "able", "ible", "al", "ial"}
_WORD_RE = re.compile(r"\b\w+\b")
_VOWEL_RE = re.compile(r"[aeiou]")


class HierarchicalHyperbolicTokenizer(HyperbolicTokenizer):
    """reference hierarchical_hyperbolic_merge.py:41-513."""

    # (threshold, step budget, merge quota below which an empty candidate list widens the threshold by 1.2)
    # of the three phases, reference :294-300, :343-349, :385-391.  Class attributes so that tests can shorten them.
    phase_thresholds = (0.05, 0.1, 0.2)
    phase_steps = (2000, 5000, 10000)
    phase_quotas = (500, 2000, 5000)

    def __init__(self, vocab: List[str], embeddings: torch.nn.Parameter, corpus_path: Optional[str] = None,
                 curvature: float = 1.0, merge_threshold: float = 0.05, lr: float = 1e-3,
                 device: Optional[torch.device] = None, max_vocab_size: int = 100000,
                 use_approximate_search: bool = True, language: str = "english",
                 semantics: Optional[str] = None):
        super().__init__(vocab=vocab, embeddings=embeddings, curvature=curvature, merge_threshold=merge_threshold,
                         lr=lr, device=device, max_vocab_size=max_vocab_size,
                         use_approximate_search=use_approximate_search, semantics=semantics)
        self.language = language
        self.token_frequencies: Dict[str, int] = {}
        self.common_morphemes: Set[str] = set()
        self.common_words: Set[str] = set()
        if corpus_path:
            self._compute_corpus_statistics(corpus_path)

    # ---- corpus statistics (host; reference :108-154) ---------------------------------------------------------
    def _compute_corpus_statistics(self, corpus_path: str) -> None:
        words: Counter = Counter()
        grams: Counter = Counter()
        with open(corpus_path, "r", encoding="utf-8") as f:
            for line in f:
                found = _WORD_RE.findall(line.lower())
                words.update(found)
                for w in found:
                    for n in range(2, min(6, len(w) + 1)):          # character n-grams, n = 2..5
                        grams.update(w[k:k + n] for k in range(len(w) - n + 1))
        self.token_frequencies = dict(words)
        gram_cut = np.percentile(list(grams.values()), 80)
        self.common_morphemes = {g for g, cnt in grams.items() if cnt >= gram_cut}
        word_cut = np.percentile(list(words.values()), 70)
        self.common_words = {w for w, cnt in words.items() if cnt >= word_cut}
        logger.info(f"Identified {len(self.common_morphemes)} common morphemes and "
                    f"{len(self.common_words)} common words")

    # ---- string heuristics (reference :156-225) -------------------------------------------------------------
    def _is_potential_morpheme(self, token: str) -> bool:
        if token in self.common_morphemes:
            return True
        if NLTK_AVAILABLE:
            if token in _PREFIXES or token in _SUFFIXES:
                return True
            if len(token) > 2 and any(wordnet.synsets(token, pos=pos)
                                      for pos in (wordnet.NOUN, wordnet.VERB, wordnet.ADJ, wordnet.ADV)):
                return True
        if 2 <= len(token) <= 5:
            return sum(1 for word in self.common_words if token in word) >= 5
        return False

    def _is_valid_word(self, token: str) -> bool:
        if token in self.common_words:
            return True
        if NLTK_AVAILABLE and wordnet.synsets(token):
            return True
        return len(token) >= 3 and _VOWEL_RE.search(token) is not None

    def _reweight(self, candidates: List[Candidate], accept: Callable[[str], bool], factor: float) -> List[Candidate]:
        """Candidates whose merged string passes `accept` get their distance scaled by `factor` (a Python-float
        product, as in the reference); all candidates are kept, in order."""
        vocab = self.vocab
        return [(i, j, dist * factor) if accept(vocab[i] + vocab[j]) else (i, j, dist) for i, j, dist in candidates]

    def _filter_morphologically_valid(self, candidates: List[Candidate]) -> List[Candidate]:
        """reference :227-251."""
        return self._reweight(candidates, self._is_potential_morpheme, 0.8)

    def _filter_word_valid(self, candidates: List[Candidate]) -> List[Candidate]:
        """reference :253-277."""
        return self._reweight(candidates, self._is_valid_word, 0.7)

    # ---- the three phases (reference :279-428) --------------------------------------------------------------------
    def _run_phase(self, phase: int, choose: Callable[[List[Candidate], int], Optional[Candidate]],
                   target_vocab_size: Optional[int], threshold_cap: Optional[float] = None) -> bool:
        """One phase: up to `phase_steps[phase]` steps; an empty candidate list widens the threshold (x1.2, costing
        the step) while the phase is under its quota (and under `threshold_cap`), else ends the phase; `choose`
        returning None ends the phase.  Returns True when the target vocabulary size was reached."""
        self.merge_threshold = self.phase_thresholds[phase]
        merged = 0
        for _ in range(self.phase_steps[phase]):
            candidates = self._find_merge_candidates()
            if not candidates:
                if merged < self.phase_quotas[phase] and (threshold_cap is None or self.merge_threshold < threshold_cap):
                    self.merge_threshold *= 1.2
                    continue
                break
            pick = choose(candidates, merged)
            if pick is None:
                break
            self._merge_tokens(pick[0], pick[1])
            merged += 1
            if target_vocab_size and self.current_vocab_size >= target_vocab_size:
                return True
        logger.info(f"Completed Phase {phase + 1} with {merged} merges. Vocabulary size: {self.current_vocab_size}")
        return False

    def _hierarchical_merge_strategy(self, target_vocab_size: Optional[int] = None) -> None:
        vocab = self.vocab

        def first_min(cands: List[Candidate]) -> Optional[Candidate]:
            return min(cands, key=lambda c: c[2]) if cands else None       # the first minimum, like the reference

        def characters(cands: List[Candidate], merged: int) -> Optional[Candidate]:
            short = [c for c in cands if len(vocab[c[0]]) <= 2 and len(vocab[c[1]]) <= 2]
            if not short and merged < self.phase_quotas[0]:
                short = [c for c in cands if len(vocab[c[0]]) <= 3 and len(vocab[c[1]]) <= 3]
            return first_min(short)

        def morphemes(cands: List[Candidate], merged: int) -> Optional[Candidate]:
            return first_min(self._filter_morphologically_valid(cands))

        def words(cands: List[Candidate], merged: int) -> Optional[Candidate]:
            return first_min(self._filter_word_valid(cands))

        if self._run_phase(0, characters, target_vocab_size):
            return
        if self._run_phase(1, morphemes, target_vocab_size):
            return
        self._run_phase(2, words, target_vocab_size, threshold_cap=1.0)
        logger.info(f"Final vocabulary size: {self.current_vocab_size}")

    def optimize_merges(self, steps: int = 10000, log_every: int = 1000, hierarchical: bool = True,
                        target_vocab_size: Optional[int] = None) -> None:
        """reference :430-448."""
        if hierarchical:
            self._hierarchical_merge_strategy(target_vocab_size)
        else:
            super().optimize_merges(steps, log_every)

    # ---- persistence (reference :450-513) ---------------------------------------------------------------------------
    def save(self, path: str) -> None:
        super().save(path)
        with open(f"{path}/hierarchical_data.json", "w") as f:
            json.dump({"language": self.language, "common_morphemes": list(self.common_morphemes),
                       "common_words": list(self.common_words)}, f)

    @classmethod
    def load(cls, path: str, device: Optional[torch.device] = None) -> "HierarchicalHyperbolicTokenizer":
        tokenizer = super().load(path, device)
        try:
            with open(f"{path}/hierarchical_data.json", "r") as f:
                data = json.load(f)
            tokenizer.language = data.get("language", "english")
            tokenizer.common_morphemes = set(data.get("common_morphemes", []))
            tokenizer.common_words = set(data.get("common_words", []))
        except FileNotFoundError:
            logger.warning("Hierarchical data file not found")
        return tokenizer
