"""Seeded synthetic inputs for benchmarks and size-scaled tests.

`synthetic_embeddings` follows the reference's init law
(scripts/train_hyperbolic_tokenizer.py:64-109: randn*scale tangent vector at the origin,
exponential map, projection) in closed form: for the origin o = (1, 0, ..., 0) and tangent
(0, t), exp_o(t) = (cosh|t|, sinh|t| t/|t|), then x0 <- sqrt(1 + |xs|^2).  The values are inputs,
not results: both the GPU arm and the CPU arm consume the same tensor.
"""
from __future__ import annotations

import numpy as np
import torch


def synthetic_embeddings(n: int, d: int, scale: float = 0.01, seed: int = 42, device="cpu") -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    t = torch.randn((n, d), generator=g, dtype=torch.float32) * scale
    r = torch.sqrt(torch.clamp((t * t).sum(-1, keepdim=True), min=1e-8))
    xs = torch.sinh(r) * (t / r)
    x0 = torch.sqrt(1.0 + (xs * xs).sum(-1, keepdim=True))
    return torch.cat([x0, xs], dim=-1).to(device)


def synthetic_vocab(n: int):
    return [f"w{k}" for k in range(n)]


def synthetic_corpus(n_bytes: int, seed: int = 0, words_per_line: int = 20) -> np.ndarray:
    """ASCII corpus of config 4: lines of `words_per_line` random lower-case words (length 2..10)
    joined by single spaces.  Returned as a uint8 array of exactly n_bytes (last line cut)."""
    rng = np.random.default_rng(seed)
    out = np.empty(n_bytes, dtype=np.uint8)
    pos = 0
    while pos < n_bytes:
        m = min(1 << 24, n_bytes - pos)
        blk = rng.integers(97, 123, size=m, dtype=np.uint8)
        lens = rng.integers(2, 11, size=m // 3 + 8)
        ends = np.cumsum(lens + 1) - 1
        ends = ends[ends < m]
        blk[ends] = 32
        blk[ends[words_per_line - 1::words_per_line]] = 10
        out[pos:pos + m] = blk
        pos += m
    return out


_ENGLISH_LETTERS = b"etaoinshrdlcumwfgypbvkjxqz"
_ENGLISH_FREQ = (12.7, 9.06, 8.17, 7.51, 6.97, 6.75, 6.33, 6.09, 5.99, 4.25, 4.03, 2.78, 2.76, 2.41, 2.36, 2.23, 2.02,
                 1.97, 1.93, 1.49, 0.98, 0.77, 0.15, 0.15, 0.10, 0.07)


def english_corpus(n_bytes: int, seed: int = 0, words_per_line: int = 20) -> np.ndarray:
    """Like `synthetic_corpus`, but the letters follow English letter frequencies (first-order model): the same 27
    symbols + line break, with a Zipf-like bigram distribution instead of a flat one -- rare letters (0.1 % of the
    text) and counters that pass 255, which the uniform stream never produces."""
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(_ENGLISH_LETTERS, np.uint8)
    p = np.asarray(_ENGLISH_FREQ, np.float64)
    # inverse CDF through a 65 536-entry table: frequencies quantised to 2^-16, one gather per byte
    table = letters[np.searchsorted(np.cumsum(p / p.sum()), (np.arange(65536) + 0.5) / 65536.0).clip(0, len(letters) - 1)]
    out = table[rng.integers(0, 65536, size=n_bytes, dtype=np.uint16)]
    lens = rng.integers(2, 9, size=n_bytes // 3 + 8)
    ends = np.cumsum(lens + 1) - 1
    ends = ends[ends < n_bytes]
    out[ends] = 32
    out[ends[words_per_line - 1::words_per_line]] = 10
    return out
