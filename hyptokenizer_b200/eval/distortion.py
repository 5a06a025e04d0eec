"""The distance part of the reference's ``scripts/eval_hierarchy.py`` `compute_distortion`.

The reference samples node pairs from a WordNet graph on the host (networkx shortest paths, :118-137; out of
scope) and then calls `distance(...).item()` once per pair (:143-156).  Here the pairs are re-scored in one launch
of the K3 kernel; the ratio statistics are the reference's numpy calls on the same values.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from .._lib import SEM, check, ptr, stream_ptr
from ..embedding import lorentz_model as LM


def pair_distances(embeddings: torch.Tensor, idx_i: Sequence[int], idx_j: Sequence[int], curvature: float = 1.0,
                   semantics: Optional[str] = None) -> np.ndarray:
    """distance(embeddings[i], embeddings[j]) for every listed pair: what eval_hierarchy.py:148-152 computes one
    `.item()` at a time.  `embeddings` is a float32 CUDA tensor (n, D)."""
    if not embeddings.is_cuda or embeddings.dtype != torch.float32 or embeddings.dim() != 2:
        raise RuntimeError("pair_distances needs a float32 CUDA tensor (n, D); there is no CPU fallback")
    E = embeddings.detach().contiguous()
    dev = E.device
    _lib.check_device(dev)
    ii = torch.as_tensor(np.asarray(idx_i, dtype=np.int32)).to(dev)
    jj = torch.as_tensor(np.asarray(idx_j, dtype=np.int32)).to(dev)
    if ii.shape != jj.shape or ii.dim() != 1:
        raise ValueError("idx_i and idx_j must be 1-D and of equal length")
    n = ii.numel()
    if n and (int(ii.min()) < 0 or int(jj.min()) < 0 or int(ii.max()) >= E.shape[0] or int(jj.max()) >= E.shape[0]):
        raise IndexError("pair index out of range")
    out = torch.empty(n, dtype=torch.float32, device=dev)
    name = LM.get_semantics() if semantics is None else semantics
    with torch.cuda.device(dev):
        check(_lib.lib().hyp_rescore_pairs(ptr(E), E.stride(0), ptr(ii), ptr(jj), ptr(out), None, n, E.shape[1],
                                           float(curvature), SEM[name], stream_ptr()))
    return out.cpu().numpy()


def distortion_ratios(embeddings: torch.Tensor, pairs: Sequence[Tuple[int, int, float]], curvature: float = 1.0,
                      semantics: Optional[str] = None) -> Tuple[np.ndarray, Dict[str, float]]:
    """eval_hierarchy.py:139-170 for already sampled `(i, j, graph_distance)` triples: embedding distance over
    graph distance per pair, and the reference's summary statistics."""
    ii = [p[0] for p in pairs]
    jj = [p[1] for p in pairs]
    graph = np.array([p[2] for p in pairs], dtype=np.float64)
    # `.item()` of an fp32 distance is a Python float: the division happens in double, as in the reference
    ratios_array = pair_distances(embeddings, ii, jj, curvature, semantics).astype(np.float64) / graph
    stats = {
        "mean": float(np.mean(ratios_array)),
        "median": float(np.median(ratios_array)),
        "min": float(np.min(ratios_array)),
        "max": float(np.max(ratios_array)),
        "std": float(np.std(ratios_array)),
        "num_pairs": len(ratios_array),
    }
    return ratios_array, stats
