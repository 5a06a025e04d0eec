"""Drop-in for the loss functions of the reference's ``multimodal/contrastive_loss.py``.

The reference fills the B x B text-to-image distance matrix one row at a time (:38-45: B `distance` calls on
expanded rows) and lets autograd unroll that; here the matrix is ONE `batch_distance` call (same values: both
kernels form <x,y> in the reference's summation order) whose backward is a tile kernel plus two GEMMs.
The two-tower projector of the same file (:132-248) is ordinary torch.nn code and stays out of scope.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from ..embedding.lorentz_model import batch_distance, distance


def hyperbolic_contrastive_loss(z_text: torch.Tensor, z_img: torch.Tensor, temp: float = 0.07,
                                reduction: str = "mean", semantics: Optional[str] = None) -> torch.Tensor:
    """reference multimodal/contrastive_loss.py:17-60 (curvature fixed at 1.0 there)."""
    batch_size = z_text.size(0)
    text_to_img_dist = batch_distance(z_text, z_img, c=1.0, semantics=semantics)
    similarities = -text_to_img_dist / temp
    labels = torch.arange(batch_size, device=z_text.device)
    loss_t2i = F.cross_entropy(similarities, labels, reduction=reduction)
    loss_i2t = F.cross_entropy(similarities.t(), labels, reduction=reduction)
    return (loss_t2i + loss_i2t) / 2.0


def hyperbolic_triplet_loss(anchor: torch.Tensor, positive: torch.Tensor, negative: torch.Tensor,
                            margin: float = 1.0, reduction: str = "mean",
                            semantics: Optional[str] = None) -> torch.Tensor:
    """reference multimodal/contrastive_loss.py:63-95."""
    d_pos = distance(anchor, positive, c=1.0, semantics=semantics)
    d_neg = distance(anchor, negative, c=1.0, semantics=semantics)
    losses = F.relu(d_pos - d_neg + margin)
    if reduction == "mean":
        return losses.mean()
    if reduction == "sum":
        return losses.sum()
    return losses


class HyperbolicInfoNCE(torch.nn.Module):
    """reference multimodal/contrastive_loss.py:98-129."""

    def __init__(self, temperature: float = 0.07, semantics: Optional[str] = None):
        super().__init__()
        self.temperature = temperature
        self.semantics = semantics

    def forward(self, z1: torch.Tensor, z2: torch.Tensor) -> torch.Tensor:
        return hyperbolic_contrastive_loss(z1, z2, temp=self.temperature, semantics=self.semantics)
