"""Losses of the reference's ``multimodal`` package that sit directly on the Lorentz distance kernels."""
from .contrastive_loss import HyperbolicInfoNCE, hyperbolic_contrastive_loss, hyperbolic_triplet_loss

__all__ = ["HyperbolicInfoNCE", "hyperbolic_contrastive_loss", "hyperbolic_triplet_loss"]
