"""Evaluation helpers of the reference whose inner loop is a Lorentz distance call."""
from .distortion import distortion_ratios, pair_distances

__all__ = ["distortion_ratios", "pair_distances"]
