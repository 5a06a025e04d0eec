"""Shared test helpers: bit-pattern <-> tensor conversion, synthetic inputs."""
import numpy as np
import torch


def from_bits(bits, *shape) -> torch.Tensor:
    a = np.asarray(bits, dtype=np.uint32).view(np.float32)
    t = torch.from_numpy(a.copy())
    return t.reshape(*shape) if shape else t


def to_bits(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().contiguous().numpy()
    return np.asarray(t, dtype=np.float32).view(np.uint32).reshape(-1)


def fbits(x: float) -> int:
    return int(np.array([x], dtype=np.float32).view(np.uint32)[0])


def same_bits(a, b) -> bool:
    """Bit equality, except that any NaN equals any NaN (payloads are not part of the contract)."""
    a, b = to_bits(a), to_bits(b)
    if a.shape != b.shape:
        return False
    fa, fb = a.view(np.float32), b.view(np.float32)
    return bool(np.all((a == b) | (np.isnan(fa) & np.isnan(fb))))


def ulp_diff(a, b) -> np.ndarray:
    """|a-b| in units in the last place of fp32 (monotone integer mapping); NaN==NaN -> 0."""
    a, b = to_bits(a).astype(np.int64), to_bits(b).astype(np.int64)
    fa = a.astype(np.uint32).view(np.float32)
    fb = b.astype(np.uint32).view(np.float32)
    ma = np.where(a & 0x80000000, 0x80000000 - a, a)
    mb = np.where(b & 0x80000000, 0x80000000 - b, b)
    d = np.abs(ma - mb)
    d[np.isnan(fa) & np.isnan(fb)] = 0
    return d
