"""Bit-exact numpy emulation of the ATen CPU fp32 reductions the reference inherits.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference never writes a reduction loop itself: `torch.sum(xs*ys, dim=-1)`
(embedding/lorentz_model.py:25,85,163) and `torch.norm(xs, dim=-1)` (:53) run
ATen's vectorised CPU kernels, whose summation ORDER decides the last bit of the
pre-clamp Minkowski product.  The CUDA kernels follow this order (SURVEY.md
Appendix D), so `u` is reproduced bit for bit; this file states the order in
plain numpy so the tests can check the kernels (and torch on the current host)
against it.

Observed with torch 2.11 CPU (8 fp32 lanes, ILP 4):

sum:   products rounded to fp32 first; lane-vectors k=0..N//8-1 accumulated into
       4 interleaved partial vectors (k mod 4) in order; partials folded
       ((p0+p1)+p2)+p3; scalar tail summed first from 0, then lanes 0..7 added in
       order.  N < 8: same 4-partial scheme on scalars.
norm:  8 lane accumulators acc += x*x (product rounded, no FMA), lanes folded
       sequentially 0..7, tail in groups of 4 (rounded products) then a <=3-element
       remainder with FMA, then sqrt.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _fold4(vals):
    """ATen row_sum with ILP 4 over a list of scalars or equal-length vectors."""
    n = len(vals)
    zero = np.zeros_like(np.asarray(vals[0], dtype=F32))
    part = [zero.copy() for _ in range(4)]
    for r in range(n // 4):
        for k in range(4):
            part[k] = (part[k] + vals[4 * r + k]).astype(F32)
    for i in range(4 * (n // 4), n):
        part[0] = (part[0] + vals[i]).astype(F32)
    acc = (part[0] + part[1]).astype(F32)
    acc = (acc + part[2]).astype(F32)
    acc = (acc + part[3]).astype(F32)
    return acc


def sum_fp32(p: np.ndarray) -> np.float32:
    """torch.sum over a contiguous fp32 vector `p` of already-rounded products."""
    p = np.asarray(p, dtype=F32)
    n = p.shape[0]
    if n == 0:
        return F32(0)
    if n < 8:
        return F32(_fold4([p[i] for i in range(n)]))
    vs = n // 8
    lanes = _fold4([p[8 * k:8 * k + 8] for k in range(vs)])
    acc = F32(0)
    for k in range(8 * vs, n):
        acc = F32(acc + p[k])
    for lane in range(8):
        acc = F32(acc + lanes[lane])
    return acc


def mdot_fp32(x: np.ndarray, y: np.ndarray) -> np.float32:
    """embedding/lorentz_model.py:25 in fp32 with ATen's order:
    fl(fl(x0*y0) - sum_fp32(fl(xs*ys)))."""
    x = np.asarray(x, dtype=F32)
    y = np.asarray(y, dtype=F32)
    t = F32(x[0] * y[0])
    s = sum_fp32((x[1:] * y[1:]).astype(F32))
    return F32(t - s)


def norm_fp32(x: np.ndarray) -> np.float32:
    """torch.norm(x, dim=-1) of a contiguous fp32 vector (embedding/lorentz_model.py:53)."""
    x = np.asarray(x, dtype=F32)
    n = x.shape[0]
    vs = n // 8
    acc = np.zeros(8, dtype=F32)
    for k in range(vs):
        d = x[8 * k:8 * k + 8]
        acc = (acc + (d * d).astype(F32)).astype(F32)
    b = acc[0]
    for lane in range(1, 8):
        b = F32(b + acc[lane])
    k = 8 * vs
    while n - k >= 4:
        for q in range(4):
            b = F32(b + F32(x[k + q] * x[k + q]))
        k += 4
    while k < n:
        b = F32(np.float64(b) + np.float64(x[k]) * np.float64(x[k]))  # FMA
        k += 1
    return F32(np.sqrt(b))


def host_matches_torch(trials: int = 64, dims=(3, 5, 50, 100), seed: int = 0) -> bool:
    """True when torch's CPU kernels on THIS host reduce in the order stated above.
    The lane count is a property of the ATen dispatch of the machine that produced
    the golden vectors; tests that depend on it skip (not fail) elsewhere."""
    import torch
    g = torch.Generator().manual_seed(seed)
    for d in dims:
        for t in range(trials):
            a = torch.randn(d + 1, generator=g) * (0.01 if t % 2 else 1.0)
            b = torch.randn(d + 1, generator=g)
            ref = (a[0] * b[0] - torch.sum(a[1:] * b[1:], dim=-1)).item()
            if F32(ref) != mdot_fp32(a.numpy(), b.numpy()):
                return False
            nr = torch.norm(a[1:], dim=-1).item()
            if F32(nr) != norm_fp32(a[1:].numpy()):
                return False
    return True
