"""Oracle restatement of the reference's Lorentz-model math (torch CPU eager, fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function names the
reference lines it follows (paths relative to /root/reference).  The op ORDER is
kept exactly (separately rounded products, ATen ``sum``, then one multiply and one
subtract) because the GPU kernels are checked bit for bit on the pre-clamp value.

Two semantics:

``reference``  the shipped code: ``distance``/``log_map`` negate a (+,-,...,-)
               Minkowski product and clamp at 1.0f, so every distance is 0.0 and
               every log map is NaN.
``lorentz``    the three-function correction (no negation, ``where`` instead of
               the 0*NaN mask) that makes the geometry non-degenerate.
"""
from __future__ import annotations

import torch

SEMANTICS = ("reference", "lorentz")
# `1.0 + 1e-8` is a Python double; torch.clamp on an fp32 tensor rounds it to 1.0f.
_CLAMP_MIN = 1.0 + 1e-8


def _sgn(semantics: str) -> float:
    if semantics not in SEMANTICS:
        raise ValueError(f"semantics must be one of {SEMANTICS}")
    return -1.0 if semantics == "reference" else 1.0


def minkowski_dot(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """embedding/lorentz_model.py:14-25 -- signature (+,-,...,-)."""
    time_part = x[..., 0] * y[..., 0]
    space_part = torch.sum(x[..., 1:] * y[..., 1:], dim=-1)
    return time_part - space_part


def minkowski_norm(x: torch.Tensor) -> torch.Tensor:
    """embedding/lorentz_model.py:28-38."""
    return torch.sqrt(torch.clamp(minkowski_dot(x, x), min=1e-8))


def project_to_hyperboloid(x: torch.Tensor, c: float = 1.0) -> torch.Tensor:
    """embedding/lorentz_model.py:41-56 -- x0 <- sqrt(1 + c*|xs|^2), xs kept."""
    r = torch.norm(x[..., 1:], dim=-1, keepdim=True)
    x0 = torch.sqrt(1.0 + c * r * r)
    return torch.cat([x0, x[..., 1:]], dim=-1)


def lorentz_to_klein(x: torch.Tensor, c: float = 1.0) -> torch.Tensor:
    """embedding/lorentz_model.py:59-70."""
    return x[..., 1:] / x[..., 0:1]


def exp_map(x: torch.Tensor, v: torch.Tensor, c: float = 1.0) -> torch.Tensor:
    """embedding/lorentz_model.py:73-93 (spatial-only norm, `c` ignored)."""
    sq = torch.sum(v[..., 1:] * v[..., 1:], dim=-1, keepdim=True)
    vn = torch.sqrt(torch.clamp(sq, min=1e-8))
    small = (vn < 1e-6).to(v.dtype)          # never true: vn >= 1e-4
    direction = v / (vn + small)
    direction = small * torch.zeros_like(direction) + (1 - small) * direction
    return torch.cosh(vn) * x + torch.sinh(vn) * direction


def log_map(x: torch.Tensor, y: torch.Tensor, c: float = 1.0,
            semantics: str = "reference") -> torch.Tensor:
    """embedding/lorentz_model.py:96-119 (`c` ignored).

    reference: u=clamp(-<x,y>); coef via the 0*NaN mask; coef*(y + <x,y> x).
    lorentz  : SURVEY Appendix B: u=clamp(+<x,y>); where(); coef*(y - <x,y> x).
    """
    m = minkowski_dot(x, y)
    if semantics == "reference":
        u = torch.clamp(-m, min=_CLAMP_MIN)
        coef = torch.acosh(u) / torch.sqrt(u * u - 1)
        coef = torch.clamp(coef, max=1e4)
        bad = ((coef != coef) | (coef > 1e4)).to(coef.dtype)
        coef = bad * torch.ones_like(coef) + (1 - bad) * coef
        return coef.unsqueeze(-1) * (y + m.unsqueeze(-1) * x)
    _sgn(semantics)
    u = torch.clamp(m, min=_CLAMP_MIN)
    coef = torch.clamp(torch.acosh(u) / torch.sqrt(u * u - 1), max=1e4)
    coef = torch.where(torch.isnan(coef) | (coef > 1e4), torch.ones_like(coef), coef)
    return coef.unsqueeze(-1) * (y - m.unsqueeze(-1) * x)


def distance(x: torch.Tensor, y: torch.Tensor, c: float = 1.0,
             semantics: str = "reference") -> torch.Tensor:
    """embedding/lorentz_model.py:122-138: acosh(clamp(sgn*<x,y>, 1.0f)) / sqrt(c)."""
    s = _sgn(semantics)
    m = minkowski_dot(x, y)
    u = torch.clamp(-m if s < 0 else m, min=_CLAMP_MIN)
    c_t = torch.tensor(c, device=x.device, dtype=x.dtype)
    return torch.acosh(u) / torch.sqrt(c_t)


def gram_u(x: torch.Tensor, y: torch.Tensor, row_chunk: int = 64) -> torch.Tensor:
    """Pre-clamp, pre-sign pairwise Minkowski products, op order of
    embedding/lorentz_model.py:155-166 (broadcast product, sum over the last
    axis, time term minus space term).  Chunked over rows of `x` only to bound
    the (n, m, d) temporary; chunking does not change any value."""
    out = torch.empty((x.shape[0], y.shape[0]), dtype=x.dtype)
    yt = y.unsqueeze(0)
    for a in range(0, x.shape[0], row_chunk):
        xa = x[a:a + row_chunk].unsqueeze(1)
        t = xa[..., 0] * yt[..., 0]
        s = torch.sum(xa[..., 1:] * yt[..., 1:], dim=-1)
        out[a:a + row_chunk] = t - s
    return out


def batch_distance(x: torch.Tensor, y: torch.Tensor, c: float = 1.0,
                   semantics: str = "reference", row_chunk: int = 64) -> torch.Tensor:
    """embedding/lorentz_model.py:141-178 (full (n, m) matrix)."""
    s = _sgn(semantics)
    m = gram_u(x, y, row_chunk)
    u = torch.clamp(-m if s < 0 else m, min=_CLAMP_MIN)
    c_t = torch.tensor(c, device=x.device, dtype=x.dtype)
    return torch.acosh(u) / torch.sqrt(c_t)


def midpoint(xi: torch.Tensor, xj: torch.Tensor, len_i: int, len_j: int,
             c: float = 1.0, semantics: str = "reference",
             project: bool = True) -> torch.Tensor:
    """tokenizer/hyperbolic_merge.py:323-340: weighted geodesic point between two
    rows (1-D tensors of D = d+1).  `project=False` is the un-projected variant of
    tokenizer/frequency_aware_hyperbolic_merge.py:139-141 (returns shape (1, D))."""
    w_j = len_j / (len_i + len_j)
    a = xi.unsqueeze(0)
    b = xj.unsqueeze(0)
    v = log_map(a, b, c, semantics) * w_j
    m = exp_map(a, v, c)
    if not project:
        return m
    return project_to_hyperboloid(m[0], c)


def initialize_embeddings(n: int, d: int, c: float = 1.0, scale: float = 0.01) -> torch.Tensor:
    """scripts/train_hyperbolic_tokenizer.py:64-109: randn*scale tangent at the
    origin -> exp_map row by row (1-D operands) -> project.  Consumes the global
    torch CPU generator exactly as the script does (one randn of (n, d))."""
    tangent = torch.zeros((n, d + 1), dtype=torch.float32)
    tangent[:, 1:] = torch.randn((n, d), dtype=torch.float32) * scale
    origin = torch.zeros(d + 1, dtype=torch.float32)
    origin[0] = 1.0
    emb = torch.zeros((n, d + 1), dtype=torch.float32)
    for r in range(n):
        emb[r] = exp_map(origin, tangent[r], c)
    return project_to_hyperboloid(emb, c)
