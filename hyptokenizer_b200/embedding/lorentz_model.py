"""Drop-in for the reference's ``embedding/lorentz_model.py`` on B200.

Same function names, argument order, broadcasting (`...` leading dims, last dim D = d+1)
and return shapes as the reference; every function runs a hand-written sm_100a kernel
through the C ABI (include/hyptok_b200.h).  CUDA fp32 tensors only -- no CPU fallback.
`distance` and `batch_distance` are differentiable with respect to x and y (what the reference's
multimodal losses need); the other functions return detached results (the merge
loop never differentiates).

`semantics` selects the arithmetic:
  "reference"  the shipped code, bit-faithful: distance == 0.0, log_map == NaN
  "lorentz"    corrected geometry
The module-level default is "reference" (drop-in); change it with `set_semantics`.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from .. import _lib
from .._lib import SEM, check, ptr, require_cuda, stream_ptr

_default_semantics = "reference"


def set_semantics(name: str) -> None:
    global _default_semantics
    if name not in SEM:
        raise ValueError(f"semantics must be one of {tuple(SEM)}")
    _default_semantics = name


def get_semantics() -> str:
    return _default_semantics


def _sem(semantics: Optional[str]) -> int:
    name = _default_semantics if semantics is None else semantics
    if name not in SEM:
        raise ValueError(f"semantics must be one of {tuple(SEM)}")
    return SEM[name]


def _curv(c) -> float:
    # the reference does torch.tensor(c) and so accepts Parameters too (lorentz_model.py:137)
    return float(c.detach().item()) if isinstance(c, torch.Tensor) else float(c)


def _rows(t: torch.Tensor, bshape: torch.Size, D: int) -> Tuple[torch.Tensor, int]:
    """(n, D) view/copy of `t` broadcast to bshape, plus its row stride (0 = one row for all)."""
    if t.shape[-1] != D:
        raise ValueError(f"last dimension mismatch: {t.shape[-1]} vs {D}")
    lead = t.shape[:-1]
    if math.prod(lead) == 1 and math.prod(bshape) != 1:
        return t.reshape(1, D).contiguous(), 0
    te = t.expand(*bshape, D) if tuple(lead) != tuple(bshape) else t
    return te.reshape(-1, D).contiguous(), D


def _no_grad_path(name: str, *tensors) -> None:
    """The reference computes these functions with torch ops, so autograd flows through them.  Here only `distance`
    and `batch_distance` carry a backward kernel (the merge loop never differentiates anything else); the others return
    detached results.  Failing loudly beats a training loop that silently stops learning."""
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        raise RuntimeError(f"hyptokenizer_b200.embedding.lorentz_model.{name} has no backward pass: its result is detached. "
                           "Call it under torch.no_grad() / on .detach()ed inputs, or differentiate through `distance` / "
                           "`batch_distance`, which do carry gradients.")


def _pair(x: torch.Tensor, y: torch.Tensor):
    dev = require_cuda(x, y)
    _lib.check_device(dev)
    D = x.shape[-1]
    bshape = torch.broadcast_shapes(x.shape[:-1], y.shape[:-1])
    xr, ldx = _rows(x.detach(), bshape, D)
    yr, ldy = _rows(y.detach(), bshape, D)
    return dev, D, bshape, math.prod(bshape), xr, ldx, yr, ldy


def minkowski_dot(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """reference embedding/lorentz_model.py:14-25."""
    _no_grad_path("minkowski_dot", x, y)
    dev, D, bshape, n, xr, ldx, yr, ldy = _pair(x, y)
    out = torch.empty(n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().hyp_minkowski_dot(ptr(xr), ldx, ptr(yr), ldy, ptr(out), n, D, stream_ptr()))
    return out.reshape(bshape)


def minkowski_norm(x: torch.Tensor) -> torch.Tensor:
    """reference embedding/lorentz_model.py:28-38 (sqrt/clamp are exact elementwise ops)."""
    return torch.sqrt(torch.clamp(minkowski_dot(x, x), min=1e-8))


def project_to_hyperboloid(x: torch.Tensor, c: float = 1.0) -> torch.Tensor:
    """reference embedding/lorentz_model.py:41-56."""
    _no_grad_path("project_to_hyperboloid", x, c)
    dev = require_cuda(x)
    _lib.check_device(dev)
    D = x.shape[-1]
    xr = x.detach().reshape(-1, D).contiguous()
    out = torch.empty_like(xr)
    with torch.cuda.device(dev):
        check(_lib.lib().hyp_project(ptr(xr), D, ptr(out), D, xr.shape[0], D, _curv(c), stream_ptr()))
    return out.reshape(x.shape)


def lorentz_to_klein(x: torch.Tensor, c: float = 1.0) -> torch.Tensor:
    """reference embedding/lorentz_model.py:59-70 (a plain IEEE division)."""
    require_cuda(x)
    return x[..., 1:] / x[..., 0:1]


def exp_map(x: torch.Tensor, v: torch.Tensor, c: float = 1.0) -> torch.Tensor:
    """reference embedding/lorentz_model.py:73-93."""
    _no_grad_path("exp_map", x, v)
    dev, D, bshape, n, xr, ldx, vr, ldv = _pair(x, v)
    out = torch.empty((n, D), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().hyp_exp_map(ptr(xr), ldx, ptr(vr), ldv, ptr(out), D, n, D, stream_ptr()))
    return out.reshape(*bshape, D)


def log_map(x: torch.Tensor, y: torch.Tensor, c: float = 1.0, semantics: Optional[str] = None) -> torch.Tensor:
    """reference embedding/lorentz_model.py:96-119."""
    _no_grad_path("log_map", x, y)
    dev, D, bshape, n, xr, ldx, yr, ldy = _pair(x, y)
    out = torch.empty((n, D), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().hyp_log_map(ptr(xr), ldx, ptr(yr), ldy, ptr(out), D, n, D, _sem(semantics), stream_ptr()))
    return out.reshape(*bshape, D)


def _distance_fwd(x: torch.Tensor, y: torch.Tensor, c: float, sem: int) -> torch.Tensor:
    dev, D, bshape, n, xr, ldx, yr, ldy = _pair(x, y)
    out = torch.empty(n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().hyp_distance(ptr(xr), ldx, ptr(yr), ldy, ptr(out), n, D, c, sem, stream_ptr()))
    return out.reshape(bshape)


class _Distance(torch.autograd.Function):
    """distance with the gradient torch autograd derives for lorentz_model.py:122-138."""

    @staticmethod
    def forward(ctx, x, y, c, sem):
        ctx.save_for_backward(x, y)
        ctx.c, ctx.sem = c, sem
        return _distance_fwd(x, y, c, sem)

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        dev, D, bshape, n, xr, ldx, yr, ldy = _pair(x, y)
        gr = g.detach().to(torch.float32).expand(bshape).reshape(-1).contiguous()
        need_x, need_y = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gx = torch.empty((n, D), dtype=torch.float32, device=dev) if need_x else None
        gy = torch.empty((n, D), dtype=torch.float32, device=dev) if need_y else None
        with torch.cuda.device(dev):
            check(_lib.lib().hyp_distance_backward(ptr(xr), ldx, ptr(yr), ldy, ptr(gr), ptr(gx) if need_x else None,
                                                   ptr(gy) if need_y else None, n, D, ctx.c, ctx.sem, stream_ptr()))
        # rows that were broadcast receive the sum of their copies' gradients
        gx = gx.reshape(*bshape, D).sum_to_size(x.shape) if need_x else None
        gy = gy.reshape(*bshape, D).sum_to_size(y.shape) if need_y else None
        return gx, gy, None, None


def distance(x: torch.Tensor, y: torch.Tensor, c: float = 1.0, semantics: Optional[str] = None) -> torch.Tensor:
    """reference embedding/lorentz_model.py:122-138.  Differentiable in x and y."""
    if torch.is_grad_enabled() and (x.requires_grad or y.requires_grad):
        return _Distance.apply(x, y, _curv(c), _sem(semantics))
    return _distance_fwd(x, y, _curv(c), _sem(semantics))


def _batch_distance_fwd(x: torch.Tensor, y: torch.Tensor, c: float, sem: int) -> torch.Tensor:
    dev = require_cuda(x, y)
    _lib.check_device(dev)
    if x.dim() != 2 or y.dim() != 2 or x.shape[1] != y.shape[1]:
        raise ValueError("batch_distance expects (n1, D) and (n2, D)")
    xr, yr = x.detach().contiguous(), y.detach().contiguous()
    n1, D = xr.shape
    n2 = yr.shape[0]
    out = torch.empty((n1, n2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().hyp_batch_distance(ptr(xr), D, n1, ptr(yr), D, n2, ptr(out), n2, D, c, sem, stream_ptr()))
    return out


class _BatchDistance(torch.autograd.Function):
    """batch_distance with autograd: W = g * d(dist)/d<x_i,y_j> from the tile kernel, then the row gradients are
    two library GEMMs against the sign-flipped operands (d<x,y>/dx = (y0, -ys))."""

    @staticmethod
    def forward(ctx, x, y, c, sem):
        ctx.save_for_backward(x, y)
        ctx.c, ctx.sem = c, sem
        return _batch_distance_fwd(x, y, c, sem)

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        xr, yr = x.detach().contiguous(), y.detach().contiguous()
        n1, D = xr.shape
        n2 = yr.shape[0]
        gr = g.detach().to(torch.float32).contiguous()
        w = torch.empty((n1, n2), dtype=torch.float32, device=xr.device)
        if n1 and n2:
            with torch.cuda.device(xr.device):
                check(_lib.lib().hyp_batch_distance_backward_coef(ptr(xr), D, n1, ptr(yr), D, n2, ptr(gr), n2, ptr(w), n2,
                                                                  D, ctx.c, ctx.sem, stream_ptr()))
        sig = torch.full((D,), -1.0, dtype=torch.float32, device=xr.device)
        sig[0] = 1.0
        gx = (w @ (yr * sig)) if ctx.needs_input_grad[0] else None
        gy = (w.t() @ (xr * sig)) if ctx.needs_input_grad[1] else None
        return gx, gy, None, None


def batch_distance(x: torch.Tensor, y: torch.Tensor, c: float = 1.0, semantics: Optional[str] = None) -> torch.Tensor:
    """reference embedding/lorentz_model.py:141-178: (n1, D) x (n2, D) -> (n1, n2), without the
    (n1, n2, d) temporary.  Differentiable in x and y."""
    if torch.is_grad_enabled() and (x.requires_grad or y.requires_grad):
        return _BatchDistance.apply(x, y, _curv(c), _sem(semantics))
    return _batch_distance_fwd(x, y, _curv(c), _sem(semantics))


def batch_distance_optimized(x: torch.Tensor, y: torch.Tensor, c: float = 1.0,
                             semantics: Optional[str] = None) -> torch.Tensor:
    """reference embedding/lorentz_model.py:181-210 (dead code there); same kernel here."""
    return batch_distance(x, y, c, semantics)


def parallel_transport(v: torch.Tensor, x: torch.Tensor, y: torch.Tensor, c: float = 1.0) -> torch.Tensor:
    """reference embedding/lorentz_model.py:213-228 (off the hot path; formula kept as shipped)."""
    xy = -minkowski_dot(x, y).unsqueeze(-1)
    coef = minkowski_dot(y, v).unsqueeze(-1) / (1 - xy)
    return v + coef * (x + y)


def riemannian_gradient(euclidean_grad: torch.Tensor, x: torch.Tensor, c: float = 1.0) -> torch.Tensor:
    """reference embedding/lorentz_model.py:231-244 (off the hot path)."""
    return euclidean_grad + minkowski_dot(x, euclidean_grad).unsqueeze(-1) * x
