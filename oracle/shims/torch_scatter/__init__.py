"""Pure-torch stand-in for the `torch_scatter` wheel (TEST INFRASTRUCTURE ONLY).

The reference layer (`/root/reference/experiments/optimized_layers.py:8,225-240`,
`/root/reference/experiments/layers.py:7,203-212`) calls
``torch_scatter.scatter(src, index, dim, out, dim_size, reduce)``.  The wheel is
not installable offline, so the oracle puts this directory on ``sys.path`` and
lets the reference source run unmodified on top of it.

Semantics restated from the published behaviour of torch-scatter 2.0.9
(SURVEY.md App. A-5):

* ``sum``  : zero-initialised add.
* ``mean`` : sum divided by ``clamp(count, min=1)``.
* ``min`` / ``max`` : value of the extremum; segments with no element yield 0;
  ties are broken in favour of the FIRST element in ``index`` order (strict
  comparison while scanning), and the gradient is routed to that single element
  (it is NOT split between ties, unlike ``torch.scatter_reduce``).

Only ``dim == 0`` (or the equivalent negative dim) is supported - that is the
only way the reference calls it (``node_dim = 0``).
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

__all__ = ["scatter", "scatter_add", "scatter_sum", "scatter_mean", "scatter_min", "scatter_max"]


def _check_dim(src: Tensor, dim: int) -> None:
    if dim < 0:
        dim += src.dim()
    if dim != 0:
        raise NotImplementedError("oracle shim: torch_scatter.scatter only restated for dim=0")


def _dim_size(index: Tensor, dim_size: Optional[int]) -> int:
    if dim_size is not None:
        return int(dim_size)
    return int(index.max()) + 1 if index.numel() > 0 else 0


def _expand_index(index: Tensor, src: Tensor) -> Tensor:
    # torch_scatter broadcasts a 1-D index over the trailing dims of src.
    if index.dim() == src.dim():
        return index.expand_as(src)
    view = [-1] + [1] * (src.dim() - 1)
    return index.view(view).expand_as(src)


def scatter_sum(src: Tensor, index: Tensor, dim: int = 0, out: Optional[Tensor] = None,
                dim_size: Optional[int] = None) -> Tensor:
    _check_dim(src, dim)
    n = _dim_size(index, dim_size)
    idx = _expand_index(index, src)
    if out is None:
        out = src.new_zeros((n,) + tuple(src.shape[1:]))
        return out.scatter_add(0, idx, src)
    return out.scatter_add_(0, idx, src)


scatter_add = scatter_sum


def scatter_mean(src: Tensor, index: Tensor, dim: int = 0, out: Optional[Tensor] = None,
                 dim_size: Optional[int] = None) -> Tensor:
    _check_dim(src, dim)
    n = _dim_size(index, dim_size)
    total = scatter_sum(src, index, 0, None, n)
    idx1 = index if index.dim() == 1 else index[(slice(None),) + (0,) * (index.dim() - 1)]
    count = torch.zeros(n, dtype=src.dtype, device=src.device)
    count.scatter_add_(0, idx1, torch.ones(src.shape[0], dtype=src.dtype, device=src.device))
    count = count.clamp_(min=1)
    count = count.view([-1] + [1] * (src.dim() - 1))
    if src.is_floating_point():
        return total / count
    return torch.div(total, count, rounding_mode="floor")


class _ScatterArg(torch.autograd.Function):
    """min/max with single, first-wins argument and single-element gradient."""

    @staticmethod
    def forward(ctx, src: Tensor, index: Tensor, n: int, is_max: bool):
        e = src.shape[0]
        flat = src.reshape(e, -1)
        f = flat.shape[1]
        idx = index.view(-1, 1).expand(e, f)
        red = "amax" if is_max else "amin"
        ext = flat.new_zeros((n, f)).scatter_reduce(0, idx, flat, reduce=red, include_self=False)
        # first position (in index order) attaining the extremum of its segment
        hit = flat == ext.gather(0, idx)
        pos = torch.arange(e, device=src.device).view(-1, 1).expand(e, f)
        cand = torch.where(hit, pos, torch.full_like(pos, e))
        arg = torch.full((n, f), e, dtype=torch.long, device=src.device)
        arg = arg.scatter_reduce(0, idx, cand, reduce="amin", include_self=True)
        out = torch.where(arg < e, ext, torch.zeros_like(ext))  # empty segment -> 0
        ctx.save_for_backward(arg)
        ctx.src_shape = src.shape
        ctx.mark_non_differentiable(arg)
        return out.view((n,) + tuple(src.shape[1:])), arg.view((n,) + tuple(src.shape[1:]))

    @staticmethod
    def backward(ctx, grad_out: Tensor, _grad_arg):
        (arg,) = ctx.saved_tensors
        e = ctx.src_shape[0]
        g = grad_out.reshape(arg.shape)
        grad_src = g.new_zeros((e + 1, arg.shape[1]))  # row e swallows empty segments
        grad_src.scatter_(0, arg, g)
        return grad_src[:e].view(ctx.src_shape), None, None, None


def scatter_min(src: Tensor, index: Tensor, dim: int = 0, out: Optional[Tensor] = None,
                dim_size: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    _check_dim(src, dim)
    assert out is None
    if index.dim() != 1:
        index = index[(slice(None),) + (0,) * (index.dim() - 1)]
    return _ScatterArg.apply(src, index, _dim_size(index, dim_size), False)


def scatter_max(src: Tensor, index: Tensor, dim: int = 0, out: Optional[Tensor] = None,
                dim_size: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    _check_dim(src, dim)
    assert out is None
    if index.dim() != 1:
        index = index[(slice(None),) + (0,) * (index.dim() - 1)]
    return _ScatterArg.apply(src, index, _dim_size(index, dim_size), True)


def scatter(src: Tensor, index: Tensor, dim: int = -1, out: Optional[Tensor] = None,
            dim_size: Optional[int] = None, reduce: str = "sum") -> Tensor:
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == "min":
        return scatter_min(src, index, dim, out, dim_size)[0]
    if reduce == "max":
        return scatter_max(src, index, dim, out, dim_size)[0]
    raise ValueError(f"unknown reduce {reduce!r}")
