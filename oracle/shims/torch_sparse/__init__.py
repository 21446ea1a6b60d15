"""Pure-torch stand-in for the `torch_sparse` wheel (TEST INFRASTRUCTURE ONLY).

The reference uses `torch_sparse.SparseTensor`, `matmul`, `fill_diag`, `sum`
and `mul` (`/root/reference/experiments/optimized_layers.py:9,16,173,254,269-276`,
`/root/reference/experiments/layers.py:8,225`,
`/root/reference/experiments/utils.py:9,107-113`).  The wheel is not installable
offline; this module restates the published semantics of torch-sparse 0.6.13
 so the reference source can be imported verbatim:

* storage is COO sorted by (row, col) + a lazily built ``rowptr`` (CSR) and the
  ``csr2csc`` permutation; duplicates are kept;
* ``matmul(A, x, reduce)``: every nnz contributes ``value * x[col]`` (``x[col]``
  when the matrix has no values); ``mean`` divides by the row's nnz count
  (min 1), NOT by the value sum; ``min``/``max``: empty row -> 0, first nnz wins
  ties, gradient goes to that single nnz;
* ``fill_diag(A, v)``: drop every diagonal entry, insert one per ``i <
  min(rows, cols)``, keep the matrix sorted; values only if the matrix had them.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from torch_scatter import _ScatterArg  # first-wins arg-extremum with single-element gradient

__all__ = ["SparseTensor", "matmul", "fill_diag", "sum", "mul", "set_diag", "remove_diag"]


class _Storage:
    """The subset of `SparseStorage` accessors the reference touches."""

    def __init__(self, owner: "SparseTensor"):
        self._o = owner

    def row(self) -> Tensor:
        return self._o._row

    def col(self) -> Tensor:
        return self._o._col

    def value(self) -> Optional[Tensor]:
        return self._o._value

    def rowptr(self) -> Tensor:
        return self._o._get_rowptr()

    def rowcount(self) -> Tensor:
        rp = self._o._get_rowptr()
        return rp[1:] - rp[:-1]

    def csr2csc(self) -> Tensor:
        return self._o._get_csr2csc()

    def colptr(self) -> Tensor:
        return self._o._get_colptr()

    def sparse_sizes(self) -> Tuple[int, int]:
        return self._o._sizes


class SparseTensor:
    def __init__(self, row: Optional[Tensor] = None, rowptr: Optional[Tensor] = None,
                 col: Optional[Tensor] = None, value: Optional[Tensor] = None,
                 sparse_sizes: Optional[Tuple[int, int]] = None, is_sorted: bool = False):
        assert col is not None
        if row is None:
            assert rowptr is not None
            counts = rowptr[1:] - rowptr[:-1]
            row = torch.repeat_interleave(torch.arange(counts.numel(), device=col.device), counts)
            is_sorted = True
        if sparse_sizes is None:
            m = int(row.max()) + 1 if row.numel() else 0
            n = int(col.max()) + 1 if col.numel() else 0
            sparse_sizes = (m, n)
        m, n = int(sparse_sizes[0]), int(sparse_sizes[1])
        if not is_sorted and row.numel() > 0:
            key = row * n + col
            if not bool((key[1:] >= key[:-1]).all()):
                perm = torch.argsort(key, stable=True)
                row, col = row[perm], col[perm]
                if value is not None:
                    value = value[perm]
        self._row, self._col, self._value = row, col, value
        self._sizes = (m, n)
        self._rowptr = rowptr if (rowptr is not None and rowptr.numel() == m + 1) else None
        self._csr2csc: Optional[Tensor] = None
        self._colptr: Optional[Tensor] = None
        self.storage = _Storage(self)

    # ---- constructors --------------------------------------------------------------
    @classmethod
    def from_edge_index(cls, edge_index: Tensor, edge_attr: Optional[Tensor] = None,
                        sparse_sizes: Optional[Tuple[int, int]] = None, is_sorted: bool = False):
        return cls(row=edge_index[0], col=edge_index[1], value=edge_attr,
                   sparse_sizes=sparse_sizes, is_sorted=is_sorted)

    # ---- lazily cached structure ---------------------------------------------------
    def _get_rowptr(self) -> Tensor:
        if self._rowptr is None:
            counts = torch.bincount(self._row, minlength=self._sizes[0])
            rp = torch.zeros(self._sizes[0] + 1, dtype=torch.long, device=self._row.device)
            torch.cumsum(counts, 0, out=rp[1:])
            self._rowptr = rp
        return self._rowptr

    def _get_csr2csc(self) -> Tensor:
        if self._csr2csc is None:
            self._csr2csc = torch.argsort(self._col * self._sizes[0] + self._row, stable=True)
        return self._csr2csc

    def _get_colptr(self) -> Tensor:
        if self._colptr is None:
            counts = torch.bincount(self._col, minlength=self._sizes[1])
            cp = torch.zeros(self._sizes[1] + 1, dtype=torch.long, device=self._col.device)
            torch.cumsum(counts, 0, out=cp[1:])
            self._colptr = cp
        return self._colptr

    # ---- accessors -----------------------------------------------------------------
    def coo(self):
        return self._row, self._col, self._value

    def csr(self):
        return self._get_rowptr(), self._col, self._value

    def csc(self):
        perm = self._get_csr2csc()
        v = self._value[perm] if self._value is not None else None
        return self._get_colptr(), self._row[perm], v

    def has_value(self) -> bool:
        return self._value is not None

    def sparse_sizes(self) -> Tuple[int, int]:
        return self._sizes

    def sparse_size(self, dim: int) -> int:
        return self._sizes[dim]

    def size(self, dim: int) -> int:
        return self._sizes[dim]

    def sizes(self):
        return list(self._sizes)

    def nnz(self) -> int:
        return int(self._col.numel())

    def device(self):
        return self._col.device

    def dim(self) -> int:
        return 2

    # ---- functional updates (never mutate) -----------------------------------------
    def _like(self, row, col, value, sorted_=True) -> "SparseTensor":
        return SparseTensor(row=row, col=col, value=value, sparse_sizes=self._sizes, is_sorted=sorted_)

    def set_value(self, value: Optional[Tensor], layout: Optional[str] = None) -> "SparseTensor":
        out = self._like(self._row, self._col, value)
        out._rowptr, out._csr2csc, out._colptr = self._rowptr, self._csr2csc, self._colptr
        return out

    def fill_value(self, fill_value: float, dtype=None) -> "SparseTensor":
        v = torch.full((self.nnz(),), fill_value, dtype=dtype or torch.get_default_dtype(),
                       device=self._col.device)
        return self.set_value(v)

    def to(self, *args, **kwargs) -> "SparseTensor":
        idx_kwargs = {k: v for k, v in kwargs.items() if k == "device"}
        dev = [a for a in args if isinstance(a, (str, torch.device))]
        row = self._row.to(*dev, **idx_kwargs)
        col = self._col.to(*dev, **idx_kwargs)
        val = self._value.to(*args, **kwargs) if self._value is not None else None
        return self._like(row, col, val)

    def t(self) -> "SparseTensor":
        perm = self._get_csr2csc()
        v = self._value[perm] if self._value is not None else None
        return SparseTensor(row=self._col[perm], col=self._row[perm], value=v,
                            sparse_sizes=(self._sizes[1], self._sizes[0]), is_sorted=True)

    def coalesce(self, reduce: str = "sum") -> "SparseTensor":
        key = self._row * self._sizes[1] + self._col
        uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
        row = torch.div(uniq, self._sizes[1], rounding_mode="floor")
        col = uniq - row * self._sizes[1]
        val = None
        if self._value is not None:
            val = torch.zeros((uniq.numel(),) + tuple(self._value.shape[1:]),
                              dtype=self._value.dtype, device=self._value.device)
            val.index_add_(0, inv, self._value)
        return self._like(row, col, val)

    def to_symmetric(self, reduce: str = "sum") -> "SparseTensor":
        row = torch.cat([self._row, self._col])
        col = torch.cat([self._col, self._row])
        val = torch.cat([self._value, self._value]) if self._value is not None else None
        n = max(self._sizes)
        return SparseTensor(row=row, col=col, value=val, sparse_sizes=(n, n)).coalesce(reduce)

    def to_dense(self) -> Tensor:
        v = self._value if self._value is not None else torch.ones(self.nnz())
        out = torch.zeros(self._sizes, dtype=v.dtype)
        out.index_put_((self._row, self._col), v, accumulate=True)
        return out

    # ---- ops -----------------------------------------------------------------------
    def matmul(self, other: Tensor, reduce: str = "sum") -> Tensor:
        return matmul(self, other, reduce)

    def __matmul__(self, other: Tensor) -> Tensor:
        return matmul(self, other, "sum")

    def sum(self, dim: Optional[int] = None) -> Tensor:
        return sum(self, dim)

    def mul(self, other: Tensor) -> "SparseTensor":
        return mul(self, other)

    def fill_diag(self, fill_value: float, k: int = 0) -> "SparseTensor":
        return fill_diag(self, fill_value, k)

    def __repr__(self) -> str:
        return f"SparseTensor(nnz={self.nnz()}, sparse_sizes={self._sizes}, has_value={self.has_value()})"


def matmul(src: SparseTensor, other: Tensor, reduce: str = "sum") -> Tensor:
    """CSR SpMM, rows = output index (torch-sparse `spmm`)."""
    row, col, value = src.coo()
    m = src.size(0)
    gathered = other.index_select(0, col)
    if value is not None:
        gathered = gathered * value.view([-1] + [1] * (other.dim() - 1)).to(gathered.dtype)
    if reduce in ("sum", "add", "mean"):
        out = other.new_zeros((m,) + tuple(other.shape[1:]))
        out = out.index_add(0, row, gathered)
        if reduce == "mean":
            cnt = src.storage.rowcount().clamp(min=1).to(other.dtype)
            out = out / cnt.view([-1] + [1] * (other.dim() - 1))
        return out
    if reduce in ("min", "max"):
        out, _ = _ScatterArg.apply(gathered, row, m, reduce == "max")
        return out
    raise ValueError(f"unknown reduce {reduce!r}")


def remove_diag(src: SparseTensor, k: int = 0) -> SparseTensor:
    row, col, value = src.coo()
    keep = row != (col - k)
    v = value[keep] if value is not None else None
    return SparseTensor(row=row[keep], col=col[keep], value=v, sparse_sizes=src.sparse_sizes(), is_sorted=True)


def fill_diag(src: SparseTensor, fill_value: float, k: int = 0) -> SparseTensor:
    assert k == 0, "oracle shim: only the main diagonal is restated"
    base = remove_diag(src)
    row, col, value = base.coo()
    m, n = src.sparse_sizes()
    d = torch.arange(min(m, n), dtype=row.dtype, device=row.device)
    new_row = torch.cat([row, d])
    new_col = torch.cat([col, d])
    new_val = None
    if value is not None:
        dv = torch.full((d.numel(),) + tuple(value.shape[1:]), fill_value, dtype=value.dtype, device=value.device)
        new_val = torch.cat([value, dv])
    # unique position of each diagonal entry inside its (already sorted) row
    perm = torch.argsort(new_row * n + new_col, stable=True)
    new_val = new_val[perm] if new_val is not None else None
    return SparseTensor(row=new_row[perm], col=new_col[perm], value=new_val, sparse_sizes=(m, n), is_sorted=True)


set_diag = fill_diag


def sum(src: SparseTensor, dim: Optional[int] = None) -> Tensor:  # noqa: A001 (mirrors torch_sparse.sum)
    row, col, value = src.coo()
    v = value if value is not None else torch.ones(src.nnz(), device=col.device)
    if dim is None:
        return v.sum()
    if dim < 0:
        dim += 2
    if dim == 1:
        return torch.zeros(src.size(0), dtype=v.dtype, device=v.device).index_add_(0, row, v)
    if dim == 0:
        return torch.zeros(src.size(1), dtype=v.dtype, device=v.device).index_add_(0, col, v)
    raise ValueError(dim)


def mul(src: SparseTensor, other: Tensor) -> SparseTensor:
    row, col, value = src.coo()
    if other.dim() == 2 and other.size(0) == src.size(0) and other.size(1) == 1:
        scale = other.view(-1)[row]      # row-wise broadcast
    elif other.dim() == 2 and other.size(0) == 1 and other.size(1) == src.size(1):
        scale = other.view(-1)[col]      # column-wise broadcast
    else:
        raise ValueError("oracle shim: mul expects an [M,1] or [1,N] dense operand")
    v = value * scale if value is not None else scale
    return src.set_value(v)
