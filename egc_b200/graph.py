"""Graph ingestion for the EGConv hot path: the reference's `edge_index` / `SparseTensor` inputs
become a device-resident int32 CSR (+ lazily the CSC used by backward), with PyG's self-loop and
`gcn_norm` semantics reproduced by the kernels in `csrc/graph_build.cu`.

Reference behaviour mirrored here (all in /root/reference/experiments/optimized_layers.py):
  * symnorm + Tensor input      -> gcn_norm(edge_index, num_nodes=N, add_self_loops)          :128-141
  * symnorm + SparseTensor      -> gcn_norm(adj_t): fill_value(1) / fill_diag / D^-1/2 A D^-1/2 :143-156
  * no symnorm, add_self_loops  -> add_remaining_self_loops(edge_index)  (no num_nodes!)      :159-166
                                   fill_diag(adj_t, 1.0)                                       :168-175
and `SparseTensor(row=, col=, value=, sparse_sizes=, is_sorted=)` as constructed by
/root/reference/experiments/utils.py:107-109.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import RowPlan, check, ptr


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ws(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


class SparseTensor:
    """Minimal stand-in for `torch_sparse.SparseTensor` (the container only - no compute).

    Stores a matrix sorted by (row, col); for the layer, rows are TARGET nodes (`adj_t`).
    Supports the constructor forms the reference uses plus `.csr()`, `.coo()`, `.sparse_sizes()`,
    `.set_value()`, `.to()`, `.t()`, `.to_symmetric()` and `.storage.rowptr()/.csr2csc()` no-ops.
    """

    def __init__(self, row: Optional[Tensor] = None, rowptr: Optional[Tensor] = None, col: Optional[Tensor] = None,
                 value: Optional[Tensor] = None, sparse_sizes: Optional[Tuple[int, int]] = None,
                 is_sorted: bool = False):
        if col is None:
            raise ValueError("SparseTensor needs `col`")
        if row is None:
            if rowptr is None:
                raise ValueError("SparseTensor needs `row` or `rowptr`")
            counts = rowptr[1:] - rowptr[:-1]
            row = torch.repeat_interleave(torch.arange(counts.numel(), device=col.device), counts)
            is_sorted = True
        row, col = row.long(), col.long()
        if sparse_sizes is None:
            m = int(row.max()) + 1 if row.numel() else 0
            n = int(col.max()) + 1 if col.numel() else 0
            sparse_sizes = (m, n)
        m, n = int(sparse_sizes[0]), int(sparse_sizes[1])
        if not is_sorted and row.numel() > 0:
            perm = torch.argsort(row * n + col, stable=True)
            row, col = row[perm], col[perm]
            value = value[perm] if value is not None else None
        self._row, self._col, self._value, self._sizes = row, col, value, (m, n)
        self._rowptr = rowptr.long() if (rowptr is not None and rowptr.numel() == m + 1) else None
        self.storage = self     # `adj_t.storage.rowptr()` / `.csr2csc()` cache warm-ups are accepted

    # -- structure -------------------------------------------------------------------------
    def rowptr(self) -> Tensor:
        if self._rowptr is None:
            counts = torch.bincount(self._row, minlength=self._sizes[0])
            self._rowptr = torch.cat([counts.new_zeros(1), torch.cumsum(counts, 0)])
        return self._rowptr

    def csr2csc(self) -> Tensor:
        return torch.argsort(self._col * self._sizes[0] + self._row, stable=True)

    def csr(self):
        return self.rowptr(), self._col, self._value

    def coo(self):
        return self._row, self._col, self._value

    def has_value(self) -> bool:
        return self._value is not None

    def sparse_sizes(self) -> Tuple[int, int]:
        return self._sizes

    def sparse_size(self, dim: int) -> int:
        return self._sizes[dim]

    size = sparse_size

    def nnz(self) -> int:
        return int(self._col.numel())

    @property
    def device(self):
        return self._col.device

    # -- functional updates ----------------------------------------------------------------
    def set_value(self, value: Optional[Tensor], layout: Optional[str] = None) -> "SparseTensor":
        out = SparseTensor(row=self._row, col=self._col, value=value, sparse_sizes=self._sizes, is_sorted=True)
        out._rowptr = self._rowptr
        return out

    def to(self, device) -> "SparseTensor":
        v = self._value.to(device) if self._value is not None else None
        out = SparseTensor(row=self._row.to(device), col=self._col.to(device), value=v, sparse_sizes=self._sizes,
                           is_sorted=True)
        out._rowptr = self._rowptr.to(device) if self._rowptr is not None else None
        return out

    def cuda(self) -> "SparseTensor":
        return self.to("cuda")

    def t(self) -> "SparseTensor":
        perm = self.csr2csc()
        v = self._value[perm] if self._value is not None else None
        return SparseTensor(row=self._col[perm], col=self._row[perm], value=v,
                            sparse_sizes=(self._sizes[1], self._sizes[0]), is_sorted=True)

    def to_symmetric(self) -> "SparseTensor":
        """A | A^T with duplicate entries merged (values summed), as `adj_t.to_symmetric()` does in
        /root/reference/experiments/mag/configs.py:85."""
        n = max(self._sizes)
        row = torch.cat([self._row, self._col])
        col = torch.cat([self._col, self._row])
        key, inv = torch.unique(row * n + col, sorted=True, return_inverse=True)
        val = None
        if self._value is not None:
            v2 = torch.cat([self._value, self._value])
            val = torch.zeros(key.numel(), dtype=v2.dtype, device=v2.device).index_add_(0, inv, v2)
        r = torch.div(key, n, rounding_mode="floor")
        return SparseTensor(row=r, col=key - r * n, value=val, sparse_sizes=(n, n), is_sorted=True)

    def __repr__(self) -> str:
        return f"SparseTensor(nnz={self.nnz()}, sparse_sizes={self._sizes}, has_value={self.has_value()})"


def to_sparse_tensor(edge_index: Tensor, num_nodes: int) -> SparseTensor:
    """`adj_t` of an edge list as the reference's `ToSparseTensor` transform builds it
    (/root/reference/experiments/utils.py:89-115): rows = targets, entries sorted by (target, source), duplicates kept,
    edge attributes dropped, `rowptr` pre-computed.  Runs on whatever device `edge_index` lives on."""
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("edge_index must have shape [2, E]")
    n = int(num_nodes)
    src, dst = edge_index[0].long(), edge_index[1].long()
    perm = torch.argsort(dst * n + src, stable=True)                      # ref :93 `(col * N + row).argsort()`
    adj_t = SparseTensor(row=dst[perm], col=src[perm], value=None, sparse_sizes=(n, n), is_sorted=True)   # ref :107-109
    adj_t.rowptr()                                                        # ref :111-113 fill_cache
    return adj_t


class _Plan:
    """Device arrays + the ctypes view of an `egc_row_plan`."""

    def __init__(self, ptr_array: Tensor, n_rows: int, n_long: int, n_chunks: int):
        dev = ptr_array.device
        self.n_long, self.n_chunks = n_long, n_chunks
        self.struct = RowPlan(n_long, n_chunks, None, None, None, None)
        if n_long > 0:
            lib = _lib.load()
            self.long_rows = torch.empty(n_long, dtype=torch.int32, device=dev)
            self.long_chunk_ptr = torch.empty(n_long + 1, dtype=torch.int32, device=dev)
            self.chunk_row = torch.empty(n_chunks, dtype=torch.int32, device=dev)
            self.chunk_begin = torch.empty(n_chunks, dtype=torch.int32, device=dev)
            nbytes = lib.egc_plan_build_workspace_bytes(n_rows)
            ws = _ws(nbytes, dev)
            check(lib.egc_plan_build(ptr(ptr_array), n_rows, n_long, n_chunks, ptr(self.long_rows),
                                     ptr(self.long_chunk_ptr), ptr(self.chunk_row), ptr(self.chunk_begin),
                                     ptr(ws), nbytes, _stream()), "egc_plan_build")
            self.struct = RowPlan(n_long, n_chunks, self.long_rows.data_ptr(), self.long_chunk_ptr.data_ptr(),
                                  self.chunk_row.data_ptr(), self.chunk_begin.data_ptr())


def _raise_on_flags(flags: int) -> None:
    if flags & 1:
        raise IndexError("egc_b200: node index out of range in the graph input")
    if flags & 2:
        raise ValueError("egc_b200: CSR rows must be sorted by column (SparseTensor invariant)")


class GraphStructure:
    """Prepared, device-resident graph for one EGConv configuration (what the reference caches in
    `_cached_edge_index` / `_cached_adj_t`, optimized_layers.py:71-72)."""

    def __init__(self):
        self.n_dst = self.n_src = self.nnz = 0
        self.max_deg = 0
        self.rowptr = self.col = None
        self.val_sym = self.val_lin = self.deg = self.dis = None
        self.plan: Optional[_Plan] = None
        # CSC side (built on first backward)
        self.colptr = self.rowidx = self.csr2csc = None
        self.csc_val_sym = self.csc_val_lin = None
        self.csc_plan: Optional[_Plan] = None

    @property
    def device(self):
        return self.rowptr.device

    # -- builders --------------------------------------------------------------------------
    @staticmethod
    def _empty(n_dst: int, n_src: int, dev) -> "GraphStructure":
        """A graph without target rows (zero-node batch, a node type without nodes): no kernel ever runs on it."""
        g = GraphStructure()
        g.n_dst, g.n_src = n_dst, n_src
        g.rowptr = torch.zeros(n_dst + 1, dtype=torch.int32, device=dev)
        g.col = torch.zeros(0, dtype=torch.int32, device=dev)
        g.plan = _Plan(g.rowptr, n_dst, 0, 0)
        return g

    @staticmethod
    def from_edge_index(edge_index: Tensor, num_nodes: int, symnorm: bool, add_self_loops: bool,
                        expect: Optional[dict] = None) -> "GraphStructure":
        """`expect` (fixed-shape mini-batches, see `egc_b200.batch.pad_batch`): {"nnz": ..} - the caller KNOWS the size of
        the prepared graph (and that no row or column exceeds EGC_CHUNK_EDGES entries), so the one device-to-host read
        of graph preparation is skipped and the whole build can be captured in a CUDA graph; `verify()` compares the
        device-side counters with the expectation afterwards."""
        if not edge_index.is_cuda:
            raise RuntimeError("egc_b200: edge_index must live on a CUDA device (there is no CPU path)")
        if edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise ValueError("edge_index must have shape [2, E]")
        lib = _lib.load()
        ei = edge_index.long().contiguous()
        dev, n_edges = ei.device, ei.size(1)
        if num_nodes == 0:
            if n_edges:
                raise IndexError("egc_b200: node index out of range in the graph input")
            return GraphStructure._empty(0, 0, dev)
        loops = _lib.LOOPS_NONE if not add_self_loops else (_lib.LOOPS_ALL_NODES if symnorm else _lib.LOOPS_UP_TO_MAX_ID)
        g = GraphStructure()
        g.n_dst = g.n_src = int(num_nodes)
        g.rowptr = torch.empty(num_nodes + 1, dtype=torch.int32, device=dev)
        col_cap = torch.empty(n_edges + num_nodes, dtype=torch.int32, device=dev)
        meta = torch.empty(_lib.META_SLOTS, dtype=torch.int32, device=dev)
        nbytes = lib.egc_csr_from_edges_workspace_bytes(n_edges, num_nodes)
        ws = _ws(nbytes, dev)
        with torch.cuda.device(dev):
            check(lib.egc_csr_from_edges(ptr(ei[0]), ptr(ei[1]), n_edges, num_nodes, loops, ptr(g.rowptr), ptr(col_cap),
                                         ptr(meta), ptr(ws), nbytes, _stream()), "egc_csr_from_edges")
            g._finish(meta, col_cap, None, symnorm, keep_values=False, expect=expect)
        return g

    @staticmethod
    def from_csr(rowptr: Tensor, col: Tensor, value: Optional[Tensor], n_src: int, symnorm: bool,
                 add_self_loops: bool) -> "GraphStructure":
        if not col.is_cuda:
            raise RuntimeError("egc_b200: the adjacency must live on a CUDA device (there is no CPU path)")
        lib = _lib.load()
        dev = col.device
        rowptr64, col64 = rowptr.to(dev).long().contiguous(), col.long().contiguous()
        val_in = value.to(dev, torch.float32).contiguous() if value is not None else None
        n_dst, nnz_in = rowptr64.numel() - 1, col64.numel()
        if n_dst == 0:
            return GraphStructure._empty(0, int(n_src), dev)
        cap = nnz_in + (min(n_dst, n_src) if add_self_loops else 0)
        g = GraphStructure()
        g.n_dst, g.n_src = int(n_dst), int(n_src)
        g.rowptr = torch.empty(n_dst + 1, dtype=torch.int32, device=dev)
        col_cap = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
        val_cap = torch.empty(max(cap, 1), dtype=torch.float32, device=dev) if val_in is not None else None
        meta = torch.empty(_lib.META_SLOTS, dtype=torch.int32, device=dev)
        nbytes = lib.egc_csr_fill_diag_workspace_bytes(n_dst)
        ws = _ws(nbytes, dev)
        with torch.cuda.device(dev):
            check(lib.egc_csr_fill_diag(ptr(rowptr64), ptr(col64), ptr(val_in), n_dst, n_src, int(add_self_loops),
                                        ptr(g.rowptr), ptr(col_cap), ptr(val_cap), ptr(meta), ptr(ws), nbytes,
                                        _stream()), "egc_csr_fill_diag")
            g._finish(meta, col_cap, val_cap, symnorm, keep_values=not symnorm)
        return g

    @staticmethod
    def from_prepared(rowptr: Tensor, col: Tensor, n_src: int, val_sym: Optional[Tensor] = None,
                      val_lin: Optional[Tensor] = None, device=None) -> "GraphStructure":
        """Adopt an already prepared CSR (self-loops and weights decided elsewhere, e.g. the local
        block of a row-partitioned graph whose column ids live in the [own | halo] space)."""
        dev = torch.device(device) if device is not None else col.device
        if dev.type != "cuda":
            raise RuntimeError("egc_b200: graphs live on a CUDA device (there is no CPU path)")
        lib = _lib.load()
        rowptr64, col64 = rowptr.to(dev).long().contiguous(), col.to(dev).long().contiguous()
        n_dst, nnz = rowptr64.numel() - 1, col64.numel()
        g = GraphStructure()
        g.n_dst, g.n_src = int(n_dst), int(n_src)
        g.rowptr = torch.empty(n_dst + 1, dtype=torch.int32, device=dev)
        col_cap = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        meta = torch.empty(_lib.META_SLOTS, dtype=torch.int32, device=dev)
        nbytes = lib.egc_csr_fill_diag_workspace_bytes(n_dst)
        ws = _ws(nbytes, dev)
        with torch.cuda.device(dev):
            check(lib.egc_csr_fill_diag(ptr(rowptr64), ptr(col64), None, n_dst, n_src, 0, ptr(g.rowptr), ptr(col_cap),
                                        None, ptr(meta), ptr(ws), nbytes, _stream()), "egc_csr_fill_diag")
            g._finish(meta, col_cap, None, symnorm=False, keep_values=False)
        g.val_sym = val_sym.to(dev, torch.float32).contiguous() if val_sym is not None else None
        g.val_lin = val_lin.to(dev, torch.float32).contiguous() if val_lin is not None else None
        return g

    def _finish(self, meta: Tensor, col_cap: Tensor, val_cap: Optional[Tensor], symnorm: bool, keep_values: bool,
                expect: Optional[dict] = None):
        lib = _lib.load()
        if expect is None:
            m = meta.cpu().tolist()                # the one host sync of graph preparation
            _raise_on_flags(m[_lib.META_ERRFLAGS])
        else:                                      # trusted shape: no host read, checked later by verify()
            m = [0] * _lib.META_SLOTS
            m[_lib.META_NNZ] = int(expect["nnz"])
            self._expected = [(meta, {_lib.META_NNZ: int(expect["nnz"]), _lib.META_N_LONG: 0, _lib.META_N_CHUNKS: 0,
                                      _lib.META_ERRFLAGS: 0}, "CSR")]
        self.nnz, self.max_deg = m[_lib.META_NNZ], m[_lib.META_MAX_DEG]
        self.n_loops = m[_lib.META_N_LOOPS]
        self.col = col_cap[: self.nnz]
        values = val_cap[: self.nnz] if val_cap is not None else None
        dev = self.rowptr.device
        if symnorm:
            if self.n_dst != self.n_src:
                raise ValueError("symnorm needs a square adjacency")
            self.deg = torch.empty(self.n_dst, dtype=torch.float32, device=dev)
            self.dis = torch.empty(self.n_dst, dtype=torch.float32, device=dev)
            self.val_sym = torch.empty(max(self.nnz, 1), dtype=torch.float32, device=dev)[: self.nnz]
            check(lib.egc_symnorm_weights(ptr(self.rowptr), ptr(self.col), ptr(values), self.n_dst, ptr(self.deg),
                                          ptr(self.dis), ptr(self.val_sym), _stream()), "egc_symnorm_weights")
        if keep_values and values is not None:
            self.val_lin = values
        self.plan = _Plan(self.rowptr, self.n_dst, m[_lib.META_N_LONG], m[_lib.META_N_CHUNKS])

    def verify(self) -> None:
        """Fixed-shape graphs (`expect=`): compare the counters the build kernels left on the device with what the caller
        promised (synchronises).  Raises if the graph has a different nnz, an id out of range or a row / column too
        long for the plan-free kernels - the results computed from it are then invalid."""
        for meta, want, what in getattr(self, "_expected", []):
            got = meta.cpu().tolist()
            _raise_on_flags(got[_lib.META_ERRFLAGS])
            for k, v in want.items():
                if got[k] != v:
                    raise ValueError(f"egc_b200: fixed-shape graph does not match its declared shape ({what}: counter {k} is "
                                     f"{got[k]}, declared {v}) - pad the batch with egc_b200.pad_batch or drop `expect`")

    def ensure_csc(self) -> None:
        """CSC view + CSC-ordered weights + its chunk plan (backward only; cached)."""
        if self.colptr is not None:
            return
        lib = _lib.load()
        dev = self.device
        with torch.cuda.device(dev):
            self.colptr = torch.empty(self.n_src + 1, dtype=torch.int32, device=dev)
            self.rowidx = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)[: self.nnz]
            self.csr2csc = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)[: self.nnz]
            meta = torch.empty(_lib.META_SLOTS, dtype=torch.int32, device=dev)
            nbytes = lib.egc_csr_transpose_workspace_bytes(self.nnz, self.n_dst, self.n_src)
            ws = _ws(nbytes, dev)
            check(lib.egc_csr_transpose(ptr(self.rowptr), ptr(self.col), self.nnz, self.n_dst, self.n_src,
                                        ptr(self.colptr), ptr(self.rowidx), ptr(self.csr2csc), ptr(meta), ptr(ws),
                                        nbytes, _stream()), "egc_csr_transpose")
            for name in ("val_sym", "val_lin"):
                v = getattr(self, name)
                if v is not None and self.nnz > 0:
                    out = torch.empty_like(v)
                    check(lib.egc_permute_f32(ptr(v), ptr(self.csr2csc), self.nnz, ptr(out), _stream()),
                          "egc_permute_f32")
                    setattr(self, "csc_" + name, out)
            if getattr(self, "_expected", None):       # fixed-shape graph: no long columns by declaration, no host read
                self._expected.append((meta, {_lib.META_N_LONG: 0, _lib.META_N_CHUNKS: 0}, "CSC"))
                m = [0] * _lib.META_SLOTS
            else:
                m = meta.cpu().tolist()
            self.csc_plan = _Plan(self.colptr, self.n_src, m[_lib.META_N_LONG], m[_lib.META_N_CHUNKS])

    def source_ids(self, arg: Tensor) -> Tensor:
        """Map nnz positions (egc_aggregate_fwd `arg_out`) to source node ids (-1 stays -1)."""
        safe = arg.clamp(min=0).long()
        return torch.where(arg >= 0, self.col.long()[safe], torch.full_like(safe, -1))


def adjacency_to_csr(adj) -> Tuple[Tensor, Tensor, Optional[Tensor], int]:
    """(rowptr, col, value|None, n_src) from any supported SparseTensor-like object."""
    if isinstance(adj, SparseTensor):
        rowptr, col, value = adj.csr()
        return rowptr, col, value, adj.sparse_size(1)
    if isinstance(adj, Tensor) and adj.layout == torch.sparse_csr:
        return adj.crow_indices(), adj.col_indices(), adj.values(), adj.size(1)
    if hasattr(adj, "csr") and hasattr(adj, "sparse_sizes"):        # a real torch_sparse.SparseTensor
        rowptr, col, value = adj.csr()
        return rowptr, col, value, adj.sparse_sizes()[1]
    raise TypeError(f"unsupported adjacency type {type(adj)!r}")
